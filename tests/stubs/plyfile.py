"""Stand-in for `plyfile` (PLY I/O of scene/gaussian_model.py:278-345 and scene/dataset_readers.py:124-146): the
synthetic-scene harness never reads or writes PLY files."""


class _Unavailable:
    def __init__(self, *a, **k):
        raise RuntimeError("plyfile is not installed in this image (test stub)")

    @classmethod
    def describe(cls, *a, **k):
        raise RuntimeError("plyfile is not installed in this image (test stub)")

    @classmethod
    def read(cls, *a, **k):
        raise RuntimeError("plyfile is not installed in this image (test stub)")


class PlyData(_Unavailable):
    pass


class PlyElement(_Unavailable):
    pass
