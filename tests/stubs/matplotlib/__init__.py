"""Stand-in for `matplotlib` (utils/general_utils.py:17 imports it for the `colorize` debugging helper only)."""


class _Cm:
    @staticmethod
    def get_cmap(*a, **k):
        raise RuntimeError("matplotlib is not installed in this image (test stub)")


cm = _Cm()
