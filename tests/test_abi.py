"""CPU: the C-ABI shared library loads, exports every symbol include/ibgs_b200.h declares, and its struct
layouts agree with the ctypes mirror.  No compute entry point is exercised here (no GPU)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ibgs_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ibgs_[a-z0-9_]+)\s*\(", src)) - {"ibgs_alloc_fn"})


def test_library_exports_every_declared_symbol():
    from ibgs_b200 import _native as N
    names = _declared_functions()
    assert len(names) >= 24
    for n in names:
        assert hasattr(N.lib, n), f"{n} declared in include/ibgs_b200.h but not exported"
    assert sorted(N.EXPORTS) == names
    assert N.lib.ibgs_abi_version() == 5


def test_struct_layouts_match_header(tmp_path):
    from ibgs_b200 import _native as N
    probes = [("sizeof(IbgsView)", C.sizeof(N.IbgsView)), ("sizeof(IbgsForwardArgs)", C.sizeof(N.IbgsForwardArgs)),
              ("sizeof(IbgsBackwardArgs)", C.sizeof(N.IbgsBackwardArgs)),
              ("offsetof(IbgsForwardArgs, alloc)", N.IbgsForwardArgs.alloc.offset),
              ("offsetof(IbgsBackwardArgs, dL_dmeans3D)", N.IbgsBackwardArgs.dL_dmeans3D.offset),
              ("offsetof(IbgsView, bg)", N.IbgsView.bg.offset),
              ("sizeof(IbgsPrologueArgs)", C.sizeof(N.IbgsPrologueArgs)),
              ("offsetof(IbgsPrologueArgs, d_offset)", N.IbgsPrologueArgs.d_offset.offset),
              ("sizeof(IbgsDepthBatchArgs)", C.sizeof(N.IbgsDepthBatchArgs)),
              ("offsetof(IbgsDepthBatchArgs, viewmatrices)", N.IbgsDepthBatchArgs.viewmatrices.offset),
              ("offsetof(IbgsDepthBatchArgs, num_rendered)", N.IbgsDepthBatchArgs.num_rendered.offset),
              ("offsetof(IbgsDepthBatchArgs, alloc_user)", N.IbgsDepthBatchArgs.alloc_user.offset),
              ("sizeof(IbgsSsimArgs)", C.sizeof(N.IbgsSsimArgs)),
              ("offsetof(IbgsSsimArgs, dL_dmap_scale)", N.IbgsSsimArgs.dL_dmap_scale.offset),
              ("offsetof(IbgsSsimArgs, dL_dimg2)", N.IbgsSsimArgs.dL_dimg2.offset),
              ("sizeof(IbgsAdamGroup)", C.sizeof(N.IbgsAdamGroup)), ("sizeof(IbgsAdamArgs)", C.sizeof(N.IbgsAdamArgs)),
              ("offsetof(IbgsAdamArgs, groups)", N.IbgsAdamArgs.groups.offset),
              ("offsetof(IbgsAdamArgs, step)", N.IbgsAdamArgs.step.offset),
              ("offsetof(IbgsAdamArgs, zero_grads)", N.IbgsAdamArgs.zero_grads.offset),
              ("sizeof(IbgsColorFeatArgs)", C.sizeof(N.IbgsColorFeatArgs)),
              ("offsetof(IbgsColorFeatArgs, warped)", N.IbgsColorFeatArgs.warped.offset),
              ("offsetof(IbgsColorFeatArgs, d_b2)", N.IbgsColorFeatArgs.d_b2.offset),
              ("IBGS_MAX_DEPTH_BATCH", N.MAX_DEPTH_BATCH), ("IBGS_ADAM_MAX_GROUPS", N.ADAM_MAX_GROUPS),
              ("IBGS_MAX_SRC", N.MAX_SRC), ("IBGS_MAX_BUFFER_LENGTH", N.MAX_BUFFER_LENGTH)]
    prog = tmp_path / "sz.c"
    body = "".join(f'printf("%zu\\n", (size_t)({expr}));' for expr, _ in probes)
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ibgs_b200.h"\nint main(){' + body + 'return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    for (expr, want), g in zip(probes, got):
        assert g == want, (expr, g, want)


def test_sort_bits_follow_getHigherMsb():
    # 256^2 -> 41, 1080p -> 45, 1237x822 -> 44, 4K -> 47 (rasterizer_impl.cu:152-167,449)
    from ibgs_b200 import _native as N
    for (w, h), bits in (((256, 256), 41), ((1920, 1080), 45), ((1237, 822), 44), ((3840, 2160), 47)):
        tiles = ((w + 15) // 16) * ((h + 15) // 16)
        assert N.lib.ibgs_sort_bits(tiles) == bits


def test_state_layout_is_aligned_and_ordered():
    from ibgs_b200 import _native as N
    for which, count, aux in ((N.IBGS_BUF_GEOM, 12345, 0), (N.IBGS_BUF_IMAGE, 1920 * 1080, 8160),
                              (N.IBGS_BUF_BINNING, 99999, 0), (N.IBGS_BUF_SCRATCH, 99999, 0)):
        offs, total = N.state_layout(which, count, aux)
        assert all(o % 256 == 0 for o in offs) and offs == sorted(offs) and total >= offs[-1]
    offs, total = N.state_layout(N.IBGS_BUF_GEOM, 1000)
    assert offs[1] - offs[0] == 64 * 1000   # one 64-byte record per Gaussian


def test_argument_errors_without_gpu():
    from ibgs_b200 import _native as N
    assert N.lib.ibgs_forward(None, None) == -1 and "NULL" in N.last_error()
    a = N.IbgsForwardArgs()
    a.P = -5
    assert N.lib.ibgs_forward(C.byref(a), None) == -1
    a.P = 0
    assert N.lib.ibgs_forward(C.byref(a), None) == 0          # P == 0 -> nothing rendered (rasterize_points.cu:101)
    b = N.IbgsBackwardArgs()
    b.P = 0
    assert N.lib.ibgs_backward(C.byref(b), None) == 0
    assert N.lib.ibgs_mark_visible(0, None, None, None, None, None) == 0
    assert N.lib.ibgs_dist2(0, None, None, None, 0, None) == 0
    assert N.lib.ibgs_dist2(-1, None, None, None, 0, None) == -1
    assert N.lib.ibgs_dist2_scratch_bytes(1000) > 0
    with pytest.raises(RuntimeError):
        N.check(-1, "x")


def test_missing_library_fails_loudly(tmp_path):
    # the product path must not fall back to anything when the CUDA library is absent
    code = ("import sys, importlib; sys.path.insert(0, %r)\n"
            "import ibgs_b200._native as N\n" % ROOT)
    env = dict(os.environ)
    script = tmp_path / "t.py"
    script.write_text("import os, sys\nsys.path.insert(0, %r)\nimport ibgs_b200\n"
                      "import ibgs_b200.build as b\n"
                      "import importlib.util\n"
                      "spec = importlib.util.spec_from_file_location('nat', os.path.join(%r, 'ibgs_b200', '_native.py'))\n"
                      "src = open(spec.origin).read().replace('\"_lib\", \"libibgs_b200.so\"', '\"_lib\", \"nope.so\"')\n"
                      "ns = {'__name__': 'nat', '__file__': spec.origin}\n"
                      "try:\n    exec(compile(src, spec.origin, 'exec'), ns)\n    print('LOADED')\n"
                      "except ImportError as e:\n    print('IMPORTERROR', e)\n" % (ROOT, ROOT))
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, env=env).stdout
    assert "IMPORTERROR" in out and "no CPU" in out
