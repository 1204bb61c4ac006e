"""`simple_knn._C` replacement: distCUDA2 (reference submodules/simple-knn/spatial.cu:15-26, ext.cpp:15-17)."""
import ctypes as C

import torch

from .. import _native as N


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    """Mean squared distance of every point to its 3 nearest neighbours; points [P,3] CUDA -> float32 [P]."""
    if not points.is_cuda:
        raise RuntimeError("distCUDA2: points must be a CUDA tensor (there is no CPU path)")
    device = points.device
    pts = points.contiguous()
    if pts.dtype != torch.float32:
        pts = pts.float()
    P = pts.size(0)
    means = torch.zeros((P,), dtype=torch.float32, device=device)  # torch::full({P}, 0.0), spatial.cu:21
    if P == 0:
        return means
    nbytes = int(N.lib.ibgs_dist2_scratch_bytes(P))
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        N.check(N.lib.ibgs_dist2(P, pts.data_ptr(), means.data_ptr(), scratch.data_ptr(), nbytes,
                                 C.c_void_p(stream)), "ibgs_dist2")
    return means
