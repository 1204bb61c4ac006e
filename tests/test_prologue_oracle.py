"""CPU: the numpy oracle of the parameter prologue (oracle/prologue_oracle.py) against torch autograd of the
reference's own expressions in float64 -- this pins the oracle the GPU test then uses as its checker."""
import numpy as np
import torch

import prologue_ref as PR
from oracle import prologue_oracle as O

NAMES = ("opacity", "scales", "rotations", "shs", "all_map")
IN = ("xyz", "opacity_raw", "scaling_raw", "rotation_raw", "fdc", "frest", "normal_raw", "offset")


def test_oracle_matches_torch_autograd_float64():
    p = PR.random_params(500, K=9, seed=1, dtype=torch.float64)
    leaves = {k: p[k].clone().requires_grad_(True) for k in IN}
    outs = PR.torch_prologue(*[leaves[k] for k in IN], p["V"], p["cam"])
    fw = O.forward(*[p[k].numpy() for k in IN], p["V"].numpy(), p["cam"].numpy())
    for n, o in zip(NAMES, outs):
        assert np.allclose(fw[n], o.detach().numpy(), rtol=1e-12, atol=1e-12), n
    g = torch.Generator().manual_seed(5)
    cots = [torch.randn(o.shape, generator=g, dtype=torch.float64) for o in outs]
    torch.autograd.backward(list(outs), cots)
    d = O.backward(fw, *[c.numpy() for c in cots])
    name = dict(fdc="features_dc", frest="features_rest")
    for k in IN:
        want = leaves[k].grad.numpy()
        got = d[name.get(k, k)].reshape(want.shape)
        assert np.allclose(got, want, rtol=1e-9, atol=1e-12), k


def test_oracle_flips_normals_towards_the_camera():
    p = PR.random_params(200, seed=2, dtype=torch.float64)
    fw = O.forward(*[p[k].numpy() for k in IN], p["V"].numpy(), p["cam"].numpy())
    c = fw["_cache"]
    assert ((c["ng"] * (p["cam"].numpy()[None] - p["xyz"].numpy())).sum(-1) >= 0).all()
    assert (c["sgn"] < 0).any() and (c["sgn"] > 0).any()
    assert (fw["all_map"][:, 4] >= 0).all() and (fw["all_map"][:, 3] == 1).all()


IN_SA = IN[:6]


def test_smallest_axis_oracle_matches_torch_autograd_float64():
    """learnt_normal=False: the oracle's hand-derived backward through quaternion_to_matrix / gather / flip against
    torch autograd of the reference's expressions (scene/gaussian_model.py:149-161)."""
    p = PR.random_params(400, K=4, seed=7, dtype=torch.float64)
    leaves = {k: p[k].clone().requires_grad_(True) for k in IN_SA}
    outs = PR.torch_prologue_smallest_axis(*[leaves[k] for k in IN_SA], p["V"], p["cam"])
    fw = O.forward_smallest_axis(*[p[k].numpy() for k in IN_SA], p["V"].numpy(), p["cam"].numpy())
    for n, o in zip(NAMES, outs):
        assert np.allclose(fw[n], o.detach().numpy(), rtol=1e-12, atol=1e-12), n
    assert len(set(fw["_cache"]["idx"].tolist())) == 3          # all three axes occur as the shortest one
    assert (fw["_cache"]["sgn"] < 0).any() and (fw["_cache"]["sgn"] > 0).any()
    assert np.allclose(np.linalg.norm(fw["_cache"]["nh"], axis=1), 1.0, atol=1e-12)
    g = torch.Generator().manual_seed(9)
    cots = [torch.randn(o.shape, generator=g, dtype=torch.float64) for o in outs]
    torch.autograd.backward(list(outs), cots)
    d = O.backward_smallest_axis(fw, *[c.numpy() for c in cots])
    name = dict(fdc="features_dc", frest="features_rest")
    for k in IN_SA:
        want = leaves[k].grad.numpy()
        got = d[name.get(k, k)].reshape(want.shape)
        assert np.allclose(got, want, rtol=1e-9, atol=1e-12), k
