"""CPU: the Adam oracle against torch.optim.Adam (the reference's optimizer, scene/gaussian_model.py:240) over
several steps with the reference's settings (lr per group, eps=1e-15) and a learning-rate change mid-run."""
import numpy as np
import torch

from oracle import adam_oracle as AO


def test_oracle_matches_torch_adam():
    g = torch.Generator().manual_seed(0)
    shapes = {"xyz": (50, 3), "f_rest": (50, 8, 3), "opacity": (50, 1)}
    lrs = {"xyz": 1.6e-4, "f_rest": 0.0025 / 20, "opacity": 0.05}
    ps = {k: torch.randn(s, generator=g, dtype=torch.float64).requires_grad_(True) for k, s in shapes.items()}
    opt = torch.optim.Adam([{"params": [ps[k]], "lr": lrs[k], "name": k} for k in shapes], lr=0.0, eps=1e-15)
    mine = {k: (ps[k].detach().numpy().copy(), np.zeros(shapes[k]), np.zeros(shapes[k])) for k in shapes}
    for step in range(1, 8):
        if step == 4:
            opt.param_groups[0]["lr"] = lrs["xyz"] = 9e-5   # update_learning_rate_offset (gaussian_model.py:251-262)
        for k in shapes:
            gr = torch.randn(shapes[k], generator=g, dtype=torch.float64) * (10.0 ** (step % 3 - 2))
            ps[k].grad = gr.clone()
            p, m, v = mine[k]
            mine[k] = AO.adam_step(p, gr.numpy(), m, v, step, lrs[k])
        opt.step()
        for k in shapes:
            assert np.abs(mine[k][0] - ps[k].detach().numpy()).max() < 1e-13, (step, k)
            st = opt.state[ps[k]]
            assert np.abs(mine[k][1] - st["exp_avg"].numpy()).max() < 1e-15
            assert np.abs(mine[k][2] - st["exp_avg_sq"].numpy()).max() < 1e-15
