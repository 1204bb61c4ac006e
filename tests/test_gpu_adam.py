"""GPU parity of the one-launch Adam (csrc/adam.cu through ibgs_b200.optim -> C ABI) against the numpy oracle and
torch.optim.Adam on the same device, with the reference's eight parameter groups (scene/gaussian_model.py:227-240)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GROUPS = (("xyz", 3, 1.6e-4), ("f_dc", 3, 0.0025), ("f_rest", 24, 0.0025 / 20), ("opacity", 1, 0.05),
          ("scaling", 3, 0.005), ("rotation", 4, 0.001), ("normal", 3, 0.001), ("offset", 1, 1.6e-5))


def _init(P, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {n: torch.randn((P, w), generator=g).cuda() for n, w, _ in GROUPS}


@pytest.mark.parametrize("P", [1, 1023, 4097])
def test_matches_oracle_and_torch_adam(P):
    from ibgs_b200.optim import ArenaAdam
    from oracle import adam_oracle as AO
    init = _init(P)
    lrs = {n: lr for n, _, lr in GROUPS}
    opt = ArenaAdam(init, lrs)
    ref_p = {n: init[n].clone().requires_grad_(True) for n in init}
    ref = torch.optim.Adam([{"params": [ref_p[n]], "lr": lrs[n], "name": n} for n in init], lr=0.0, eps=1e-15)
    orc = {n: (init[n].cpu().numpy().astype(np.float64), np.zeros((P, w)), np.zeros((P, w))) for n, w, _ in GROUPS}
    g = torch.Generator().manual_seed(1)
    for step in range(1, 7):
        if step == 3:   # update_learning_rate_offset changes two groups every iteration
            for grp_list in (opt.param_groups, ref.param_groups):
                for grp in grp_list:
                    if grp["name"] in ("xyz", "offset"):
                        grp["lr"] *= 0.7
            lrs["xyz"] *= 0.7
            lrs["offset"] *= 0.7
        for n, w, _ in GROUPS:
            gr = torch.randn((P, w), generator=g) * (10.0 ** (step % 4 - 3))
            # autograd-style accumulation into the arena-backed .grad (two "views" of a batch)
            opt.params[n].grad.add_(gr.cuda())
            opt.params[n].grad.add_(gr.cuda())
            ref_p[n].grad = (2 * gr).cuda()
            orc[n] = AO.adam_step(*orc[n][:1], (2 * gr).numpy(), *orc[n][1:], step, lrs[n])
        opt.step(zero_grads=True)
        ref.step()
        assert float(opt.flat_grads.abs().max()) == 0.0          # fused zero_grad
        for n, w, _ in GROUPS:
            ours = opt.params[n].detach().cpu().numpy().astype(np.float64)
            # float32 arithmetic against the float64 oracle; against torch's float32 Adam the difference is rounding
            assert np.abs(ours - orc[n][0]).max() <= 2e-6 * max(1.0, np.abs(orc[n][0]).max()), (step, n)
            assert np.abs(ours - ref_p[n].detach().cpu().numpy()).max() <= 2e-6, (step, n)
            st = opt.state(n)
            assert np.abs(st["exp_avg"].cpu().numpy() - orc[n][1]).max() <= 1e-6 * max(1.0, np.abs(orc[n][1]).max())
            assert st["step"] == step


def test_grad_scale_partial_groups_and_errors():
    from ibgs_b200.optim import ArenaAdam
    init = _init(100)
    lrs = {n: lr for n, _, lr in GROUPS}
    a, b = ArenaAdam(init, lrs), ArenaAdam(init, lrs)
    g = torch.Generator().manual_seed(2)
    for n, w, _ in GROUPS:
        gr = torch.randn((100, w), generator=g).cuda()
        a.params[n].grad.copy_(gr)
        b.params[n].grad.copy_(4 * gr)
    a.step()
    b.step(grad_scale=0.25)                       # mean over a batch of 4 views
    assert torch.allclose(a.flat_params, b.flat_params, rtol=0, atol=1e-7)
    assert float(a.flat_grads.abs().max()) > 0    # zero_grads defaults to off
    a.zero_grad()
    assert float(a.flat_grads.abs().max()) == 0.0
    a.params["xyz"].grad = None
    with pytest.raises(RuntimeError):
        a.step()
    with pytest.raises(RuntimeError):
        ArenaAdam({"x": torch.zeros(3)}, {"x": 0.1})    # CPU tensor: no fallback
    # moments survive a rebuild (densification path)
    c = ArenaAdam.from_state({n: b.params[n].detach() for n in init}, lrs,
                             {n: b.state(n)["exp_avg"] for n in init}, {n: b.state(n)["exp_avg_sq"] for n in init},
                             step=b.step_count)
    assert torch.equal(c.exp_avg, b.exp_avg) and torch.equal(c.flat_params, b.flat_params) and c.step_count == 1


def test_densification_surgery_matches_the_reference_helpers():
    """prune / extend / reset_state against the reference's optimizer surgery applied to torch.optim.Adam
    (scene/gaussian_model.py:362-438, re-stated inline on the same state dict), then one more step on both."""
    from ibgs_b200.optim import ArenaAdam
    P = 200
    init = _init(P, seed=3)
    lrs = {n: lr for n, _, lr in GROUPS}
    opt = ArenaAdam(init, lrs)
    ref_p = {n: torch.nn.Parameter(init[n].clone()) for n in init}
    ref = torch.optim.Adam([{"params": [ref_p[n]], "lr": lrs[n], "name": n} for n in init], lr=0.0, eps=1e-15)
    g = torch.Generator().manual_seed(9)

    def step_both():
        for grp in ref.param_groups:
            n = grp["name"]
            gr = torch.randn(tuple(grp["params"][0].shape), generator=g).cuda()
            grp["params"][0].grad = gr
            opt.params[n].grad.copy_(gr)
        ref.step()
        opt.step(zero_grads=True)

    def check():
        for grp in ref.param_groups:
            n, p = grp["name"], grp["params"][0]
            assert torch.allclose(opt.params[n].detach(), p.detach(), rtol=0, atol=5e-6), n
            # float32 moments: torch's multi-tensor kernels round a few operations differently (~1e-5 relative)
            assert torch.allclose(opt.state(n)["exp_avg"], ref.state[p]["exp_avg"], rtol=1e-3, atol=1e-6), n
            assert torch.allclose(opt.state(n)["exp_avg_sq"], ref.state[p]["exp_avg_sq"], rtol=1e-3, atol=1e-8), n

    step_both(); step_both(); check()
    # prune (_prune_optimizer)
    keep = (torch.rand(P, generator=g) > 0.3).cuda()
    for grp in ref.param_groups:
        p = grp["params"][0]
        st = ref.state.pop(p)
        st["exp_avg"], st["exp_avg_sq"] = st["exp_avg"][keep], st["exp_avg_sq"][keep]
        grp["params"][0] = torch.nn.Parameter(p.detach()[keep])
        ref.state[grp["params"][0]] = st
    opt = opt.prune(keep)
    check(); step_both(); check()
    # extend (cat_tensors_to_optimizer)
    new = {n: torch.randn((17, w), generator=g).cuda() for n, w, _ in GROUPS}
    for grp in ref.param_groups:
        p, ext = grp["params"][0], new[grp["name"]]
        st = ref.state.pop(p)
        st["exp_avg"] = torch.cat((st["exp_avg"], torch.zeros_like(ext)), dim=0)
        st["exp_avg_sq"] = torch.cat((st["exp_avg_sq"], torch.zeros_like(ext)), dim=0)
        grp["params"][0] = torch.nn.Parameter(torch.cat((p.detach(), ext), dim=0))
        ref.state[grp["params"][0]] = st
    opt = opt.extend(new)
    assert opt.step_count == 3 and opt.params["xyz"].shape[0] == int(keep.sum()) + 17
    check(); step_both(); check()
    # reset_state (replace_tensor_to_optimizer: the opacity reset)
    newop = torch.full_like(opt.params["opacity"].detach(), -2.0)
    for grp in ref.param_groups:
        if grp["name"] == "opacity":
            p = grp["params"][0]
            st = ref.state.pop(p)
            st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(newop), torch.zeros_like(newop)
            grp["params"][0] = torch.nn.Parameter(newop.clone())
            ref.state[grp["params"][0]] = st
    opt.reset_state("opacity", newop)
    check(); step_both(); check()
