"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the optimizer step the reference runs:
torch.optim.Adam(l, lr=0.0, eps=1e-15) (scene/gaussian_model.py:240), i.e. torch/optim/adam.py `_single_tensor_adam`
with amsgrad=False, weight_decay=0, maximize=False.  torch is a third-party dependency of the reference (not under
/root/reference; this image: torch 2.11.0): its published algorithm is restated here and PINNED against
torch.optim.Adam itself run on the CPU (tests/test_adam_oracle.py).  Only tests/ may import this."""
import numpy as np


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-15, dtype=np.float64):
    """One update; returns (p, m, v).  `step` is the 1-based count of this update."""
    p, g, m, v = (np.asarray(x, dtype) for x in (p, g, m, v))
    m = m + (g - m) * dtype(1 - beta1)                 # exp_avg.lerp_(grad, 1 - beta1)
    v = v * dtype(beta2) + dtype(1 - beta2) * g * g    # exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = np.sqrt(v) / dtype(np.sqrt(bc2)) + dtype(eps)
    p = p - dtype(lr / bc1) * (m / denom)              # param.addcdiv_(exp_avg, denom, value=-step_size)
    return p, m, v
