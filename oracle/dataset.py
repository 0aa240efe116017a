"""Oracle (test infrastructure): analytic convection-diffusion target and forcing.

Restates ``data/diffusion_dataset.py:5-38`` of the reference: the Gaussian pulse
``u = exp(-100((x-.5)^2+(y-.5)^2)) exp(-t)`` and its closed-form forcing
``r = u_t + v_x u_x + v_y u_y - D (u_xx + u_yy)`` with D = 0.01, v = (1, 1).
"""

import torch

D_DEFAULT = 0.01
VX_DEFAULT = 1.0
VY_DEFAULT = 1.0


def u_exact(txy):
    t, x, y = txy[:, 0:1], txy[:, 1:2], txy[:, 2:3]
    return torch.exp(-100.0 * ((x - 0.5) ** 2 + (y - 0.5) ** 2)) * torch.exp(-t)


def forcing(txy, diffusion=D_DEFAULT, v_x=VX_DEFAULT, v_y=VY_DEFAULT):
    x, y = txy[:, 1:2], txy[:, 2:3]
    u = u_exact(txy)
    u_t = -u
    u_x = -200.0 * (x - 0.5) * u
    u_y = -200.0 * (y - 0.5) * u
    u_xx = (40000.0 * (x - 0.5) ** 2 - 400.0) * u
    u_yy = (40000.0 * (y - 0.5) ** 2 - 400.0) * u
    return u_t + v_x * u_x + v_y * u_y - diffusion * (u_xx + u_yy)


def box_points(lo, hi, count, generator=None, dtype=torch.float32):
    """``Sampler.sample`` (``:12-19``): uniform points in the box [lo, hi] (degenerate faces ok)."""
    lo = torch.as_tensor(lo, dtype=dtype).reshape(1, -1)
    hi = torch.as_tensor(hi, dtype=dtype).reshape(1, -1)
    rnd = torch.rand(count, lo.shape[1], generator=generator, dtype=dtype)
    return lo + (hi - lo) * rnd


# boxes used by trainer/diffusion_train.py:9-25 (only bc1 is ever sampled)
IC_BOX = ([0.0, 0.0, 0.0], [0.0, 1.0, 1.0])
BC1_BOX = ([0.0, 0.0, 0.0], [1.0, 0.0, 1.0])
DOM_BOX = ([0.0, 0.0, 0.0], [1.0, 1.0, 1.0])
