"""PDE residual operators (drop-in for reference nn/pde.py).

``diffusion_operator`` keeps the reference signature and return value ``(u, residual)``.  For the
B200 ``DVPDESolver`` it dispatches to ONE fused Taylor-mode kernel that carries
``u, u_t, u_x, u_y, u_xx, u_yy`` together (no nested autograd); both outputs stay connected to
every model parameter through a hand-written adjoint kernel.  Any other ``nn.Module`` takes the
generic formulation of the reference (five ``autograd.grad(create_graph=True)`` calls).
"""

import torch


def _diffusion_coeffs(sigma_t, sigma_x, sigma_y, D, v_x, v_y):
    # r = u_t/st + v_x u_x/sx + v_y u_y/sy - D (u_xx/sx^2 + u_yy/sy^2)   (reference :60-71)
    return (1.0 / sigma_t, v_x / sigma_x, v_y / sigma_y,
            -D / (sigma_x * sigma_x), -D / (sigma_y * sigma_y))


def diffusion_operator(model, t, x, y, sigma_t=1.0, sigma_x=1.0, sigma_y=1.0,
                       D=0.01, v_x=1.0, v_y=1.0):
    t.requires_grad = True
    x.requires_grad = True
    y.requires_grad = True
    fused = getattr(model, "taylor_residual", None)
    if fused is not None:
        return fused(torch.cat((t, x, y), 1),
                     _diffusion_coeffs(sigma_t, sigma_x, sigma_y, D, v_x, v_y))
    u = model(torch.cat((t, x, y), 1))
    ones = torch.ones_like(u)

    def grad(out, wrt):
        return torch.autograd.grad(out, wrt, ones, create_graph=True)[0]

    u_t = grad(u, t) / sigma_t
    u_x = grad(u, x) / sigma_x
    u_y = grad(u, y) / sigma_y
    u_xx = grad(u_x, x) / sigma_x
    u_yy = grad(u_y, y) / sigma_y
    residual = u_t + v_x * u_x + v_y * u_y - D * (u_xx + u_yy)
    return u, residual
