"""PDE residual operators (drop-in for reference nn/pde.py).

Every operator keeps the reference signature and return value.  For the B200 ``DVPDESolver`` they
dispatch to ONE fused Taylor-mode kernel that carries ``u`` with its first and pure second
derivatives together (no nested autograd); the outputs stay connected to every model parameter
through a hand-written adjoint kernel.  Any other ``nn.Module`` takes the generic formulation of the
reference (nested ``autograd.grad(create_graph=True)`` calls).

``diffusion_operator`` (reference :53-72) is the train-step hot path.  ``wave_operator`` (:41-52),
``klein_gordon_operator`` (:26-40) and ``helmholtz_operator`` (:73-95) act on two-input solvers
(``classic_network=[2, H, 1]``): their two coordinates ride in the kernel's two second-derivative
slots, the linear part of the residual is the kernel's coefficient vector and the ``u`` / ``u**k``
terms are added on the (differentiable) outputs.  ``navier_stokes_2D_operator`` (:2-25) needs three
outputs and products of streams: a three-output solver returns the six streams of every output as
differentiable tensors (``taylor_streams_grad``) and the products are plain tensor ops.
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
        # The fused route returns (u, r) connected to every PARAMETER; the coordinates are detached
        # on purpose: d(u, r)/d(t, x, y) would need third derivatives that the six Taylor streams
        # do not carry, so ``autograd.grad(r, x)`` fails loudly ("not used in the graph") instead
        # of returning silent zeros.  ``model.forward(X)`` (value mode) does give d u / d X.
        return fused(torch.cat((t, x, y), 1).detach(),
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


def _grad(out, wrt):
    return torch.autograd.grad(out, wrt, torch.ones_like(out), create_graph=True)[0]


def _two_input_residual(model, a, b, c_aa, c_bb):
    """(u, c_aa u_aa + c_bb u_bb) for a two-input model: fused kernel when available."""
    a.requires_grad_(True)
    b.requires_grad_(True)
    fused = getattr(model, "taylor_residual", None)
    if fused is not None:
        # the two coordinates sit in the kernel's x / y slots (the ones with second derivatives)
        # (coordinates detached: see diffusion_operator)
        return fused(torch.cat((a, b), 1).detach(), (0.0, 0.0, 0.0, c_aa, c_bb))
    u = model(torch.cat((a, b), 1))
    u_aa = _grad(_grad(u, a), a)
    u_bb = _grad(_grad(u, b), b)
    return u, c_aa * u_aa + c_bb * u_bb


def wave_operator(model, t, x, sigma_t=1.0, sigma_x=1.0):
    """u_tt - c^2 u_xx with c = 2 (reference nn/pde.py:41-52; the sigmas are accepted and unused
    there as well).  Returns (u, residual)."""
    c = 2
    return _two_input_residual(model, t, x, 1.0, -float(c ** 2))


def klein_gordon_operator(fluid_model, t, x, x_min=0.0, x_max=1.0):
    """u_tt + alpha u_xx + beta u + gamma u^k with (alpha, beta, gamma, k) = (-1, 0, 1, 3)
    (reference nn/pde.py:26-40).  Returns (u, residual)."""
    alpha, beta, gamma, k = -1.0, 0.0, 1.0, 3
    u, lin = _two_input_residual(fluid_model, t, x, 1.0, alpha)
    return u, lin + beta * u + gamma * u ** k


def helmholtz_operator(fluid_model, x1, x2):
    """u_x1x1 + u_x2x2 + lambda u with lambda = 1 (reference nn/pde.py:73-95).  Returns the list
    [u, residual] like the reference."""
    LAMBDA = 1.0
    u, lin = _two_input_residual(fluid_model, x1, x2, 1.0, 1.0)
    return [u, lin + LAMBDA * u]


def navier_stokes_2D_operator(model, t, x, y, min_x=0, max_x=1):
    """[continuity, f_u, f_v] of the incompressible 2-D Navier-Stokes equations for a three-output
    model (u, v, p) (reference nn/pde.py:2-25).  A B200 ``DVPDESolver`` with
    ``classic_network=[3, H, 3]`` takes the fused route: the six Taylor streams of each output come
    from the forward kernel (``model.taylor_streams_grad``: 12 derivatives + 3 values, no nested
    autograd), the stream products ``u u_x + v u_y`` are ordinary differentiable tensor ops on them,
    and the adjoint kernels take the per-stream cotangents back to every parameter.  Any other
    module uses the reference's nested-autograd formulation."""
    mu, density = 0.00345, 1056.0
    t.requires_grad = True
    x.requires_grad = True
    y.requires_grad = True
    fused = getattr(model, "taylor_streams_grad", None)
    if fused is not None:
        if getattr(model, "n_outputs", 1) != 3:
            raise ValueError("navier_stokes_2D_operator needs a three-output model (u, v, p): "
                             "classic_network=[3, H, 3]")
        S = fused(torch.cat((t, x, y), 1).detach())           # (B, 3, 6)
        col = lambda o, c: S[:, o, c:c + 1]
        u, u_t, u_x, u_y, u_xx, u_yy = (col(0, c) for c in range(6))
        v, v_t, v_x, v_y, v_xx, v_yy = (col(1, c) for c in range(6))
        p_x, p_y = col(2, 2), col(2, 3)
    else:
        uvp = model(torch.cat((t, x, y), 1))
        u, v, p = uvp[:, 0:1], uvp[:, 1:2], uvp[:, 2:3]
        u_t, u_x, u_y = _grad(u, t), _grad(u, x), _grad(u, y)
        v_t, v_x, v_y = _grad(v, t), _grad(v, x), _grad(v, y)
        p_x, p_y = _grad(p, x), _grad(p, y)
        u_xx, u_yy = _grad(u_x, x), _grad(u_y, y)
        v_xx, v_yy = _grad(v_x, x), _grad(v_y, y)
    continuity = u_x + v_y
    f_u = u_t + (u * u_x + v * u_y) + p_x / density - mu * (u_xx + u_yy)
    f_v = v_t + (u * v_x + v * v_y) + p_y / density - mu * (v_xx + v_yy)
    return [continuity, f_u, f_v]
