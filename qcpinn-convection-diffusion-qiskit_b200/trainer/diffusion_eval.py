"""Evaluation of a trained convection-diffusion solver on a regular space-time grid (the
"EVALUATION" block of reference trainer/diffusion_hybrid_trainer.py:125-185):
``num_points``^3 grid over the unit cube through ``diffusion_operator`` (the same six-stream fused
forward as training, no backward), then relative L2 errors (percent) of ``u`` and of the residual
against the analytic solution / forcing.  Plotting is out of scope."""

import torch

from ..data.diffusion_dataset import r, u
from ..nn.pde import diffusion_operator


def evaluation_grid(num_points=20, device="cpu"):
    """(num_points^3, 3) points (t, x, y), t slowest -- ``torch.meshgrid(indexing='ij')`` order."""
    axis = torch.linspace(0.0, 1.0, num_points, dtype=torch.float32, device=device)
    t, x, y = torch.meshgrid(axis, axis, axis, indexing="ij")
    return torch.stack((t.flatten(), x.flatten(), y.flatten()), dim=1)


def evaluate(model, num_points=20):
    """Returns dict(error_u, error_f, u_pred, f_pred, X_star); errors are percentages like the
    reference's ``Relative L2 error_u`` / ``error_f`` log lines (which this also logs)."""
    X_star = evaluation_grid(num_points, model.device)
    u_pred, f_pred = diffusion_operator(model, X_star[:, 0:1], X_star[:, 1:2], X_star[:, 2:3])
    u_pred, f_pred = u_pred.detach(), f_pred.detach()
    u_ref, f_ref = u(X_star), r(X_star)
    error_u = (torch.linalg.norm(u_ref - u_pred) / torch.linalg.norm(u_ref)).item() * 100.0
    error_f = (torch.linalg.norm(f_ref - f_pred) / torch.linalg.norm(f_ref + 1e-9)).item() * 100.0
    model.logger.print("Relative L2 error_u: {:.2e}".format(error_u))
    model.logger.print("Relative L2 error_f: {:.2e}".format(error_f))
    return {"error_u": error_u, "error_f": error_f, "u_pred": u_pred, "f_pred": f_pred,
            "X_star": X_star}
