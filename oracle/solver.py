"""Oracle (test infrastructure): DVPDESolver, diffusion residual and one train step on the CPU.

Restates ``nn/DVPDESolver.py:28-51,69-76,81-110`` (pre MLP -> quantum layer -> cast/transposes ->
post MLP), ``nn/pde.py:53-72`` (five nested ``autograd.grad(create_graph=True)`` calls) and
``trainer/diffusion_train.py:30-49,81-90`` (weighted MSE loss, backward, clip, Adam) of the
reference, on top of :mod:`oracle.circuits`.  PARITY UNPINNED (see package docstring).

Precision modes
---------------
``"f64"``    everything float64 / complex128 (the 1e-10 parity target of the CUDA f64 path)
``"f32"``    everything float32 / complex64   (context for the 1e-5 target of the CUDA f32 path)
``"mixed"``  reference-faithful: float32 MLPs and gate angles, complex128 state, float64 expvals
             cast back to float32 (``nn/DVPDESolver.py:96``)
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch

from . import circuits, dataset

WEIGHT_NAMES = ("w1", "b1", "w2", "b2", "theta", "w3", "b3", "w4", "b4")


def _dtypes(mode):
    if mode == "f64":
        return torch.float64, torch.complex128
    if mode == "f32":
        return torch.float32, torch.complex64
    if mode == "mixed":
        return torch.float32, torch.complex128
    raise ValueError(mode)


def init_weights(n, layers, ansatz, hidden=50, in_dim=3, out_dim=1, seed=0):
    """Reference initialisation: xavier-normal + zero bias on the pre MLP only
    (``nn/DVPDESolver.py:69-76``), torch-default Linear init on the post MLP, xavier-normal on the
    (L, P) angle matrix (``nn/DVQuantumLayer.py:216-244``).  Construction order follows
    ``DVPDESolver.__init__`` so ``torch.manual_seed(seed)`` reproduces a reference model's stream.
    """
    torch.manual_seed(seed)
    pre1 = torch.nn.Linear(in_dim, hidden)
    pre2 = torch.nn.Linear(hidden, n)
    post1 = torch.nn.Linear(n, hidden)
    post2 = torch.nn.Linear(hidden, out_dim)
    theta = torch.empty(layers, circuits.params_per_layer(ansatz, n))
    torch.nn.init.xavier_normal_(theta)
    for lin in (pre1, pre2):
        torch.nn.init.xavier_normal_(lin.weight)
        torch.nn.init.zeros_(lin.bias)
    w = {
        "w1": pre1.weight, "b1": pre1.bias, "w2": pre2.weight, "b2": pre2.bias,
        "theta": theta,
        "w3": post1.weight, "b3": post1.bias, "w4": post2.weight, "b4": post2.bias,
    }
    return {k: v.detach().clone() for k, v in w.items()}


@dataclass
class OracleSolver:
    """Functional restatement of ``DVPDESolver`` (weights live in ``self.w``)."""

    n: int
    layers: int
    ansatz: str
    encoding: str = "angle"
    haar_seed: int | None = None
    mode: str = "f64"
    w: dict = field(default_factory=dict)

    def __post_init__(self):
        self.rdtype, self.cdtype = _dtypes(self.mode)
        self.haar = circuits.haar_for(self.haar_seed, self.n)

    def set_weights(self, weights, requires_grad=True):
        self.w = {
            k: weights[k].detach().to(self.rdtype).clone().requires_grad_(requires_grad)
            for k in WEIGHT_NAMES
        }
        return self

    # -- modules -----------------------------------------------------------------------------
    def pre(self, x):
        w = self.w
        return torch.tanh(x @ w["w1"].T + w["b1"]) @ w["w2"].T + w["b2"]

    def post(self, q):
        w = self.w
        return torch.tanh(q @ w["w3"].T + w["b3"]) @ w["w4"].T + w["b4"]

    def quantum(self, z):
        """(B, n) -> (n, B), like ``DVQuantumLayer.forward``."""
        return circuits.quantum_layer(
            z, self.w["theta"], self.ansatz, self.n, self.encoding, self.haar, self.cdtype
        )

    def forward(self, x):
        """``DVPDESolver.forward`` (``nn/DVPDESolver.py:81-110``): (B,3) -> (B,1)."""
        if x.dim() != 2:
            raise ValueError(f"Expected 2D input tensor, got shape {x.shape}")
        z = self.pre(x.to(self.rdtype))
        q = self.quantum(z).to(self.rdtype)  # (n, B); the float32 cast of :96 in mixed mode
        feats = q.T if (q.shape[0] == self.n and q.dim() == 2) else q
        return self.post(feats.reshape(-1, self.n))


def diffusion_operator(model, t, x, y, sigma_t=1.0, sigma_x=1.0, sigma_y=1.0,
                       D=0.01, v_x=1.0, v_y=1.0):
    """``nn/pde.py:53-72`` with nested autograd: returns (u, residual), both (N,1)."""
    t = t.requires_grad_(True)
    x = x.requires_grad_(True)
    y = y.requires_grad_(True)
    u = model.forward(torch.cat((t, x, y), 1))
    ones = torch.ones_like(u)

    def g(out, inp):
        return torch.autograd.grad(out, inp, ones, create_graph=True)[0]

    u_t = g(u, t) / sigma_t
    u_x = g(u, x) / sigma_x
    u_y = g(u, y) / sigma_y
    u_xx = g(u_x, x) / sigma_x
    u_yy = g(u_y, y) / sigma_y
    residual = u_t + v_x * u_x + v_y * u_y - D * (u_xx + u_yy)
    return u, residual


def diffusion_streams(model, X):
    """All six Taylor streams (u, u_t, u_x, u_y, u_xx, u_yy) as (N,6), for stream-level parity."""
    cols = [X[:, i:i + 1].detach().clone().requires_grad_(True) for i in range(3)]
    u = model.forward(torch.cat(cols, 1))
    ones = torch.ones_like(u)
    first = [torch.autograd.grad(u, c, ones, create_graph=True)[0] for c in cols]
    u_xx = torch.autograd.grad(first[1], cols[1], ones, create_graph=True)[0]
    u_yy = torch.autograd.grad(first[2], cols[2], ones, create_graph=True)[0]
    return torch.cat([u] + first + [u_xx, u_yy], dim=1)


def mse(a, b):
    return ((a - b) ** 2).mean()


def loss_terms(model, batches):
    """``objective_fn`` (``trainer/diffusion_train.py:30-49``).

    ``batches`` = dict(X_ic, u_ic, X_bc, u_bc, X_res, r_res).  Returns (loss, loss_r, loss_bc, loss_ic).
    """
    rd = model.rdtype
    u_bc = model.forward(batches["X_bc"].to(rd))
    u_ic = model.forward(batches["X_ic"].to(rd))
    xr = batches["X_res"].to(rd)
    _, r_pred = diffusion_operator(
        model, xr[:, 0:1].clone(), xr[:, 1:2].clone(), xr[:, 2:3].clone()
    )
    loss_r = mse(r_pred, batches["r_res"].to(rd))
    loss_bc = mse(u_bc, batches["u_bc"].to(rd))
    loss_ic = mse(u_ic, batches["u_ic"].to(rd))
    loss = 2.0 * loss_r + 4.0 * loss_bc + 2.0 * loss_ic
    return loss, loss_r, loss_bc, loss_ic


def loss_and_grads(model, batches):
    """Loss and d loss / d weight for every weight tensor (dict keyed like ``WEIGHT_NAMES``)."""
    for v in model.w.values():
        v.grad = None
    loss, lr_, lbc, lic = loss_terms(model, batches)
    loss.backward()
    grads = {k: (v.grad.detach().clone() if v.grad is not None else torch.zeros_like(v))
             for k, v in model.w.items()}
    return {"loss": loss.detach(), "loss_r": lr_.detach(), "loss_bc": lbc.detach(),
            "loss_ic": lic.detach()}, grads


def make_batches(n_res, seed=1234, dtype=torch.float32):
    """Synthetic step inputs as the trainer draws them (``trainer/diffusion_train.py:34-36``)."""
    g = torch.Generator().manual_seed(seed)
    x_ic = dataset.box_points(*dataset.IC_BOX, n_res // 3, g, dtype)
    x_bc = dataset.box_points(*dataset.BC1_BOX, n_res // 3, g, dtype)
    x_res = dataset.box_points(*dataset.DOM_BOX, n_res, g, dtype)
    return {
        "X_ic": x_ic, "u_ic": dataset.u_exact(x_ic),
        "X_bc": x_bc, "u_bc": dataset.u_exact(x_bc),
        "X_res": x_res, "r_res": dataset.forcing(x_res),
    }


class OracleTrainer:
    """Full reference step on the CPU: loss, backward, clip(1.0), Adam, ReduceLROnPlateau, item()."""

    def __init__(self, model: OracleSolver, lr=0.005):
        self.model = model
        self.params = [model.w[k] for k in WEIGHT_NAMES]
        self.opt = torch.optim.Adam(self.params, lr=lr)
        self.sched = torch.optim.lr_scheduler.ReduceLROnPlateau(
            self.opt, mode="min", factor=0.9, patience=1000)
        self.loss_history = []

    def step(self, batches):
        self.opt.zero_grad()
        loss, *_ = loss_terms(self.model, batches)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.params, max_norm=1)
        self.opt.step()
        self.sched.step(loss.detach())
        self.loss_history.append(loss.item())
        return self.loss_history[-1]


def flops_per_point(n, layers, ansatz, haar, hidden=50):
    """SURVEY.md section 8(d) algorithmic flop model: 20 * (F_fwd + F_mlp)."""
    m = 2 ** n
    per_layer = {
        "cascade": 27 * n, "layered": 40 * n, "sim_circ_15": 28 * n,
        "cross_mesh": 40 * n + 3 * n * (n - 1), "farhi": 20 * (n - 1),
        "alternate": 40 * (len(range(n - 1)[::2]) + len(range(n)[1::2])),
    }[ansatz]
    f_fwd = (14 * n + layers * per_layer + (60 if haar else 0) + 14 + (3 + n)) * m
    f_mlp = 2 * hidden * (2 * n + 4)
    return 20 * (f_fwd + f_mlp), f_fwd, f_mlp
