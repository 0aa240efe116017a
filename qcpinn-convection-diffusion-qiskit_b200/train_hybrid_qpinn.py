"""Single-file hybrid QPINN trainer (drop-in for reference train_hybrid_qpinn.py) on the B200 kernels.

The reference ships a second, self-contained entry point onto the same circuit family: argparse
flags (``:50-109``), the pure-diffusion problem ``u_t = D (u_xx + u_yy)`` with the analytic
solution ``sin(pi x) sin(pi y) exp(-2 pi^2 D t)`` (``:115-135``), four zero Dirichlet faces
(``:160-205``), a one-layer circuit with a 1-D angle vector ``randn(P) * 0.1`` and the Haar blocks
always on for n >= 4 (``:396-424``), xavier-normal on every Linear (``:593-599``), plateau patience
500 (``:583-585``) and its own checkpoint dicts (``:722-735``).  This module keeps those names,
defaults and file formats; the model is the fused ``DVPDESolver`` machinery underneath, so the
residual ``u_t - D (u_xx + u_yy)`` comes out of one Taylor-mode kernel call.

Differences, all forced by the stack: no IBM-hardware branch (``--use-ibm`` raises), no matplotlib
plots (``evaluate`` writes ``evaluation.json`` with the same relative L2 error instead).

    python -m qcpinn_b200.train_hybrid_qpinn --num-qubits 4 --ansatz cascade --epochs 200
"""

from __future__ import annotations

import argparse
import json
import math
import os
import time
from datetime import datetime

import numpy as np
import torch
import torch.nn as nn

from .nn.DVPDESolver import DVPDESolver, _RankConsistentPlateau
from .nn.DVQuantumLayer import DVQuantumLayer
from .program import ANSATZ_PARAM_COUNT
from .utils.logger import Logging

ANSATZ_CHOICES = ["cascade", "layered", "alternate", "farhi", "sim_circ_15", "cross_mesh"]


def parse_args(argv=None):
    """Same flags and defaults as reference train_hybrid_qpinn.py:50-109 (+ ``--dtype``)."""
    p = argparse.ArgumentParser(description="Hybrid Quantum PINN Trainer for 2D PDEs",
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("--device", type=str, default="auto", choices=["auto", "cuda", "cpu"],
                   help="Compute device (auto-detects CUDA if available)")
    p.add_argument("--use-ibm", action="store_true",
                   help="Use IBM Quantum hardware instead of simulator (not available here)")
    p.add_argument("--ibm-token", type=str, default=None, help="IBM Quantum API token")
    p.add_argument("--ibm-backend", type=str, default="ibm_torino", help="IBM Quantum backend name")
    p.add_argument("--ibm-instance", type=str, default=None, help="IBM Quantum instance (optional)")
    p.add_argument("--num-qubits", type=int, default=4, help="Number of qubits in quantum circuit")
    p.add_argument("--ansatz", type=str, default="cascade", choices=ANSATZ_CHOICES,
                   help="Quantum circuit ansatz type")
    p.add_argument("--encoding", type=str, default="angle", choices=["angle", "amplitude"],
                   help="Input encoding method")
    p.add_argument("--shots", type=int, default=1024, help="Measurement shots for hardware execution")
    p.add_argument("--epochs", type=int, default=5000, help="Number of training epochs")
    p.add_argument("--batch-size", type=int, default=64, help="Training batch size")
    p.add_argument("--lr", type=float, default=0.005, help="Learning rate")
    p.add_argument("--seed", type=int, default=42, help="Random seed for reproducibility")
    p.add_argument("--hidden-dim", type=int, default=50, help="Hidden dimension of classical layers")
    p.add_argument("--print-every", type=int, default=100, help="Print loss every N epochs")
    p.add_argument("--output-dir", type=str, default="./outputs", help="Base output directory")
    p.add_argument("--diffusion-coef", type=float, default=0.01, help="Diffusion coefficient D")
    p.add_argument("--dtype", type=str, default="float64", choices=["float64", "float32"],
                   help="Arithmetic of the CUDA kernels (not in the reference)")
    return p.parse_args(argv)


# -- analytic solution and samplers (reference :115-205) -------------------------------------------
def analytical_solution(t, x, y, D=0.01):
    return np.sin(np.pi * x) * np.sin(np.pi * y) * np.exp(-2 * np.pi ** 2 * D * t)


def analytical_solution_torch(X, D=0.01):
    t, x, y = X[:, 0:1], X[:, 1:2], X[:, 2:3]
    return torch.sin(torch.pi * x) * torch.sin(torch.pi * y) * torch.exp(-2 * torch.pi ** 2 * D * t)


class DataSampler:
    """Uniform points in a box ``coords`` = [[min...], [max...]] with targets ``func(X[, D])``."""

    def __init__(self, coords, func, device="cpu", D=0.01):
        self.coords, self.func, self.device, self.D = coords, func, device, D
        self.dim = coords.shape[1]

    def sample(self, N):
        rnd = torch.rand(N, self.dim, device=self.device)
        X = self.coords[0:1, :] + (self.coords[1:2, :] - self.coords[0:1, :]) * rnd
        y = self.func(X, self.D) if self.D is not None else self.func(X)
        return X, y


def _zeros_target(X, D=None):
    return torch.zeros((X.shape[0], 1), device=X.device)


def create_samplers(device, D=0.01):
    """(ics_sampler, [four boundary faces], residual sampler, domain box), reference :160-205."""
    def box(lo, hi):
        return torch.tensor([lo, hi], dtype=torch.float32, device=device)

    ics = DataSampler(box([0.0, 0.0, 0.0], [0.0, 1.0, 1.0]), analytical_solution_torch, device, D)
    faces = [([0.0, 0.0, 0.0], [1.0, 0.0, 1.0]), ([0.0, 1.0, 0.0], [1.0, 1.0, 1.0]),
             ([0.0, 0.0, 0.0], [1.0, 1.0, 0.0]), ([0.0, 0.0, 1.0], [1.0, 1.0, 1.0])]
    bcs = [DataSampler(box(lo, hi), _zeros_target, device, None) for lo, hi in faces]
    dom = box([0.0, 0.0, 0.0], [1.0, 1.0, 1.0])
    return ics, bcs, DataSampler(dom, _zeros_target, device, None), dom


# -- model (reference :396-626) ------------------------------------------------------------------------
class QuantumLayer(DVQuantumLayer):
    """One-layer circuit with a flat angle vector (reference :396-424): ``params`` is (P,) drawn as
    ``randn * 0.1``, the Haar blocks are on for n >= 4, ``alternate`` has no wrap-around pair, an
    unknown ansatz name falls back to ``cascade`` (``:217-226,381-390``)."""

    program_variant = "single_file"

    def __init__(self, num_qubits, ansatz_type="cascade", encoding="angle", use_ibm=False,
                 ibm_token=None, ibm_backend=None, ibm_instance=None, shots=1024, seed=42,
                 dtype="float64"):
        if ansatz_type not in ANSATZ_PARAM_COUNT:
            ansatz_type = "cascade"
        super().__init__({"num_qubits": num_qubits, "num_quantum_layers": 1, "q_ansatz": ansatz_type,
                          "problem": "diffusion", "encoding": encoding, "shots": shots, "seed": seed,
                          "use_ibm_hardware": bool(use_ibm), "dtype": dtype})
        self.ansatz_type = ansatz_type
        count = self.params.shape[1]
        self.params = nn.Parameter(torch.randn(count) * 0.1)

    def forward(self, x):
        return super().forward(x).T          # (B, n) like reference :500-508


class HybridQPINN(DVPDESolver):
    """``HybridQPINN(args, device)`` of reference :540-626 (argparse namespace in, same submodule
    and state-dict names out) on the fused kernels."""

    def __init__(self, args, device):
        self.hidden_dim = args.hidden_dim
        cfg = {"batch_size": args.batch_size, "epochs": args.epochs, "lr": args.lr,
               "num_qubits": args.num_qubits, "num_quantum_layers": 1,
               "classic_network": [3, args.hidden_dim, 1], "q_ansatz": args.ansatz,
               "problem": "diffusion", "print_every": args.print_every, "solver": "DV",
               "encoding": args.encoding, "seed": args.seed, "shots": args.shots,
               "use_ibm_hardware": bool(getattr(args, "use_ibm", False)),
               "dtype": getattr(args, "dtype", "float64")}
        self._ns = args
        super().__init__(cfg, _NullLogger(), device=device)
        self.args = args                     # the reference keeps the namespace here

    def _build_quantum_layer(self, cfg):
        a = self._ns
        return QuantumLayer(num_qubits=a.num_qubits, ansatz_type=a.ansatz, encoding=a.encoding,
                            use_ibm=getattr(a, "use_ibm", False), shots=a.shots, seed=a.seed,
                            dtype=cfg["dtype"])

    def _make_scheduler(self):
        return _RankConsistentPlateau(self.optimizer, mode="min", factor=0.9, patience=500)

    def _initialize_weights(self):
        for m in self.modules():             # every Linear, not only the preprocessor
            if isinstance(m, nn.Linear):
                nn.init.xavier_normal_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)


class _NullLogger(Logging):
    """The single-file trainer prints to stdout and has no log directory of its own."""

    def __init__(self):
        self.dir = None

    def print(self, *a):
        pass

    def get_output_dir(self):
        return os.getcwd()


def diffusion_operator(model, t, x, y, D=0.01):
    """(u, u_t - D (u_xx + u_yy)), reference :633-667."""
    t = t.requires_grad_(True)
    x = x.requires_grad_(True)
    y = y.requires_grad_(True)
    fused = getattr(model, "taylor_residual", None)
    if fused is not None:
        return fused(torch.cat([t, x, y], dim=1), (1.0, 0.0, 0.0, -D, -D))
    u = model(torch.cat([t, x, y], dim=1))
    ones = torch.ones_like(u)
    g = lambda out, wrt: torch.autograd.grad(out, wrt, ones, create_graph=True, retain_graph=True)[0]  # noqa: E731
    return u, g(u, t) - D * (g(g(u, x), x) + g(g(u, y), y))


def train(model, args, ics_sampler, bc_samplers, res_sampler, output_dir):
    """Training loop of reference :674-748: ``bs // 3`` initial, ``4 x bs // 12`` boundary and ``bs``
    residual points, loss ``2 MSE_res + 4 MSE_bc + 2 MSE_ic``, clip 1.0, Adam, plateau scheduler,
    ``checkpoint.pth`` every ``print_every`` epochs, ``model.pth`` at the end."""
    print("\n" + "=" * 60 + "\nTRAINING HYBRID QUANTUM PINN\n" + "=" * 60)
    D, bs = args.diffusion_coef, args.batch_size
    t0, times = time.time(), []
    many = getattr(model, "forward_many", None)
    for epoch in range(args.epochs + 1):
        te = time.time()
        model.optimizer.zero_grad()
        X_ics, u_ics = ics_sampler.sample(bs // 3)
        X_res, _ = res_sampler.sample(bs)
        parts = [s.sample(bs // 12) for s in bc_samplers]
        X_bc = torch.cat([p[0] for p in parts], dim=0)
        u_bc = torch.cat([p[1] for p in parts], dim=0)
        coeffs = (1.0, 0.0, 0.0, -D, -D)
        if many is not None:
            u_ics_pred, u_bc_pred, (_, residual) = many([(X_ics, None), (X_bc, None), (X_res, coeffs)])
        else:
            u_ics_pred, u_bc_pred = model(X_ics), model(X_bc)
            _, residual = diffusion_operator(model, X_res[:, 0:1], X_res[:, 1:2], X_res[:, 2:3], D)
        loss_ics = model.loss_fn(u_ics_pred, u_ics)
        loss_bc = model.loss_fn(u_bc_pred, u_bc)
        loss_res = model.loss_fn(residual, torch.zeros_like(residual))
        loss = 2.0 * loss_res + 4.0 * loss_bc + 2.0 * loss_ics
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
        model.optimizer.step()
        value = loss.item()
        model.scheduler.step(value)
        model.loss_history.append(value)
        times.append(time.time() - te)
        if epoch % args.print_every == 0 or epoch == 0:
            total = time.time() - t0
            eta = sum(times) / len(times) * (args.epochs - epoch)
            lr = float(model.optimizer.param_groups[0]["lr"])
            pct = 100 * epoch / args.epochs if args.epochs else 100.0
            print(f"Epoch {epoch:5d}/{args.epochs} [{pct:5.1f}%] | Loss: {value:.2e} | "
                  f"Res: {loss_res.item():.2e} | BC: {loss_bc.item():.2e} | IC: {loss_ics.item():.2e} | "
                  f"LR: {lr:.2e} | Time: {total:.1f}s | ETA: {eta:.1f}s")
            if epoch > 0:
                torch.save({"epoch": epoch, "model_state_dict": model.state_dict(),
                            "optimizer_state_dict": model.optimizer.state_dict(), "loss": value,
                            "loss_history": model.loss_history},
                           os.path.join(output_dir, "checkpoint.pth"))
    print(f"\nTraining completed in {time.time() - t0:.1f}s")
    torch.save(model.state_dict(), os.path.join(output_dir, "model.pth"))
    return model


def evaluate(model, args, dom_coords, output_dir):
    """Relative L2 error on the 20 x 20 slice t = 0.5 (reference :755-780); written to
    ``evaluation.json`` (the reference's matplotlib figures are not produced)."""
    print("\n" + "=" * 60 + "\nEVALUATION\n" + "=" * 60)
    device, D, n = model.device, args.diffusion_coef, 20
    xs = torch.linspace(0, 1, n, device=device)
    Xm, Ym = torch.meshgrid(xs, xs, indexing="ij")
    X_eval = torch.stack([torch.full_like(Xm.flatten(), 0.5), Xm.flatten(), Ym.flatten()], dim=1)
    model.eval()
    with torch.no_grad():
        u_pred = model(X_eval)
    u_ref = analytical_solution_torch(X_eval, D)
    error = float(torch.linalg.norm(u_ref - u_pred) / torch.linalg.norm(u_ref))
    print(f"Relative L2 Error at t=0.5: {error * 100:.4f}%")
    with open(os.path.join(output_dir, "evaluation.json"), "w") as f:
        json.dump({"t": 0.5, "grid": n, "relative_l2_error": error,
                   "max_abs_error": float((u_ref - u_pred).abs().max()),
                   "final_loss": model.loss_history[-1] if model.loss_history else math.nan}, f)
    return error


def main(argv=None):
    args = parse_args(argv)
    if args.use_ibm:
        raise NotImplementedError("--use-ibm (remote QPU execution) is outside the B200 hot path")
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu") if args.device == "auto" \
        else torch.device(args.device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    print("\n" + "=" * 60 + "\nHYBRID QUANTUM PINN TRAINER\n" + "=" * 60)
    for k, v in (("Device", device), ("Qubits", args.num_qubits), ("Ansatz", args.ansatz),
                 ("Encoding", args.encoding), ("Epochs", args.epochs), ("Batch size", args.batch_size),
                 ("Learning rate", args.lr), ("Diffusion coefficient", args.diffusion_coef)):
        print(f"{k}: {v}")
    output_dir = os.path.join(args.output_dir, datetime.now().strftime("%Y-%m-%d_%H-%M-%S"))
    os.makedirs(output_dir, exist_ok=True)
    print(f"Output directory: {output_dir}")
    with open(os.path.join(output_dir, "config.txt"), "w") as f:
        for key, value in vars(args).items():
            f.write(f"{key}: {'****' if key == 'ibm_token' and value else value}\n")
    ics, bcs, res, dom = create_samplers(device, D=args.diffusion_coef)
    model = HybridQPINN(args, device).to(device)
    print(f"Total parameters: {sum(p.numel() for p in model.parameters())}")
    print(f"Quantum parameters: {model.quantum_layer.params.numel()}")
    model = train(model, args, ics, bcs, res, output_dir)
    error = evaluate(model, args, dom, output_dir)
    print("\n" + "=" * 60 + "\nTRAINING COMPLETE\n" + "=" * 60)
    print(f"Final L2 Error: {error * 100:.4f}%\nResults saved to: {output_dir}")
    return error


if __name__ == "__main__":
    main()
