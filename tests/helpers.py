"""Shared helpers of the GPU parity tests (the oracle is the checker, never the thing measured)."""

import numpy as np
import torch

import qcpinn_b200 as qb
from oracle import circuits as oc
from oracle import solver as osolver

F = qb.functional

# parity bars from BASELINE.json north_star: complex128 path 1e-10 relative, complex64 path 1e-5
TOL = {torch.float64: 1e-10, torch.float32: 1e-5}


def rel_err(got, want):
    got = torch.as_tensor(got).detach().double().cpu()
    want = torch.as_tensor(want).detach().double().cpu()
    scale = max(float(want.abs().max()), 1e-30)
    return float((got - want).abs().max()) / scale


def make_case(ansatz, n, layers, encoding="angle", haar_seed=None, hidden=50, seed=0):
    w = osolver.init_weights(n, layers, ansatz, hidden=hidden, seed=seed)
    # perturb biases so every gradient path is exercised (reference init zeroes the pre biases)
    g = torch.Generator().manual_seed(seed + 17)
    for k in ("b1", "b2"):
        w[k] = 0.1 * torch.randn(w[k].shape, generator=g)
    if encoding == "amplitude":
        w["b2"] = w["b2"] + 0.5   # keep |f| away from 0
    oracle = osolver.OracleSolver(n, layers, ansatz, encoding, haar_seed, "f64").set_weights(w)
    prog = qb.program.compile_program(ansatz, n, layers, haar_seed)
    return w, oracle, prog


def device_weights(w, dtype, device="cuda", requires_grad=False):
    return {k: v.to(device=device, dtype=dtype).clone().requires_grad_(requires_grad)
            for k, v in w.items()}


def mlp_list(dw):
    return [dw[k] for k in F.MLP_NAMES]


def points(n_pts, seed=3):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n_pts, 3, generator=g, dtype=torch.float64)
