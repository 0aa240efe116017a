"""Opt-in double-differentiable evaluation of the compiled gate program with plain torch ops.

The CUDA kernels give first-order adjoints only (the residual's second derivatives ride forward in
Taylor mode), so code that differentiates the model output ITSELF with respect to the coordinates
-- ``torch.autograd.grad(u, t, create_graph=True)``, what the reference's ``nn/pde.py:60-70`` does
literally -- cannot run on them.  ``args["diff_mode"] = "autograd"`` routes ``DVQuantumLayer`` /
``DVPDESolver`` through this module instead: a batched statevector simulation of the same gate
table (:mod:`..program`) on the module's device, differentiable to any order by the autograd
engine like ``default.qubit`` + ``diff_method="backprop"`` (reference nn/DVQuantumLayer.py:143-149).
It is a slow path for experiments (gradient-enhanced terms, adaptive sampling, mixed second
derivatives), orders of magnitude slower than the kernels, never selected implicitly, and not a
fallback: the default mode still fails loudly without the CUDA library.

State layout: (B, 2, ..., 2) complex, wire 0 = first qubit axis (PennyLane's convention, most
significant bit of the basis index).
"""

from __future__ import annotations

import math

import torch

from ..program import CNOT, CRX, CRZ, CZ, HAD, RX, RY, RY_IN, RZ, RZ_IN, U4, CircuitProgram


def _rot(kind: int, theta: torch.Tensor, cdtype) -> torch.Tensor:
    """2x2 matrix of RX / RY / RZ; theta is 0-d (shared angle) or (B,) (per sample)."""
    c, s = torch.cos(theta / 2), torch.sin(theta / 2)
    zero = torch.zeros_like(c)
    if kind == RX:
        rows = ((torch.complex(c, zero), torch.complex(zero, -s)),
                (torch.complex(zero, -s), torch.complex(c, zero)))
    elif kind == RY:
        rows = ((torch.complex(c, zero), torch.complex(-s, zero)),
                (torch.complex(s, zero), torch.complex(c, zero)))
    else:
        rows = ((torch.complex(c, -s), torch.complex(zero, zero)),
                (torch.complex(zero, zero), torch.complex(c, s)))
    return torch.stack([torch.stack(r, dim=-1) for r in rows], dim=-2).to(cdtype)


def _apply_1q(state, m, wire):
    st = state.movedim(1 + wire, -1)
    if m.dim() == 2:
        st = st @ m.transpose(0, 1)
    else:
        shape = st.shape
        st = (st.reshape(shape[0], -1, 2) @ m.transpose(1, 2)).reshape(shape)
    return st.movedim(-1, 1 + wire)


def _apply_controlled(state, m, control, target):
    keep, hit = state.unbind(dim=1 + control)
    # after unbind the target axis moved down by one when it sat behind the control axis
    t_axis = target if target < control else target - 1
    hit = _apply_1q(hit, m, t_axis)
    return torch.stack((keep, hit), dim=1 + control)


def _apply_u4(state, u, w0, w1):
    u = u.reshape(2, 2, 2, 2)                                   # out0 out1 in0 in1
    st = torch.tensordot(state, u, dims=([1 + w0, 1 + w1], [2, 3]))
    return st.movedim((-2, -1), (1 + w0, 1 + w1))


def run_layer(program: CircuitProgram, encoding: str, x: torch.Tensor, theta: torch.Tensor,
              cdtype=torch.complex128) -> torch.Tensor:
    """(B, n) features, (L, P) angles -> (n, B) expectation values <Z_i> (real dtype of cdtype).
    Circuit = encoding -> gate table -> measurement, reference nn/DVQuantumLayer.py:176-214."""
    n = program.n_qubits
    rdtype = torch.float64 if cdtype == torch.complex128 else torch.float32
    x = x.to(rdtype)
    flat = theta.reshape(-1).to(rdtype)
    batch = x.shape[0]
    if program.reupload:
        state = torch.zeros((batch,) + (2,) * n, dtype=cdtype, device=x.device)
        state[(slice(None),) + (0,) * n] = 1.0
    elif encoding == "amplitude":
        feats = x
        if feats.shape[1] < 2 ** n:
            feats = torch.nn.functional.pad(feats, (0, 2 ** n - feats.shape[1]))
        feats = feats / torch.linalg.norm(feats, dim=1, keepdim=True)
        state = feats.to(cdtype).reshape((batch,) + (2,) * n)
    else:
        state = torch.zeros((batch,) + (2,) * n, dtype=cdtype, device=x.device)
        state[(slice(None),) + (0,) * n] = 1.0
        for i in range(n):
            state = _apply_1q(state, _rot(RX, x[:, i], cdtype), i)
    consts = torch.as_tensor(program.consts, dtype=cdtype, device=x.device)
    xflip = torch.tensor([[0, 1], [1, 0]], dtype=cdtype, device=x.device)
    had = torch.tensor([[1, 1], [1, -1]], dtype=cdtype, device=x.device) / math.sqrt(2.0)
    for kind, a, b, p in program.ops.tolist():
        if kind in (RX, RY, RZ):
            state = _apply_1q(state, _rot(kind, flat[p], cdtype), a)
        elif kind in (CRX, CRZ):
            state = _apply_controlled(state, _rot(RX if kind == CRX else RZ, flat[p], cdtype), a, b)
        elif kind == CNOT:
            state = _apply_controlled(state, xflip, a, b)
        elif kind == HAD:
            state = _apply_1q(state, had, a)
        elif kind == U4:
            state = _apply_u4(state, consts[p], a, b)
        elif kind in (RY_IN, RZ_IN):          # per-sample rotation by (p / 4) * x[:, b]
            state = _apply_1q(state, _rot(RY if kind == RY_IN else RZ, 0.25 * p * x[:, b], cdtype), a)
        elif kind == CZ:
            zflip = torch.tensor([[1, 0], [0, -1]], dtype=cdtype, device=x.device)
            state = _apply_controlled(state, zflip, a, b)
        else:
            raise ValueError(f"unknown gate kind {kind}")
    probs = (state.real ** 2 + state.imag ** 2).reshape(batch, -1)
    index = torch.arange(2 ** n, device=x.device)
    signs = torch.stack([1.0 - 2.0 * ((index >> (n - 1 - i)) & 1).to(rdtype) for i in range(n)])
    return signs @ probs.T                                          # (n, B)
