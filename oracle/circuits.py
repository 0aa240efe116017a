"""Oracle (test infrastructure): gate-by-gate batched statevector simulation of the DV circuit.

Restates ``nn/DVQuantumLayer.py:176-214`` (circuit order) and ``:246-371`` (the six ansaetze) of the
reference using PennyLane's published gate conventions (wire 0 = most significant bit of the basis
index).  Every function is differentiable with plain torch autograd to any order, which is what
``default.qubit`` + ``diff_method="backprop"`` gives the reference.  PARITY UNPINNED (see package
docstring).
"""

from __future__ import annotations

import math

import numpy as np
import torch

ANSATZ_NAMES = ("layered", "alternate", "cascade", "farhi", "sim_circ_15", "cross_mesh")
# data re-uploading family of reference hybrid_testing/CG_HQPINN_IBMtest_16qubits.py:217-235 (its own
# entry point: no separate encoding stage, no Haar blocks, no final Hadamard)
REUPLOAD_NAMES = ("cz_melt",)


def params_per_layer(ansatz: str, n: int) -> int:
    """Trainable angles per layer, reference ``nn/DVQuantumLayer.py:25-78``."""
    if ansatz == "layered":
        return 4 * n
    if ansatz in ("alternate", "alternate_flat"):
        return 4 * n - 4
    if ansatz == "cascade":
        return 3 * n
    if ansatz == "farhi":
        return 2 * n - 2
    if ansatz == "sim_circ_15":
        return 2 * n
    if ansatz == "cross_mesh":
        return 4 * n + n * (n - 1)
    if ansatz == "cz_melt":
        return 3 * n
    raise ValueError("Parameters are not initialized. Check the q_ansatz value.")


# ----------------------------------------------------------------------------------------------
# gate matrices (PennyLane definitions)
# ----------------------------------------------------------------------------------------------

def _cplx(re, im):
    return torch.complex(re, im)


def rx_matrix(theta, cdtype):
    """RX(t) = [[c, -i s], [-i s, c]], c = cos(t/2), s = sin(t/2).  theta: () or (B,)."""
    c = torch.cos(theta / 2)
    s = torch.sin(theta / 2)
    z = torch.zeros_like(c)
    row0 = torch.stack([_cplx(c, z), _cplx(z, -s)], dim=-1)
    row1 = torch.stack([_cplx(z, -s), _cplx(c, z)], dim=-1)
    return torch.stack([row0, row1], dim=-2).to(cdtype)


def ry_matrix(theta, cdtype):
    """RY(t) = [[c, -s], [s, c]]."""
    c = torch.cos(theta / 2)
    s = torch.sin(theta / 2)
    z = torch.zeros_like(c)
    row0 = torch.stack([_cplx(c, z), _cplx(-s, z)], dim=-1)
    row1 = torch.stack([_cplx(s, z), _cplx(c, z)], dim=-1)
    return torch.stack([row0, row1], dim=-2).to(cdtype)


def rz_matrix(theta, cdtype):
    """RZ(t) = diag(exp(-i t/2), exp(+i t/2))."""
    c = torch.cos(theta / 2)
    s = torch.sin(theta / 2)
    z = torch.zeros_like(c)
    zero = _cplx(z, z)
    row0 = torch.stack([_cplx(c, -s), zero], dim=-1)
    row1 = torch.stack([zero, _cplx(c, s)], dim=-1)
    return torch.stack([row0, row1], dim=-2).to(cdtype)


def hadamard_matrix(cdtype):
    h = 1.0 / math.sqrt(2.0)
    return torch.tensor([[h, h], [h, -h]], dtype=cdtype)


def controlled(u2, cdtype):
    """|0><0| (x) I + |1><1| (x) U on wires [control, target] (control = high bit)."""
    eye = torch.eye(2, dtype=cdtype)
    top = torch.cat([eye, torch.zeros(2, 2, dtype=cdtype)], dim=1)
    bot = torch.cat([torch.zeros(2, 2, dtype=cdtype), u2.to(cdtype)], dim=1)
    return torch.cat([top, bot], dim=0)


def cnot_matrix(cdtype):
    return torch.tensor(
        [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=cdtype
    )


def haar_unitaries(seed):
    """Reference ``nn/DVQuantumLayer.py:203-207``: two fixed 4x4 Haar unitaries from scipy."""
    from scipy.stats import unitary_group

    u1 = unitary_group.rvs(4, random_state=np.random.RandomState(seed))
    u2 = unitary_group.rvs(4, random_state=np.random.RandomState(seed + 1))
    return np.asarray(u1, dtype=np.complex128), np.asarray(u2, dtype=np.complex128)


# ----------------------------------------------------------------------------------------------
# state manipulation
# ----------------------------------------------------------------------------------------------

def apply_matrix(state, u, wires, n):
    """Apply ``u`` ((2^k,2^k) shared or (B,2^k,2^k) per sample) to ``wires`` of ``state`` (B,2^n).

    ``u``'s row/column index is ``sum_j bit(wires[j]) << (k-1-j)`` (first listed wire = high bit),
    PennyLane's ``QubitUnitary`` convention.
    """
    b = state.shape[0]
    k = len(wires)
    psi = state.reshape((b,) + (2,) * n)
    src = [1 + w for w in wires]
    dst = list(range(1, 1 + k))
    psi = torch.movedim(psi, src, dst)
    shp = psi.shape
    psi = psi.reshape(b, 2 ** k, -1)
    if u.dim() == 2:
        psi = torch.einsum("ij,bjr->bir", u, psi)
    else:
        psi = torch.einsum("bij,bjr->bir", u, psi)
    psi = torch.movedim(psi.reshape(shp), dst, src)
    return psi.reshape(b, 2 ** n)


def expval_z(state, n):
    """<Z_i> for every wire -> (n, B) real (reference output orientation, ``:154,214``)."""
    b = state.shape[0]
    prob = (state.real ** 2 + state.imag ** 2).reshape((b,) + (2,) * n)
    outs = []
    for i in range(n):
        axes = tuple(a for a in range(1, n + 1) if a != 1 + i)
        marg = prob.sum(dim=axes) if axes else prob
        outs.append(marg[:, 0] - marg[:, 1])
    return torch.stack(outs)


# ----------------------------------------------------------------------------------------------
# ansaetze: each returns a flat list of ops ("RX"|"RY"|"RZ"|"CRX"|"CRZ"|"CNOT", wires, param_idx)
# ----------------------------------------------------------------------------------------------

def ansatz_ops(ansatz: str, n: int):
    """Gate list of ONE layer, in application order; param_idx indexes that layer's angle row."""
    ops = []
    p = 0

    def rot(kind, wire):
        nonlocal p
        ops.append((kind, (wire,), p))
        p += 1

    def crot(kind, c, t):
        nonlocal p
        ops.append((kind, (c, t), p))
        p += 1

    def cnot(c, t):
        ops.append(("CNOT", (c, t), None))

    if ansatz == "layered":  # :246-262
        for q in range(n):
            rot("RZ", q)
            rot("RX", q)
        for q in range(n):
            cnot(q, (q + 1) % n)
        for q in range(n):
            rot("RX", q)
            rot("RZ", q)
    elif ansatz == "alternate":  # :264-285 (over-indexes its 4n-4 angles for even n)
        def tdcnot(c, t):
            rot("RY", c)
            rot("RY", t)
            cnot(c, t)
            rot("RZ", c)
            rot("RZ", t)

        for i in list(range(n - 1))[::2]:
            tdcnot(i, (i + 1) % n)
        for i in list(range(n))[1::2]:
            tdcnot(i, (i + 1) % n)
    elif ansatz == "alternate_flat":  # train_hybrid_qpinn.py:273-295: no wrap-around pair
        for i in list(range(0, n - 1, 2)) + list(range(1, n - 1, 2)):
            rot("RY", i)
            rot("RY", i + 1)
            cnot(i, i + 1)
            rot("RZ", i)
            rot("RZ", i + 1)
    elif ansatz == "cascade":  # :287-305
        for q in range(n):
            rot("RX", q)
        for q in range(n):
            rot("RZ", q)
        crot("CRX", n - 1, 0)
        for q in reversed(range(1, n)):
            crot("CRX", q - 1, q)
    elif ansatz == "farhi":  # :307-324
        for q in range(n - 1):
            cnot(n - 1, q)
            rot("RX", n - 1)
            cnot(n - 1, q)
        for q in range(n - 1):
            cnot(n - 1, q)
            rot("RZ", n - 1)
            cnot(n - 1, q)
    elif ansatz == "sim_circ_15":  # :326-346
        for q in range(n):
            rot("RY", q)
        for q in reversed(range(n)):
            cnot(q, (q + 1) % n)
        for q in range(n):
            rot("RY", q)
        for q in range(n):
            c = (q + n - 1) % n
            cnot(c, (c + 3) % n)
    elif ansatz == "cross_mesh":  # :348-371
        for q in range(n):
            rot("RX", q)
        for q in range(n):
            rot("RZ", q)
        for i in range(n - 1, -1, -1):
            for j in range(n - 1, -1, -1):
                if j != i:
                    crot("CRZ", i, j)
        for q in range(n):
            rot("RX", q)
        for q in range(n):
            rot("RZ", q)
    else:
        raise ValueError("Parameters are not initialized. Check the q_ansatz value.")
    return ops


_ROT = {"RX": rx_matrix, "RY": ry_matrix, "RZ": rz_matrix}


def apply_ansatz_layer(state, ansatz, n, row, cdtype):
    """One layer with angle row ``row`` (P,).  ``row[idx]`` raises IndexError when over-indexed."""
    for kind, wires, idx in ansatz_ops(ansatz, n):
        if kind == "CNOT":
            state = apply_matrix(state, cnot_matrix(cdtype), list(wires), n)
        elif kind in _ROT:
            state = apply_matrix(state, _ROT[kind](row[idx], cdtype), list(wires), n)
        elif kind == "CRX":
            state = apply_matrix(state, controlled(rx_matrix(row[idx], cdtype), cdtype), list(wires), n)
        elif kind == "CRZ":
            state = apply_matrix(state, controlled(rz_matrix(row[idx], cdtype), cdtype), list(wires), n)
        else:  # pragma: no cover
            raise AssertionError(kind)
    return state


def check_param_row(ansatz, n, row):
    """The reference's per-builder length checks (``:247,265,309,327,351``)."""
    want = params_per_layer(ansatz, n)
    if ansatz in ("layered", "alternate", "alternate_flat"):
        assert row is not None and len(row) == want
    elif ansatz in ("farhi", "sim_circ_15", "cross_mesh"):
        if row is None or len(row) != want:
            raise ValueError(f"Expected {want} parameters but got {tuple(row.shape)}")


def encode(x, n, encoding, cdtype):
    """Reference ``:177-182``.  x: (B, n) real -> (B, 2^n) complex."""
    b = x.shape[0]
    if encoding == "amplitude":
        pad = torch.zeros(b, 2 ** n - x.shape[1], dtype=x.dtype)
        full = torch.cat([x, pad], dim=1)
        full = full / torch.sqrt((full ** 2).sum(dim=1, keepdim=True))
        return torch.complex(full, torch.zeros_like(full)).to(cdtype)
    state = torch.zeros(b, 2 ** n, dtype=cdtype)
    state[:, 0] = 1.0
    for i in range(n):
        state = apply_matrix(state, rx_matrix(x[:, i], cdtype), [i], n)
    return state


def cz_melt_state(x, params, n, cdtype=torch.complex128):
    """Reference ``hybrid_testing/CG_HQPINN_IBMtest_16qubits.py:217-235``: RY(x_i) on wire i; per
    layer and wire RZ(0.5 x_{(i+layer) % n}) then Rot(phi, theta, omega) = RZ(omega) RY(theta)
    RZ(phi) with (phi, theta, omega) = weights[layer, i, :]; CZ on even pairs, odd pairs and
    (n-1, 0).  ``params``: (L, 3n) = the reference's (L, n, 3) weights flattened."""
    b = x.shape[0]
    state = torch.zeros(b, 2 ** n, dtype=cdtype)
    state[:, 0] = 1.0
    cz = torch.diag(torch.tensor([1, 1, 1, -1], dtype=cdtype))
    for i in range(n):
        state = apply_matrix(state, ry_matrix(x[:, i], cdtype), [i], n)
    for layer in range(params.shape[0]):
        row = params[layer]
        for i in range(n):
            state = apply_matrix(state, rz_matrix(0.5 * x[:, (i + layer) % n], cdtype), [i], n)
            state = apply_matrix(state, rz_matrix(row[3 * i], cdtype), [i], n)
            state = apply_matrix(state, ry_matrix(row[3 * i + 1], cdtype), [i], n)
            state = apply_matrix(state, rz_matrix(row[3 * i + 2], cdtype), [i], n)
        pairs = [(i, i + 1) for i in range(0, n - 1, 2)] + [(i, i + 1) for i in range(1, n - 1, 2)]
        for a, c in pairs + [(n - 1, 0)]:
            state = apply_matrix(state, cz, [a, c], n)
    return state


def final_state(x, params, ansatz, n, encoding="angle", haar=None, cdtype=torch.complex128):
    """State just before measurement: encoding -> L x ansatz -> [Haar] -> H(last wire)."""
    if ansatz == "cz_melt":
        return cz_melt_state(x, params, n, cdtype)
    state = encode(x, n, encoding, cdtype)
    for layer in range(params.shape[0]):
        row = params[layer]
        check_param_row(ansatz, n, row)
        state = apply_ansatz_layer(state, ansatz, n, row, cdtype)
    if haar is not None:
        u1, u2 = haar
        state = apply_matrix(state, torch.as_tensor(u1).to(cdtype), [0, 1], n)
        state = apply_matrix(state, torch.as_tensor(u2).to(cdtype), [2, 3], n)
    if n > 0:
        state = apply_matrix(state, hadamard_matrix(cdtype), [n - 1], n)
    return state


def quantum_layer(x, params, ansatz, n, encoding="angle", haar=None, cdtype=torch.complex128):
    """``DVQuantumLayer.forward`` (batch branch, ``:151-154``): (B,n) -> (n,B) real."""
    return expval_z(final_state(x, params, ansatz, n, encoding, haar, cdtype), n)


def haar_for(args_seed, n):
    """Haar rule ``:88-94``: only if a seed is given and n >= 4."""
    if args_seed is None or n < 4:
        return None
    return haar_unitaries(args_seed)
