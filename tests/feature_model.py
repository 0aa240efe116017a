"""Test-only numpy mirror of the batch-shared maths the CUDA library implements
(qcp_plan.cu: prepare_kernel / theta_grad_kernel).  It interprets a compiled gate program with the
ORACLE's gate matrices and rebuilds V, O_i = V^dag Z_i V and the feature matrix C, so the CPU suite
can check the circuit compiler and the Heisenberg-picture derivation without a GPU."""

import numpy as np
import torch

from oracle import circuits as oc

KIND = ["RX", "RY", "RZ", "CRX", "CRZ", "CNOT", "H", "U4"]


def program_unitary(program, theta):
    """V (2^n x 2^n complex128) of the gate table applied to the computational basis."""
    n = program.n_qubits
    m = 2 ** n
    cd = torch.complex128
    state = torch.eye(m, dtype=cd)           # row a = basis state a (batch dimension)
    th = torch.as_tensor(theta, dtype=torch.float64).reshape(-1)
    for kind, a, b, p in program.ops.tolist():
        name = KIND[kind]
        if name in ("RX", "RY", "RZ"):
            u = {"RX": oc.rx_matrix, "RY": oc.ry_matrix, "RZ": oc.rz_matrix}[name](th[p], cd)
            state = oc.apply_matrix(state, u, [a], n)
        elif name == "CRX":
            state = oc.apply_matrix(state, oc.controlled(oc.rx_matrix(th[p], cd), cd), [a, b], n)
        elif name == "CRZ":
            state = oc.apply_matrix(state, oc.controlled(oc.rz_matrix(th[p], cd), cd), [a, b], n)
        elif name == "CNOT":
            state = oc.apply_matrix(state, oc.cnot_matrix(cd), [a, b], n)
        elif name == "H":
            state = oc.apply_matrix(state, oc.hadamard_matrix(cd), [a], n)
        else:
            state = oc.apply_matrix(state, torch.as_tensor(program.consts[p]).to(cd), [a, b], n)
    return state.numpy().T                    # column a = V|a>


def observables(V, n):
    m = 2 ** n
    out = []
    for i in range(n):
        z = np.array([1.0 - 2.0 * ((k >> (n - 1 - i)) & 1) for k in range(m)])
        out.append(V.conj().T @ (z[:, None] * V))
    return out


_PAULI = {
    0: np.eye(2, dtype=complex),
    1: np.array([[0, -1j], [1j, 0]]),
    2: np.array([[1, 0], [0, -1]], dtype=complex),
}


def feature_matrix_angle(O, n):
    """C[i, s] = tr(P_s O_i) / 2^n with s = sum_j s_j 3^(n-1-j), s_j in {I, Y, Z}."""
    F = 3 ** n
    C = np.zeros((n, F))
    for s in range(F):
        P = np.array([[1.0 + 0j]])
        for j in range(n):
            P = np.kron(P, _PAULI[(s // 3 ** (n - 1 - j)) % 3])
        for i in range(n):
            C[i, s] = np.real(np.trace(P @ O[i])) / 2 ** n
    return C


def feature_matrix_amplitude(O, n):
    C = np.zeros((n, n * (n + 1) // 2))
    s = 0
    for a in range(n):
        for b in range(a, n):
            for i in range(n):
                C[i, s] = (1.0 if a == b else 2.0) * np.real(O[i][a, b])
            s += 1
    return C


def features_angle(z):
    """phi(z) = kron_j (1, -sin z_j, cos z_j);  z: (B, n) -> (B, 3^n)."""
    z = np.asarray(z, dtype=np.float64)
    out = np.ones((z.shape[0], 1))
    for j in range(z.shape[1]):
        b = np.stack([np.ones(z.shape[0]), -np.sin(z[:, j]), np.cos(z[:, j])], axis=1)
        out = (out[:, :, None] * b[:, None, :]).reshape(z.shape[0], -1)
    return out


def features_amplitude(f):
    f = np.asarray(f, dtype=np.float64)
    n = f.shape[1]
    nrm = (f ** 2).sum(axis=1)
    cols = [f[:, a] * f[:, b] / nrm for a in range(n) for b in range(a, n)]
    return np.stack(cols, axis=1)


def feature_tensor_angle_fast(O, n):
    """Same C as :func:`feature_matrix_angle`, by n axis contractions instead of 3^n Kronecker
    products: view O_i as a (row bits, column bits) tensor and contract every qubit's (row, column)
    pair with K[t, r, c] = P_t[c, r] / 2 (so that sum_{r,c} K[t,r,c] O[r,c] = tr(P_t O) / 2).
    O(n 4^n) work per observable -- the transform a device prepare step would run for n = 10
    (DESIGN.md "what comes next", item 1).  Returns (n, 3, ..., 3) with one axis per qubit."""
    K = np.stack([_PAULI[t].T / 2.0 for t in range(3)])            # [t, r, c]
    out = []
    for i in range(n):
        T = np.asarray(O[i]).reshape((2,) * (2 * n))               # axes: r_0..r_{n-1}, c_0..c_{n-1}
        for j in range(n):
            # qubit j's row axis is always the first remaining one, its column axis n - j later
            T = np.tensordot(K, np.moveaxis(T, [0, n - j], [0, 1]), axes=([1, 2], [0, 1]))
            T = np.moveaxis(T, 0, -1)                               # finished trit axes go last
        out.append(np.real(T))
    return np.stack(out)


def factored_contraction(C, z, n_a):
    """<Z_i>(z) = phiA(z_A)^T C_i phiB(z_B) with the features split after the first ``n_a`` qubits:
    C (n, 3, ..., 3) -> (n, 3^n_a, 3^(n-n_a)); the (B x 3^nB) x (3^nB x n 3^nA) product is the GEMM
    of the planned tensor-core engine, the phiA dot its epilogue."""
    n = C.shape[0]
    Cm = C.reshape(n, 3 ** n_a, 3 ** (n - n_a))
    pa, pb = features_angle(z[:, :n_a]), features_angle(z[:, n_a:])
    T = np.einsum("pb,iab->pia", pb, Cm)                           # the GEMM
    return np.einsum("pa,pia->ip", pa, T)                          # the epilogue
