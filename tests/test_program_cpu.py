"""CPU checks of the host-side circuit compiler and of the Heisenberg-picture derivation the CUDA
library implements (feature matrix C with <Z_i>(z) = C phi(z)), against the oracle."""

import numpy as np
import pytest
import torch

import feature_model as fm
import qcpinn_b200 as qb
from oracle import circuits as oc

P = qb.program


@pytest.mark.parametrize("ansatz", oc.ANSATZ_NAMES)
@pytest.mark.parametrize("n", [2, 3, 4])
def test_program_matches_oracle_for_both_encodings(ansatz, n):
    if ansatz == "alternate" and n % 2 == 0:
        pytest.skip("reference over-indexes (separate test)")
    if ansatz == "sim_circ_15" and n == 3:
        pytest.skip("CNOT(c, (c+3)%3) acts on one wire: invalid in the reference as well")
    g = torch.Generator().manual_seed(n)
    for layers in (1, 2):
        for seed in (None, 1):
            prog = P.compile_program(ansatz, n, layers, seed)
            assert prog.n_theta == layers * oc.params_per_layer(ansatz, n)
            th = torch.randn(layers, prog.params_per_layer, generator=g, dtype=torch.float64)
            V = fm.program_unitary(prog, th)
            assert np.abs(V.conj().T @ V - np.eye(2 ** n)).max() < 1e-12
            O = fm.observables(V, n)
            z = torch.randn(5, n, generator=g, dtype=torch.float64)
            haar = oc.haar_for(seed, n)
            want = oc.quantum_layer(z, th, ansatz, n, "angle", haar).numpy()
            got = fm.feature_matrix_angle(O, n) @ fm.features_angle(z.numpy()).T
            assert np.abs(got - want).max() < 1e-12
            want = oc.quantum_layer(z, th, ansatz, n, "amplitude", haar).numpy()
            got = fm.feature_matrix_amplitude(O, n) @ fm.features_amplitude(z.numpy()).T
            assert np.abs(got - want).max() < 1e-12


@pytest.mark.parametrize("ansatz,n,layers,n_a", [("cross_mesh", 6, 1, 3), ("layered", 5, 2, 3),
                                                  ("sim_circ_15", 7, 1, 4), ("cascade", 6, 1, 3)])
def test_factored_feature_contraction_matches_oracle(ansatz, n, layers, n_a):
    """Specification of the planned n > 4 feature engine: the Pauli
    coefficients from n axis contractions equal the Kronecker definition, and
    <Z_i>(z) = phiA^T C_i phiB (one GEMM over the B-half features + a dot with the A-half) equals
    the gate-by-gate statevector oracle."""
    g = torch.Generator().manual_seed(10 * n + layers)
    prog = P.compile_program(ansatz, n, layers, 1)
    th = torch.randn(layers, prog.params_per_layer, generator=g, dtype=torch.float64)
    O = fm.observables(fm.program_unitary(prog, th), n)
    C = fm.feature_tensor_angle_fast(O, n)
    assert C.shape == (n,) + (3,) * n
    if n <= 5:
        assert np.abs(C.reshape(n, -1) - fm.feature_matrix_angle(O, n)).max() < 1e-13
    z = torch.randn(7, n, generator=g, dtype=torch.float64)
    want = oc.quantum_layer(z, th, ansatz, n, "angle", oc.haar_for(1, n)).numpy()
    got = fm.factored_contraction(C, z.numpy(), n_a)
    assert np.abs(got - want).max() < 1e-12


def test_gate_table_layout():
    prog = P.compile_program("cascade", 4, 1, 1)
    ops = prog.ops.tolist()
    assert ops[0] == [P.RX, 0, -1, 0] and ops[8] == [P.CRX, 3, 0, 8] and ops[11] == [P.CRX, 0, 1, 11]
    assert ops[-3:] == [[P.U4, 0, 1, 0], [P.U4, 2, 3, 1], [P.HAD, 3, -1, -1]]
    assert prog.consts.shape == (2, 4, 4) and prog.ops.dtype == np.int32
    # Haar only with a seed AND n >= 4 (reference nn/DVQuantumLayer.py:88-94)
    assert P.compile_program("cascade", 3, 1, 1).consts.shape[0] == 0
    assert P.compile_program("cascade", 4, 1, None).consts.shape[0] == 0
    # second layer indexes the second row of the (L, P) angle matrix
    two = P.compile_program("layered", 4, 2, None)
    assert two.ops[two.ops[:, 3] >= 0][:, 3].tolist() == list(range(32))


def test_error_paths_match_reference():
    with pytest.raises(ValueError, match="Parameters are not initialized"):
        P.compile_program("nope", 4, 1)
    with pytest.raises(IndexError):           # alternate, even n
        P.compile_program("alternate", 4, 1)
    assert P.compile_program("alternate", 5, 1).n_theta == 16


def test_haar_pair_is_scipy_reproducible():
    u1, u2 = P.haar_pair(1)
    r1, r2 = oc.haar_unitaries(1)
    assert np.array_equal(u1, r1) and np.array_equal(u2, r2)


def test_generic_pde_operators_on_a_plain_module():
    """nn/pde.py operators with an ordinary torch module (no fused path): closed-form check on
    u = sin(a) * exp(b) (+ two more outputs for Navier-Stokes)."""
    import torch

    from qcpinn_b200.nn import pde

    class Analytic(torch.nn.Module):
        def forward(self, X):
            return torch.sin(X[:, 0:1]) * torch.exp(X[:, 1:2])

    a = torch.rand(7, 1, dtype=torch.float64)
    b = torch.rand(7, 1, dtype=torch.float64)
    u = torch.sin(a) * torch.exp(b)
    m = Analytic()
    uw, rw = pde.wave_operator(m, a.clone(), b.clone())
    assert torch.allclose(uw, u) and torch.allclose(rw, -u - 4.0 * u)
    uk, rk = pde.klein_gordon_operator(m, a.clone(), b.clone())
    assert torch.allclose(rk, -u - u + u ** 3)
    uh, rh = pde.helmholtz_operator(m, a.clone(), b.clone())
    assert torch.allclose(rh, -u + u + u)

    class ThreeOut(torch.nn.Module):
        def forward(self, X):
            t, x, y = X[:, 0:1], X[:, 1:2], X[:, 2:3]
            return torch.cat((t * x * y, t + x * x, x * y), 1)

    t, x, y = (torch.rand(5, 1, dtype=torch.float64) for _ in range(3))
    cont, fu, fv = pde.navier_stokes_2D_operator(ThreeOut(), t.clone(), x.clone(), y.clone())
    uu, vv = t * x * y, t + x * x
    assert torch.allclose(cont, t * y + 0 * x)
    assert torch.allclose(fu, x * y + uu * t * y + vv * t * x + y / 1056.0)
    assert torch.allclose(fv, 1 + uu * 2 * x + x / 1056.0 - 0.00345 * 2.0)


def test_single_file_trainer_cli_defaults_and_flat_alternate():
    """reference train_hybrid_qpinn.py:50-109 flag defaults; :273-295 the non-wrapping alternate."""
    import pytest

    from qcpinn_b200 import train_hybrid_qpinn as th
    from qcpinn_b200.program import CNOT, compile_program

    ns = th.parse_args([])
    assert (ns.device, ns.num_qubits, ns.ansatz, ns.encoding, ns.shots) == ("auto", 4, "cascade", "angle", 1024)
    assert (ns.epochs, ns.batch_size, ns.lr, ns.seed, ns.hidden_dim) == (5000, 64, 0.005, 42, 50)
    assert (ns.print_every, ns.output_dir, ns.diffusion_coef, ns.use_ibm) == (100, "./outputs", 0.01, False)
    with pytest.raises(NotImplementedError):
        th.main(["--use-ibm"])
    flat = compile_program("alternate", 4, 1, None, variant="single_file")
    pairs = [(a, b) for k, a, b, _ in flat.ops.tolist() if k == CNOT]
    assert pairs == [(0, 1), (2, 3), (1, 2)] and flat.n_theta == 12
    with pytest.raises(IndexError):
        compile_program("alternate", 4, 1, None)            # the DVQuantumLayer variant over-indexes


def _check_plan(ansatz, n, layers, seed, dtype_code):
    import ctypes

    import numpy as np

    import qcpinn_b200 as qb
    from qcpinn_b200.program import compile_program

    lib = qb._lib.load()
    prog = compile_program(ansatz, n, layers, seed)
    return _check_ops(n, dtype_code, prog.ops, prog.consts, prog.n_theta)


def _check_ops(n, dtype_code, ops, consts, n_theta):
    import ctypes

    from qcpinn_b200 import _lib

    lib = _lib.load()
    ops = np.ascontiguousarray(ops, dtype=np.int32)
    n_consts = len(consts)
    consts = np.ascontiguousarray(consts, dtype=np.complex128).view(np.float64)
    theta = np.random.default_rng(7).normal(size=max(n_theta, 1))
    err, nphys, nsw = ctypes.c_double(), ctypes.c_int(), ctypes.c_int()
    rc = lib.qcp_debug_check_plan(
        n, dtype_code, ops.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ops.shape[0],
        consts.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n_consts,
        theta.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n_theta,
        ctypes.byref(err), ctypes.byref(nphys), ctypes.byref(nsw))
    assert rc == 0, lib.qcp_last_error()
    return err.value, nphys.value, nsw.value


def test_statevector_planners_reproduce_the_logical_circuit():
    """Engine R (register layouts, SWAP / PERM, phase tables) and engine T (sweeps, evolving memory
    bit map, tile-bit controls) planners, checked on the CPU: the planned physical program applied
    to a random statevector equals the logical gate list, for every ansatz, both dtypes' layouts
    and qubit counts across both engines.  No CUDA device involved."""
    cases = []
    for ansatz in ("cascade", "layered", "farhi", "sim_circ_15", "cross_mesh"):
        for n, layers, seed in ((5, 1, None), (6, 2, 1), (8, 1, 1), (10, 2, None), (11, 1, 1), (13, 1, None)):
            cases.append((ansatz, n, layers, seed))
    for n in (5, 7, 9, 11, 13):
        cases.append(("alternate", n, 1, 1))
    cases.append(("sim_circ_15", 16, 2, None))          # BASELINE config 4
    cases.append(("cascade", 16, 1, 1))
    for ansatz, n, layers, seed in cases:
        for dtype_code in (0, 1):                          # QCP_F32 (LB = 5) / QCP_F64 (LB = 4)
            err, nphys, nsw = _check_plan(ansatz, n, layers, seed, dtype_code)
            assert err < 1e-12, (ansatz, n, layers, seed, dtype_code, err)
            assert nphys > 0
            tiled = n > 10
            assert (nsw > 0) == tiled


def test_tile_planner_gate_orders(monkeypatch):
    """Engine T plans sweeps either in program order or with the dependency-aware list scheduler
    (QCP_TILE_ORDER=program|dag; default = the one with fewer sweeps).  Both orders must reproduce
    the logical circuit, the default is never worse than either, and BASELINE config 4 needs at
    most half the program-order sweeps."""
    cases = [("sim_circ_15", 16, 2, None), ("sim_circ_15", 12, 2, 1), ("layered", 16, 1, None),
             ("layered", 12, 2, 1), ("cascade", 12, 2, None), ("farhi", 13, 2, None),
             ("cross_mesh", 13, 1, None), ("cross_mesh", 11, 2, 1), ("alternate", 13, 1, 1)]
    for ansatz, n, layers, seed in cases:
        for dtype_code in (0, 1):
            sweeps = {}
            for order in ("program", "dag", None):
                if order is None:
                    monkeypatch.delenv("QCP_TILE_ORDER", raising=False)
                else:
                    monkeypatch.setenv("QCP_TILE_ORDER", order)
                err, _, nsw = _check_plan(ansatz, n, layers, seed, dtype_code)
                assert err < 1e-12, (ansatz, n, layers, seed, dtype_code, order, err)
                sweeps[order] = nsw
            assert sweeps[None] == min(sweeps["program"], sweeps["dag"]), (ansatz, n, sweeps)
            if (ansatz, n, layers) == ("sim_circ_15", 16, 2):
                assert 2 * sweeps[None] <= sweeps["program"] + 1, sweeps


def test_planners_on_random_gate_programs(monkeypatch):
    """Random gate lists (all eight gate kinds, random wires, Haar-style two-qubit blocks) through
    both planners, both gate orders, with and without the diagonal-block tables of engine T: the
    dependency rules of the list scheduler (controls / RZ / CRZ / tables commute, everything else
    keeps its order) must hold for circuits that look nothing like the six ansaetze."""
    from scipy.stats import unitary_group

    rng = np.random.default_rng(2024)
    for trial in range(40):
        n = int(rng.integers(5, 14))
        n_gates = int(rng.integers(20, 70))
        ops, consts, n_theta = [], [], 0
        for _ in range(n_gates):
            kind = int(rng.choice([P.RX, P.RY, P.RZ, P.CRX, P.CRZ, P.CNOT, P.HAD, P.U4],
                                  p=[.17, .17, .14, .12, .14, .14, .06, .06]))
            a, b = (int(v) for v in rng.choice(n, size=2, replace=False))
            if kind in (P.RX, P.RY, P.RZ):
                ops.append([kind, a, -1, n_theta]); n_theta += 1
            elif kind in (P.CRX, P.CRZ):
                ops.append([kind, a, b, n_theta]); n_theta += 1
            elif kind == P.CNOT:
                ops.append([kind, a, b, -1])
            elif kind == P.HAD:
                ops.append([kind, a, -1, -1])
            else:
                ops.append([kind, a, b, len(consts)])
                consts.append(unitary_group.rvs(4, random_state=np.random.RandomState(trial * 100 + len(consts))))
        consts = np.array(consts, dtype=np.complex128).reshape(-1, 4, 4)
        for dtype_code in (0, 1):
            for order in ("program", "dag"):
                for diag in ("1", "0"):
                    monkeypatch.setenv("QCP_TILE_ORDER", order)
                    monkeypatch.setenv("QCP_REG_ORDER", order)
                    monkeypatch.setenv("QCP_TILE_DIAG", diag)
                    err, nphys, _ = _check_ops(n, dtype_code, ops, consts, n_theta)
                    assert err < 1e-11, (trial, n, dtype_code, order, diag, err)
                    assert nphys > 0
