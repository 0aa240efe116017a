"""Pin the CPU oracle with hand-derivable known-answer tests,
finite differences, invariants and the committed golden fixtures.  The reference ships no tests
or vectors for this path and cannot run here (PennyLane absent): "parity unpinned" means these
KATs are what stands between the oracle and PennyLane's published gate conventions."""

import glob
import math
import os

import numpy as np
import pytest
import torch

from oracle import circuits as oc
from oracle import dataset as od
from oracle import solver as osolver

torch.set_default_dtype(torch.float32)
D = torch.float64
# oracle-generated fixtures only (refstub_* / pennylane_* have their own tests: test_reference_fixtures.py)
GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt"))
                if not os.path.basename(p).startswith(("refstub_", "pennylane_")))


def _x(rows):
    return torch.tensor(rows, dtype=D)


def test_kat1_zero_params_cascade_reports_cos_and_zero_on_last_wire():
    x = _x([[0.3, 1.1, -0.7, 2.0], [0.1, 0.2, 0.3, 0.4]])
    q = oc.quantum_layer(x, torch.zeros(1, 12, dtype=D), "cascade", 4)
    want = torch.cos(x).T.clone()
    want[3] = 0.0        # H on the last wire turns <Z> into <X> of RX(x)|0> = 0
    assert torch.allclose(q, want, atol=1e-14)


def test_kat2_rx_composes_with_encoding():
    x = _x([[0.3, 1.1, -0.7, 2.0]])
    p = torch.zeros(1, 12, dtype=D)
    p[0, 0] = 0.5
    q = oc.quantum_layer(x, p, "cascade", 4)
    assert abs(q[0, 0].item() - math.cos(0.3 + 0.5)) < 1e-14


def test_kat3_rz_on_last_wire():
    x = _x([[0.3, 1.1, -0.7, 2.0]])
    p = torch.zeros(1, 12, dtype=D)
    p[0, 7] = 0.9
    q = oc.quantum_layer(x, p, "cascade", 4)
    assert abs(q[3, 0].item() - math.sin(2.0) * math.sin(0.9)) < 1e-14


def test_kat4_crx_control_target_order():
    p = torch.zeros(1, 6, dtype=D)
    p[0, 4] = 0.77                                  # CRX(wires=[n-1, 0])
    q = oc.quantum_layer(_x([[0.0, math.pi]]), p, "cascade", 2)
    assert torch.allclose(q[:, 0], _x([math.cos(0.77), 0.0]), atol=1e-14)
    p = torch.zeros(1, 6, dtype=D)
    p[0, 5] = 0.77                                  # CRX(wires=[0, 1])
    q = oc.quantum_layer(_x([[math.pi, 0.0]]), p, "cascade", 2)
    assert torch.allclose(q[:, 0], _x([-1.0, 0.0]), atol=1e-14)


def test_kat5_amplitude_embedding_pads_and_normalises():
    st = oc.encode(_x([[3.0, 4.0]]), 2, "amplitude", torch.complex128)
    assert torch.allclose(st.real, _x([[0.6, 0.8, 0.0, 0.0]]), atol=1e-15)
    assert float(st.imag.abs().max()) == 0.0


def test_kat6_haar_unitaries_are_scipys():
    u1, u2 = oc.haar_unitaries(1)
    assert abs(u1[0, 0] - (0.67931743 - 0.0721112j)) < 1e-7
    for u in (u1, u2):
        assert np.abs(u.conj().T @ u - np.eye(4)).max() < 1e-12
    assert oc.haar_for(1, 3) is None and oc.haar_for(None, 4) is None


@pytest.mark.parametrize("ansatz", oc.ANSATZ_NAMES)
@pytest.mark.parametrize("enc", ["angle", "amplitude"])
def test_norm_and_bounds(ansatz, enc):
    n = 5 if ansatz == "alternate" else 4
    g = torch.Generator().manual_seed(1)
    th = torch.randn(2, oc.params_per_layer(ansatz, n), generator=g, dtype=D)
    x = torch.randn(7, n, generator=g, dtype=D)
    st = oc.final_state(x, th, ansatz, n, enc, oc.haar_for(1, n))
    assert float(((st.abs() ** 2).sum(1) - 1).abs().max()) < 1e-12
    q = oc.expval_z(st, n)
    assert q.shape == (n, 7) and float(q.abs().max()) <= 1 + 1e-12


def test_param_counts_and_error_paths():
    assert [oc.params_per_layer(a, 4) for a in oc.ANSATZ_NAMES] == [16, 12, 12, 6, 8, 28]
    with pytest.raises(ValueError):
        oc.params_per_layer("nope", 4)
    with pytest.raises(IndexError):   # alternate over-indexes its row for even n
        oc.quantum_layer(torch.zeros(1, 4, dtype=D), torch.zeros(1, 12, dtype=D), "alternate", 4)
    with pytest.raises(ValueError):
        oc.quantum_layer(torch.zeros(1, 4, dtype=D), torch.zeros(1, 5, dtype=D), "farhi", 4)


def test_taylor_streams_match_finite_differences():
    w = osolver.init_weights(4, 1, "cascade", seed=2)
    m = osolver.OracleSolver(4, 1, "cascade", "angle", 1, "f64").set_weights(w, requires_grad=False)
    X = torch.rand(5, 3, dtype=D)
    s = osolver.diffusion_streams(m, X)
    h = 1e-4

    def u(Xp):
        return m.forward(Xp)[:, 0]

    for d in range(3):
        e = torch.zeros(3, dtype=D)
        e[d] = h
        fd1 = (u(X + e) - u(X - e)) / (2 * h)
        assert torch.allclose(s[:, 1 + d], fd1, atol=1e-7)
        if d > 0:
            fd2 = (u(X + e) - 2 * u(X) + u(X - e)) / h ** 2
            assert torch.allclose(s[:, 3 + d], fd2, atol=1e-5)


def test_theta_gradient_matches_finite_differences():
    w = osolver.init_weights(4, 1, "layered", seed=4)
    m = osolver.OracleSolver(4, 1, "layered", "angle", None, "f64").set_weights(w)
    b = osolver.make_batches(6, seed=3, dtype=D)
    _, grads = osolver.loss_and_grads(m, b)
    h = 1e-6
    for idx in (0, 5, 11):
        wp = {k: v.clone() for k, v in w.items()}
        wm = {k: v.clone() for k, v in w.items()}
        wp["theta"] = wp["theta"].double(); wm["theta"] = wm["theta"].double()
        wp["theta"].view(-1)[idx] += h
        wm["theta"].view(-1)[idx] -= h
        lp = osolver.loss_terms(osolver.OracleSolver(4, 1, "layered", mode="f64").set_weights(wp), b)[0]
        lm = osolver.loss_terms(osolver.OracleSolver(4, 1, "layered", mode="f64").set_weights(wm), b)[0]
        fd = (lp - lm).item() / (2 * h)
        assert abs(fd - grads["theta"].view(-1)[idx].item()) < 1e-6 * max(1.0, abs(fd))


def test_dataset_forcing_follows_the_reference_closed_form():
    """The reference's u_xx / u_yy use (40000 d^2 - 400) u where calculus gives (40000 d^2 - 200) u
    (data/diffusion_dataset.py:30-33).  The forcing is only a regression target, so the quirk is
    reproduced: r_reference = r_true + D * 400 * u."""
    X = torch.rand(9, 3, dtype=D).requires_grad_(True)
    u = od.u_exact(X)
    g = torch.autograd.grad(u.sum(), X, create_graph=True)[0]
    uxx = torch.autograd.grad(g[:, 1].sum(), X, retain_graph=True)[0][:, 1]
    uyy = torch.autograd.grad(g[:, 2].sum(), X)[0][:, 2]
    r_true = g[:, 0] + g[:, 1] + g[:, 2] - 0.01 * (uxx + uyy)
    want = (r_true + 0.01 * 400.0 * u[:, 0]).detach()
    assert torch.allclose(od.forcing(X.detach())[:, 0], want, atol=1e-12)


def test_flop_model_reproduces_baseline_md():
    assert osolver.flops_per_point(4, 1, "cascade", False)[0] == 83_200
    assert osolver.flops_per_point(4, 1, "cascade", True)[0] == 102_400
    assert osolver.flops_per_point(4, 1, "layered", False)[0] == 99_840
    assert osolver.flops_per_point(10, 2, "cross_mesh", False)[1] == 1_543_168
    assert osolver.flops_per_point(16, 2, "sim_circ_15", False)[1] == 75_563_008


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_oracle_reproduces_golden_fixture(path):
    gold = torch.load(path, weights_only=False)
    m = gold["meta"]
    model = osolver.OracleSolver(m["n"], m["layers"], m["ansatz"], m["encoding"], m["haar_seed"],
                                 "f64").set_weights(gold["weights"])
    with torch.no_grad():
        q = model.quantum(gold["z"])
    assert torch.allclose(q, gold["q"], atol=1e-13)
    assert torch.allclose(osolver.diffusion_streams(model, gold["batches"]["X_res"]).detach(),
                          gold["streams"], rtol=1e-11, atol=1e-12)
    terms, grads = osolver.loss_and_grads(model, gold["batches"])
    for k, v in gold["terms"].items():
        assert abs(terms[k].item() - v.item()) < 1e-12 * max(1.0, abs(v.item()))
    for k, v in gold["grads"].items():
        assert torch.allclose(grads[k], v, rtol=1e-10, atol=1e-12), k


def test_golden_directory_is_populated():
    assert len(GOLDEN) >= 9
