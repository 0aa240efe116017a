"""The reference's OWN ``nn/DVQuantumLayer.py`` / ``nn/DVPDESolver.py`` / ``nn/pde.py``, imported
unmodified from /root/reference and executed over ``tests/pennylane_stub.py`` (PennyLane itself is
not installable here), against this repository's gate compiler and oracle.

What this pins: the gate order, parameter indexing, control/target order, Haar rule, (n, B) output
orientation, error behaviour and the solver / residual plumbing of the reference source.  What it
cannot pin: PennyLane's internal conventions (the stub restates the published ones) -- that needs
``tests/golden/pennylane_*.pt`` (see ``make_pennylane_golden.py``).
CPU only; skipped where the reference checkout is absent (GPU boxes).
"""

import os
import sys

import numpy as np
import pytest
import torch

import pennylane_stub
import qcpinn_b200 as qb
from oracle import circuits as oc
from oracle import solver as osolver

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
P = qb.program
KIND = {"RX": P.RX, "RY": P.RY, "RZ": P.RZ, "CRX": P.CRX, "CRZ": P.CRZ, "CNOT": P.CNOT,
        "Hadamard": P.HAD, "QubitUnitary": P.U4}


@pytest.fixture(scope="module")
def ref():
    """The reference modules, imported from their checkout with the stand-in ``pennylane``."""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_pennylane_golden as gen

    saved = {k: v for k, v in sys.modules.items()
             if k.split(".")[0] in ("nn", "utils", "data", "trainer", "pennylane")}
    real_pennylane = not pennylane_stub.install()
    mods = gen.import_reference(REF)
    yield mods + (real_pennylane,)
    for k in list(sys.modules):
        if k.split(".")[0] in ("nn", "utils", "data", "trainer"):
            del sys.modules[k]
    if not real_pennylane:
        sys.modules.pop("pennylane", None)
    sys.modules.update(saved)
    if REF in sys.path:
        sys.path.remove(REF)


def _args(ansatz, n, layers, seed, enc="None"):
    a = {"batch_size": 8, "epochs": 1, "lr": 0.005, "print_every": 100, "num_qubits": n,
         "num_quantum_layers": layers, "classic_network": [3, 50, 1], "q_ansatz": ansatz,
         "problem": "diffusion", "solver": "DV", "encoding": enc}
    if seed is not None:
        a["seed"] = seed
    return a


CASES = [(a, n, L, s) for a in oc.ANSATZ_NAMES for n in (2, 3, 4, 5, 6) for L in (1, 2)
         for s in (None, 1)
         if not (a == "alternate" and n % 2 == 0) and not (a == "sim_circ_15" and n == 3)
         and not (a == "sim_circ_15" and n == 2)]


@pytest.mark.parametrize("ansatz,n,layers,seed", CASES)
def test_reference_gate_tape_equals_compiled_program(ref, ansatz, n, layers, seed):
    """Gate for gate: what the reference's circuit function emits == program.compile_program."""
    dvq, *_, real = ref
    if real:
        pytest.skip("tape inspection needs the stub")
    layer = dvq.DVQuantumLayer(_args(ansatz, n, layers, seed)).double()
    x = torch.randn(3, n, dtype=torch.float64)
    layer(x)
    tape = pennylane_stub.QNode.last_tape
    prog = P.compile_program(ansatz, n, layers, seed)
    # the encoding: RX(x[:, i]) on wire i, in wire order (reference nn/DVQuantumLayer.py:182)
    enc, body = tape[:n], tape[n:]
    for i, op in enumerate(enc):
        assert op.name == "RX" and op.wires == [i] and torch.equal(op.param, x[:, i])
    assert len(body) == prog.ops.shape[0]
    flat = layer.params.detach().reshape(-1)
    n_const = 0
    for op, (kind, a, b, p) in zip(body, prog.ops.tolist()):
        assert KIND[op.name] == kind, (op, kind)
        assert op.wires == ([a] if b < 0 else [a, b]), (op, a, b)
        if p >= 0 and kind != P.U4:
            assert float(op.param.detach()) == float(flat[p]), (op, p)       # same angle slot
        if kind == P.U4:
            assert np.array_equal(op.matrix, prog.consts[p])
            n_const += 1
    assert n_const == (2 if (seed is not None and n >= 4) else 0)


@pytest.mark.parametrize("ansatz,n,layers,seed", CASES)
@pytest.mark.parametrize("enc", ["None", "amplitude"])
def test_reference_layer_output_equals_oracle(ref, ansatz, n, layers, seed, enc):
    dvq = ref[0]
    layer = dvq.DVQuantumLayer(_args(ansatz, n, layers, seed, enc)).double()
    g = torch.Generator().manual_seed(7 * n + layers)
    x = torch.randn(5, n, generator=g, dtype=torch.float64) + (0.7 if enc == "amplitude" else 0.0)
    out = layer(x)
    assert out.shape == (n, 5) and out.dtype == torch.float64
    want = oc.quantum_layer(x, layer.params.detach(), ansatz, n,
                            "amplitude" if enc == "amplitude" else "angle", oc.haar_for(seed, n))
    assert float((out - want).abs().max()) < 1e-12
    # gradient w.r.t. the angles through the reference's own graph vs the oracle's
    out.sum().backward()
    th = layer.params.detach().clone().requires_grad_(True)
    oc.quantum_layer(x, th, ansatz, n, "amplitude" if enc == "amplitude" else "angle",
                     oc.haar_for(seed, n)).sum().backward()
    assert float((layer.params.grad - th.grad).abs().max()) < 1e-11


def test_reference_error_behaviour_matches(ref):
    dvq = ref[0]
    with pytest.raises(ValueError, match="Parameters are not initialized"):
        dvq.DVQuantumLayer(_args("nope", 4, 1, None))
    layer = dvq.DVQuantumLayer(_args("alternate", 4, 1, None))
    with pytest.raises(IndexError):
        layer(torch.randn(2, 4))
    with pytest.raises(IndexError):
        P.compile_program("alternate", 4, 1)


def test_reference_init_statistics_match(ref):
    """xavier_normal_ on (L, P): same generator stream as this package's layer and the oracle."""
    dvq = ref[0]
    for ansatz, n, layers in (("cascade", 4, 1), ("cross_mesh", 5, 2), ("farhi", 6, 3)):
        torch.manual_seed(11)
        want = dvq.DVQuantumLayer(_args(ansatz, n, layers, None)).params.detach().clone()
        torch.manual_seed(11)
        ours = qb.DVQuantumLayer(_args(ansatz, n, layers, None))
        assert torch.equal(ours.params.detach().cpu(), want)


@pytest.mark.parametrize("ansatz,n,layers,seed,enc", [
    ("cascade", 4, 1, None, "None"), ("cascade", 4, 1, 1, "amplitude"), ("layered", 4, 2, 1, "None"),
    ("cross_mesh", 5, 1, 1, "None"), ("sim_circ_15", 6, 1, None, "None"), ("farhi", 3, 2, None, "None"),
    ("alternate", 5, 1, 1, "None")])
def test_reference_solver_and_residual_equal_oracle(ref, tmp_path, ansatz, n, layers, seed, enc):
    """DVPDESolver.forward + nn/pde.py:diffusion_operator + the objective's gradients, reference
    code vs the oracle in the reference's own ("mixed") precision."""
    dvq, dvs, pde, logger_mod, _ = ref
    torch.manual_seed(5)
    model = dvs.DVPDESolver(_args(ansatz, n, layers, seed, enc), logger_mod.Logging(str(tmp_path)),
                            device=torch.device("cpu"))
    model.draw_quantum_circuit_flag = False
    with torch.no_grad():
        model.preprocessor[2].bias.add_(0.4)
    pre, post = model.preprocessor, model.postprocessor
    params = {"w1": pre[0].weight, "b1": pre[0].bias, "w2": pre[2].weight, "b2": pre[2].bias,
              "theta": model.quantum_layer.params, "w3": post[0].weight, "b3": post[0].bias,
              "w4": post[2].weight, "b4": post[2].bias}
    oracle = osolver.OracleSolver(n, layers, ansatz, "amplitude" if enc == "amplitude" else "angle",
                                  seed, "mixed").set_weights({k: v.detach() for k, v in params.items()})
    b = osolver.make_batches(12, seed=9)
    X = b["X_res"]
    out = model(X)
    assert out.shape == (12, 1) and out.dtype == torch.float32
    rel = lambda got, want: float((got.double() - want.double()).abs().max()
                                  / want.double().abs().max().clamp_min(1e-30))
    assert rel(out, oracle.forward(X)) < 2e-6
    t, x, y = (X[:, i:i + 1].clone() for i in range(3))
    u_ref, r_ref = pde.diffusion_operator(model, t, x, y)
    u_o, r_o = osolver.diffusion_operator(oracle, X[:, 0:1].clone(), X[:, 1:2].clone(), X[:, 2:3].clone())
    assert rel(u_ref, u_o) < 2e-6 and rel(r_ref, r_o) < 2e-5
    loss = 2 * model.loss_fn(r_ref, b["r_res"]) + 4 * model.loss_fn(model(b["X_bc"]), b["u_bc"]) \
        + 2 * model.loss_fn(model(b["X_ic"]), b["u_ic"])
    model.zero_grad()
    loss.backward()
    terms, grads = osolver.loss_and_grads(oracle, b)
    assert rel(loss, terms["loss"]) < 2e-6
    for k, p in params.items():
        assert rel(p.grad, grads[k]) < 5e-5, k
    with pytest.raises(ValueError, match="Expected 2D input"):
        model(torch.zeros(3))


# ---------------------------------------------------------------------------------------------------
# data re-uploading circuit family: reference hybrid_testing/CG_HQPINN_IBMtest_16qubits.py
# ---------------------------------------------------------------------------------------------------
CZ_SCRIPT = os.path.join(REF, "hybrid_testing", "CG_HQPINN_IBMtest_16qubits.py")


@pytest.fixture(scope="module")
def cz_script(ref):
    import importlib.util

    spec = importlib.util.spec_from_file_location("reference_cg_hqpinn_16q", CZ_SCRIPT)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod          # dataclasses resolve cls.__module__ through sys.modules
    spec.loader.exec_module(mod)
    yield mod
    sys.modules.pop(spec.name, None)


@pytest.mark.parametrize("n,layers", [(2, 1), (4, 1), (5, 3), (6, 2), (8, 2)])
def test_reference_reupload_layer_equals_program_and_oracle(ref, cz_script, n, layers):
    """The reference's own ``make_quantum_layer`` (RY encoding, RZ(0.5 x) re-upload inside every
    layer, Rot, CZ brick + ring; :217-253), executed unmodified: its gate tape equals
    ``compile_program("cz_melt")`` gate for gate (incl. which input feeds which re-upload gate and
    the 0.5 scale) and its outputs / gradients equal the oracle extension."""
    if ref[-1]:
        pytest.skip("tape inspection needs the stub")
    import pennylane as qml

    layer = cz_script.make_quantum_layer(qml.device("default.qubit", wires=n), n, layers, "backprop").double()
    g = torch.Generator().manual_seed(3 * n + layers)
    x = torch.randn(4, n, generator=g, dtype=torch.float64).requires_grad_(True)
    out = layer(x)                                                    # (B, n), one sample at a time
    w = layer.layer.weights                                           # (L, n, 3)
    assert out.shape == (4, n) and tuple(w.shape) == (layers, n, 3)
    tape = pennylane_stub.QNode.last_tape                             # the last sample's circuit
    prog = P.compile_program("cz_melt", n, layers)
    assert prog.n_theta == w.numel() and prog.reupload
    name = {P.RY: "RY", P.RZ: "RZ", P.CZ: "CZ", P.RY_IN: "RY", P.RZ_IN: "RZ"}
    assert len(tape) == prog.ops.shape[0]
    flat = w.detach().reshape(-1)
    for op, (kind, a, b, p) in zip(tape, prog.ops.tolist()):
        assert op.name == name[kind], (op, kind)
        assert op.wires == ([a, b] if kind == P.CZ else [a])
        if kind in (P.RY_IN, P.RZ_IN):
            assert abs(float(op.param.detach()) - 0.25 * p * float(x[-1, b])) < 1e-15
        elif kind != P.CZ:
            assert float(op.param.detach()) == float(flat[p])
    xo = x.detach().clone().requires_grad_(True)
    wo = w.detach().reshape(layers, 3 * n).clone().requires_grad_(True)
    want = oc.quantum_layer(xo, wo, "cz_melt", n)                    # (n, B)
    assert float((out.T - want).abs().max()) < 1e-12
    cot = torch.randn(4, n, generator=g, dtype=torch.float64)
    (out * cot).sum().backward()
    (want * cot.T).sum().backward()
    assert float((x.grad - xo.grad).abs().max()) < 1e-11
    assert float((w.grad.reshape(layers, 3 * n) - wo.grad).abs().max()) < 1e-11
