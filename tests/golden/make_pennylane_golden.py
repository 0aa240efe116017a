"""Generate ``pennylane_<case>.pt`` fixtures from the UNMODIFIED reference implementation.

    python tests/golden/make_pennylane_golden.py [--reference /root/reference] [--out tests/golden]

This is the script that PINS parity: it imports the reference's own ``nn/DVQuantumLayer.py``,
``nn/DVPDESolver.py`` and ``nn/pde.py`` (PennyLane ``default.qubit``, ``diff_method="backprop"``)
from a read-only checkout, loads the very weights / inputs of the nine oracle cases of
``make_golden.py`` into them and records what the reference computes:

* ``q``        ``DVQuantumLayer.forward(z)`` -- (n, B) float64 expectation values, with the layer
               cast to float64 (``layer.double()``; module API, no source change), so the 1e-10
               bar of the complex128 path can be checked against PennyLane itself;
* ``u``, ``streams``, ``residual`` -- ``DVPDESolver.forward`` and the nested ``autograd.grad`` chain
               of ``nn/pde.py:53-72`` (u, u_t, u_x, u_y, u_xx, u_yy) on the residual batch;
* ``terms``, ``grads`` -- the trainer's objective (``trainer/diffusion_train.py:30-49``) and
               d loss / d weight for all nine weight tensors.
  The solver-level records carry the reference's own precision: float32 MLPs and angles, complex128
  state, float32 cast of the expectation values (``nn/DVPDESolver.py:96``) = the oracle's "mixed"
  mode, so they are compared at the float32 bar.

With ``--stub`` the same script runs over ``tests/pennylane_stub.py`` (a tape recorder + small
simulator standing in for the PennyLane API) and writes ``refstub_<case>.pt`` instead: those pin
the reference's own circuit-construction / solver / residual code, executed unmodified, but not
``default.qubit`` (``meta["simulator"]`` says which).  That is what can be generated in the build
container, and those files ARE committed.

The real thing needs PennyLane (0.44.x, pulled in by ``pennylane-qiskit==0.44.1``,
``requirements-dev.txt:1``).  PennyLane is NOT installable in the build container or on the GPU
boxes (no network, no wheel), so the fixtures cannot be produced there: run this script on any
machine with ``pip install pennylane torch scipy`` and commit the resulting files.  Until then
``tests/test_pennylane_golden.py`` reports "parity unpinned" (skip) and DESIGN.md says so.
``qiskit_ibm_runtime`` / ``matplotlib`` are imported at module top by the reference but used only
on the IBM-hardware / drawing paths; when missing they are replaced by empty stand-in modules.
"""

import argparse
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

FORMAT_VERSION = 1


def pennylane_available():
    try:
        import pennylane  # noqa: F401
        return True
    except Exception:
        return False


def _stand_in(name, **attrs):
    if name in sys.modules:
        return
    try:
        __import__(name)
        return
    except Exception:
        pass
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    parent, _, child = name.rpartition(".")
    if parent:
        _stand_in(parent)
        setattr(sys.modules[parent], child, mod)


def import_reference(reference_root):
    """Import the reference's own modules (``nn.*``, ``utils.logger``) from its checkout."""
    for name in list(sys.modules):
        if name.split(".")[0] in ("nn", "utils", "data", "trainer"):
            del sys.modules[name]          # make sure no alias of this repo's package shadows them
    sys.path.insert(0, reference_root)
    _stand_in("qiskit_ibm_runtime", QiskitRuntimeService=object)
    _stand_in("matplotlib")
    _stand_in("matplotlib.pyplot")
    import importlib

    dvq = importlib.import_module("nn.DVQuantumLayer")
    dvs = importlib.import_module("nn.DVPDESolver")
    pde = importlib.import_module("nn.pde")
    logger = importlib.import_module("utils.logger")
    for mod in (dvq, dvs, pde, logger):
        assert os.path.abspath(mod.__file__).startswith(os.path.abspath(reference_root)), mod.__file__
    return dvq, dvs, pde, logger


def _args(ansatz, n, layers, enc, seed):
    args = {"batch_size": 12, "epochs": 1, "lr": 0.005, "print_every": 100, "num_qubits": n,
            "num_quantum_layers": layers, "classic_network": [3, 50, 1], "q_ansatz": ansatz,
            "problem": "diffusion", "solver": "DV",
            "encoding": "amplitude" if enc == "amplitude" else "None"}
    if seed is not None:
        args["seed"] = seed
    return args


def _load_weights(model, w):
    pre, post = model.preprocessor, model.postprocessor
    pairs = [(pre[0].weight, "w1"), (pre[0].bias, "b1"), (pre[2].weight, "w2"), (pre[2].bias, "b2"),
             (model.quantum_layer.params, "theta"), (post[0].weight, "w3"), (post[0].bias, "b3"),
             (post[2].weight, "w4"), (post[2].bias, "b4")]
    with torch.no_grad():
        for p, k in pairs:
            p.copy_(w[k].to(p.dtype))
    return {k: p for p, k in pairs}


def build_case(mods, tmp_dir, name, ansatz, n, layers, enc, seed):
    """One fixture: reference outputs on the inputs / weights of the oracle fixture ``name``."""
    dvq, dvs, pde, logger_mod = mods
    oracle_fix = torch.load(os.path.join(HERE, name + ".pt"), weights_only=False)
    w, z, batches = oracle_fix["weights"], oracle_fix["z"], oracle_fix["batches"]
    args = _args(ansatz, n, layers, enc, seed)

    # (a) stand-alone layer in float64
    layer = dvq.DVQuantumLayer(dict(args)).double()
    with torch.no_grad():
        layer.params.copy_(w["theta"].double())
        q = layer(z.double())
    q = torch.as_tensor(q).detach().double()

    # (b) solver, residual operator and the trainer's objective in the reference's own precision
    model = dvs.DVPDESolver(dict(args), logger_mod.Logging(tmp_dir), device=torch.device("cpu"))
    model.draw_quantum_circuit_flag = False
    params = _load_weights(model, w)
    f32 = {k: v.float() for k, v in batches.items()}
    cols = [f32["X_res"][:, i:i + 1].clone().requires_grad_(True) for i in range(3)]
    u = model(torch.cat(cols, 1))
    ones = torch.ones_like(u)
    first = [torch.autograd.grad(u, c, ones, create_graph=True)[0] for c in cols]
    u_xx = torch.autograd.grad(first[1], cols[1], ones, create_graph=True)[0]
    u_yy = torch.autograd.grad(first[2], cols[2], ones, create_graph=True)[0]
    streams = torch.cat([u] + first + [u_xx, u_yy], dim=1).detach().double()

    t_r, x_r, y_r = (f32["X_res"][:, i:i + 1].clone() for i in range(3))
    model.zero_grad()
    u_bc = model.forward(f32["X_bc"])
    u_ic = model.forward(f32["X_ic"])
    _, r_pred = pde.diffusion_operator(model, t_r, x_r, y_r)
    loss_r = model.loss_fn(r_pred, f32["r_res"])
    loss_bc = model.loss_fn(u_bc, f32["u_bc"])
    loss_ic = model.loss_fn(u_ic, f32["u_ic"])
    loss = 2.0 * loss_r + 4.0 * loss_bc + 2.0 * loss_ic
    loss.backward()
    grads = {k: (p.grad.detach().double() if p.grad is not None else torch.zeros_like(p).double())
             for k, p in params.items()}
    import pennylane as qml

    simulator = "pennylane_stub" if qml.__version__.endswith("stub") else "pennylane default.qubit"
    return {
        "format": FORMAT_VERSION,
        "meta": dict(oracle_fix["meta"], pennylane=qml.__version__, torch=torch.__version__,
                     simulator=simulator,
                     source="reference nn/DVQuantumLayer.py + nn/DVPDESolver.py + nn/pde.py, "
                            "default.qubit backprop"),
        "weights": w, "z": z, "batches": batches,
        "q": q, "streams": streams, "residual": r_pred.detach().double(),
        "terms": {"loss": loss.detach().double(), "loss_r": loss_r.detach().double(),
                  "loss_bc": loss_bc.detach().double(), "loss_ic": loss_ic.detach().double()},
        "grads": grads,
    }


REUPLOAD_CASES = [("reupload6_l2", 6, 2), ("reupload4_l1", 4, 1)]


def build_reupload_case(reference_root, name, n, layers):
    """The data re-uploading layer of reference hybrid_testing/CG_HQPINN_IBMtest_16qubits.py:217-253
    (its own ``make_quantum_layer``, unmodified): outputs (B, n) and the gradients of a fixed linear
    functional with respect to the inputs and the (L, n, 3) weights, in float64."""
    import importlib.util

    import pennylane as qml

    path = os.path.join(reference_root, "hybrid_testing", "CG_HQPINN_IBMtest_16qubits.py")
    spec = importlib.util.spec_from_file_location("reference_cg_hqpinn_16q_gen", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    torch.manual_seed(100 + n)
    layer = mod.make_quantum_layer(qml.device("default.qubit", wires=n), n, layers, "backprop").double()
    g = torch.Generator().manual_seed(7 + n)
    x = torch.randn(5, n, generator=g, dtype=torch.float64).requires_grad_(True)
    cot = torch.randn(5, n, generator=g, dtype=torch.float64)
    out = layer(x)
    (out * cot).sum().backward()
    weights = layer.layer.weights
    simulator = "pennylane_stub" if qml.__version__.endswith("stub") else "pennylane default.qubit"
    return {"format": FORMAT_VERSION,
            "meta": {"ansatz": "cz_melt", "n": n, "layers": layers, "simulator": simulator,
                     "pennylane": qml.__version__,
                     "source": "reference hybrid_testing/CG_HQPINN_IBMtest_16qubits.py make_quantum_layer"},
            "x": x.detach().clone(), "cot": cot, "theta": weights.detach().reshape(layers, 3 * n).clone(),
            "q": out.detach().clone(), "grad_x": x.grad.detach().clone(),
            "grad_theta": weights.grad.detach().reshape(layers, 3 * n).clone()}


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=HERE)
    ap.add_argument("--stub", action="store_true",
                    help="execute the reference over tests/pennylane_stub.py -> refstub_<case>.pt")
    ns = ap.parse_args(argv)
    prefix = "pennylane_"
    if ns.stub:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import pennylane_stub

        if not pennylane_stub.install():
            print("real PennyLane is importable: run without --stub")
            return 4
        prefix = "refstub_"
    if not pennylane_available():
        print("PennyLane is not importable here: no fixture written (parity stays unpinned). "
              "Run this script where `pip install pennylane` works.")
        return 3
    if not os.path.isdir(ns.reference):
        print(f"reference checkout {ns.reference} not found")
        return 2
    from make_golden import CASES
    import tempfile

    mods = import_reference(ns.reference)
    with tempfile.TemporaryDirectory() as tmp:
        for case in CASES:
            fix = build_case(mods, tmp, *case)
            path = os.path.join(ns.out, prefix + case[0] + ".pt")
            torch.save(fix, path)
            print("wrote", path)
    for name, n, layers in REUPLOAD_CASES:
        path = os.path.join(ns.out, prefix + name + ".pt")
        torch.save(build_reupload_case(ns.reference, name, n, layers), path)
        print("wrote", path)
    return 0


if __name__ == "__main__":
    sys.exit(main())
