"""Generate the golden fixtures in this directory from the CPU oracle.

    python tests/golden/make_golden.py

The reference itself cannot be imported (PennyLane / Qiskit are not installable here), so these
vectors pin the ORACLE (and through it the CUDA path), not PennyLane: see oracle/__init__.py
("parity unpinned").  Each ``*.pt`` holds the inputs, weights and float64 outputs of one small case:
expectation values of the stand-alone layer, the six Taylor streams, the residual, the trainer loss
terms and every parameter gradient.
"""

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import solver as osolver  # noqa: E402

CASES = [
    # name, ansatz, n, layers, encoding, haar_seed
    ("cascade4_quickstart", "cascade", 4, 1, "angle", None),
    ("cascade4_haar", "cascade", 4, 1, "angle", 1),
    ("layered4", "layered", 4, 1, "angle", None),
    ("cross_mesh4_l2", "cross_mesh", 4, 2, "angle", 1),
    ("farhi4_l2", "farhi", 4, 2, "angle", None),
    ("sim15_4", "sim_circ_15", 4, 1, "angle", 1),
    ("alternate3_l2", "alternate", 3, 2, "angle", None),
    ("cascade4_amplitude", "cascade", 4, 1, "amplitude", 1),
    ("layered2", "layered", 2, 1, "angle", None),
]


def build(name, ansatz, n, layers, enc, seed):
    w = osolver.init_weights(n, layers, ansatz, seed=3)
    g = torch.Generator().manual_seed(21)
    w["b1"] = 0.1 * torch.randn(w["b1"].shape, generator=g)
    w["b2"] = 0.1 * torch.randn(w["b2"].shape, generator=g) + (0.5 if enc == "amplitude" else 0.0)
    model = osolver.OracleSolver(n, layers, ansatz, enc, seed, "f64").set_weights(w)
    z = torch.randn(6, n, generator=g, dtype=torch.float64) + (0.8 if enc == "amplitude" else 0.0)
    batches = osolver.make_batches(12, seed=5, dtype=torch.float64)
    with torch.no_grad():
        q = model.quantum(z)
    streams = osolver.diffusion_streams(model, batches["X_res"]).detach()
    terms, grads = osolver.loss_and_grads(model, batches)
    return {
        "meta": {"ansatz": ansatz, "n": n, "layers": layers, "encoding": enc, "haar_seed": seed},
        "weights": w, "z": z, "q": q, "batches": batches, "streams": streams,
        "terms": {k: v.double() for k, v in terms.items()},
        "grads": {k: v.double() for k, v in grads.items()},
    }


if __name__ == "__main__":
    for case in CASES:
        torch.save(build(*case), os.path.join(HERE, case[0] + ".pt"))
        print("wrote", case[0])
