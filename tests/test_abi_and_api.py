"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, the Python surface mirrors the reference's, and the product path fails loudly without
CUDA (no oracle / CPU fallback)."""

import ast
import os
import re

import pytest
import torch

import qcpinn_b200 as qb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TRAINER = "/root/reference/trainer/diffusion_train.py"

ARGS = {
    "batch_size": 64, "epochs": 3, "lr": 0.005, "seed": 1, "print_every": 100,
    "num_qubits": 4, "num_quantum_layers": 1, "classic_network": [3, 50, 1],
    "q_ansatz": "cascade", "problem": "diffusion", "solver": "DV", "encoding": "None",
}


@pytest.fixture
def logger(tmp_path):
    lg = qb.Logging(str(tmp_path))
    yield lg
    lg.close()


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "qcpinn_b200.h")).read()
    declared = set(re.findall(r"\b(qcp_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = qb._lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/qcpinn_b200.h but not exported"
    assert declared == set(qb._lib.SIGNATURES), "ctypes table and header disagree"
    assert lib.qcp_version() >= 100


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "qcpinn-convection-diffusion-qiskit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_cuda(logger):
    with pytest.raises(RuntimeError, match="CUDA"):
        qb._lib.require_cuda()
    prog = qb.program.compile_program("cascade", 4, 1)
    with pytest.raises(RuntimeError):
        qb.functional.Plan(prog, 0, torch.float64, 50, torch.device("cpu"))
    model = qb.DVPDESolver(dict(ARGS), logger, device=torch.device("cpu"))
    with pytest.raises(RuntimeError):
        model.forward(torch.rand(4, 3))
    with pytest.raises(RuntimeError):
        qb.diffusion_operator(model, torch.rand(4, 1), torch.rand(4, 1), torch.rand(4, 1))


def test_solver_surface_matches_reference(logger):
    model = qb.DVPDESolver(dict(ARGS), logger, device=torch.device("cpu"))
    assert sorted(model.state_dict()) == sorted([
        "preprocessor.0.weight", "preprocessor.0.bias", "preprocessor.2.weight",
        "preprocessor.2.bias", "postprocessor.0.weight", "postprocessor.0.bias",
        "postprocessor.2.weight", "postprocessor.2.bias", "quantum_layer.params"])
    assert sum(p.numel() for p in model.parameters()) == 705 + 12
    assert model.quantum_layer.params.shape == (1, 12) and model.quantum_layer.params.dtype == torch.float32
    assert float(model.preprocessor[0].bias.detach().abs().max()) == 0.0          # zero-bias init, pre only
    assert model.quantum_layer.haar_seed1 == 1 and model.quantum_layer.haar_seed2 == 2
    assert isinstance(model.optimizer, torch.optim.Adam)
    assert model.scheduler.factor == 0.9 and model.scheduler.patience == 1000
    assert isinstance(model.loss_fn, torch.nn.MSELoss) and model.loss_history == []
    for key in ("batch_size", "num_qubits", "epochs", "classic_network", "lr"):
        bad = {k: v for k, v in ARGS.items() if k != key}
        with pytest.raises(KeyError):
            qb.DVPDESolver(bad, logger, device=torch.device("cpu"))


def test_non_2d_input_is_logged_then_raised(logger):
    model = qb.DVPDESolver(dict(ARGS), logger, device=torch.device("cpu"))
    with pytest.raises(ValueError, match="Expected 2D input tensor"):
        model.forward(torch.rand(3))
    log = open(os.path.join(logger.get_output_dir(), "output.log")).read()
    assert "Forward pass failed" in log


def test_quantum_layer_constructor_contract():
    for ansatz, count in [("layered", 16), ("alternate", 12), ("cascade", 12), ("farhi", 6),
                          ("sim_circ_15", 8), ("cross_mesh", 28)]:
        layer = qb.DVQuantumLayer({"num_qubits": 4, "num_quantum_layers": 2, "q_ansatz": ansatz,
                                   "problem": "diffusion"})
        assert layer.params.shape == (2, count)
        assert list(layer.state_dict()) == ["params"]
        assert layer.haar_seed1 is None                      # no "seed" key => no Haar block
    with pytest.raises(ValueError, match="Parameters are not initialized"):
        qb.DVQuantumLayer({"num_qubits": 4, "num_quantum_layers": 1, "q_ansatz": "x",
                           "problem": "diffusion"})
    with pytest.raises(KeyError):
        qb.DVQuantumLayer({"num_qubits": 4, "num_quantum_layers": 1, "q_ansatz": "cascade"})
    with pytest.raises(NotImplementedError):
        qb.DVQuantumLayer({"num_qubits": 4, "num_quantum_layers": 1, "q_ansatz": "cascade",
                           "problem": "diffusion", "use_ibm_hardware": True})
    even = qb.DVQuantumLayer({"num_qubits": 4, "num_quantum_layers": 1, "q_ansatz": "alternate",
                              "problem": "diffusion"})      # constructs, like the reference...
    with pytest.raises(IndexError):                          # ...and over-indexes when compiled
        even.program


def test_checkpoint_roundtrip(logger, tmp_path):
    model = qb.DVPDESolver(dict(ARGS), logger, device=torch.device("cpu"))
    model.loss_history = [1.0, 0.5]
    path = str(tmp_path / "m.pth")
    model.save_state(path)
    state = qb.DVPDESolver.load_state(path)
    assert set(state) == {"args", "classic_network", "quantum_params", "preprocessor",
                          "quantum_layer", "postprocessor", "optimizer", "scheduler",
                          "loss_history", "log_path"}
    other = qb.DVPDESolver(dict(ARGS), logger, device=torch.device("cpu")).restore(state)
    for a, b in zip(model.parameters(), other.parameters()):
        assert torch.equal(a, b)
    assert other.loss_history == [1.0, 0.5]


def test_sampler_and_targets():
    from qcpinn_b200.data.diffusion_dataset import Sampler, r, training_boxes, u

    boxes = training_boxes("cpu")
    x, y = Sampler(3, boxes["ics"], u, device="cpu").sample(10)
    assert x.shape == (10, 3) and y.shape == (10, 1) and float(x[:, 0].abs().max()) == 0.0
    x, y = Sampler(3, boxes["bc1"], u, device="cpu").sample(10)
    assert float(x[:, 1].abs().max()) == 0.0
    x, y = Sampler(3, boxes["dom"], r, device="cpu").sample(0)
    assert x.shape == (0, 3) and y.shape == (0, 1)


def test_logging_formats_like_reference(tmp_path):
    lg = qb.Logging(str(tmp_path), experiment_name="exp")
    lg.print("a: ", 1.5, " b")
    lg.print("single")
    lg.close()
    assert lg.get_output_dir().endswith("_exp")
    lines = open(os.path.join(lg.get_output_dir(), "output.log")).read().splitlines()
    assert lines == ["a: 1.5000e+00 b", "single"]


def test_reference_aliases_resolve_to_this_package():
    qb.install_reference_aliases(force=True)
    import importlib

    pde = importlib.import_module("nn.pde")
    ds = importlib.import_module("data.diffusion_dataset")
    assert pde.diffusion_operator is qb.diffusion_operator
    assert {"Sampler", "u", "r"} <= set(dir(ds))
    assert importlib.import_module("utils.logger").Logging is qb.Logging


@pytest.mark.skipif(not os.path.exists(REF_TRAINER), reason="reference checkout not present")
def test_everything_the_reference_trainer_touches_exists(logger):
    """Static drop-in check: every import and every ``model.<attr>`` used by the UNMODIFIED
    reference trainer/diffusion_train.py resolves against this package."""
    tree = ast.parse(open(REF_TRAINER).read())
    model = qb.DVPDESolver(dict(ARGS), logger, device=torch.device("cpu"))
    qb.install_reference_aliases(force=True)
    import importlib

    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module.split(".")[0] in ("data", "nn", "utils"):
            mod = importlib.import_module(node.module)
            for alias in node.names:
                assert hasattr(mod, alias.name), f"{node.module}.{alias.name}"
        if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id == "model":
            assert hasattr(model, node.attr), f"model.{node.attr}"
    for key in ("print_every", "solver"):
        assert key in model.args
