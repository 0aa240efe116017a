"""GPU tests through the reference-shaped Python surface (DVQuantumLayer / DVPDESolver /
diffusion_operator / trainer), the golden fixtures, edge cases and full-size properties."""

import glob
import math
import os

import pytest
import torch

import qcpinn_b200 as qb
from helpers import F, TOL, device_weights, make_case, mlp_list, points, rel_err
from oracle import solver as osolver

pytestmark = pytest.mark.gpu
# oracle-generated fixtures only (refstub_* / pennylane_* have their own tests: test_reference_fixtures.py)
GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt"))
                if not os.path.basename(p).startswith(("refstub_", "pennylane_")))
DEV = torch.device("cuda", 0)

ARGS = {
    "batch_size": 64, "epochs": 4, "lr": 0.005, "seed": 1, "print_every": 2,
    "num_qubits": 4, "num_quantum_layers": 1, "classic_network": [3, 50, 1],
    "q_ansatz": "cascade", "problem": "diffusion", "solver": "DV", "encoding": "None",
}


def _model(tmp_path, **over):
    torch.manual_seed(0)
    args = dict(ARGS, **over)
    return qb.DVPDESolver(args, qb.Logging(str(tmp_path)), device=DEV)


def _weights_of(model):
    pre, post = model.preprocessor, model.postprocessor
    return {"w1": pre[0].weight, "b1": pre[0].bias, "w2": pre[2].weight, "b2": pre[2].bias,
            "theta": model.quantum_layer.params,
            "w3": post[0].weight, "b3": post[0].bias, "w4": post[2].weight, "b4": post[2].bias}


def _oracle_of(model, mode="f64"):
    a = model.args
    w = {k: v.detach().cpu() for k, v in _weights_of(model).items()}
    return osolver.OracleSolver(a["num_qubits"], a["num_quantum_layers"], a["q_ansatz"],
                                "amplitude" if a.get("encoding") == "amplitude" else "angle",
                                a.get("seed"), mode).set_weights(w)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_cuda_reproduces_golden_fixture(path, dtype):
    gold = torch.load(path, weights_only=False)
    m = gold["meta"]
    prog = qb.program.compile_program(m["ansatz"], m["n"], m["layers"], m["haar_seed"])
    plan = F.Plan(prog, F.encoding_code(m["encoding"]), dtype, 50, DEV)
    dw = device_weights(gold["weights"], dtype, DEV, requires_grad=True)
    tol = TOL[dtype]
    q = F.layer_apply(plan, gold["z"].to(DEV, dtype), dw["theta"])
    assert rel_err(q, gold["q"]) < tol
    b = {k: v.to(DEV, dtype) for k, v in gold["batches"].items()}
    coeffs = (1.0, 1.0, 1.0, -0.01, -0.01)
    streams = F.solver_streams(plan, b["X_res"], dw["theta"], mlp_list(dw), coeffs)[2]
    for c in range(6):
        assert rel_err(streams[:, c], gold["streams"][:, c]) < tol, c
    # the trainer's objective, assembled from the three fused calls
    u_bc = F.solver_value(plan, b["X_bc"], dw["theta"], mlp_list(dw))
    u_ic = F.solver_value(plan, b["X_ic"], dw["theta"], mlp_list(dw))
    _, r = F.solver_residual(plan, b["X_res"], dw["theta"], mlp_list(dw), coeffs)
    mse = lambda a, t: ((a - t) ** 2).mean()
    loss = 2 * mse(r, b["r_res"]) + 4 * mse(u_bc, b["u_bc"]) + 2 * mse(u_ic, b["u_ic"])
    loss.backward()
    assert abs(loss.item() - gold["terms"]["loss"].item()) < tol * abs(gold["terms"]["loss"].item())
    for k, g in gold["grads"].items():
        assert rel_err(dw[k].grad, g) < tol, k


def test_quantum_layer_module_matches_oracle():
    layer = qb.DVQuantumLayer({"num_qubits": 4, "num_quantum_layers": 2, "q_ansatz": "cross_mesh",
                               "problem": "diffusion", "seed": 1}).to(DEV)
    x = torch.randn(19, 4, device=DEV)
    out = layer(x)
    assert out.shape == (4, 19) and out.dtype == torch.float64      # (n, B) float64 like PennyLane
    from oracle import circuits as oc
    want = oc.quantum_layer(x.cpu().double(), layer.params.detach().cpu().double(), "cross_mesh",
                            4, "angle", oc.haar_for(1, 4))
    assert rel_err(out, want) < 1e-10
    out.sum().backward()
    assert layer.params.grad is not None and layer.params.grad.shape == (2, 28)
    single = layer(x[0])
    assert single.shape == (4,) and rel_err(single, want[:, 0]) < 1e-10


def test_solver_module_forward_and_residual_match_oracle(tmp_path):
    model = _model(tmp_path)
    oracle = _oracle_of(model)
    X = points(40).float().to(DEV)
    u = model(X)
    assert u.shape == (40, 1) and u.dtype == torch.float32
    assert rel_err(u, oracle.forward(X.cpu().double())) < 1e-6      # float32 cast of the output
    t, x, y = (X[:, i:i + 1].clone() for i in range(3))
    uu, r = qb.diffusion_operator(model, t, x, y)
    assert t.requires_grad and x.requires_grad and y.requires_grad
    Xc = X.cpu().double()
    uo, ro = osolver.diffusion_operator(oracle, Xc[:, 0:1].clone(), Xc[:, 1:2].clone(), Xc[:, 2:3].clone())
    assert rel_err(uu, uo) < 1e-6 and rel_err(r, ro) < 1e-6
    # the module reads / writes float32 tensors (float64 arithmetic inside): float32 rounding only
    assert rel_err(model.taylor_streams(X), osolver.diffusion_streams(oracle, Xc)) < 1e-6
    log = open(os.path.join(model.log_path, "output.log")).read()
    assert "The circuit used in the study:" in log and "CRX[3,0] theta[8]" in log


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_train_step_loss_and_gradients_match_oracle(tmp_path, dtype):
    """One objective of the trainer (3 model calls + weighted MSE) against the oracle's step."""
    from qcpinn_b200.trainer.diffusion_train import TrainStep

    model = _model(tmp_path, dtype=dtype)
    oracle = _oracle_of(model)
    batches = osolver.make_batches(96, seed=11)
    terms, grads = osolver.loss_and_grads(oracle, batches)
    step = TrainStep(model, 96)
    batch = tuple(batches[k].to(DEV) for k in ("X_ic", "u_ic", "X_bc", "u_bc", "X_res", "r_res"))
    loss, _, lr_, lbc, lic = step.objective(batch)
    loss.backward()
    tol = 2e-5 if dtype == "float32" else 2e-6       # outputs/loss pass through float32
    assert abs(loss.item() - terms["loss"].item()) < tol * terms["loss"].item()
    assert abs(lr_.item() - terms["loss_r"].item()) < tol * terms["loss_r"].item()
    for k, p in _weights_of(model).items():
        assert rel_err(p.grad, grads[k]) < 10 * tol, k
    assert batch[0].grad is not None      # X_ics.requires_grad_(True) in the trainer gets a grad


def test_trainer_runs_and_matches_oracle_trajectory(tmp_path):
    """Five optimisation steps on fixed batches: the loss history follows the CPU oracle."""
    from qcpinn_b200.trainer.diffusion_train import TrainStep

    model = _model(tmp_path, seed=None)
    oracle = _oracle_of(model, "f64")
    otr = osolver.OracleTrainer(oracle, lr=0.005)
    step = TrainStep(model, 48)
    for i in range(5):
        b = osolver.make_batches(48, seed=100 + i)
        want = otr.step(b)
        got = step(tuple(b[k].to(DEV) for k in ("X_ic", "u_ic", "X_bc", "u_bc", "X_res", "r_res")))
        assert abs(got - want) < 5e-5 * abs(want), (i, got, want)
    assert len(model.loss_history) == 5


def test_train_function_end_to_end(tmp_path):
    from qcpinn_b200.trainer import diffusion_train

    model = _model(tmp_path)
    before = [p.detach().clone() for p in model.parameters()]
    diffusion_train.train(model, batch_size=128)
    assert len(model.loss_history) == model.epochs + 1
    assert all(torch.isfinite(torch.tensor(model.loss_history)))
    assert any(not torch.equal(a, b) for a, b in zip(before, model.parameters()))
    assert os.path.exists(os.path.join(model.log_path, "model.pth"))      # print_every checkpoint
    log = open(os.path.join(model.log_path, "output.log")).read()
    assert "Starting training for 4 epochs" in log and "Epoch: 4/4" in log and "Training completed" in log
    state = qb.DVPDESolver.load_state(os.path.join(model.log_path, "model.pth"))
    assert state["quantum_layer"]["params"].shape == (1, 12)


@pytest.mark.parametrize("batch", [0, 1, 31, 32, 33, 127, 129, 1000])
def test_ragged_and_empty_batches(batch):
    w, oracle, prog = make_case("cascade", 4, 1, "angle", 1)
    plan = F.Plan(prog, 0, torch.float64, 50, DEV)
    dw = device_weights(w, torch.float64, DEV, requires_grad=True)
    X = points(max(batch, 1), seed=batch)[:batch]
    u, r = F.solver_residual(plan, X.to(DEV), dw["theta"], mlp_list(dw), (1, 1, 1, -0.01, -0.01))
    assert u.shape == (batch, 1) and r.shape == (batch, 1)
    (u.sum() + 2 * r.sum()).backward()
    if batch == 0:
        assert float(dw["w1"].grad.abs().max()) == 0.0 and float(dw["theta"].grad.abs().max()) == 0.0
        return
    uo, ro = osolver.diffusion_operator(oracle, X[:, 0:1].clone(), X[:, 1:2].clone(), X[:, 2:3].clone())
    (uo.sum() + 2 * ro.sum()).backward()
    assert rel_err(r, ro) < 1e-10
    for k in osolver.WEIGHT_NAMES:
        assert rel_err(dw[k].grad, oracle.w[k].grad) < 1e-10, k


def test_other_hidden_widths_and_pde_coefficients():
    w, oracle, prog = make_case("layered", 3, 1, "angle", None, hidden=17)
    plan = F.Plan(prog, 0, torch.float64, 17, DEV)
    dw = device_weights(w, torch.float64, DEV, requires_grad=True)
    X = points(21)
    kw = dict(sigma_t=2.0, sigma_x=0.5, sigma_y=4.0, D=0.3, v_x=-1.5, v_y=0.25)
    uo, ro = osolver.diffusion_operator(oracle, X[:, 0:1].clone(), X[:, 1:2].clone(), X[:, 2:3].clone(), **kw)
    from qcpinn_b200.nn.pde import _diffusion_coeffs
    coeffs = _diffusion_coeffs(kw["sigma_t"], kw["sigma_x"], kw["sigma_y"], kw["D"], kw["v_x"], kw["v_y"])
    u, r = F.solver_residual(plan, X.to(DEV), dw["theta"], mlp_list(dw), coeffs)
    assert rel_err(r, ro) < 1e-10 and rel_err(u, uo) < 1e-10


def test_full_size_properties_4m_points():
    """BASELINE config 5 size (4 194 304 points): size-independent properties.
    (a) the residual kernel is deterministic and finite; (b) gradients are additive over a
    partition of the batch (what data-parallel sharding relies on); (c) a strided sample agrees
    with the oracle."""
    n_pts = 4_194_304
    w, oracle, prog = make_case("cascade", 4, 1, "angle", None)
    plan = F.Plan(prog, 0, torch.float64, 50, DEV)
    dw = device_weights(w, torch.float64, DEV)
    g = torch.Generator(device=DEV).manual_seed(5)
    X = torch.rand(n_pts, 3, device=DEV, dtype=torch.float64, generator=g)
    gr = torch.rand(n_pts, device=DEV, dtype=torch.float64, generator=g) / n_pts
    coeffs = (1.0, 1.0, 1.0, -0.01, -0.01)
    theta = dw["theta"].reshape(-1)
    mlp = mlp_list(dw)
    plan.prepare(theta)
    u1, r1, _ = plan.solver_forward(X, mlp, F.MODE_RESIDUAL, coeffs)
    u2, r2, _ = plan.solver_forward(X, mlp, F.MODE_RESIDUAL, coeffs)
    assert torch.equal(r1, r2) and torch.equal(u1, u2) and bool(torch.isfinite(r1).all())
    full, _ = plan.solver_backward(X, mlp, theta, None, gr, F.MODE_RESIDUAL, coeffs)
    full = [v.clone() for v in full]
    half = n_pts // 2
    a, _ = plan.solver_backward(X[:half].contiguous(), mlp, theta, None, gr[:half].contiguous(),
                                F.MODE_RESIDUAL, coeffs)
    a = [v.clone() for v in a]
    b, _ = plan.solver_backward(X[half:].contiguous(), mlp, theta, None, gr[half:].contiguous(),
                                F.MODE_RESIDUAL, coeffs)
    for f_, a_, b_ in zip(full, a, b):
        assert rel_err(a_ + b_, f_) < 1e-10
    idx = torch.arange(0, n_pts, n_pts // 64, device=DEV)
    Xs = X[idx].cpu()
    _, ro = osolver.diffusion_operator(oracle, Xs[:, 0:1].clone(), Xs[:, 1:2].clone(), Xs[:, 2:3].clone())
    assert rel_err(r1[idx], ro[:, 0]) < 1e-10


def test_abi_error_reporting():
    lib = qb._lib.load()
    import ctypes
    handle = ctypes.c_void_p()
    ops = (ctypes.c_int32 * 4)(0, 9, -1, 0)            # wire 9 on a 4-qubit plan
    rc = lib.qcp_plan_create(ctypes.byref(handle), 4, 0, 1, 50, ops, 1, None, 0, 1)
    assert rc != 0 and b"invalid" in lib.qcp_last_error()
    rc = lib.qcp_plan_create(ctypes.byref(handle), 17, 0, 1, 50, ops, 0, None, 0, 0)
    assert rc != 0 and b"qubits" in lib.qcp_last_error()
    rc = lib.qcp_plan_create(ctypes.byref(handle), 1, 0, 1, 50, ops, 0, None, 0, 0)
    assert rc != 0 and b"qubits" in lib.qcp_last_error()
    prog = qb.program.compile_program("cascade", 4, 1)
    plan = F.Plan(prog, 0, torch.float64, 50, DEV)
    with pytest.raises(RuntimeError, match="qcp_prepare"):
        plan.layer_forward(torch.zeros(3, 4, dtype=torch.float64, device=DEV))
    with pytest.raises(ValueError):
        plan.prepare(torch.zeros(5, dtype=torch.float64, device=DEV))
    with pytest.raises(RuntimeError, match="plan on"):
        plan._t(torch.zeros(3))


def test_fused_sampler_matches_torch_expressions():
    """Sampler.sample's CUDA fast path (one kernel) vs the module's torch expressions for u / r."""
    from qcpinn_b200.data.diffusion_dataset import Sampler, r, training_boxes, u

    boxes = training_boxes(DEV)
    for func, box in ((u, "ics"), (u, "bc1"), (r, "dom")):
        torch.manual_seed(3)
        X, y = Sampler(3, boxes[box], func, device=DEV).sample(4097)
        torch.manual_seed(3)
        rnd = torch.rand(4097, 3, device=DEV)
        lo, hi = boxes[box][0:1], boxes[box][1:2]
        Xt = lo + (hi - lo) * rnd
        assert torch.equal(X, Xt)
        assert y.shape == (4097, 1)
        assert rel_err(y, func(Xt.double())) < 1e-5     # float32 evaluation of exp(-100 d^2)
    # generic callables still take the torch path
    X, y = Sampler(3, boxes["dom"], lambda p: p.sum(1, keepdim=True), device=DEV).sample(5)
    assert torch.allclose(y, X.sum(1, keepdim=True))


def test_evaluation_grid_matches_oracle(tmp_path):
    """Reference evaluation block (20^3 grid, no backward, rel-L2 in percent)."""
    from qcpinn_b200.trainer.diffusion_eval import evaluate, evaluation_grid

    model = _model(tmp_path)
    out = evaluate(model, num_points=8)
    X = evaluation_grid(8)
    assert out["X_star"].shape == (512, 3) and torch.equal(out["X_star"].cpu(), X)
    assert torch.equal(X[1], torch.tensor([0.0, 0.0, 1.0 / 7.0]))        # y fastest, t slowest
    oracle = _oracle_of(model)
    Xd = X.double()
    uo, ro = osolver.diffusion_operator(oracle, Xd[:, 0:1].clone(), Xd[:, 1:2].clone(), Xd[:, 2:3].clone())
    from oracle import dataset as od
    eu = float(torch.linalg.norm(od.u_exact(Xd) - uo) / torch.linalg.norm(od.u_exact(Xd))) * 100
    ef = float(torch.linalg.norm(od.forcing(Xd) - ro) / torch.linalg.norm(od.forcing(Xd) + 1e-9)) * 100
    assert abs(out["error_u"] - eu) < 1e-3 * eu and abs(out["error_f"] - ef) < 1e-3 * ef
    log = open(os.path.join(model.log_path, "output.log")).read()
    assert "Relative L2 error_u:" in log and "Relative L2 error_f:" in log


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_forward_many_equals_separate_calls(tmp_path, dtype):
    """One autograd node for the three model calls (shared reduction) == three separate nodes."""
    from qcpinn_b200.trainer.diffusion_train import TrainStep

    batches = osolver.make_batches(200, seed=21)
    batch = [batches[k].to(DEV) for k in ("X_ic", "u_ic", "X_bc", "u_bc", "X_res", "r_res")]
    grads = {}
    for fuse in (True, False):
        model = _model(tmp_path, dtype=dtype, fuse_model_calls=fuse, cuda_graph=False)
        step = TrainStep(model, 200)
        loss, *_ = step.objective(tuple(t.clone() for t in batch))
        loss.backward()
        grads[fuse] = (loss.item(), [p.grad.clone() for p in model.parameters()])
    assert abs(grads[True][0] - grads[False][0]) < 1e-6 * abs(grads[False][0])
    tol = 1e-9 if dtype == "float64" else 2e-5
    for a, b in zip(grads[True][1], grads[False][1]):
        assert rel_err(a, b) < max(tol, 2e-6)       # grads are float32 tensors


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_fused_step_equals_autograd_route(tmp_path, dtype):
    """TrainStep's graph-free fast path (train_step_grads + qcp_mse_seed + qcp_clip_grads) against
    the autograd route (objective -> loss.backward -> clip_grad_norm_): same loss terms, same
    clipped gradients, same parameters after Adam."""
    from qcpinn_b200.trainer.diffusion_train import DIFFUSION_COEFFS, TrainStep

    batches = osolver.make_batches(96, seed=21)
    batch = tuple(batches[k].to(DEV) for k in ("X_ic", "u_ic", "X_bc", "u_bc", "X_res", "r_res"))
    ma = _model(tmp_path / "a", dtype=dtype)
    mf = _model(tmp_path / "f", dtype=dtype)
    mf.load_state_dict(ma.state_dict())
    sa = TrainStep(ma, 96, use_graph=False)
    sa.fuse_step = False
    sf = TrainStep(mf, 96, use_graph=False)
    assert sf.fuse_step and mf.supports_fused_step()

    # gradients before clipping
    loss, _, lr_, lbc, lic = sa.objective(tuple(t.clone() for t in batch))
    loss.backward()
    flat, numel = mf.train_step_grads(batch, DIFFUSION_COEFFS)
    want = torch.cat([p.grad.reshape(-1) for p in ma.parameters()])
    assert rel_err(flat[:numel], want) < 1e-5
    got_terms = flat[numel:numel + 4].cpu()
    for g, w in zip(got_terms, (loss, lr_, lbc, lic)):
        assert abs(g.item() - w.item()) < 1e-5 * abs(w.item())
    for p in ma.parameters():
        p.grad = None

    # full steps: parameters stay together
    for _ in range(3):
        va = sa(tuple(t.clone() for t in batch))
        vf = sf(batch)
        assert abs(va - vf) < 2e-5 * abs(va)
    for pa, pf in zip(ma.parameters(), mf.parameters()):
        assert rel_err(pf, pa) < 1e-4


def test_side_stream_overlap_is_bit_identical(tmp_path, monkeypatch):
    """train_step_grads runs the IC/BC chain on a side stream (joined before the partial reduction):
    same bits as the serialised order (QCP_OVERLAP=0), eagerly and under repeated calls."""
    from qcpinn_b200.trainer.diffusion_train import DIFFUSION_COEFFS

    batches = osolver.make_batches(3000, seed=5)
    batch = tuple(batches[k].to(DEV) for k in ("X_ic", "u_ic", "X_bc", "u_bc", "X_res", "r_res"))
    model = _model(tmp_path)
    outs = {}
    for mode in ("0", "1", "1", "0"):
        monkeypatch.setenv("QCP_OVERLAP", mode)
        flat, numel = model.train_step_grads(batch, DIFFUSION_COEFFS)
        torch.cuda.synchronize()
        outs.setdefault(mode, []).append(flat[:numel + 4].clone())
    ref = outs["0"][0]
    assert torch.isfinite(ref).all() and ref.abs().sum() > 0
    for got in outs["0"][1:] + outs["1"]:
        assert torch.equal(got, ref)


def _two_input_model(tmp_path, dtype):
    torch.manual_seed(3)
    args = dict(ARGS, classic_network=[2, 50, 1], dtype=dtype)
    return qb.DVPDESolver(args, qb.Logging(str(tmp_path)), device=DEV)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", ["wave", "klein_gordon", "helmholtz"])
def test_two_input_operators_match_oracle(tmp_path, name, dtype):
    """wave / Klein-Gordon / Helmholtz residuals (reference nn/pde.py:26-52,73-95) of a two-input
    solver: fused Taylor kernel vs nested autograd on the CPU oracle, values and every gradient."""
    from qcpinn_b200.nn import pde

    model = _two_input_model(tmp_path, dtype)
    oracle = _oracle_of(model)
    g = torch.Generator().manual_seed(17)
    a = torch.rand(33, 1, generator=g, dtype=torch.float64)
    b = torch.rand(33, 1, generator=g, dtype=torch.float64)
    cu = torch.randn(33, 1, generator=g, dtype=torch.float64)
    cr = torch.randn(33, 1, generator=g, dtype=torch.float64)

    ao, bo = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    uo = oracle.forward(torch.cat((ao, bo), 1))
    ones = torch.ones_like(uo)
    d = lambda out, wrt: torch.autograd.grad(out, wrt, ones, create_graph=True)[0]   # noqa: E731
    u_aa, u_bb = d(d(uo, ao), ao), d(d(uo, bo), bo)
    ro = {"wave": u_aa - 4.0 * u_bb,
          "klein_gordon": u_aa - u_bb + uo ** 3,
          "helmholtz": u_aa + u_bb + uo}[name]
    ((uo * cu).sum() + (ro * cr).sum()).backward()

    op = {"wave": pde.wave_operator, "klein_gordon": pde.klein_gordon_operator,
          "helmholtz": pde.helmholtz_operator}[name]
    ud, rd = op(model, a.to(DEV, torch.float32), b.to(DEV, torch.float32))
    ((ud * cu.to(DEV, torch.float32)).sum() + (rd * cr.to(DEV, torch.float32)).sum()).backward()
    tol = 2e-5 if dtype == "float32" else 5e-6        # module tensors / outputs are float32
    assert rel_err(ud, uo) < tol and rel_err(rd, ro) < tol
    for k, p in _weights_of(model).items():
        assert rel_err(p.grad, oracle.w[k].grad) < 10 * tol, k
    assert _weights_of(model)["w1"].grad.shape == (50, 2)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_navier_stokes_operator_on_the_fused_kernels(tmp_path, dtype):
    """reference nn/pde.py:2-25 with a three-output solver (u, v, p): continuity / momentum
    residuals (12 derivatives, products of streams) and every parameter gradient against the
    reference's own nested-autograd formulation evaluated on the CPU oracle."""
    from qcpinn_b200.nn import pde

    model = _model(tmp_path, dtype=dtype, classic_network=[3, 50, 3], seed=1)
    assert model.n_outputs == 3
    with torch.no_grad():
        model.preprocessor[2].bias.add_(0.3)
    oracle = _oracle_of(model)
    g = torch.Generator().manual_seed(5)
    X = torch.rand(37, 3, generator=g, dtype=torch.float64)
    coef = [torch.randn(37, 1, generator=g, dtype=torch.float64) for _ in range(3)]

    want = pde.navier_stokes_2D_operator(lambda Z: oracle.forward(Z), X[:, 0:1].clone(),
                                         X[:, 1:2].clone(), X[:, 2:3].clone())
    sum(((w_ * c).sum() for w_, c in zip(want, coef))).backward()

    Xd = X.to(DEV, torch.float32)
    uvp = model(Xd)
    assert uvp.shape == (37, 3) and uvp.dtype == torch.float32
    assert rel_err(uvp, oracle.forward(Xd.cpu().double())) < 2e-6
    got = pde.navier_stokes_2D_operator(model, Xd[:, 0:1].clone(), Xd[:, 1:2].clone(), Xd[:, 2:3].clone())
    sum(((g_ * c.to(DEV, torch.float32)).sum() for g_, c in zip(got, coef))).backward()
    tol = 3e-5 if dtype == "float32" else 5e-6        # module tensors / outputs are float32
    for a, b in zip(got, want):
        assert a.shape == (37, 1) and rel_err(a, b) < tol
    for k, p in _weights_of(model).items():
        assert rel_err(p.grad, oracle.w[k].grad) < 10 * tol, k
    # a one-output solver is rejected with the reference's requirement spelled out
    with pytest.raises(ValueError, match="three-output"):
        pde.navier_stokes_2D_operator(_model(tmp_path), Xd[:, 0:1], Xd[:, 1:2], Xd[:, 2:3])
    # the streams themselves, output by output
    S = model.taylor_streams_grad(Xd)
    assert S.shape == (37, 3, 6)


def test_single_file_trainer_entry_point(tmp_path):
    """reference train_hybrid_qpinn.py on the fused kernels: flags, 1-D angle vector, four boundary
    faces, pure-diffusion residual against the CPU oracle, checkpoint files."""
    from qcpinn_b200 import train_hybrid_qpinn as th

    ns = th.parse_args(["--num-qubits", "4", "--ansatz", "alternate", "--epochs", "3",
                        "--batch-size", "48", "--print-every", "2", "--output-dir", str(tmp_path),
                        "--device", "cuda"])
    assert ns.lr == 0.005 and ns.seed == 42 and ns.hidden_dim == 50 and ns.diffusion_coef == 0.01
    torch.manual_seed(ns.seed)
    model = th.HybridQPINN(ns, DEV)
    assert model.quantum_layer.params.shape == (12,)              # 4n - 4, no wrap-around pair
    assert model.quantum_layer.haar_seed1 == 42 and model.quantum_layer.haar_seed2 == 43
    assert set(model.state_dict()) >= {"preprocessor.0.weight", "quantum_layer.params",
                                       "postprocessor.2.bias"}
    # residual u_t - D (u_xx + u_yy) against nested autograd on the oracle (flat alternate program)
    from oracle import circuits as oc
    w = {k: v.detach().cpu() for k, v in _weights_of(model).items()}
    w["theta"] = w["theta"].reshape(1, -1)
    oracle = osolver.OracleSolver(4, 1, "alternate_flat", "angle", 42, "f64").set_weights(w)
    X = points(9, seed=5)
    uo, ro = osolver.diffusion_operator(oracle, X[:, 0:1].clone(), X[:, 1:2].clone(),
                                        X[:, 2:3].clone(), D=0.01, v_x=0.0, v_y=0.0)
    Xd = X.to(DEV, torch.float32)
    ud, rd = th.diffusion_operator(model, Xd[:, 0:1], Xd[:, 1:2], Xd[:, 2:3], D=0.01)
    assert rel_err(ud, uo) < 5e-6 and rel_err(rd, ro) < 5e-6
    # samplers: four faces, zero targets
    ics, bcs, res, dom = th.create_samplers(DEV, D=0.01)
    assert len(bcs) == 4 and bcs[1].sample(5)[0][:, 1].eq(1).all() and bcs[3].sample(5)[0][:, 2].eq(1).all()
    assert res.sample(6)[1].abs().sum() == 0
    out = tmp_path / "run"
    out.mkdir()
    th.train(model, ns, ics, bcs, res, str(out))
    assert len(model.loss_history) == 4 and all(math.isfinite(v) for v in model.loss_history)
    ck = torch.load(out / "checkpoint.pth", weights_only=False)
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict", "loss", "loss_history"}
    assert (out / "model.pth").exists()
    err = th.evaluate(model, ns, dom, str(out))
    assert math.isfinite(err) and (out / "evaluation.json").exists()


def test_prefetched_host_batches_give_the_same_steps(tmp_path):
    """TrainStep.prefetch (side-stream H2D into staging slots) vs feeding the host batch directly:
    identical loss trajectory, in eager and in graph-replay steps."""
    from qcpinn_b200.trainer.diffusion_train import TrainStep

    def run(prefetch):
        model = _model(tmp_path / ("p" if prefetch else "d"))
        step = TrainStep(model, 48)
        batches = []
        for i in range(6):
            b = osolver.make_batches(48, seed=300 + i)
            batches.append(tuple(b[k].float().contiguous().pin_memory()
                                 for k in ("X_ic", "u_ic", "X_bc", "u_bc", "X_res", "r_res")))
        out = []
        if prefetch:
            step.prefetch(batches[0])
        for i, hb in enumerate(batches):
            if prefetch and i + 1 < len(batches):
                step.prefetch(batches[i + 1])
            out.append(step(hb))
        return out

    a, b = run(False), run(True)
    assert len(a) == 6 and all(abs(x - y) <= 1e-6 * abs(x) for x, y in zip(a, b)), (a, b)


def test_eager_forward_between_graph_replays_sees_fresh_parameters(tmp_path):
    """A replayed CUDA graph updates the parameters behind Python's back; an eager forward between
    replays (periodic validation) must re-cast the weights and re-run qcp_prepare every time."""
    from qcpinn_b200.trainer.diffusion_train import TrainStep

    model = _model(tmp_path)
    step = TrainStep(model, 96)
    X = points(33).float().to(DEV)
    assert step.use_graph
    for round_ in range(3):
        for _ in range(5):
            step()
        assert round_ == 0 or step.steady()
        got = model(X)                                           # cached plan of the module
        fresh = F.Plan(model.quantum_layer.program, 0, torch.float64, 50, DEV)
        want = F.solver_value(fresh, X, model.quantum_layer.params.detach(),
                              [t.detach() for t in model._mlp_tensors()])
        assert rel_err(got, want) < 1e-6, round_
        u, r = qb.diffusion_operator(model, X[:, 0:1].clone(), X[:, 1:2].clone(), X[:, 2:3].clone())
        _, r_want = F.solver_residual(fresh, X, model.quantum_layer.params.detach(),
                                      [t.detach() for t in model._mlp_tensors()],
                                      (1.0, 1.0, 1.0, -0.01, -0.01))
        assert rel_err(r, r_want) < 1e-6, round_


def test_restore_on_cuda_keeps_graph_step_invariants(tmp_path):
    """restore() of our own checkpoint (read with the default map_location='cpu') and of a
    reference-style one (float lr, no capturable / fused): the lr stays the device tensor the
    captured step reads, and a graph-captured TrainStep still trains and follows lr reductions."""
    from qcpinn_b200.trainer.diffusion_train import TrainStep

    model = _model(tmp_path)
    step = TrainStep(model, 48)
    for _ in range(5):
        step()
    path = os.path.join(str(tmp_path), "ckpt.pth")
    model.save_state(path)
    state = qb.DVPDESolver.load_state(path)                      # CPU tensors, CPU lr tensor
    ref_style = dict(state)
    ref_opt = {"state": {k: {kk: (float(vv) if kk == "step" else vv) for kk, vv in v.items()}
                         for k, v in state["optimizer"]["state"].items()},
               "param_groups": [dict(g, lr=float(g["lr"]), capturable=False, fused=None, foreach=None)
                                for g in state["optimizer"]["param_groups"]]}
    ref_style["optimizer"] = ref_opt
    for st in (state, ref_style):
        m2 = _model(tmp_path)
        lr_tensor = m2.optimizer.param_groups[0]["lr"]
        m2.restore(st)
        g = m2.optimizer.param_groups[0]
        assert g["lr"] is lr_tensor and lr_tensor.is_cuda and g["capturable"] and g["fused"]
        assert abs(float(lr_tensor) - 0.005) < 1e-9
        for p_a, p_b in zip(model.parameters(), m2.parameters()):
            assert torch.equal(p_a, p_b)
        s2 = TrainStep(m2, 48)
        torch.manual_seed(77)
        losses = [s2() for _ in range(6)]                         # 3 eager, capture, replays
        assert s2.steady() and all(math.isfinite(v) for v in losses)
        before = [p.detach().clone() for p in m2.parameters()]
        for grp in m2.optimizer.param_groups:                    # what ReduceLROnPlateau does
            grp["lr"].fill_(0.0) if isinstance(grp["lr"], torch.Tensor) else None
        s2()
        # lr = 0 reached the captured Adam kernel: parameters did not move
        assert all(torch.equal(a, b) for a, b in zip(before, m2.parameters()))


def test_residual_outputs_are_not_connected_to_the_coordinates(tmp_path):
    """d(u, r)/d(t, x, y) on the fused route fails loudly instead of returning silent zeros."""
    model = _model(tmp_path)
    X = points(9).float().to(DEV)
    t, x, y = (X[:, i:i + 1].clone() for i in range(3))
    u, r = qb.diffusion_operator(model, t, x, y)
    with pytest.raises(RuntimeError, match="not have been used in the graph"):
        torch.autograd.grad(r.sum(), x)
    Xv = X.clone().requires_grad_(True)                          # value mode does give du/dX
    (gx,) = torch.autograd.grad(model(Xv).sum(), Xv)
    assert gx.shape == X.shape and float(gx.abs().max()) > 0


def test_opt_in_autograd_mode_is_differentiable_to_any_order(tmp_path):
    """args["diff_mode"] = "autograd": the reference's literal nested-autograd usage
    (nn/pde.py:60-70: autograd.grad(u, t, create_graph=True), second derivatives, backward through
    all of it) runs on the module; values, residual and gradients agree with the kernel mode and
    with the CPU oracle; mixed second derivatives -- which the six Taylor streams do not carry --
    are available here."""
    model_k = _model(tmp_path)
    model_a = _model(tmp_path, diff_mode="autograd")
    for pa, pk in zip(model_a.parameters(), model_k.parameters()):
        assert torch.equal(pa, pk)
    assert model_a.taylor_residual is None and not model_a.supports_fused_step()
    oracle = _oracle_of(model_k, "mixed")
    X = points(21).float().to(DEV)
    t, x, y = (X[:, i:i + 1].clone().requires_grad_(True) for i in range(3))
    u = model_a(torch.cat((t, x, y), 1))
    ones = torch.ones_like(u)
    u_x = torch.autograd.grad(u, x, ones, create_graph=True)[0]
    u_xy = torch.autograd.grad(u_x, y, ones, create_graph=True)[0]            # a MIXED derivative
    u_xx = torch.autograd.grad(u_x, x, ones, create_graph=True)[0]
    assert rel_err(u, model_k(X)) < 1e-6
    S = model_k.taylor_streams(X)
    assert rel_err(u_x, S[:, 2:3]) < 1e-5 and rel_err(u_xx, S[:, 4:5]) < 1e-4
    Xc = X.cpu()
    to, xo, yo = (Xc[:, i:i + 1].clone().requires_grad_(True) for i in range(3))
    uo = oracle.forward(torch.cat((to, xo, yo), 1))
    uo_x = torch.autograd.grad(uo, xo, torch.ones_like(uo), create_graph=True)[0]
    uo_xy = torch.autograd.grad(uo_x, yo, torch.ones_like(uo), create_graph=True)[0]
    assert rel_err(u_xy, uo_xy) < 1e-4
    # the generic formulation of nn.pde on the autograd-mode module == the fused operator
    ua, ra = qb.diffusion_operator(model_a, X[:, 0:1].clone(), X[:, 1:2].clone(), X[:, 2:3].clone())
    uk, rk = qb.diffusion_operator(model_k, X[:, 0:1].clone(), X[:, 1:2].clone(), X[:, 2:3].clone())
    assert rel_err(ra, rk) < 1e-4
    (ra ** 2).mean().backward()
    (rk ** 2).mean().backward()
    for pa, pk in zip(model_a.parameters(), model_k.parameters()):
        # (the residual does not depend on the output bias: autograd leaves its .grad unset)
        ga = pa.grad if pa.grad is not None else torch.zeros_like(pa)
        if float(pk.grad.abs().max()) == 0.0:
            assert float(ga.abs().max()) == 0.0
        else:
            assert rel_err(ga, pk.grad) < 2e-3         # float32 nested autograd vs float64 kernels
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        qb.DVQuantumLayer(dict(ARGS, diff_mode="autograd"))(torch.zeros(2, 4))
    with pytest.raises(ValueError, match="diff_mode"):
        qb.DVQuantumLayer(dict(ARGS, diff_mode="nope"))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("batch", [0, 1, 31, 127, 128, 129, 255, 511, 513, 1000, 4097])
def test_value_mode_ragged_batches_with_several_points_per_thread(batch, dtype):
    """Value mode (IC / BC points) runs the post / pre adjoints with four and the contraction
    adjoint with two points per thread: every chunk boundary (128 x K points per block iteration),
    the saved tanh / sin / cos rows and grad_X against the oracle, for sizes around those boundaries."""
    w, oracle, prog = make_case("cascade", 4, 1, "angle", 1)
    plan = F.Plan(prog, 0, dtype, 50, DEV)
    dw = device_weights(w, dtype, DEV, requires_grad=True)
    X = points(max(batch, 1), seed=batch + 1)[:batch]
    Xd = X.to(DEV, dtype).requires_grad_(True)
    u = F.solver_value(plan, Xd, dw["theta"], mlp_list(dw))
    assert u.shape == (batch, 1)
    g = torch.Generator().manual_seed(batch)
    cot = torch.randn(batch, 1, generator=g, dtype=torch.float64)
    (u * cot.to(DEV, dtype)).sum().backward()
    if batch == 0:
        assert float(dw["w3"].grad.abs().max()) == 0.0
        return
    Xo = X.clone().requires_grad_(True)
    uo = oracle.forward(Xo)
    (uo * cot).sum().backward()
    tol = TOL[dtype]
    assert rel_err(u, uo) < tol
    assert rel_err(Xd.grad, Xo.grad) < 20 * tol
    for k in osolver.WEIGHT_NAMES:
        assert rel_err(dw[k].grad, oracle.w[k].grad) < 20 * tol, k
