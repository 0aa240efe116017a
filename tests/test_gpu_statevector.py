"""GPU parity of the per-sample statevector engine (n >= 5 qubits: BASELINE configs 3 and 4 shapes)
against the CPU oracle: stand-alone layer, solver value mode, Taylor streams, residual and every
gradient.  Bars as in test_gpu_parity.py (1e-10 float64, 1e-5 float32)."""

import pytest
import torch

import qcpinn_b200 as qb
from helpers import F, TOL, device_weights, make_case, mlp_list, points, rel_err
from oracle import solver as osolver

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)

CASES = [
    # ansatz, n, layers, encoding, haar_seed
    ("cascade", 5, 1, "angle", 1),
    ("layered", 5, 2, "angle", None),
    ("alternate", 5, 1, "angle", 1),
    ("farhi", 6, 1, "angle", None),
    ("sim_circ_15", 6, 2, "angle", 1),
    ("cross_mesh", 5, 1, "angle", None),
    ("cascade", 6, 1, "amplitude", 1),
    ("cross_mesh", 5, 2, "amplitude", None),
    # larger registers-resident shapes (engine R: lane/local swaps, diagonal blocks, Haar hops);
    # float64 n = 10 runs with two warps per stream vector
    ("cross_mesh", 10, 2, "angle", None),
    ("sim_circ_15", 8, 1, "angle", 1),
    ("layered", 9, 1, "angle", None),
    ("cascade", 7, 1, "amplitude", None),
    ("farhi", 10, 1, "angle", 1),
    ("alternate", 7, 1, "angle", None),
    # tiled engine (engine T: HBM slab swept through registers, controls on tile-index bits)
    ("cross_mesh", 11, 1, "angle", None),
    ("cascade", 12, 1, "angle", 1),
    ("layered", 11, 1, "amplitude", None),
    ("sim_circ_15", 13, 1, "angle", None),
    ("farhi", 11, 1, "angle", 1),
    ("alternate", 11, 1, "angle", None),
    ("layered", 10, 1, "angle", 1),
    ("cascade", 10, 1, "amplitude", 1),
]
DTYPES = [torch.float64, torch.float32]


def _ids(c):
    return "-".join(map(str, c))


@pytest.mark.parametrize("dtype", DTYPES, ids=["f64", "f32"])
@pytest.mark.parametrize("case", CASES, ids=_ids)
def test_layer_forward_backward(case, dtype):
    ansatz, n, layers, enc, seed = case
    w, oracle, prog = make_case(ansatz, n, layers, enc, seed)
    g = torch.Generator().manual_seed(5)
    z = torch.randn(9, n, generator=g, dtype=torch.float64) + (0.8 if enc == "amplitude" else 0.0)
    cot = torch.randn(n, 9, generator=g, dtype=torch.float64)
    zo = z.clone().requires_grad_(True)
    qo = oracle.quantum(zo)
    (qo * cot).sum().backward()

    plan = F.Plan(prog, F.encoding_code(enc), dtype, 50, DEV)
    assert plan.engine == ("register" if n <= 10 else "tiled")    # float64 at n = 10: warp pairs
    zd = z.to(DEV, dtype).requires_grad_(True)
    th = w["theta"].to(DEV, dtype).requires_grad_(True)
    qd = F.layer_apply(plan, zd, th)
    (qd * cot.to(DEV, dtype)).sum().backward()
    tol = TOL[dtype]
    assert rel_err(qd, qo) < tol
    assert rel_err(zd.grad, zo.grad) < tol
    assert rel_err(th.grad, oracle.w["theta"].grad) < tol


@pytest.mark.parametrize("dtype", DTYPES, ids=["f64", "f32"])
@pytest.mark.parametrize("case", CASES, ids=_ids)
def test_residual_streams_and_gradients(case, dtype):
    ansatz, n, layers, enc, seed = case
    w, oracle, prog = make_case(ansatz, n, layers, enc, seed)
    X = points(7, seed=11)
    g = torch.Generator().manual_seed(13)
    cu = torch.randn(7, 1, generator=g, dtype=torch.float64)
    cr = torch.randn(7, 1, generator=g, dtype=torch.float64)
    coeffs = (1.0, 1.0, 1.0, -0.01, -0.01)
    streams_o = osolver.diffusion_streams(oracle, X).detach()
    for v in oracle.w.values():
        v.grad = None
    uo, ro = osolver.diffusion_operator(
        oracle, X[:, 0:1].clone(), X[:, 1:2].clone(), X[:, 2:3].clone())
    ((uo * cu).sum() + (ro * cr).sum()).backward()

    plan = F.Plan(prog, F.encoding_code(enc), dtype, 50, DEV)
    dw = device_weights(w, dtype, requires_grad=True)
    Xd = X.to(DEV, dtype)
    _, _, streams_d = F.solver_streams(plan, Xd, dw["theta"], mlp_list(dw), coeffs)
    ud, rd = F.solver_residual(plan, Xd, dw["theta"], mlp_list(dw), coeffs)
    ((ud * cu.to(DEV, dtype)).sum() + (rd * cr.to(DEV, dtype)).sum()).backward()
    tol = TOL[dtype]
    for c, name in enumerate(["u", "u_t", "u_x", "u_y", "u_xx", "u_yy"]):
        assert rel_err(streams_d[:, c], streams_o[:, c]) < tol, name
    assert rel_err(rd, ro) < tol
    for k in osolver.WEIGHT_NAMES:
        assert rel_err(dw[k].grad, oracle.w[k].grad) < tol, k


@pytest.mark.parametrize("dtype", DTYPES, ids=["f64", "f32"])
def test_value_mode_and_input_gradient(dtype):
    w, oracle, prog = make_case("cascade", 5, 1, "angle", 1)
    X = points(11)
    Xo = X.clone().requires_grad_(True)
    uo = oracle.forward(Xo)
    uo.sum().backward()
    plan = F.Plan(prog, 0, dtype, 50, DEV)
    dw = device_weights(w, dtype, requires_grad=True)
    Xd = X.to(DEV, dtype).requires_grad_(True)
    ud = F.solver_value(plan, Xd, dw["theta"], mlp_list(dw))
    ud.sum().backward()
    tol = TOL[dtype]
    assert rel_err(ud, uo) < tol and rel_err(Xd.grad, Xo.grad) < tol
    for k in osolver.WEIGHT_NAMES:
        assert rel_err(dw[k].grad, oracle.w[k].grad) < tol, k


@pytest.mark.parametrize("dtype", DTYPES, ids=["f64", "f32"])
def test_config3_shape_cross_mesh_10q(dtype):
    """BASELINE config 3 shape: cross_mesh, 10 qubits, 2 layers (small batch for the CPU oracle)."""
    w, oracle, prog = make_case("cross_mesh", 10, 2, "angle", None)
    X = points(4, seed=2)
    coeffs = (1.0, 1.0, 1.0, -0.01, -0.01)
    for v in oracle.w.values():
        v.grad = None
    uo, ro = osolver.diffusion_operator(
        oracle, X[:, 0:1].clone(), X[:, 1:2].clone(), X[:, 2:3].clone())
    (uo.sum() + ro.sum()).backward()
    plan = F.Plan(prog, 0, dtype, 50, DEV)
    dw = device_weights(w, dtype, requires_grad=True)
    ud, rd = F.solver_residual(plan, X.to(DEV, dtype), dw["theta"], mlp_list(dw), coeffs)
    (ud.sum() + rd.sum()).backward()
    tol = TOL[dtype]
    assert rel_err(ud, uo) < tol and rel_err(rd, ro) < tol
    for k in osolver.WEIGHT_NAMES:
        assert rel_err(dw[k].grad, oracle.w[k].grad) < tol, k


def test_config4_shape_sim_circ_15_16q():
    """BASELINE config 4 shape: sim_circ_15, 16 qubits, 2 layers; state in the global workspace."""
    w, oracle, prog = make_case("sim_circ_15", 16, 2, "angle", None)
    X = points(2, seed=4)
    coeffs = (1.0, 1.0, 1.0, -0.01, -0.01)
    for v in oracle.w.values():
        v.grad = None
    uo, ro = osolver.diffusion_operator(
        oracle, X[:, 0:1].clone(), X[:, 1:2].clone(), X[:, 2:3].clone())
    (uo.sum() + ro.sum()).backward()
    plan = F.Plan(prog, 0, torch.float64, 50, DEV)
    dw = device_weights(w, torch.float64, requires_grad=True)
    ud, rd = F.solver_residual(plan, X.to(DEV), dw["theta"], mlp_list(dw), coeffs)
    (ud.sum() + rd.sum()).backward()
    assert rel_err(ud, uo) < 1e-10 and rel_err(rd, ro) < 1e-10
    for k in osolver.WEIGHT_NAMES:
        assert rel_err(dw[k].grad, oracle.w[k].grad) < 1e-10, k


def test_solver_module_runs_a_train_step_at_6_qubits(tmp_path):
    from qcpinn_b200.trainer.diffusion_train import TrainStep

    torch.manual_seed(0)
    args = {"batch_size": 16, "epochs": 2, "lr": 0.005, "seed": 1, "print_every": 10,
            "num_qubits": 6, "num_quantum_layers": 1, "classic_network": [3, 50, 1],
            "q_ansatz": "layered", "problem": "diffusion", "solver": "DV", "cuda_graph": False}
    model = qb.DVPDESolver(args, qb.Logging(str(tmp_path)), device=DEV)
    step = TrainStep(model, 24)
    l0 = step()
    l1 = step()
    assert l0 == l0 and l1 == l1 and len(model.loss_history) == 2


@pytest.mark.parametrize("case", [("cross_mesh", 6, 1, "angle", None), ("sim_circ_15", 11, 1, "angle", None)],
                         ids=_ids)
def test_backward_without_saved_state_recomputes_the_forward(case, monkeypatch):
    """Engines R / T keep the final psi streams in the workspace only while they fit the memory
    budget; with the budget at zero the backward recomputes the forward -- same gradients, and the
    flag follows the workspace object when calls with and without state interleave."""
    ansatz, n, layers, enc, seed = case
    w, oracle, prog = make_case(ansatz, n, layers, enc, seed)
    X = points(6, seed=3)
    coeffs = (1.0, 1.0, 1.0, -0.01, -0.01)
    uo, ro = osolver.diffusion_operator(
        oracle, X[:, 0:1].clone(), X[:, 1:2].clone(), X[:, 2:3].clone())
    (uo.sum() + ro.sum()).backward()
    plan = F.Plan(prog, 0, torch.float64, 50, DEV)
    dw = device_weights(w, torch.float64, requires_grad=True)
    monkeypatch.setattr(F.Plan, "STATE_SAVE_BUDGET", 0.0)
    ud, rd = F.solver_residual(plan, X.to(DEV), dw["theta"], mlp_list(dw), coeffs)   # no state
    monkeypatch.setattr(F.Plan, "STATE_SAVE_BUDGET", 0.35)
    ud2, rd2 = F.solver_residual(plan, X.to(DEV), dw["theta"], mlp_list(dw), coeffs)  # with state
    (ud.sum() + rd.sum()).backward()          # backward of the FIRST call after the flag flipped
    assert rel_err(rd, ro) < 1e-10 and rel_err(rd2, ro) < 1e-10
    for k in osolver.WEIGHT_NAMES:
        assert rel_err(dw[k].grad, oracle.w[k].grad) < 1e-10, k


@pytest.mark.parametrize("n,ansatz", [(6, "layered"), (11, "sim_circ_15")])
def test_train_step_graph_replay_on_statevector_engines(tmp_path, n, ansatz):
    """Default TrainStep (CUDA-graph replay after three eager steps) on engines R and T: the
    captured step allocates nothing, and replays keep training (loss history grows, stays finite,
    parameters move) exactly like eager steps on the same batches."""
    from qcpinn_b200.trainer.diffusion_train import TrainStep

    def run(graph):
        torch.manual_seed(0)
        args = {"batch_size": 16, "epochs": 2, "lr": 0.005, "seed": 1, "print_every": 10,
                "num_qubits": n, "num_quantum_layers": 1, "classic_network": [3, 50, 1],
                "q_ansatz": ansatz, "problem": "diffusion", "solver": "DV", "cuda_graph": graph,
                "dtype": "float32"}
        model = qb.DVPDESolver(args, qb.Logging(str(tmp_path / str(graph))), device=DEV)
        step = TrainStep(model, 24)
        out = []
        for i in range(6):
            b = osolver.make_batches(24, seed=400 + i)
            out.append(step(tuple(b[k].float().to(DEV) for k in
                                  ("X_ic", "u_ic", "X_bc", "u_bc", "X_res", "r_res"))))
        return out

    eager, graph = run(False), run(True)
    assert all(v == v for v in graph)
    assert all(abs(a - b) <= 1e-4 * abs(a) for a, b in zip(eager, graph)), (eager, graph)


@pytest.mark.parametrize("ansatz,n,layers,n_pts,n_check", [
    ("cross_mesh", 10, 2, 262_144, 6),      # BASELINE config 3 at full size (engine R, float32)
    ("sim_circ_15", 16, 2, 16_384, 2),      # BASELINE config 4 at full size (engine T, float32)
])
def test_full_size_properties_configs_3_and_4(ansatz, n, layers, n_pts, n_check):
    """Size-independent properties at the full BASELINE sizes: (a) deterministic and finite;
    (b) gradients are additive over a partition of the batch (persistent grids, padding and the
    saved-state workspace do not depend on where a point sits); (c) a strided sample of points
    agrees with the CPU oracle."""
    w, oracle, prog = make_case(ansatz, n, layers, "angle", None)
    plan = F.Plan(prog, 0, torch.float32, 50, DEV)
    dw = device_weights(w, torch.float32, DEV)
    g = torch.Generator(device=DEV).manual_seed(5)
    X = torch.rand(n_pts, 3, device=DEV, dtype=torch.float32, generator=g)
    gr = torch.rand(n_pts, device=DEV, dtype=torch.float32, generator=g) / n_pts
    coeffs = (1.0, 1.0, 1.0, -0.01, -0.01)
    theta = dw["theta"].reshape(-1)
    mlp = mlp_list(dw)
    plan.prepare(theta)

    def fwd_bwd(Xb, gb):
        ws = plan.workspace(Xb.shape[0], F.MODE_RESIDUAL)
        u, r, _ = plan.solver_forward(Xb, mlp, F.MODE_RESIDUAL, coeffs, save=ws)
        grads, _ = plan.solver_backward(Xb, mlp, theta, None, gb, F.MODE_RESIDUAL, coeffs, save=ws)
        return u.clone(), r.clone(), [v.clone() for v in grads]

    u1, r1, full = fwd_bwd(X, gr)
    u2, r2, again = fwd_bwd(X, gr)
    assert torch.equal(r1, r2) and torch.equal(u1, u2) and bool(torch.isfinite(r1).all())
    half = n_pts // 2 + 7                       # ragged split
    _, _, a = fwd_bwd(X[:half].contiguous(), gr[:half].contiguous())
    _, _, b = fwd_bwd(X[half:].contiguous(), gr[half:].contiguous())
    for f_, a_, b_ in zip(full, a, b):
        assert rel_err(a_ + b_, f_) < 2e-4          # float32 sums in a different order
    idx = torch.arange(0, n_pts, n_pts // n_check, device=DEV)[:n_check]
    Xs = X[idx].cpu().double()
    uo, ro = osolver.diffusion_operator(oracle, Xs[:, 0:1].clone(), Xs[:, 1:2].clone(), Xs[:, 2:3].clone())
    assert rel_err(u1[idx], uo[:, 0]) < 1e-5 and rel_err(r1[idx], ro[:, 0]) < 1e-5


@pytest.mark.parametrize("dtype", DTYPES, ids=["f64", "f32"])
@pytest.mark.parametrize("case", [("cross_mesh", 10, 2, "angle", None), ("layered", 7, 2, "angle", 1),
                                  ("sim_circ_15", 8, 1, "angle", None)], ids=_ids)
def test_register_engine_gradients_are_bit_reproducible(case, dtype):
    """Engine R accumulates dL/dtheta and the diagonal-block sums in one global row per warp (plain
    read-modify-write, rows summed in order): repeated backward passes over the same batch give
    bit-identical gradients -- what data-parallel equivalence checks rely on."""
    ansatz, n, layers, enc, seed = case
    w, _, prog = make_case(ansatz, n, layers, enc, seed)
    plan = F.Plan(prog, F.encoding_code(enc), dtype, 50, DEV)
    assert plan.engine == "register"
    X = points(700, seed=9).to(DEV, dtype)
    runs = []
    for _ in range(3):
        dw = device_weights(w, dtype, DEV, requires_grad=True)
        u, r = F.solver_residual(plan, X, dw["theta"], mlp_list(dw), (1, 1, 1, -0.01, -0.01))
        ((u ** 2).sum() + (r ** 2).sum()).backward()
        runs.append({k: v.grad.detach().clone() for k, v in dw.items()})
    for other in runs[1:]:
        for k, g in runs[0].items():
            assert torch.equal(g, other[k]), k


@pytest.mark.parametrize("dtype", DTYPES, ids=["f64", "f32"])
@pytest.mark.parametrize("case", [("sim_circ_15", 12, 1, "angle", None), ("cross_mesh", 11, 1, "angle", None),
                                  ("layered", 11, 1, "amplitude", 1)], ids=_ids)
def test_tiled_engine_is_bit_reproducible(case, dtype):
    """Engine T: measurement sums, table cotangents and dL/dtheta are accumulated per warp and
    combined in a fixed order (diagonal-block sums: all streams of a tile on one warp), so repeated
    passes over the same batch give bit-identical outputs and gradients."""
    ansatz, n, layers, enc, seed = case
    w, _, prog = make_case(ansatz, n, layers, enc, seed)
    plan = F.Plan(prog, F.encoding_code(enc), dtype, 50, DEV)
    assert plan.engine == "tiled"
    X = points(300, seed=9).to(DEV, dtype)
    runs = []
    for _ in range(3):
        dw = device_weights(w, dtype, DEV, requires_grad=True)
        u, r = F.solver_residual(plan, X, dw["theta"], mlp_list(dw), (1, 1, 1, -0.01, -0.01))
        ((u ** 2).sum() + (r ** 2).sum()).backward()
        out = {k: v.grad.detach().clone() for k, v in dw.items()}
        out["u"], out["r"] = u.detach().clone(), r.detach().clone()
        runs.append(out)
    for other in runs[1:]:
        for k, g in runs[0].items():
            assert torch.equal(g, other[k]), k
