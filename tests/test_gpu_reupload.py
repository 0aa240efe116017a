"""GPU parity of the data re-uploading circuit family (reference
hybrid_testing/CG_HQPINN_IBMtest_16qubits.py:217-235): RY(x_i) encoding, RZ(0.5 x_j) re-upload inside
every layer, Rot, CZ brick + ring -- per-sample "jet gates" inside the program (engine L), value mode
and Taylor mode, against the oracle extension (which tests/test_reference_execution.py pins on the
reference's own, unmodified ``make_quantum_layer``).  Bars: 1e-10 float64 / 1e-5 float32."""

import pytest
import torch

import qcpinn_b200 as qb
from helpers import F, TOL, device_weights, mlp_list, points, rel_err
from oracle import circuits as oc
from oracle import solver as osolver

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)
DTYPES = [torch.float64, torch.float32]


def _case(n, layers, seed=0):
    w = osolver.init_weights(n, layers, "cz_melt", hidden=50, seed=seed)
    g = torch.Generator().manual_seed(seed + 17)
    for k in ("b1", "b2"):
        w[k] = 0.1 * torch.randn(w[k].shape, generator=g)
    w["theta"] = torch.randn(w["theta"].shape, generator=g)
    oracle = osolver.OracleSolver(n, layers, "cz_melt", "angle", None, "f64").set_weights(w)
    prog = qb.program.compile_program("cz_melt", n, layers)
    return w, oracle, prog


@pytest.mark.parametrize("dtype", DTYPES, ids=["f64", "f32"])
@pytest.mark.parametrize("n,layers,batch", [(6, 2, 9), (4, 1, 5), (5, 3, 7), (16, 2, 3), (11, 1, 4)])
def test_reupload_layer_forward_backward(n, layers, batch, dtype):
    w, oracle, prog = _case(n, layers)
    plan = F.Plan(prog, 0, dtype, 50, DEV)
    assert plan.engine == "global" and not plan.fused_engine      # jet gates run on engine L
    g = torch.Generator().manual_seed(5)
    z = torch.randn(batch, n, generator=g, dtype=torch.float64)
    cot = torch.randn(n, batch, generator=g, dtype=torch.float64)
    zo = z.clone().requires_grad_(True)
    qo = oracle.quantum(zo)
    (qo * cot).sum().backward()
    zd = z.to(DEV, dtype).requires_grad_(True)
    th = w["theta"].to(DEV, dtype).requires_grad_(True)
    q = F.layer_apply(plan, zd, th)
    assert q.shape == (n, batch)
    (q * cot.to(DEV, dtype)).sum().backward()
    tol = TOL[dtype]
    assert rel_err(q, qo) < tol
    assert rel_err(zd.grad, zo.grad) < 10 * tol
    assert rel_err(th.grad, oracle.w["theta"].grad) < 10 * tol


@pytest.mark.parametrize("dtype", DTYPES, ids=["f64", "f32"])
@pytest.mark.parametrize("n,layers,batch", [(6, 2, 11), (4, 2, 6), (16, 1, 2)])
def test_reupload_solver_taylor_streams_and_gradients(n, layers, batch, dtype):
    """The six Taylor streams THROUGH the per-sample gates (u_t .. u_yy via the jet-gate rules), the
    residual and every parameter gradient against nested autograd on the oracle."""
    w, oracle, prog = _case(n, layers, seed=1)
    plan = F.Plan(prog, 0, dtype, 50, DEV)
    dw = device_weights(w, dtype, DEV, requires_grad=True)
    X = points(batch, seed=4)
    coeffs = (1.0, 1.0, 1.0, -0.01, -0.01)
    tol = TOL[dtype]
    streams = F.solver_streams(plan, X.to(DEV, dtype), dw["theta"], mlp_list(dw), coeffs)[2]
    want = osolver.diffusion_streams(oracle, X)
    for c in range(6):
        assert rel_err(streams[:, c], want[:, c]) < 20 * tol, c
    g = torch.Generator().manual_seed(8)
    cu, cr = (torch.randn(batch, 1, generator=g, dtype=torch.float64) for _ in range(2))
    uo, ro = osolver.diffusion_operator(oracle, X[:, 0:1].clone(), X[:, 1:2].clone(), X[:, 2:3].clone())
    ((uo * cu).sum() + (ro * cr).sum()).backward()
    u, r = F.solver_residual(plan, X.to(DEV, dtype), dw["theta"], mlp_list(dw), coeffs)
    ((u * cu.to(DEV, dtype)).sum() + (r * cr.to(DEV, dtype)).sum()).backward()
    assert rel_err(u, uo) < 10 * tol and rel_err(r, ro) < 20 * tol
    for k in osolver.WEIGHT_NAMES:
        assert rel_err(dw[k].grad, oracle.w[k].grad) < 50 * tol, k


def test_reupload_module_surface(tmp_path):
    """DVQuantumLayer / DVPDESolver with q_ansatz = "cz_melt": (L, 3n) angles = the reference's
    (L, n, 3) weights flattened, no Haar blocks whatever ``seed`` says, residual operator works."""
    args = {"batch_size": 8, "epochs": 1, "lr": 0.005, "seed": 1, "print_every": 100,
            "num_qubits": 6, "num_quantum_layers": 2, "classic_network": [3, 50, 1],
            "q_ansatz": "cz_melt", "problem": "diffusion", "solver": "DV", "encoding": "None"}
    layer = qb.DVQuantumLayer(args).to(DEV)
    assert tuple(layer.params.shape) == (2, 18) and layer.program.consts.shape[0] == 0
    x = torch.randn(5, 6, device=DEV)
    out = layer(x)
    want = oc.quantum_layer(x.cpu().double(), layer.params.detach().cpu().double(), "cz_melt", 6)
    assert out.shape == (6, 5) and rel_err(out, want) < 1e-10
    torch.manual_seed(0)
    model = qb.DVPDESolver(args, qb.Logging(str(tmp_path)), device=DEV)
    X = points(7).float().to(DEV)
    u, r = qb.diffusion_operator(model, X[:, 0:1].clone(), X[:, 1:2].clone(), X[:, 2:3].clone())
    assert u.shape == (7, 1) and r.shape == (7, 1) and bool(torch.isfinite(r).all())
    (r ** 2).mean().backward()
    assert model.quantum_layer.params.grad is not None
    assert float(model.quantum_layer.params.grad.abs().max()) > 0
