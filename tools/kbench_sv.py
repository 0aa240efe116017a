"""Developer tool: time the statevector engine (n >= 5) forward + backward.
usage: python tools/kbench_sv.py n ansatz layers points [dtype]"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
import qcpinn_b200 as qb
F = qb.functional
n = int(sys.argv[1]); ansatz = sys.argv[2]; layers = int(sys.argv[3]); pts = int(sys.argv[4])
dts = sys.argv[5].split(",") if len(sys.argv) > 5 else ["f32", "f64"]
dev = torch.device("cuda", 0)
prog = qb.program.compile_program(ansatz, n, layers, None)
for name in dts:
    dt = {"f64": torch.float64, "f32": torch.float32}[name]
    plan = F.Plan(prog, 0, dt, 50, dev)
    torch.manual_seed(0)
    X = torch.rand(pts, 3, device=dev, dtype=dt)
    g = torch.rand(pts, device=dev, dtype=dt)
    mlp = [torch.randn(s, device=dev, dtype=dt) * 0.3 for s in ((50, 3), (50,), (n, 50), (n,), (50, n), (50,), (1, 50), (1,))]
    theta = torch.randn(prog.n_theta, device=dev, dtype=dt)
    plan.prepare(theta)
    c = (1.0, 1.0, 1.0, -0.01, -0.01)
    ws = plan.workspace(pts, 6)
    def run():
        plan.solver_forward(X, mlp, 6, c, save=ws)
        plan.solver_backward(X, mlp, theta, None, g, 6, c, save=ws)
    run(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); plan.solver_forward(X, mlp, 6, c, save=ws); e1.record(); torch.cuda.synchronize()
    tf = e0.elapsed_time(e1)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    tt = e0.elapsed_time(e1)
    print(f"n={n} {ansatz} L={layers} {name}: residual fwd {tf:.2f} ms, fwd+bwd {tt:.2f} ms for {pts} pts -> "
          f"{pts / (tt * 1e-3):.0f} residual-pts/s (fwd+bwd only)")
