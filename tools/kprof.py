"""Developer tool: a few forward(save)+split-backward launches for ncu (see profiles/README)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
import qcpinn_b200 as qb
F = qb.functional
pts = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dtypes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["f64", "f32"]
modes = [int(m) for m in sys.argv[3].split(",")] if len(sys.argv) > 3 else [6]
dev = torch.device("cuda", 0)
prog = qb.program.compile_program("cascade", 4, 1, None)
for name in dtypes:
    dt = {"f64": torch.float64, "f32": torch.float32}[name]
    plan = F.Plan(prog, 0, dt, 50, dev)
    torch.manual_seed(0)
    X = torch.rand(pts, 3, device=dev, dtype=dt)
    g = torch.rand(pts, device=dev, dtype=dt)
    mlp = [torch.randn(s, device=dev, dtype=dt) * 0.3 for s in ((50, 3), (50,), (4, 50), (4,), (50, 4), (50,), (1, 50), (1,))]
    theta = torch.randn(prog.n_theta, device=dev, dtype=dt)
    plan.prepare(theta)
    c = (1.0, 1.0, 1.0, -0.01, -0.01)
    for mode in modes:
        ws = plan.workspace(pts, mode)
        for _ in range(3):
            plan.solver_forward(X, mlp, mode, c if mode == 6 else None, save=ws)
            plan.solver_backward(X, mlp, theta, g if mode == 1 else None, g if mode == 6 else None, mode, c if mode == 6 else None, save=ws)
torch.cuda.synchronize()
print("ok")
