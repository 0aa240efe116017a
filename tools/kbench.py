"""Kernel micro-benchmark (developer tool): times the fused forward / backward kernels alone with
CUDA events.  usage: python tools/kbench.py [points] [ansatz]"""
import os, sys, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
import qcpinn_b200 as qb
F = qb.functional
pts = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
ansatz = sys.argv[2] if len(sys.argv) > 2 else "cascade"
dev = torch.device("cuda", 0)
prog = qb.program.compile_program(ansatz, 4, 1, None)
out = {}
only = os.environ.get("KBENCH_DTYPES", "f64,f32").split(",")
for dt, name in ((torch.float64, "f64"), (torch.float32, "f32")):
    if name not in only:
        continue
    plan = F.Plan(prog, 0, dt, 50, dev)
    torch.manual_seed(0)
    X = torch.rand(pts, 3, device=dev, dtype=dt)
    g = torch.rand(pts, device=dev, dtype=dt)
    mlp = [torch.randn(s, device=dev, dtype=dt) * 0.3 for s in ((50, 3), (50,), (4, 50), (4,), (50, 4), (50,), (1, 50), (1,))]
    theta = torch.randn(prog.n_theta, device=dev, dtype=dt)
    plan.prepare(theta)
    c = (1.0, 1.0, 1.0, -0.01, -0.01)
    def t(fn, reps=10):
        for _ in range(3): fn()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    ws6 = plan.workspace(pts, 6); ws1 = plan.workspace(pts, 1)
    o = {
        "fwd_res": t(lambda: plan.solver_forward(X, mlp, 6, c)),
        "fwd_res_save": t(lambda: plan.solver_forward(X, mlp, 6, c, save=ws6)),
        "bwd_res_fused": t(lambda: plan.solver_backward(X, mlp, theta, None, g, 6, c)),
        "fwd_val": t(lambda: plan.solver_forward(X, mlp, 1)),
        "fwd_val_save": t(lambda: plan.solver_forward(X, mlp, 1, save=ws1)),
        "bwd_val_fused": t(lambda: plan.solver_backward(X, mlp, theta, g, None, 1)),
    }
    def split6():
        plan.solver_forward(X, mlp, 6, c, save=ws6); plan.solver_backward(X, mlp, theta, None, g, 6, c, save=ws6)
    def split1():
        plan.solver_forward(X, mlp, 1, save=ws1); plan.solver_backward(X, mlp, theta, g, None, 1, save=ws1)
    o["bwd_res_split"] = t(split6) - o["fwd_res_save"]
    o["bwd_val_split"] = t(split1) - o["fwd_val_save"]
    fused = o["fwd_res"] + o["bwd_res_fused"] + (2.0 / 3.0) * (o["fwd_val"] + o["bwd_val_fused"])
    split = o["fwd_res_save"] + o["bwd_res_split"] + (2.0 / 3.0) * (o["fwd_val_save"] + o["bwd_val_split"])
    o["Mpts_fused"] = pts / fused / 1e3
    o["Mpts_split"] = pts / split / 1e3
    out[name] = o
for k, v in out.items():
    print(k, " ".join(f"{a}={b:.3f}" for a, b in v.items()))
