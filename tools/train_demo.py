"""End-to-end demo: train the reference's quick-start model (4-qubit cascade, angle encoding) with the
fused B200 train step and report the evaluation metrics of the reference's trainer script.
usage: python tools/train_demo.py [epochs] [batch]"""
import json, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
import qcpinn_b200 as qb
from qcpinn_b200.trainer import diffusion_eval, diffusion_train

epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda", 0)


def errors(ev):
    return {"error_u_pct": ev["error_u"], "error_f_pct": ev["error_f"]}


args = {"batch_size": batch, "epochs": epochs, "lr": 0.005, "seed": 1, "print_every": max(epochs // 10, 1),
        "num_qubits": 4, "num_quantum_layers": 1, "classic_network": [3, 50, 1], "q_ansatz": "cascade",
        "problem": "diffusion", "solver": "DV", "encoding": "None"}
torch.manual_seed(0)
out = os.path.join(R, "gpurun_out", "train_demo")
os.makedirs(out, exist_ok=True)
model = qb.DVPDESolver(args, qb.Logging(out), device=dev)
before = errors(diffusion_eval.evaluate(model))
t0 = time.time()
diffusion_train.train(model, nIter=epochs, batch_size=batch)
torch.cuda.synchronize()
dt = time.time() - t0
after = errors(diffusion_eval.evaluate(model))
print(json.dumps({"epochs": epochs, "batch": batch, "seconds": dt, "steps_per_s": (epochs + 1) / dt,
                  "loss_first": model.loss_history[0], "loss_last": model.loss_history[-1],
                  "eval_before": before, "eval_after": after}))
