"""Developer tool: device timeline of ONE steady-state train step (CUDA-graph replay) from CUPTI
activity records (torch.profiler; nsys is not in the image).  Prints every kernel / memcpy of the
step with its stream, start offset and duration, plus the idle gaps of the step's critical stream.

usage: python tools/timeline.py [points] [f64|f32]      (default 524288 f64)
       torchrun --nproc-per-node N tools/timeline.py [points per rank] ...   (rank 0's timeline)
"""
import os, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402  (model_args)
import qcpinn_b200 as qb  # noqa: E402
from qcpinn_b200.trainer.diffusion_train import TrainStep  # noqa: E402

from qcpinn_b200.dist import init_from_env  # noqa: E402
from qcpinn_b200.trainer.diffusion_train import _make_averager  # noqa: E402

points = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
dtype = sys.argv[2] if len(sys.argv) > 2 else "f64"
rank, world, local = init_from_env()       # under torchrun: `points` per rank, rank 0 prints its timeline
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
torch.manual_seed(0)
model = qb.DVPDESolver(bench.model_args(dtype), qb.Logging(os.path.join(tempfile.gettempdir(), f"qcp_tl{rank}")), device=dev)
step = TrainStep(model, points, _make_averager(model), host_sync=False)
torch.manual_seed(1234 + rank)
while not step.steady():
    step()
for _ in range(5):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
import json  # noqa: E402

if world > 1:
    import torch.distributed as dist

    step.close()
    del step, model
    import gc

    gc.collect()
    dist.barrier()
    dist.destroy_process_group()
    if rank != 0:
        sys.exit(0)

trace = os.path.join(tempfile.gettempdir(), "qcp_timeline.json")
prof.export_chrome_trace(trace)
ev = [e for e in json.load(open(trace))["traceEvents"]
      if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
if not ev:
    raise SystemExit("no device activity records (CUPTI unavailable?)")
# three replays of the same graph: equal thirds
per = len(ev) // 3
one, nxt = ev[per:2 * per], ev[2 * per]
t0 = one[0]["ts"]
span = nxt["ts"] - t0
print(f"# {points} points{' per rank, ' + str(world) + ' ranks (rank 0 shown)' if world > 1 else ''} {dtype}: {len(one)} device activities in the middle replay of 3 "
      f"({span:.1f} us from its first activity to the next step's)")
print("# start_us  dur_us  stream  name")
busy = []
for e in one:
    s0 = e["ts"] - t0
    busy.append((s0, s0 + e["dur"]))
    print(f"{s0:9.1f} {e['dur']:8.1f} {e.get('args', {}).get('stream', '?')!s:>6}  {e['name'][:90]}")
busy.sort()
cover, end = 0.0, 0.0
for a, b in busy:
    if b > end:
        cover += b - max(a, end)
        end = b
print(f"# device busy (any stream) {cover:.1f} us of {span:.1f} us; idle {span - cover:.1f} us")
