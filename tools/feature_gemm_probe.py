"""Developer probe for DESIGN.md "what comes next" item 1 (cfg3 as a tensor-core contraction):
times the library GEMM  [6 B x 243] x [243 x 2430]  (six Taylor streams of the 3^5 B-half features
against the 10 x 243 x 243 factored feature matrix) in float32, TF32 and error-compensated 3xTF32,
and reports points/s per GEMM pass plus the error against float64.  Not a product path: it only
bounds what a fused tcgen05 kernel can reach.  usage: python tools/feature_gemm_probe.py [B]"""
import json, sys
import torch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda", 0)
torch.manual_seed(0)
K, N = 243, 2430
A = (torch.rand(6 * B, K, device=dev) * 2 - 1)
W = torch.randn(K, N, device=dev) / K ** 0.5
ref = (A[:4096].double() @ W.double())


def split(x):   # x = hi + lo with hi exactly representable in TF32 (10-bit mantissa)
    hi = (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
    return hi, x - hi


def timed(fn, reps=10):
    for _ in range(3):
        out = fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


res = {"B": B, "flops_per_pass": 2.0 * 6 * B * K * N}
torch.backends.cuda.matmul.allow_tf32 = False
ms, out = timed(lambda: A @ W)
res["fp32"] = {"ms": ms, "tflops": res["flops_per_pass"] / ms / 1e9, "points_per_s": B / ms * 1e3,
               "max_rel_err": float((out[:4096].double() - ref).abs().max() / ref.abs().max())}
torch.backends.cuda.matmul.allow_tf32 = True
ms, out = timed(lambda: A @ W)
res["tf32"] = {"ms": ms, "tflops": res["flops_per_pass"] / ms / 1e9, "points_per_s": B / ms * 1e3,
               "max_rel_err": float((out[:4096].double() - ref).abs().max() / ref.abs().max())}
Ah, Al = split(A)
Wh, Wl = split(W)
out3 = torch.empty(6 * B, N, device=dev)


def three_pass():
    torch.matmul(Ah, Wh, out=out3)
    out3.addmm_(Ah, Wl)
    out3.addmm_(Al, Wh)
    return out3


ms, out = timed(three_pass)
res["3xtf32"] = {"ms": ms, "tflops_equiv": res["flops_per_pass"] / ms / 1e9, "points_per_s": B / ms * 1e3,
                 "max_rel_err": float((out[:4096].double() - ref).abs().max() / ref.abs().max())}
print(json.dumps(res))
