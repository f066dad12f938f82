"""Run the conv kernels at the WRN-28-10 / batch-512 layer shapes (for ncu captures and per-layer timing).
usage: python scripts/prof_conv.py [--time]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neural-backed-decision-trees_b200"))
import torch
from nbdt._b200 import conv_ops as C

DEV = "cuda:0"
SHAPES = [  # (N, H, W, Cin, Cout, k, stride, count in WRN-28-10)
    (512, 32, 32, 64, 160, 3, 1, 1), (512, 32, 32, 160, 160, 3, 1, 7), (512, 32, 32, 160, 320, 3, 2, 1),
    (512, 16, 16, 320, 320, 3, 1, 7), (512, 16, 16, 320, 640, 3, 2, 1), (512, 8, 8, 640, 640, 3, 1, 7),
    (512, 32, 32, 64, 160, 1, 1, 1), (512, 32, 32, 160, 320, 1, 2, 1), (512, 16, 16, 320, 640, 1, 2, 1),
]
timing = "--time" in sys.argv
if "--main" in sys.argv:
    SHAPES = [SHAPES[1], SHAPES[3], SHAPES[5]]
total = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
for N, H, W, cin, cout, k, s, count in SHAPES:
    x = torch.randn((N, H, W, cin), device=DEV).bfloat16()
    w = torch.randn((cout, cin, k, k), device=DEV) * 0.05
    dy = torch.randn((N, H // s, W // s, cout), device=DEV).bfloat16()
    wp, wpt = C.pack_weights(w)
    flops = 2.0 * dy.numel() * cin * k * k
    def run(kind):
        if kind == "fwd": return C.conv_fwd(x, wp, k, s)
        if kind == "dgrad": return C.conv_dgrad(dy, wpt, (H, W), k, s)
        return C.conv_wgrad(x, dy, k, s)
    for kind in ("fwd", "dgrad", "wgrad"):
        if not timing:
            run(kind); continue
        for _ in range(2): run(kind)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run(kind)
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1) / 5
        total[kind] += ms * count
        print(f"{kind:6s} N{N} {H}x{W} {cin:4d}->{cout:4d} k{k} s{s}: {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s  (x{count})")
torch.cuda.synchronize()
if timing:
    print("per-step totals (ms):", {k: round(v, 2) for k, v in total.items()}, "sum", round(sum(total.values()), 2))
