"""A/B of the convolution forward with and without BatchNorm statistics in the epilogue (CUDA events around back-to-back launches)."""
import json, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neural-backed-decision-trees_b200"))
import torch
from nbdt._b200 import conv_ops as C

dev = "cuda:0"
shapes = [(512, 32, 32, 160, 160, 3, 1), (512, 32, 32, 160, 320, 3, 2), (512, 16, 16, 320, 320, 3, 1), (512, 16, 16, 320, 640, 3, 2), (512, 8, 8, 640, 640, 3, 1), (512, 32, 32, 64, 160, 3, 1)]
for N, H, W, cin, cout, k, s in shapes:
    x = torch.randn((N, H, W, cin), device=dev).bfloat16()
    w = torch.randn((cout, cin, k, k), device=dev) / (cin * k * k) ** 0.5
    wp, _ = C.pack_weights(w, cin_pad=cin, cout_pad=cout)
    res = torch.randn((N, H // s, W // s, cout), device=dev).bfloat16()
    out = {}
    for stats in (False, True):
        for _ in range(3):
            C.conv_fwd(x, wp, k, s, res, bn_stats=stats)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            C.conv_fwd(x, wp, k, s, res, bn_stats=stats)
        e1.record()
        torch.cuda.synchronize()
        out["stats" if stats else "plain"] = e0.elapsed_time(e1) / 10
    flops = 2.0 * N * (H // s) * (W // s) * cout * cin * k * k
    print(json.dumps({"shape": [N, H, W, cin, cout, k, s], "plain_ms": round(out["plain"], 4), "stats_ms": round(out["stats"], 4),
                      "plain_tflops": round(flops / out["plain"] / 1e9, 1), "stats_tflops": round(flops / out["stats"] / 1e9, 1)}), flush=True)
