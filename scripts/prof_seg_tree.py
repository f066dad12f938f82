"""Class-major (N, C, H, W) tree kernels vs the reference's route (permute + reshape copy -> row kernels -> copy back) on
segmentation-sized tensors: CUDA events, rotating buffer sets larger than L2, one JSON line per (hierarchy, op, dtype, route).
usage: python scripts/prof_seg_tree.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "neural-backed-decision-trees_b200"))
import torch  # noqa: E402

from bench_tree import peaks  # noqa: E402
from nbdt._b200 import lib as L  # noqa: E402
from nbdt._b200 import ops  # noqa: E402
from nbdt.tree import Tree  # noqa: E402
from nbdt.utils import coerce_tensor, uncoerce_tensor  # noqa: E402

dev = torch.device("cuda:0")
pk, _ = peaks()
PKG = os.path.join(ROOT, "neural-backed-decision-trees_b200", "nbdt")
cases = [
    ("Cityscapes", dict(path_graph=os.path.join(PKG, "hierarchies", "Cityscapes", "graph-induced-HRNet-w18-v1.json"),
                        path_wnids=os.path.join(ROOT, "tests", "golden", "wnids_cityscapes.txt"), classes=[f"c{i}" for i in range(19)]), (4, 19, 512, 1024)),
    ("ADE20K", dict(hierarchy="induced-HRNet-w48"), (2, 150, 256, 512)),
]


def timed(fn, sets, iters=8):
    for s in sets[:2]:
        fn(*s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(*sets[i % len(sets)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for name, kw, shape in cases:
    tree = Tree(name, **kw)
    N, C, H, W = shape
    B = N * H * W
    for dtype in (torch.float32, torch.bfloat16):
        esz = 4 if dtype == torch.float32 else 2
        nsets = max(2, int(300e6 // (B * C * esz)) + 1)
        sets = [((2 * torch.randn(shape, device=dev)).to(dtype), torch.randint(C, (N, H, W), device=dev)) for _ in range(nsets)]
        routes = {
            ("soft_forward", "class-major kernel"): lambda x, y: ops.soft_forward_nchw(tree, x),
            ("soft_forward", "permute copy + row kernel + copy back"): lambda x, y: uncoerce_tensor(ops.soft_forward(tree, coerce_tensor(x).contiguous()), x.shape).contiguous(),
            ("soft_loss fwd+bwd", "class-major kernel"): lambda x, y: ops.tree_loss_nchw(tree, x, y, 1.0, 1.0),
            ("soft_loss fwd+bwd", "permute copy + row kernel + copy back"): lambda x, y: uncoerce_tensor(
                ops.tree_loss(tree, coerce_tensor(x).contiguous(), y.reshape(-1), "soft", 1.0, 1.0)[1], x.shape).contiguous(),
        }
        for (op, route), fn in routes.items():
            ms = timed(fn, sets)
            kernel = L.load().nbdt_tree_kernel_variant(tree.device_tree(dev).handle).decode()
            nbytes = (B * C * esz + 4 * B * C) if op == "soft_forward" else (2 * B * C * esz + 8 * B + 4)
            gbs = nbytes / (ms * 1e-3) / 1e9
            print(json.dumps({"hierarchy": name, "shape": list(shape), "rows": B, "dtype": str(dtype).split(".")[-1], "op": op, "route": route,
                              "kernel": kernel[:48], "ms": round(ms, 4), "alg_bytes": nbytes, "gbs": round(gbs, 1), "frac": round(gbs / pk["hbm_gbs"], 4)}), flush=True)
        del sets
        torch.cuda.empty_cache()
