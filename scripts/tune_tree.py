"""Sweep the code-generation knobs of the run-time-specialised tree kernels (csrc/tree_jit.cu) on the GPU.
usage: python scripts/tune_tree.py            (prints one JSON line per (hierarchy, knob setting))"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "neural-backed-decision-trees_b200"))
import torch  # noqa: E402

from bench_tree import peaks, time_point  # noqa: E402
from nbdt._b200 import lib as L  # noqa: E402
from nbdt.tree import Tree  # noqa: E402

KNOBS = ["NBDT_B200_KA_REGCAP", "NBDT_B200_KA_MINB", "NBDT_B200_KA_UNROLL", "NBDT_B200_KA_PREFETCH", "NBDT_B200_KA_MAXLD", "NBDT_B200_KA_FENCE", "NBDT_B200_KA_WARPS", "NBDT_B200_KA_EXPMUFU", "NBDT_B200_KA_ROWS", "NBDT_B200_KA_TEAMS", "NBDT_B200_KA_LOCKSTEP"]
# (regcap, minb, unroll, prefetch, maxld, fence, warps, expmufu, rows, teams, lockstep); -1 = automatic
D = (168, 0, 8, 0, 24, -1, 0, -1, 0, 0, 16)
E = [D, D[:9] + (4, 16), D[:9] + (8, 16)]
GRIDS = {100: E, 200: E, 1000: [D]}
if len(sys.argv) > 1:
    g = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
    GRIDS = {100: g, 200: g, 1000: g}
pk, src = peaks()
cases = [("CIFAR100", "induced", 100, 1 << 20), ("TinyImagenet200", "induced", 200, 1 << 19), ("Imagenet1000", "induced", 1000, 1 << 17)]
if os.environ.get("TUNE_CASES"):
    cases = [c for c in cases if str(c[2]) in os.environ["TUNE_CASES"].split(",")]
for dataset, hierarchy, C, rows in cases:
    for g in GRIDS[C]:
        for k, v in zip(KNOBS, tuple(g) + D[len(g):]):
            os.environ[k] = str(v)
        tree = Tree(dataset, hierarchy=hierarchy)
        for dtype in (torch.float32, torch.bfloat16):
            try:
                r = time_point(tree, "soft", rows, C, dtype, iters=10, warmup=2)
            except Exception as e:  # noqa: BLE001
                print(json.dumps({"dataset": dataset, "knobs": g, "error": str(e)[:200]}), flush=True)
                continue
            print(json.dumps({"dataset": dataset, "C": C, "knobs": g, "kernel": L.load().nbdt_tree_kernel_variant(tree.device_tree(torch.device("cuda:0")).handle).decode()[:8], "dtype": r["dtype"], "ms": round(r["ms"], 4),
                              "gbs": round(r["gbs"], 1), "frac": round(r["gbs"] / pk["hbm_gbs"], 4)}), flush=True)
