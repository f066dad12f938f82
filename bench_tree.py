"""Saturating row sweep of the fused tree-loss kernel: achieved HBM GB/s per (rows, classes, dtype).

Algorithmic bytes per launch: 2*B*C*sizeof(logit) + 8*B + 4  (read logits + targets,
write gradient + loss).  Buffers are larger than L2 (126 MB) where stated, otherwise L2 is flushed
between iterations.  Prints one JSON line per point; used by bench.py for the tree_roofline block.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "neural-backed-decision-trees_b200"))

import torch  # noqa: E402


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def time_point(tree, kind, rows, C, dtype, iters=20, warmup=3, group=None):
    """Average duration of `iters` back-to-back launches (one CUDA-event pair around them, so the launch path of
    Python is hidden behind the previous kernel) rotating over enough (logits, targets, grad) sets that no launch
    finds its inputs or the previous owner of its outputs in the 126 MB L2."""
    from nbdt._b200 import ops

    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(0)
    esz = torch.empty((), dtype=dtype).element_size()
    alg_bytes = 2 * rows * C * esz + 8 * rows + 4
    nsets = min(iters, max(2, int(400e6 // alg_bytes) + 2))
    sets = []
    for _ in range(nsets):
        x = (3 * torch.randn((rows, C), device=dev, generator=gen)).to(dtype)
        y = torch.randint(C, (rows,), device=dev, generator=gen)
        sets.append((x, y, (torch.empty((), dtype=torch.float32, device=dev), torch.empty_like(x))))
    for i in range(warmup):
        x, y, out = sets[i % nsets]
        ops.tree_loss(tree, x, y, kind, 1.0, 1.0, out=out)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(iters):
        x, y, out = sets[i % nsets]
        ops.tree_loss(tree, x, y, kind, 1.0, 1.0, out=out)
    e.record()
    e.synchronize()
    ms = s.elapsed_time(e) / iters
    return {"rows": rows, "C": C, "dtype": str(dtype).replace("torch.", ""), "kind": kind, "ms": ms,
            "alg_bytes": alg_bytes, "gbs": alg_bytes / ms / 1e6,
            "l2": f"{iters} back-to-back launches rotating over {nsets} buffer sets ({nsets * alg_bytes / 1e6:.0f} MB > L2)"}


def cpu_baseline_point(dataset, hierarchy, C, kind="soft", budget_s=3.0):
    """The plain-C restatement of the reference's per-node algorithm (oracle/tree_oracle.c, OpenMP over rows) on the
    host cores, on a bounded sample of the same workload: GB/s of the same algorithmic bytes (fp32 logits)."""
    import time

    import numpy as np

    sys.path.insert(0, ROOT)
    from nbdt.utils import dataset_to_default_path_wnids, hierarchy_to_path_graph
    from oracle import tree_oracle as O
    from oracle import tree_oracle_c as OC

    ct = OC.CTree(O.OracleTree(hierarchy_to_path_graph(dataset, hierarchy), dataset_to_default_path_wnids(dataset)))
    rng = np.random.default_rng(0)
    rows = 4096
    fn = OC.soft_tree_sup_loss if kind == "soft" else OC.hard_tree_sup_loss
    while True:
        x = (3 * rng.standard_normal((rows, C))).astype(np.float32)
        y = rng.integers(0, C, rows)
        fn(ct, x[:256], y[:256])
        t0 = time.time()
        fn(ct, x, y)
        dt = time.time() - t0
        if dt > budget_s / 4 or rows >= (1 << 22):
            break
        rows *= 4
    alg_bytes = 2 * rows * C * 4 + 8 * rows + 4
    return {"value": alg_bytes / dt / 1e9, "unit": "GB/s", "cores": OC.threads(), "kind": "port",
            "sample": f"{rows} rows x {C} classes fp32, {kind} loss fwd+bwd, oracle/tree_oracle.c (per-node restatement of the reference, "
                      f"fp64 arithmetic, OpenMP over rows), {dt * 1e3:.0f} ms"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--group", type=int, default=None)
    ap.add_argument("--cpu-only", action="store_true", help="only the CPU baselines (no GPU needed)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.cpu_only:
        for dataset, hierarchy, C in (("CIFAR10", "induced", 10), ("CIFAR100", "induced", 100), ("Imagenet1000", "induced", 1000)):
            print(json.dumps({"dataset": dataset, "C": C, "cpu_baseline": cpu_baseline_point(dataset, hierarchy, C)}), flush=True)
        return
    from nbdt.tree import Tree

    pk, src = peaks()
    if args.group:
        os.environ["NBDT_B200_TREE_G"] = str(args.group)
    cases = [("CIFAR10", "induced", 10), ("LookIntoPerson", "induced-HRNet-w48-cls20", 20), ("CIFAR100", "induced", 100),
             ("ADE20K", "induced-HRNet-w48", 150), ("TinyImagenet200", "induced", 200), ("Imagenet1000", "induced", 1000)]
    rows_for = {10: [1 << 18, 1 << 22], 20: [1 << 21], 100: [1 << 16, 1 << 20], 150: [1 << 19], 200: [1 << 19], 1000: [1 << 17]}
    for dataset, hierarchy, C in cases:
        tree = Tree(dataset, hierarchy=hierarchy)
        for rows in rows_for[C][-1:] if args.quick else rows_for[C]:
            for dtype in (torch.float32, torch.bfloat16):
                for kind in ("soft",) if args.quick else ("soft", "hard"):
                    r = time_point(tree, kind, rows, C, dtype)
                    r.update(dataset=dataset, frac=r["gbs"] / pk["hbm_gbs"], peak=pk["hbm_gbs"], peak_source=src)
                    if not args.no_cpu_baseline and dtype == torch.float32 and kind == "soft" and rows == rows_for[C][-1]:
                        r["cpu_baseline"] = cpu_baseline_point(dataset, hierarchy, C)
                    print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
