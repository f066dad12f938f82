"""EfficientNet-B0 + SoftNBDT (ImageNet-1k induced hierarchy) inference throughput on one B200 -- the per-GPU share of
BASELINE config 5 (batch 2048 over 8 GPUs = 256 images of 224x224 per GPU).  Synthetic images, random-init weights,
bf16 activations.  Prints one JSON line: images/sec (device-resident input, CUDA events over `--steps` forwards) and the
same through pinned-host input + D2H of the class probabilities.  Not the headline bench (bench.py)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "neural-backed-decision-trees_b200"))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    from nbdt._b200 import ops
    from nbdt.model import SoftNBDT
    from nbdt.models import efficientnet_b0

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = efficientnet_b0().to(dev).eval()
    model = SoftNBDT(dataset="Imagenet1000", model=net, hierarchy="induced-efficientnet_b7b")
    host_x = torch.randn(args.batch, 3, args.size, args.size).pin_memory()
    x = host_x.to(dev)
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            model(x)
        l0 = ops.launch_count()
        model(x)
        launches = ops.launch_count() - l0
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.steps):
            p = model(x)
        e.record()
        e.synchronize()
        ms = s.elapsed_time(e) / args.steps
        s.record()
        for _ in range(args.steps):
            p = model(host_x.to(dev, non_blocking=True)).cpu()
        e.record()
        e.synchronize()
        ms_e2e = s.elapsed_time(e) / args.steps
    print(json.dumps({"metric": "images/sec EfficientNet-B0 + SoftNBDT inference (ImageNet-1k induced hierarchy)",
                      "value": args.batch / ms * 1e3, "unit": "images/sec", "n_gpus": 1, "ms_per_step": ms,
                      "e2e": {"value": args.batch / ms_e2e * 1e3, "unit": "images/sec",
                              "h2d_bytes_per_step": host_x.numel() * 4, "d2h_bytes_per_step": args.batch * 1000 * 4},
                      "config": {"workload": f"EfficientNet-B0 + SoftNBDT, batch={args.batch} synthetic {args.size}x{args.size} bf16"},
                      "own_kernel_launches_per_forward": launches, "dtype": "bf16", "data": "synthetic"}))


if __name__ == "__main__":
    main()
