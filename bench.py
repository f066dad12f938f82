#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json: images/sec of one WideResNet-28-10 +
SoftTreeSupLoss (CIFAR10 induced hierarchy) train step, batch 512 per GPU, bf16, synthetic 32x32 (--config 2, the default);
--config 3 / 4 / 5 run BASELINE's other configs (WRN + CIFAR100; ResNet18 + HardTreeSupLoss TinyImagenet200 at a fixed
global batch of 1024; EfficientNet-B0 + SoftTreeSupLoss ImageNet-1k at 256 images per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 2|3|4|5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

One "step" = zero_grad, forward, tree-supervision loss, backward, (grad all-reduce for N>1), SGD
update -- nothing skipped.  `value` is timed with CUDA events on device-resident inputs, bracketed by
barrier + synchronize, max over ranks.  `e2e` repeats the measurement through the public API with
pinned HOST batches copied in every step and the loss read back every step.

`--impl reference` times the UNMODIFIED reference (vendored by oracle/make_ref.py into the git-ignored oracle/_ref, which
travels to the GPU box): its own nbdt.loss and, where it builds offline, its own backbone, on all host cores, same config,
each step a bounded sample of the per-GPU batch (the full batch when it fits the time budget).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "neural-backed-decision-trees_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

# BASELINE.json configs.  `batch` is per GPU (weak scaling); config 4 fixes the GLOBAL batch (strong scaling).
# flops: algorithmic FLOPs per image of one train step = 6 x forward MACs (first-layer dgrad counted, not executed).
CONFIGS = {
    2: dict(metric="images/sec SoftNBDT WRN-28-10 train step", arch="wrn28_10_cifar10", dataset="CIFAR10",
            hierarchy="induced-wrn28_10_cifar10", loss="SoftTreeSupLoss", classes=10, hw=32, batch=512, scaling="weak", flops=31.46e9,
            workload="WideResNet-28-10 + SoftTreeSupLoss CIFAR10 (induced-wrn28_10_cifar10), batch=512/GPU synthetic 32x32 bf16"),
    3: dict(metric="images/sec SoftNBDT WRN-28-10 CIFAR100 train step", arch="wrn28_10_cifar100", dataset="CIFAR100",
            hierarchy="induced-wrn28_10_cifar100", loss="SoftTreeSupLoss", classes=100, hw=32, batch=512, scaling="weak", flops=31.46e9,
            workload="WideResNet-28-10 + SoftTreeSupLoss CIFAR100 (100-leaf induced-wrn28_10_cifar100), batch=512/GPU synthetic 32x32 bf16"),
    4: dict(metric="images/sec ResNet18 + HardTreeSupLoss TinyImagenet200 train step", arch="ResNet18", dataset="TinyImagenet200",
            hierarchy="induced-ResNet18", loss="HardTreeSupLoss", classes=200, hw=64, global_batch=1024, scaling="strong", flops=13.33e9,
            workload="ResNet18 + HardTreeSupLoss TinyImagenet200 (200-leaf induced-ResNet18), GLOBAL batch=1024 synthetic 64x64 bf16, gradient all-reduce"),
    5: dict(metric="images/sec EfficientNet-B0 + SoftTreeSupLoss ImageNet-1k train step", arch="efficientnet_b0", dataset="Imagenet1000",
            hierarchy="induced", loss="SoftTreeSupLoss", classes=1000, hw=224, batch=256, scaling="weak", flops=2.31e9,
            workload="EfficientNet-B0 + SoftTreeSupLoss ImageNet-1k (1000-leaf induced hierarchy), batch=256/GPU (2048 on 8 GPUs) synthetic 224x224 bf16"),
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU implementation on the host cores
# ------------------------------------------------------------------------------------------------
def saturating_tree_point(tree, kind, classes, dtype):
    """BASELINE's second metric, measured IN THIS RUN: the fused tree-loss kernel of this config's hierarchy on a row count that
    saturates HBM (logits + gradient sets larger than the 126 MB L2, rotated so that no launch finds its data cached)."""
    from bench_tree import time_point
    from nbdt._b200 import lib as L

    esz = torch.empty((), dtype=dtype).element_size()
    rows = max(1 << 14, int(420e6 // (2 * classes * esz + 8)) // 4096 * 4096)
    r = time_point(tree, kind, rows, classes, dtype, iters=12, warmup=3)
    name = L.load().nbdt_tree_kernel_variant(tree.device_tree(torch.device("cuda", torch.cuda.current_device())).handle).decode()
    return {"rows": rows, "classes": classes, "dtype": r["dtype"], "kind": kind, "ms": r["ms"], "alg_bytes": r["alg_bytes"], "gbs": r["gbs"],
            "kernel": name, "l2": r["l2"], "measured": "in this run (CUDA events, 12 back-to-back launches)"}


def usable_cores():
    """Fixed thread policy of the CPU arm: every core this process may run on (affinity mask, cgroup CPU quota)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period))))
    except (OSError, ValueError):
        pass
    return n


def reference_workload(cfg):
    """(net, criterion, kind, note) of the CPU arm for a config.  Runs in its own process (`--impl reference`, or a subprocess
    of the b200 arm): `import nbdt` here is the UNMODIFIED reference vendored under oracle/_ref (oracle/make_ref.py), its loss
    (reference nbdt/loss.py:191-266) and, where the reference can build it offline, its backbone.  Reference train step:
    main.py:233-239."""
    import torch.nn as nn
    from oracle import ref_shims

    root = ref_shims.vendored_root()
    if root is None and ref_shims.reference_available():
        root = ref_shims.REFERENCE_ROOT
    torch.manual_seed(0)
    if root is not None:
        for m in [m for m in sys.modules if m == "nbdt" or m.startswith("nbdt.")]:
            del sys.modules[m]
        sys.path[:] = [p for p in sys.path if not p.rstrip("/").endswith("neural-backed-decision-trees_b200")]
        import contextlib

        with contextlib.redirect_stdout(sys.stderr):  # the reference prints a `stty size` complaint at import (utils.py:175-180)
            ref_shims.install(root, strict_hash=False)
            import nbdt.loss as RL
            import nbdt.models as RM

        crit = getattr(RL, cfg["loss"])(dataset=cfg["dataset"], criterion=nn.CrossEntropyLoss(), hierarchy=cfg["hierarchy"])
        if cfg["arch"] == "ResNet18":
            net, bb = RM.ResNet18(num_classes=cfg["classes"]), "reference nbdt.models.ResNet18"
        elif cfg["arch"] == "efficientnet_b0":
            import torchvision

            net, bb = torchvision.models.efficientnet_b0(num_classes=cfg["classes"]), "torchvision efficientnet_b0 (what the reference imports)"
        else:
            from oracle import torch_port as P

            net, bb = P.EagerWRN(num_classes=cfg["classes"]), "eager WRN-28-10 restatement (pytorchcv is not installed offline)"
        return net, crit, "reference", f"unmodified reference {cfg['loss']} from oracle/_ref + {bb}"
    # no vendored reference: the pinned torch-CPU port of the same per-node algorithm
    from oracle import torch_port as P
    from oracle.tree_oracle import OracleTree

    pkg = os.path.join(ROOT, "neural-backed-decision-trees_b200", "nbdt")
    hier = cfg["hierarchy"]
    tree = OracleTree(os.path.join(pkg, f"hierarchies/{cfg['dataset']}/graph-{hier}.json"), os.path.join(pkg, f"wnids/{cfg['dataset']}.txt"))
    ce = nn.CrossEntropyLoss()
    fn = P.soft_tree_sup_loss if cfg["loss"].startswith("Soft") else P.hard_tree_sup_loss
    crit = lambda out, tgt: fn(tree, ce, out, tgt)  # noqa: E731
    if cfg["arch"] == "ResNet18":
        net = P.EagerResNet18(num_classes=cfg["classes"]) if hasattr(P, "EagerResNet18") else P.EagerResNet(num_classes=cfg["classes"])
    elif cfg["arch"] == "efficientnet_b0":
        import torchvision

        net = torchvision.models.efficientnet_b0(num_classes=cfg["classes"])
    else:
        net = P.EagerWRN(num_classes=cfg["classes"])
    return net, crit, "port", "oracle/torch_port.py (pinned port of the reference's per-node PyTorch path)"


def cpu_reference_run(cfg, steps, warmup, batch, budget_s=150.0):
    """Times `steps` train steps (after `warmup`) of the reference on the host cores.  Each step is a bounded sample of the
    config's per-GPU batch: the sample batch is the full batch when (steps + warmup) of them fit in `budget_s`, else the
    largest multiple of 8 that does (probed with one small step).  Returns a dict."""
    cores = usable_cores()
    torch.set_num_threads(cores)
    net, crit, kind, note = reference_workload(cfg)
    net.train()
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    g = torch.Generator().manual_seed(0)

    def make(b):
        return torch.randn(b, 3, cfg["hw"], cfg["hw"], generator=g), torch.randint(cfg["classes"], (b,), generator=g)

    def one(x, y):  # reference main.py:233-239
        opt.zero_grad()
        loss = crit(net(x), y)
        loss.backward()
        opt.step()
        return float(loss)

    probe_b = min(batch, 16)
    x, y = make(probe_b)
    one(x, y)  # first call pays allocator / thread-pool start-up
    t0 = time.time()
    one(x, y)
    rate = probe_b / max(time.time() - t0, 1e-3)  # images/s at a small batch (a lower bound for larger ones)
    fit = int(rate * budget_s / max(1, steps + warmup))
    b = batch if fit >= batch else max(8, fit // 8 * 8)
    x, y = make(b)
    for _ in range(warmup):
        one(x, y)
    times = []
    for _ in range(steps):
        t0 = time.time()
        one(x, y)
        times.append(time.time() - t0)
    total = sum(times)
    sample = (f"{steps} timed train steps (+{warmup} warm-up) of batch {b} of {batch} (fp32, torch {torch.__version__} CPU, {cores} threads, "
              f"{total:.1f} s): {note}; zero_grad/forward/loss/backward/SGD as in reference main.py:233-239")
    return {"value": b * steps / total, "cores": cores, "kind": kind, "sample": sample, "ms_per_step": total / steps * 1e3, "batch": b}


def run_reference(args, rank):
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    batch = cfg.get("batch") or max(1, cfg["global_batch"] // max(1, args.gpus))
    r = cpu_reference_run(cfg, args.steps, max(1, min(args.warmup, 2)), batch, budget_s=args.cpu_budget)
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": r["value"], "unit": "images/sec", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": cfg["scaling"],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "config": args.config},
        "cpu_baseline": {"value": r["value"], "unit": "images/sec", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampling of SM clock / throttle reasons while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.out = None

    def __enter__(self):
        try:
            self.out = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=self.out, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        if self.proc is None or self.out is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.out.flush()
        rows = []
        with open(self.out.name) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) >= 7:
                    rows.append(parts)
        os.unlink(self.out.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(r[col].lower().startswith("active") for r in rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(float(r[2]) for r in rows)}


def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist
    import torch.nn as nn
    from nbdt._b200 import ops
    from nbdt._b200 import parallel
    import nbdt.loss as NL
    import nbdt.models as NM

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback); "
                         "use --impl reference for the CPU arm")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = CONFIGS[args.config]
    if cfg["scaling"] == "strong":
        if cfg["global_batch"] % world:
            raise SystemExit(f"config {args.config}: global batch {cfg['global_batch']} is not divisible by {world} GPUs")
        B = cfg["global_batch"] // world  # fixed global batch, split over the ranks
    else:
        B = args.batch or cfg["batch"]    # per GPU (weak scaling)
    NC, HW = cfg["classes"], cfg["hw"]
    from nbdt._b200 import conv_ops, nn_ops
    from nbdt._b200.optim import FusedSGD

    if os.environ.get("NBDT_B200_WGRAD_STREAM", "1") != "0":
        nn_ops.enable_wgrad_overlap(True)  # weight-gradient kernels on a side stream; joined before all-reduce / SGD
    torch.manual_seed(0)
    factory = getattr(NM, cfg["arch"])
    net = (factory(num_classes=NC) if cfg["arch"] in ("ResNet18", "efficientnet_b0") else factory()).to(dev)
    net.train()
    crit = getattr(NL, cfg["loss"])(dataset=cfg["dataset"], criterion=nn.CrossEntropyLoss(), hierarchy=cfg["hierarchy"])
    kind = "soft" if cfg["loss"].startswith("Soft") else "hard"
    crit.batch_rows = B * world  # the mean runs over the global batch; shards add up under all-reduce(sum)
    if world > 1:
        parallel.broadcast_parameters(net)
    opt = FusedSGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    # gradient all-reduce: ONE collective on the flat buffer after backward.  NBDT_B200_ALLREDUCE_OVERLAP=1 sends it in pieces during
    # backward instead (eager launches only: hooks issuing NCCL inside a CUDA-graph capture hang the watchdog).  Measured at N=2
    # (profiles/r02_bench_n2_*.json): 18.92 ms/step for the single collective, 19.50 ms for four overlapped pieces -- NCCL's CTAs take
    # SMs from the persistent convolution kernels for longer than the 0.5 ms of exposed all-reduce they hide.
    overlapped = world > 1 and os.environ.get("NBDT_B200_ALLREDUCE_OVERLAP", "0") == "1"
    if overlapped:
        args.graph = False
    reducer = None
    if world > 1:
        reducer = (parallel.OverlappedGradientAllReduce(opt, world, pieces=int(os.environ.get("NBDT_B200_ALLREDUCE_PIECES", "4")),
                                                        bf16=os.environ.get("NBDT_B200_ALLREDUCE_BF16", "0") == "1")
                   if overlapped else parallel.FlatGradientAllReduce(opt, world))

    g = torch.Generator().manual_seed(1234 + rank)
    host_x = torch.randn(B, 3, HW, HW, generator=g).pin_memory()
    host_y = torch.randint(NC, (B,), generator=g).pin_memory()
    dev_x = host_x.to(dev, non_blocking=True)
    dev_y = host_y.to(dev, non_blocking=True)
    tree_ev = []

    def fwd_bwd(x, y, time_tree=False):
        opt.zero_grad()
        out = net(x)
        if time_tree:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        loss = crit(out, y)  # fused kernel: loss AND dloss/dlogits in this one launch
        if time_tree:
            e1.record()
            tree_ev.append((e0, e1))
        loss.backward()
        nn_ops.join_wgrad_streams(dev)  # weight-gradient kernels run on a side stream: join before capture ends / SGD
        opt.flush_grads()               # packed convolution gradients -> flat gradient buffer, one launch (inside the captured region)
        if overlapped:
            reducer.finish()  # the pieces were sent from autograd hooks while backward ran; send the head, join the comm stream
        return loss

    def finish():
        if reducer is not None and not overlapped:
            reducer.reduce()
        opt.step()

    def step(x, y, time_tree=False):
        loss = fwd_bwd(x, y, time_tree)
        finish()
        return loss

    # CUDA graph of zero_grad + forward + loss + backward (~420 launches) on static buffers; the gradient all-reduce and
    # the fused SGD launch stay eager.  Falls back to eager launches if capture is not possible.
    graph, graph_note = None, "disabled"
    static_x, static_y, static_loss = dev_x.clone(), dev_y.clone(), None

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # (no reference to a loss tensor may survive into the graph capture below: a live autograd graph keeps its AccumulateGrad
    # nodes on the default stream, and synchronising with that stream is illegal while capturing)
    loss_first = float(step(dev_x, dev_y).detach())
    for _ in range(max(args.warmup, 3) - 1):
        step(dev_x, dev_y)
    launches0 = ops.launch_count()
    step(dev_x, dev_y)
    launches_per_step = ops.launch_count() - launches0
    if args.graph:
        import gc

        gc.collect()  # drop any autograd graph of the warm-up steps before capturing
        try:
            torch.cuda.synchronize(dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                step(static_x, static_y)
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = fwd_bwd(static_x, static_y)
            graph_note = "fwd+loss+bwd captured"
        except Exception as e:  # noqa: BLE001
            graph, graph_note = None, f"capture failed: {type(e).__name__}: {str(e)[:80]}"
            torch.cuda.synchronize(dev)

    def run_step(x=None, y=None):
        """One full train step through the product path (graph replay when captured)."""
        if graph is None:
            return step(dev_x if x is None else x, dev_y if y is None else y)
        if x is not None:
            static_x.copy_(x, non_blocking=True)
            static_y.copy_(y, non_blocking=True)
        graph.replay()
        finish()
        return static_loss

    for _ in range(2):
        run_step()
    with ClockSampler(local_rank) as clocks:
        total_ms = timed(run_step, args.steps)
    launches = launches_per_step * args.steps
    clock_summary = clocks.summary()
    torch.cuda.synchronize(dev)
    # per-kernel CUDA-event timing of the same step (eager launches, events around every conv / loss launch)
    prof_steps = min(args.steps, 3)
    conv_ops.PROFILE = []
    overlap, nn_ops._OVERLAP_WGRAD = nn_ops._OVERLAP_WGRAD, False  # per-launch durations: no co-running side-stream kernels
    for _ in range(prof_steps):
        step(dev_x, dev_y, time_tree=True)
    torch.cuda.synchronize(dev)
    nn_ops._OVERLAP_WGRAD = overlap
    conv_prof, conv_ops.PROFILE = conv_ops.PROFILE, None
    nn_ops._OVERLAP_WGRAD = False  # as above: no weight-gradient kernels co-running on the side stream
    try:
        hbm_rows = hbm_kernel_profile(lambda: step(dev_x, dev_y), dev, load_peaks()[0]["hbm_gbs"])
    except Exception as e:  # noqa: BLE001
        hbm_rows = [{"error": f"{type(e).__name__}: {str(e)[:120]}"}]
    nn_ops._OVERLAP_WGRAD = overlap
    # The loss launch at the BASELINE batch is a few microseconds: time it as 50 launches inside one CUDA graph (event pairs
    # around single eager launches measure the Python launch path once the GPU has caught up with the CPU).
    tree_ms = sum(a.elapsed_time(b) for a, b in tree_ev) / max(len(tree_ev), 1)
    try:
        tl = torch.randn(B, NC, device=dev)
        tout = (torch.empty((), dtype=torch.float32, device=dev), torch.empty_like(tl))
        ops.tree_loss(crit.tree, tl, dev_y, kind, 1.0, 1.0, out=tout)
        torch.cuda.synchronize(dev)
        tg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(tg):
            for _ in range(50):
                ops.tree_loss(crit.tree, tl, dev_y, kind, 1.0, 1.0, out=tout)
        tg.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tg.replay()
        e1.record()
        e1.synchronize()
        tree_ms = e0.elapsed_time(e1) / 50  # memset of the loss scalar + kernel
    except Exception:  # noqa: BLE001  (keep the eager number)
        torch.cuda.synchronize(dev)
    conv_stats = {}
    for ckind, flops, e0, e1 in conv_prof:
        st = conv_stats.setdefault(ckind, [0.0, 0.0, 0])
        st[0] += flops
        st[1] += e0.elapsed_time(e1)
        st[2] += 1

    # end to end through the public API: pinned host batch in, loss value out, every step
    # Every step copies its batch from pinned host memory and its loss back to the host; the host reads loss k after it has
    # enqueued step k+1 (what an asynchronous logger does), so the GPU never waits for the Python launch path.
    host_loss = torch.empty(2, dtype=torch.float32).pin_memory()
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"i": 0, "last": None}

    def e2e_step():
        i = e2e_state["i"]
        x = host_x.to(dev, non_blocking=True)
        y = host_y.to(dev, non_blocking=True)
        loss = run_step(x, y)
        host_loss[i % 2].copy_(loss.detach().reshape(()), non_blocking=True)
        loss_ev[i % 2].record()
        if i > 0:
            loss_ev[(i - 1) % 2].synchronize()
            e2e_state["last"] = float(host_loss[(i - 1) % 2])
        e2e_state["i"] = i + 1

    e2e_step()
    e2e_ms = timed(e2e_step, args.steps)  # the closing barrier + synchronize of timed() drains the last read

    loss_last = float(run_step().detach())
    if not (loss_last == loss_last and abs(loss_last) < 1e4):  # NaN / blow-up: the timed work was not a valid train step
        raise SystemExit(f"bench.py: training diverged on the synthetic batch (loss {loss_first} -> {loss_last})")
    if rank != 0:
        return
    peaks, peak_src = load_peaks()
    ms_per_step = total_ms / args.steps
    value = B * world / (ms_per_step / 1e3)
    e2e_value = B * world / (e2e_ms / args.steps / 1e3)
    esz = 4  # the classifier GEMM emits fp32 logits
    tree_bytes = 2 * B * NC * esz + 8 * B + 4
    try:
        sat = saturating_tree_point(crit.tree, kind, NC, torch.float32)
        sat["frac"] = sat["gbs"] / peaks["hbm_gbs"]
    except Exception as e:  # noqa: BLE001
        sat = {"error": f"{type(e).__name__}: {str(e)[:120]}"}
    from nbdt._b200 import lib as L_

    tree_kernel = L_.load().nbdt_tree_kernel_variant(crit.tree.device_tree(dev).handle).decode()
    conv_flops = sum(v[0] for v in conv_stats.values())
    conv_ms = sum(v[1] for v in conv_stats.values())
    conv_tflops = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    peak_tf = peaks["bf16_tflops_sustained"]
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_conv_traffic.json")
    if os.path.exists(tpath):  # DRAM bytes per launch from the committed `ncu --set full` capture of the same kernels
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = tj["mean_bytes_per_launch"], tj["source"]
    line = {
        "metric": cfg["metric"], "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": cfg["scaling"],
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": cfg["workload"], "config": args.config, "global_batch": B * world, "per_gpu_batch": B, "parallelism": f"dp{world}",
                   "l2": "working set (activations ~5 GB/step) >> L2; no flush needed",
                   "backbone": parallel.BACKBONE_STATUS, "optimizer": "fused SGD momentum 0.9 wd 5e-4 (fp32 master weights)",
                   "cuda_graph": graph_note,
                   "allreduce": ("none (1 GPU)" if world == 1 else (f"fp32 flat gradient in {len(reducer.bounds) + 1} pieces issued during backward "
                                 f"(boundaries at elements {reducer.bounds})" if overlapped else "one fp32 all-reduce after backward"))},
        "clocks": clock_summary,
        "e2e": {"value": e2e_value, "unit": "images/sec", "h2d_bytes_per_step": host_x.numel() * 4 + host_y.numel() * 8,
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "roofline": {"kernel": "tap_gemm2h / tap_gemm2 / tap_gemm / wgrad kernels (tcgen05 implicit-GEMM conv fwd + dgrad + wgrad, every conv launch of a step)",
                     "bound": "tensor", "achieved": conv_tflops, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": conv_tflops / peak_tf, "flops": "algorithmic: 2*N*Ho*Wo*Cout*Cin*k*k with the layers' real channel counts (padding channels not counted)",
                     "step_frac": cfg["flops"] * B * world / (ms_per_step / 1e3) / 1e12 / world / peak_tf, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src + " (bf16_tflops_sustained)",
                     "per_kind": {k: {"tflops": v[0] / (v[1] * 1e-3) / 1e12, "ms_per_step": v[1] / prof_steps,
                                      "launches_per_step": v[2] / prof_steps} for k, v in conv_stats.items()},
                     "conv_share_of_step": (conv_ms / prof_steps) / ms_per_step,
                     "timed_with": f"CUDA events around each launch, {prof_steps} eager steps right after the timed region"},
        "tree_roofline": {"kernel": tree_kernel + f" ({kind} loss fwd+bwd; memset of the loss scalar + kernel, 50 launches in a CUDA graph)", "bound": "hbm",
                          "achieved": tree_bytes / (tree_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                          "frac": tree_bytes / (tree_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "launch_us": tree_ms * 1e3,
                          "note": f"{tree_bytes} B per launch at the BASELINE batch: latency-bound; `saturating` is the same kernel on rows that fill HBM",
                          "saturating": sat},
        "hbm_kernels": {"rows": hbm_rows, "bound": "hbm", "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "note": "memory-bound kernels of one eager step, CUDA events around each C-ABI call (device held back by a spin kernel while "
                                "the host enqueues); bytes = algorithmic tensor traffic from the call's size arguments"},
        "step_tflops": cfg["flops"] * B * world / (ms_per_step / 1e3) / 1e12,
        "loss": {"first_step": loss_first, "after_timed_steps": loss_last},  # sanity: the same synthetic batch is being fitted
    }
    if world == 1 and not args.no_cpu_baseline:
        # the reference needs `import nbdt` to be ITS package: a subprocess (this one has the product's nbdt loaded)
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", str(args.config), "--steps", "2",
                                  "--warmup", "1", "--cpu-budget", "40"], capture_output=True, text=True, timeout=600)
            ref = json.loads(out.stdout.strip().splitlines()[-1])
            line["cpu_baseline"] = ref["cpu_baseline"]
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "images/sec", "cores": usable_cores(), "kind": "reference",
                                    "sample": f"failed: {type(e).__name__}: {str(e)[:100]}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
# HBM roofline of the memory-bound kernels of a step, measured at the C ABI: every call of the functions below is bracketed by
# CUDA events for one eager step; the algorithmic bytes come from the call's own size arguments (bf16 activations = 2 B).
# ---------------------------------------------------------------------------------------------------------------------
def _has(a):
    return getattr(a, "value", a) is not None


def _ceil_div(a, b):
    return (a + b - 1) // b


HBM_KERNELS = {
    # name: (what, bytes(args))
    "nbdt_bn_stats": ("BatchNorm batch statistics (read x)", lambda a: a[1] * a[2] * 2),
    "nbdt_bn_stats_apply": ("BatchNorm forward: statistics pass + apply pass (+ activation, + residual)", lambda a: a[3] * a[4] * 2 * (3 + _has(a[1]))),
    "nbdt_bn_apply": ("BatchNorm apply + activation (+ residual)", lambda a: a[3] * a[4] * 2 * (2 + _has(a[1]))),
    "nbdt_bn_backward": ("BatchNorm backward: reduce pass + apply pass", lambda a: a[7] * a[8] * 2 * (5 + 2 * _has(a[1]) + _has(a[4]) + _has(a[6]))),
    "nbdt_bn_backward_silu": ("BatchNorm + SiLU backward: reduce pass + apply pass", lambda a: a[5] * a[6] * 2 * (5 + _has(a[3]))),
    "nbdt_avgpool_fwd": ("global average pool", lambda a: a[2] * a[3] * a[4] * 2),
    "nbdt_avgpool_bwd": ("global average pool backward", lambda a: a[2] * a[3] * a[4] * 2),
    "nbdt_dwconv_fwd": ("depthwise conv forward", lambda a: a[4] * a[7] * 2 * (a[5] * a[6] + _ceil_div(a[5], a[9]) * _ceil_div(a[6], a[9]))),
    "nbdt_dwconv_dgrad": ("depthwise conv data gradient", lambda a: a[3] * a[6] * 2 * (a[4] * a[5] + _ceil_div(a[4], a[8]) * _ceil_div(a[5], a[8]))),
    "nbdt_dwconv_wgrad": ("depthwise conv weight gradient (reads x and dy)", lambda a: a[3] * a[6] * 2 * (a[4] * a[5] + _ceil_div(a[4], a[8]) * _ceil_div(a[5], a[8]))),
    "nbdt_se_scale": ("squeeze-excitation scaling", lambda a: a[3] * a[4] * a[5] * 2 * 2),
    "nbdt_se_dscale": ("squeeze-excitation scale gradient (reads x and dy)", lambda a: a[3] * a[4] * a[5] * 2 * 2),
    "nbdt_sgd_step": ("fused SGD (read p, g, m; write p, m)", lambda a: a[3] * 4 * 5),
}


def hbm_kernel_profile(step_once, dev, hbm_peak):
    """One eager step with every HBM_KERNELS call timed by CUDA events -> per-function ms, GB/s and fraction of the HBM peak.
    The GPU is held back by a spin kernel while the host enqueues, so that the events bracket back-to-back device execution
    and not the Python launch path."""
    from nbdt._b200 import lib as L_

    lib = L_.load()
    records, originals = [], {}

    def wrap(name, fn, nbytes):
        def call(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*a)
            e1.record()
            try:
                b = int(nbytes(a))
            except Exception:  # noqa: BLE001
                b = 0
            records.append((name, b, e0, e1))
            return rc
        return call

    for name, (_, nbytes) in HBM_KERNELS.items():
        fn = getattr(lib, name, None)
        if fn is None:
            continue
        originals[name] = fn
        setattr(lib, name, wrap(name, fn, nbytes))
    try:
        torch.cuda.synchronize(dev)
        if hasattr(torch.cuda, "_sleep"):
            torch.cuda._sleep(int(2.5e8))  # ~0.13 s at 1.9 GHz: the host gets ahead of the device
        step_once()
        torch.cuda.synchronize(dev)
    finally:
        for name, fn in originals.items():
            setattr(lib, name, fn)
    out = {}
    for name, b, e0, e1 in records:
        st = out.setdefault(name, [0, 0.0, 0])
        st[0] += b
        st[1] += e0.elapsed_time(e1)
        st[2] += 1
    rows = []
    for name, (nb, ms, n) in sorted(out.items(), key=lambda kv: -kv[1][1]):
        gbs = nb / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        rows.append({"abi_call": name, "what": HBM_KERNELS[name][0], "launches_per_step": n, "ms_per_step": round(ms, 4),
                     "algorithmic_mb_per_step": round(nb / 1e6, 2), "gbs": round(gbs, 1), "frac": round(gbs / hbm_peak, 4)})
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("b200", "reference"), default="b200")
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json config (2 = the headline)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override (weak-scaling configs)")
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="seconds the whole CPU arm may take")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="launch every kernel eagerly (no CUDA graph)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "b200" and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    # Exactly ONE line on stdout: libraries that write to fd 1 from C (NCCL prints its version banner there when
    # NCCL_DEBUG is set) are sent to stderr for the whole run; the JSON line goes to the saved descriptor.
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(saved, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
