"""2-GPU data-parallel train step (NCCL): sharded batch + flat gradient all-reduce + fused SGD must equal the
single-process result of running the two shards one after the other (per-replica BatchNorm statistics, as in the
reference's DataParallel/DDP) and summing their gradients.  Skipped with fewer than 2 GPUs."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _setup_path():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, "neural-backed-decision-trees_b200"), root):
        if p not in sys.path:
            sys.path.insert(0, p)


def _make(dev):
    import torch.nn as nn
    from nbdt.loss import SoftTreeSupLoss
    from nbdt.models import ResNet18

    torch.manual_seed(0)
    net = ResNet18(num_classes=10).to(dev)
    net.train()
    crit = SoftTreeSupLoss(dataset="CIFAR10", criterion=nn.CrossEntropyLoss(), hierarchy="induced-ResNet18")
    g = torch.Generator().manual_seed(3)
    x = torch.randn(32, 3, 32, 32, generator=g)
    y = torch.randint(10, (32,), generator=g)
    return net, crit, x, y


def _worker(rank, world, port, out, overlapped=False, graph=False):
    _setup_path()
    import torch.distributed as dist
    from nbdt._b200 import parallel
    from nbdt._b200.optim import FusedSGD

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    net, crit, x, y = _make(dev)
    parallel.broadcast_parameters(net)
    opt = FusedSGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    red = parallel.OverlappedGradientAllReduce(opt, world, pieces=4) if overlapped else parallel.FlatGradientAllReduce(opt, world)
    if overlapped:
        assert len(red.bounds) >= 2, red.bounds  # pieces were found (BatchNorm biases as triggers)
    crit.batch_rows = x.shape[0]
    lo, hi = parallel.shard_rows(x.shape[0], rank, world)
    xs, ys = x[lo:hi].to(dev), y[lo:hi].to(dev)

    def fwd_bwd():
        opt.zero_grad()
        loss = crit(net(xs), ys)
        loss.backward()
        red.reduce()   # overlapped: sends the head piece and joins the communication stream
        return loss

    if graph:  # the collectives are captured together with forward + backward
        import gc

        state = {k: v.clone() for k, v in net.state_dict().items()}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            fwd_bwd()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        gc.collect()
        net.load_state_dict(state)  # the warm-up step moved the BatchNorm running statistics
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            loss = fwd_bwd()
        net.load_state_dict(state)
        g.replay()
    else:
        loss = fwd_bwd()
    grad = opt.flat_grad.clone()
    opt.step()
    total = loss.detach().clone()
    dist.all_reduce(total)
    if rank == 0:
        torch.save({"grad": grad.cpu(), "param": opt.flat_param.cpu(), "loss": total.cpu()}, out)
    dist.barrier(device_ids=[rank])
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("overlapped,graph", [(False, False), (True, False)])  # NCCL pieces issued from autograd hooks do not survive CUDA-graph capture (watchdog hang, measured): eager only
def test_two_rank_step_equals_sharded_single_process(tmp_path, overlapped, graph):
    _setup_path()
    import torch.multiprocessing as mp
    from nbdt._b200 import parallel
    from nbdt._b200.optim import FusedSGD

    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, overlapped, graph), nprocs=2, join=True)
    got = torch.load(out)

    dev = torch.device("cuda", 0)
    net, crit, x, y = _make(dev)
    opt = FusedSGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    crit.batch_rows = x.shape[0]
    opt.zero_grad()
    state = {k: v.clone() for k, v in net.state_dict().items()}
    losses = []
    for r in range(2):  # same shards, one after the other; BN running stats restart from the same state per "replica"
        lo, hi = parallel.shard_rows(x.shape[0], r, 2)
        if r == 1:
            bufs = {k: v for k, v in state.items() if "running" in k or "num_batches" in k}
            net.load_state_dict({**net.state_dict(), **bufs})
        loss = crit(net(x[lo:hi].to(dev)), y[lo:hi].to(dev))
        loss.backward()
        losses.append(loss.detach())
        opt.flush_grads()  # packed convolution gradients -> flat buffer (step() / the all-reduce do this themselves)
    grad = opt.flat_grad.clone().cpu()
    assert torch.allclose(got["loss"], (losses[0] + losses[1]).cpu(), rtol=1e-4)
    cos = torch.nn.functional.cosine_similarity(got["grad"].double(), grad.double(), dim=0)
    assert cos > 0.9999, float(cos)
    assert torch.allclose(got["grad"], grad, rtol=2e-2, atol=2e-3 * grad.abs().max().item())
