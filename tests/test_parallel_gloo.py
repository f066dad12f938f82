"""world_size-2 CPU (gloo) test of the data-parallel host logic: flat-bucket gradient all-reduce and
row sharding.  (The tree-loss kernels need no collective; their shard-additivity is a -m gpu test.)"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "neural-backed-decision-trees_b200"))
    from nbdt._b200 import parallel

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)  # different initial weights per rank -> broadcast must fix it
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.BatchNorm1d(16), torch.nn.Linear(16, 4))
    parallel.broadcast_parameters(net)
    red = parallel.GradientAllReduce(net.parameters(), world, bucket_bytes=256)  # force several buckets
    assert len(red.buckets) > 2
    g = torch.Generator().manual_seed(7)
    x = torch.randn(12, 8, generator=g)
    y = torch.randint(4, (12,), generator=g)
    lo, hi = parallel.shard_rows(12, rank, world)
    # per-rank loss normalised by the GLOBAL batch so that sum over ranks == full-batch gradient
    net.eval()  # BN per-replica statistics would differ between 1 and 2 ranks (as in the reference)
    loss = torch.nn.functional.cross_entropy(net(x[lo:hi]), y[lo:hi], reduction="sum") / 12
    loss.backward()
    red.reduce()
    flat = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    if rank == 0:
        torch.save({"grad": flat, "w": torch.cat([p.data.reshape(-1) for p in net.parameters()])}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_matches_single_process(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.BatchNorm1d(16), torch.nn.Linear(16, 4))
    assert torch.equal(got["w"], torch.cat([p.data.reshape(-1) for p in net.parameters()]))
    g = torch.Generator().manual_seed(7)
    x = torch.randn(12, 8, generator=g)
    y = torch.randint(4, (12,), generator=g)
    net.eval()
    torch.nn.functional.cross_entropy(net(x), y).backward()
    want = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.allclose(got["grad"], want, rtol=1e-5, atol=1e-7)


def test_shard_rows_partition():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "neural-backed-decision-trees_b200"))
    from nbdt._b200.parallel import shard_rows

    for total in (0, 1, 7, 512, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_rows(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


class _FlatOpt:
    """CPU stand-in for FusedSGD's flat buffers: every .grad is a view of one flat gradient tensor."""

    def __init__(self, params):
        self.param_groups = [{"params": list(params)}]
        self.flat_grad = torch.zeros(sum(p.numel() for p in self.param_groups[0]["params"]))
        off = 0
        for p in self.param_groups[0]["params"]:
            p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
            off += p.numel()

    def assert_grads_are_views(self):
        off = 0
        for p in self.param_groups[0]["params"]:
            assert p.grad.data_ptr() == self.flat_grad[off:off + p.numel()].data_ptr()
            off += p.numel()


def _mlp():
    L, B = torch.nn.Linear, torch.nn.BatchNorm1d
    return torch.nn.Sequential(L(8, 32), B(32), torch.nn.ReLU(), L(32, 32), B(32), torch.nn.ReLU(), L(32, 32), B(32), torch.nn.ReLU(),
                               L(32, 32), B(32), torch.nn.ReLU(), L(32, 4))


def _overlap_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "neural-backed-decision-trees_b200"))
    from nbdt._b200 import parallel

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = _mlp()
    net.eval()
    opt = _FlatOpt(net.parameters())
    red = parallel.OverlappedGradientAllReduce(opt, world, pieces=4)
    assert len(red.bounds) >= 2
    g = torch.Generator().manual_seed(7)
    x, y = torch.randn(12, 8, generator=g), torch.randint(4, (12,), generator=g)
    lo, hi = parallel.shard_rows(12, rank, world)
    for step in range(2):  # twice: the piece bookkeeping must reset
        opt.flat_grad.zero_()
        (torch.nn.functional.cross_entropy(net(x[lo:hi]), y[lo:hi], reduction="sum") / 12).backward()
        red.finish()
        # the pieces tile the buffer back to front, each sent exactly once
        spans = red.last_issued
        assert spans[0][1] == opt.flat_grad.numel() and spans[-1][0] == 0
        assert all(a[0] == b[1] for a, b in zip(spans, spans[1:])) and len(spans) == len(red.bounds) + 1
    if rank == 0:
        torch.save({"grad": opt.flat_grad.clone()}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_allreduce_pieces_cover_the_buffer_and_match(tmp_path):
    """OverlappedGradientAllReduce on host tensors: hooks fire back to front, the pieces tile the flat buffer, the sum equals the
    full-batch gradient."""
    out = str(tmp_path / "r0.pt")
    mp.spawn(_overlap_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    net = _mlp()
    net.eval()
    g = torch.Generator().manual_seed(7)
    x, y = torch.randn(12, 8, generator=g), torch.randint(4, (12,), generator=g)
    torch.nn.functional.cross_entropy(net(x), y).backward()
    want = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.allclose(got["grad"], want, rtol=1e-5, atol=1e-7)
