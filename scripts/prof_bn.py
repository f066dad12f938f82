"""Time the memory-bound backbone kernels (BN stats / apply / backward) at the WRN-28-10 / batch-512 activation
shapes against their algorithmic HBM bytes.  usage: python scripts/prof_bn.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neural-backed-decision-trees_b200"))
import torch
from nbdt._b200 import lib as L
from nbdt._b200.ops import _p

DEV = torch.device("cuda:0")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
lib = L.load()
st = L.stream_ptr(DEV)


def timeit(fn, iters=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / iters


for (N, H, W, C, count) in [(512, 32, 32, 160, 8), (512, 16, 16, 320, 8), (512, 8, 8, 640, 9), (512, 32, 32, 64, 1)]:
    M = N * H * W
    # several distinct buffers so consecutive iterations do not hit L2 (126 MB): rotate over >= 4 tensors
    nb = max(2, int(600e6 // (M * C * 2)) + 1)
    xs = [torch.randn((N, H, W, C), device=DEV).bfloat16() for _ in range(nb)]
    ys = [torch.empty_like(xs[0]) for _ in range(nb)]
    dys = [torch.randn((N, H, W, C), device=DEV).bfloat16() for _ in range(nb)]
    g = torch.ones(C, device=DEV); b = torch.zeros(C, device=DEV)
    ss = torch.empty((2, C), device=DEV); mi = torch.empty((2, C), device=DEV)
    sums = torch.empty((lib.nbdt_bn_scratch_floats(C),), device=DEV)
    it = [0]
    def nxt():
        it[0] = (it[0] + 1) % nb
        return it[0]
    tb = M * C * 2
    def stats():
        i = nxt(); L.check(lib.nbdt_bn_stats(_p(xs[i]), M, C, C, _p(g), _p(b), None, None, 0.1, 1e-5, _p(sums), _p(mi), _p(ss), st))
    def apply():
        i = nxt(); L.check(lib.nbdt_bn_apply(_p(xs[i]), None, _p(ys[i]), M, C, _p(ss), 1, st))
    def apply_res():
        i = nxt(); L.check(lib.nbdt_bn_apply(_p(xs[i]), _p(dys[i]), _p(ys[i]), M, C, _p(ss), 1, st))
    def bwd_act():
        i = nxt(); L.check(lib.nbdt_bn_backward(_p(xs[i]), _p(ys[i]), None, _p(dys[i]), None, _p(ys[(i + 1) % nb]), None, M, C, C, _p(g), _p(mi), _p(sums), st))
    def bwd_extra():
        i = nxt(); L.check(lib.nbdt_bn_backward(_p(xs[i]), None, _p(ss), _p(dys[i]), _p(dys[(i + 1) % nb]), _p(ys[(i + 1) % nb]), None, M, C, C, _p(g), _p(mi), _p(sums), st))
    def bwd():
        i = nxt(); L.check(lib.nbdt_bn_backward(_p(xs[i]), None, _p(ss), _p(dys[i]), None, _p(ys[(i + 1) % nb]), None, M, C, C, _p(g), _p(mi), _p(sums), st))
    for name, fn, nbytes in (("bn_stats(+finalize)", stats, tb), ("bn_apply", apply, 2 * tb), ("bn_apply+res", apply_res, 3 * tb),
                             ("bn_backward(mask recomputed)", bwd, 2 * tb + 3 * tb), ("bn_backward(mask recomputed, +skip grad)", bwd_extra, 2 * tb + 4 * tb),
                             ("bn_backward(mask from act)", bwd_act, 3 * tb + 4 * tb)):
        ms = timeit(fn)
        print(f"M={M} C={C:4d} {name:38s} {ms*1e3:8.1f} us  {nbytes/ms/1e6:7.0f} GB/s  {nbytes/ms/1e6/PEAK*100:5.1f}% of {PEAK:.0f}  (x{count}/step)", flush=True)
