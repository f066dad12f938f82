"""TEST INFRASTRUCTURE: run the code that csrc/tree_jit.cu GENERATES for a hierarchy on the CPU.

The generator emits one straight-line function per hierarchy (`ka_row`: one logits row in, loss + gradient row out).  With
-DKA_HOST the same text compiles as plain C++ against the few primitives below (exp2f for ex2.approx, 1/x for rcp.approx,
memcpy loads/stores), so the generated algebra -- DFS prefix sums, the target step, signed child probabilities, the
hard-loss path walk, k-ary nodes, register/shared-memory value placement -- is checked against the oracle without a GPU.
The GPU tests then only have to cover what differs on the device (approximate MUFU ops, launch geometry, the cubin cache).
"""
import ctypes
import hashlib
import os
import subprocess
import tempfile

import numpy as np

HARNESS = r'''
#include <cmath>
#include <cstdint>
#include <cstring>
#define KA_HOST 1
#define KA_DEV static inline
typedef unsigned int u32;
typedef unsigned short u16;
struct uint4 { u32 x, y, z, w; };
struct uint2 { u32 x, y; };
KA_DEV float ex2(float v) { return exp2f(v); }
KA_DEV float lg2(float v) { return log2f(v); }
KA_DEV float rcp(float v) { return 1.0f / v; }
KA_DEV float u2f(u32 v) { float f; memcpy(&f, &v, 4); return f; }
KA_DEV u32 f2u(float v) { u32 u; memcpy(&u, &v, 4); return u; }
KA_DEV float h2f(u32 v) {
  const u32 h = v & 0xffffu, s = (h >> 15) & 1u, e = (h >> 10) & 31u, m = h & 1023u;
  float r;
  if (e == 0) r = ldexpf((float)m, -24);
  else if (e == 31) r = m ? NAN : INFINITY;
  else r = ldexpf((float)(m | 1024u), (int)e - 25);
  return s ? -r : r;
}
KA_DEV u32 bf16_rn(float f) { u32 u = f2u(f); u += 0x7fffu + ((u >> 16) & 1u); return u >> 16; }
KA_DEV u32 f16_rn(float f) {  // round to nearest even, normal range + flush of tiny values (enough for gradients in tests)
  const u32 u = f2u(f), s = (u >> 16) & 0x8000u;
  const float a = fabsf(f);
  if (a >= 65520.f) return s | 0x7c00u;
  if (a < 6.103515625e-05f) { const float q = a * 16777216.f; return s | (u32)lrintf(q); }
  int e; const float m = frexpf(a, &e);  // a = m * 2^e, m in [0.5, 1)
  const u32 mant = (u32)lrintf(m * 2048.f);  // 11 bits incl. leading one, may round up to 2048
  u32 hh = ((u32)(e + 14) << 10) + (mant - 1024u);
  return s | hh;
}
KA_DEV u32 pack_bf16(float lo, float hi) { return bf16_rn(lo) | (bf16_rn(hi) << 16); }
KA_DEV u32 pack_f16(float lo, float hi) { return f16_rn(lo) | (f16_rn(hi) << 16); }
KA_DEV uint4 ld16(const char* p) { uint4 r; memcpy(&r, p, 16); return r; }
KA_DEV uint2 ld8(const char* p) { uint2 r; memcpy(&r, p, 8); return r; }
KA_DEV u32 ld4(const char* p) { u32 r; memcpy(&r, p, 4); return r; }
KA_DEV u32 ld2(const char* p) { u16 r; memcpy(&r, p, 2); return r; }
KA_DEV void st16(char* p, u32 a, u32 b, u32 c, u32 d) { u32 v[4] = {a, b, c, d}; memcpy(p, v, 16); }
KA_DEV void st8(char* p, u32 a, u32 b) { u32 v[2] = {a, b}; memcpy(p, v, 8); }
KA_DEV void st4(char* p, u32 a) { memcpy(p, &a, 4); }
KA_DEV void st2(char* p, u32 a) { u16 v = (u16)a; memcpy(p, &v, 2); }
typedef float* KA_XSB;
#include <pthread.h>
#include <thread>
#include <vector>
static pthread_barrier_t ka_bar;
static bool ka_bar_on = false;
static void ka_host_bar() { if (ka_bar_on) pthread_barrier_wait(&ka_bar); }
#define KA_BAR() ka_host_bar()
#include "gen.cu"
extern "C" void ka_run(const KAParams* p, float* loss) {
  static float Xs[131072];
  float total = 0.f;
  if (KA_W > 1) { pthread_barrier_init(&ka_bar, nullptr, KA_W); ka_bar_on = true; }
  for (long long row = 0; row < p->B; row += KA_R) {  // KA_TPB == 1 on the host: a "block" is the KA_R rows of one thread, its warps are host threads
    const int nrows = (int)(p->B - row < KA_R ? p->B - row : KA_R);
    ka_sweep_a(*p, row, nrows, Xs, 0);
#if KA_R == 1
    const KY yy = p->targets ? p->targets[row] : 0;
    const KB live = true;
#else
    KY yy; yy.a = p->targets ? p->targets[row] : 0; yy.b = (p->targets && nrows > 1) ? p->targets[row + 1] : 0;
    const KB live = mkb(true, nrows > 1);
#endif
    float acc[KA_W] = {0.f};
    if (KA_W == 1) {
      ka_part(0, *p, yy, live, Xs, 0u, acc[0]);
    } else {
      std::vector<std::thread> th;
      for (int w = 0; w < KA_W; ++w) th.emplace_back([&, w] { ka_part(w, *p, yy, live, Xs, 0u, acc[w]); });
      for (auto& t : th) t.join();
    }
    for (int w = 0; w < KA_W; ++w) total += acc[w];
    if (p->grad || p->class_probs) ka_sweep_b(*p, row, nrows, Xs, 0);
  }
  if (KA_W > 1) { pthread_barrier_destroy(&ka_bar); ka_bar_on = false; }
  if (loss) *loss = total;
}
'''


class KAParams(ctypes.Structure):
    _fields_ = [("x", ctypes.c_void_p), ("B", ctypes.c_longlong), ("ld", ctypes.c_longlong), ("targets", ctypes.c_void_p),
                ("class_probs", ctypes.c_void_p), ("loss", ctypes.c_void_p), ("grad", ctypes.c_void_p),
                ("w_xent", ctypes.c_float), ("w_tree", ctypes.c_float), ("inv_rows", ctypes.c_float),
                ("grad_scale", ctypes.c_float), ("hw", ctypes.c_longlong)]


def tree_desc(tree):
    from nbdt._b200 import lib as L
    from nbdt._b200.tables import flatten

    child_ptr, child_node, root = flatten(tree)
    desc = L.TreeDesc(len(tree.wnids_leaves), len(tree.inodes), root,
                      child_ptr.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                      child_node.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    return desc, (child_ptr, child_node)


def generated_source(tree, mode, dtype, vec):
    from nbdt._b200 import lib as L

    lib = L.load()
    fn = lib.nbdt_tree_jit_source
    fn.restype = ctypes.c_longlong
    fn.argtypes = [ctypes.POINTER(L.TreeDesc), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_longlong]
    desc, keep = tree_desc(tree)
    n = fn(ctypes.byref(desc), mode, dtype, vec, None, 0)
    if n < 0:
        return None
    buf = ctypes.create_string_buffer(n + 1)
    fn(ctypes.byref(desc), mode, dtype, vec, buf, n + 1)
    return buf.value.decode()


_cache = {}


def host_kernel(tree, mode, dtype, vec):
    """ctypes handle of the generated row function compiled for the host."""
    src = generated_source(tree, mode, dtype, vec)
    assert src is not None
    key = hashlib.sha1(src.encode()).hexdigest()
    if key in _cache:
        return _cache[key]
    d = tempfile.mkdtemp(prefix="ka_host_")
    with open(os.path.join(d, "gen.cu"), "w") as f:
        f.write(src)
    with open(os.path.join(d, "harness.cpp"), "w") as f:
        f.write(HARNESS)
    so = os.path.join(d, "ka.so")
    # -O1 -ffp-contract=off: fmaf() stays a real fused multiply-add, a*b+c is not contracted behind our back
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-pthread", "-shared", "-fPIC", "-o", so,
                    os.path.join(d, "harness.cpp")], check=True, cwd=d)
    lib = ctypes.CDLL(so)
    lib.ka_run.argtypes = [ctypes.POINTER(KAParams), ctypes.POINTER(ctypes.c_float)]
    _cache[key] = lib
    return lib


def to_dtype_bytes(x, dtype):
    """fp32 array -> (raw array in dtype code 0/1/2, the fp32 values it represents)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if dtype == 0:
        return x, x
    if dtype == 2:
        h = x.astype(np.float16)
        return h.view(np.uint16), h.astype(np.float32)
    u = x.view(np.uint32)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    return r, (r.astype(np.uint32) << 16).view(np.float32)


def from_dtype_bytes(raw, dtype):
    if dtype == 0:
        return raw.astype(np.float64)
    if dtype == 2:
        return raw.view(np.float16).astype(np.float64)
    return (raw.astype(np.uint32) << 16).view(np.float32).astype(np.float64)


def run_loss(tree, mode, dtype, vec, x, y, w_xent, w_tree, want_grad=True, batch_rows=None, grad_scale=1.0, hw=0):
    """hw > 0 (with vec = 0): x is given as rows (B, C) and handed to the class-major variant as (B / hw, C, hw)."""
    lib = host_kernel(tree, mode, dtype, vec)
    B, C = x.shape
    if hw:
        x = np.ascontiguousarray(x.reshape(B // hw, hw, C).transpose(0, 2, 1))
    raw, _ = to_dtype_bytes(x, dtype)
    y = np.ascontiguousarray(y, dtype=np.int64)
    grad = np.zeros_like(raw)
    p = KAParams(raw.ctypes.data, B, C, y.ctypes.data, None, None, grad.ctypes.data if want_grad else None,
                 w_xent, w_tree, 1.0 / (batch_rows or B), grad_scale, hw)
    loss = ctypes.c_float()
    lib.ka_run(ctypes.byref(p), ctypes.byref(loss))
    g = from_dtype_bytes(grad, dtype) if want_grad else None
    if hw and g is not None:
        g = np.ascontiguousarray(g.reshape(B // hw, C, hw).transpose(0, 2, 1)).reshape(B, C)
    return loss.value, g


def run_soft_forward(tree, dtype, vec, x, hw=0):
    lib = host_kernel(tree, 2, dtype, vec)
    B, C = x.shape
    if hw:
        x = np.ascontiguousarray(x.reshape(B // hw, hw, C).transpose(0, 2, 1))
    raw, _ = to_dtype_bytes(x, dtype)
    out = np.zeros(x.shape, dtype=np.float32)
    p = KAParams(raw.ctypes.data, B, C, None, out.ctypes.data, None, None, 0, 0, 1, 1, hw)
    lib.ka_run(ctypes.byref(p), None)
    if hw:
        out = np.ascontiguousarray(out.reshape(B // hw, C, hw).transpose(0, 2, 1)).reshape(B, C)
    return out
