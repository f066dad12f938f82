#!/usr/bin/env python
"""Generate hierarchy-specialised tree-loss kernels (csrc/tree_spec_gen.cu).

The table-driven kernel in tree.cu is instruction-issue bound for small hierarchies (~4000
instructions per row at C=10, profiles/r01_tree_kernel_v1_ncu.md).  For every shipped hierarchy with at
most MAX_LEAVES leaves this script emits a straight-line kernel in which the whole hierarchy is
compile-time structure: one thread per row, every node value in a register, no tables, no shared
memory, no loops, exponentials/logarithms in the log2 domain (bare ex2/lg2 MUFU ops).  tree.cu looks
the kernels up by a hash of the flattened hierarchy (nbdt_tree_desc) and falls back to the
table-driven kernel for any other tree.

Same math as tree.cu (reference nbdt/model.py:83-123,207-242; nbdt/loss.py:191-266).
Run by build.py; the generated file is committed for inspection.
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
NBDT = os.path.join(os.path.dirname(HERE), "nbdt")
MAX_LEAVES = 24

HIERARCHIES = [
    ("CIFAR10", "graph-induced.json"), ("CIFAR10", "graph-induced-ResNet10.json"),
    ("CIFAR10", "graph-induced-ResNet18.json"), ("CIFAR10", "graph-wordnet.json"),
    ("LookIntoPerson", "graph-induced-HRNet-w48-cls20.json"),
]


def flatten(path_graph, wnids):
    """Same numbering as nbdt/_b200/tables.py: leaf c -> c, inode i (sorted wnid order) -> C + i."""
    with open(path_graph) as f:
        data = json.load(f)
    succ = {n["id"]: [] for n in data["nodes"]}
    has_parent = set()
    for e in data["links"]:
        succ[e["source"]].append(e["target"])
        has_parent.add(e["target"])
    inodes = sorted(w for w in succ if succ[w])
    if wnids is None:  # datasets without a wnids file (Cityscapes): leaves in sorted order
        wnids = sorted(w for w in succ if not succ[w])
    C = len(wnids)
    cls = {w: i for i, w in enumerate(wnids)}
    idx = {w: C + i for i, w in enumerate(inodes)}
    children = [[idx[c] if c in idx else cls[c] for c in succ[w]] for w in inodes]
    root = [i for i, w in enumerate(inodes) if w not in has_parent]
    assert len(root) == 1
    return C, children, root[0]


def fnv1a(C, children, root):
    """Must match tree_hash() in tree.cu."""
    h = 0xCBF29CE484222325
    vals = [C, len(children), root]
    ptr = [0]
    for ch in children:
        ptr.append(ptr[-1] + len(ch))
    vals += ptr
    for ch in children:
        vals += ch
    for v in vals:
        for b in range(4):
            h ^= (v >> (8 * b)) & 0xFF
            h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def emit_kernel(name, C, children, root):
    N = len(children)
    NV = C + N
    parent = [-1] * NV
    for i, ch in enumerate(children):
        for c in ch:
            parent[c] = C + i
    cnt = [1] * C + [0] * N
    depth = [0] * NV
    order = [C + root]
    for v in order:
        if v >= C:
            for c in children[v - C]:
                depth[c] = depth[v] + 1
                order.append(c)
    for v in reversed(order):
        if parent[v] >= 0:
            cnt[parent[v]] += cnt[v]
    inner_bottom_up = [v for v in reversed(order) if v >= C]
    inner_top_down = [v for v in order if v >= C]

    L = []
    w = L.append
    w(f"// hierarchy {name}: {C} leaves, {N} inner nodes, max depth {max(depth[:C])}")
    w("template <typename T, int MODE>")
    w(f"__global__ void __launch_bounds__(128) {name}(const SpecParams p) {{")
    w("  const T* __restrict__ x = reinterpret_cast<const T*>(p.x);")
    w("  T* __restrict__ gout = reinterpret_cast<T*>(p.grad);")
    w("  float loss_acc = 0.f;")
    w("  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < p.B; row += (long long)gridDim.x * blockDim.x) {")
    w("    const T* xr = x + row * p.ld;")
    w(f"    float v[{C}];")
    w(f"    load_row<T, {C}>(xr, v);")
    for c in range(C):
        w(f"    const float x{c} = v[{c}] * kLog2e;")
    # plain CE statistics
    w("    float lse1 = 0.f, inv_s1 = 0.f, xy = 0.f; int y = 0;")
    for c in range(C):
        w(f"    float e{c} = 0.f;")
    w("    if (MODE != kSpecSoftFwd) {")
    w("      y = (int)p.targets[row];")
    w("      float m1 = x0;")
    for c in range(1, C):
        w(f"      m1 = fmaxf(m1, x{c});")
    for c in range(C):
        w(f"      e{c} = ex2(x{c} - m1);")
    w("      float s1 = " + " + ".join(f"e{c}" for c in range(C)) + ";")
    w("      lse1 = m1 + lg2(s1); inv_s1 = __fdividef(1.f, s1);")
    w("      xy = to_f<T>(xr[y]) * kLog2e;")
    w("    }")
    # up sweep
    def S(vv):
        return f"x{vv}" if vv < C else f"S{vv}"
    for vv in inner_bottom_up:
        w(f"    const float S{vv} = " + " + ".join(S(c) for c in children[vv - C]) + ";")
    # down sweep
    w(f"    const float LP{C + root} = 0.f;")
    for vv in inner_top_down:
        ch = children[vv - C]
        a = []
        for c in ch:
            if cnt[c] == 1:
                a.append(S(c))
            else:
                w(f"    const float a{c} = {S(c)} * {1.0 / cnt[c]!r}f;")
                a.append(f"a{c}")
        if len(ch) == 2:
            w(f"    const float lse{vv} = fmaxf({a[0]}, {a[1]}) + lg2(1.f + ex2(-fabsf({a[0]} - {a[1]})));")
        else:
            w(f"    float m{vv} = {a[0]};")
            for t in a[1:]:
                w(f"    m{vv} = fmaxf(m{vv}, {t});")
            w(f"    const float lse{vv} = m{vv} + lg2(" + " + ".join(f"ex2({t} - m{vv})" for t in a) + ");")
        for c, t in zip(ch, a):
            w(f"    const float lp{c} = {t} - lse{vv};")
            w(f"    const float LP{c} = LP{vv} + lp{c};")
    for c in range(C):
        w(f"    const float p{c} = ex2(LP{c});")
    w("    if (MODE == kSpecSoftFwd) {")
    w(f"      float o[{C}] = {{" + ", ".join(f"p{c}" for c in range(C)) + "};")
    w(f"      store_row<float, {C}>(p.class_probs + row * {C}, o);")
    w("      continue;")
    w("    }")
    # leaf stage
    for c in range(C):
        w(f"    float g{c};")
    w("    const float kx = p.w_xent * p.inv_rows, kt = p.w_tree * p.inv_rows;")
    w("    if (MODE == kSpecSoftLoss) {")
    for c in range(C):
        w(f"      const float q{c} = ex2(p{c} * kLog2e);")
    w("      const float s2 = " + " + ".join(f"q{c}" for c in range(C)) + ";")
    w("      const float inv_s2 = __fdividef(1.f, s2);")
    w("      const float py = " + "".join(f"y == {c} ? p{c} : " for c in range(C - 1)) + f"p{C - 1};")
    w("      loss_acc += kx * kLn2 * (lse1 - xy) + kt * (kLn2 * lg2(s2) - py);")
    w("      const float kg = kt * p.grad_scale;")
    for c in range(C):
        w(f"      g{c} = kg * (q{c} * inv_s2 - (y == {c} ? 1.f : 0.f)) * p{c};")
    w("    } else {")
    w("      const float LPy = " + "".join(f"y == {c} ? LP{c} : " for c in range(C - 1)) + f"LP{C - 1};")
    w("      loss_acc += kx * kLn2 * (lse1 - xy) - kt * kLn2 * LPy;")
    w("      const float kg = kt * p.grad_scale;")
    for c in range(C):
        w(f"      g{c} = (y == {c}) ? -kg : 0.f;")
    w("    }")
    w("    if (gout == nullptr) continue;")
    # up sweep of G
    def G(vv):
        return f"g{vv}" if vv < C else f"G{vv}"
    for vv in inner_bottom_up:
        w(f"    const float G{vv} = " + " + ".join(G(c) for c in children[vv - C]) + ";")
    # down sweep of A
    w(f"    const float A{C + root} = 0.f;")
    for vv in inner_top_down:
        ch = children[vv - C]
        for j, c in enumerate(ch):
            if len(ch) == 2 and j == 1:
                w(f"    const float pr{c} = 1.f - pr{ch[0]};")
            else:
                w(f"    const float pr{c} = ex2(lp{c});")
            scale = "" if cnt[c] == 1 else f" * {1.0 / cnt[c]!r}f"
            w(f"    const float A{c} = A{vv} + ({G(c)} - pr{c} * G{vv}){scale};")
    w("    const float kgx = kx * p.grad_scale;")
    w(f"    float o[{C}];")
    for c in range(C):
        w(f"    o[{c}] = A{c} + kgx * (e{c} * inv_s1 - (y == {c} ? 1.f : 0.f));")
    w(f"    store_row<T, {C}>(gout + row * p.ld, o);")
    w("  }")
    w("  block_loss_reduce(loss_acc, p.loss, MODE != kSpecSoftFwd);")
    w("}")
    return "\n".join(L)


PRELUDE = r'''// GENERATED by csrc/gen_tree_kernels.py -- do not edit.  Hierarchy-specialised straight-line tree-loss kernels.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdint>

#include "tree_spec.h"

namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2(float v) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float lg2(float v) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// One row per thread: the widest vector the row size allows (the launcher checks pointer / ld alignment).
template <typename T, int C>
__device__ __forceinline__ void load_row(const T* __restrict__ r, float (&v)[C]) {
  constexpr int kBytes = C * (int)sizeof(T);
  if constexpr (kBytes % 16 == 0) {
    const uint4* q = reinterpret_cast<const uint4*>(r);
#pragma unroll
    for (int i = 0; i < kBytes / 16; ++i) {
      const uint4 u = __ldg(q + i);
      const T* e = reinterpret_cast<const T*>(&u);
#pragma unroll
      for (int j = 0; j < 16 / (int)sizeof(T); ++j) v[i * (16 / (int)sizeof(T)) + j] = to_f<T>(e[j]);
    }
  } else if constexpr (kBytes % 8 == 0) {
    const uint2* q = reinterpret_cast<const uint2*>(r);
#pragma unroll
    for (int i = 0; i < kBytes / 8; ++i) {
      const uint2 u = __ldg(q + i);
      const T* e = reinterpret_cast<const T*>(&u);
#pragma unroll
      for (int j = 0; j < 8 / (int)sizeof(T); ++j) v[i * (8 / (int)sizeof(T)) + j] = to_f<T>(e[j]);
    }
  } else if constexpr (kBytes % 4 == 0 && sizeof(T) == 2) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(r);
#pragma unroll
    for (int i = 0; i < kBytes / 4; ++i) {
      const uint32_t u = __ldg(q + i);
      const T* e = reinterpret_cast<const T*>(&u);
      v[2 * i] = to_f<T>(e[0]);
      v[2 * i + 1] = to_f<T>(e[1]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < C; ++i) v[i] = to_f<T>(r[i]);
  }
}

template <typename T, int C>
__device__ __forceinline__ void store_row(T* __restrict__ r, const float (&v)[C]) {
  constexpr int kBytes = C * (int)sizeof(T);
  constexpr int kPer16 = 16 / (int)sizeof(T);
  if constexpr (kBytes % 16 == 0) {
#pragma unroll
    for (int i = 0; i < kBytes / 16; ++i) {
      uint4 u;
      T* e = reinterpret_cast<T*>(&u);
#pragma unroll
      for (int j = 0; j < kPer16; ++j) e[j] = from_f<T>(v[i * kPer16 + j]);
      reinterpret_cast<uint4*>(r)[i] = u;
    }
  } else if constexpr (kBytes % 8 == 0) {
#pragma unroll
    for (int i = 0; i < kBytes / 8; ++i) {
      uint2 u;
      T* e = reinterpret_cast<T*>(&u);
#pragma unroll
      for (int j = 0; j < kPer16 / 2; ++j) e[j] = from_f<T>(v[i * (kPer16 / 2) + j]);
      reinterpret_cast<uint2*>(r)[i] = u;
    }
  } else if constexpr (kBytes % 4 == 0 && sizeof(T) == 2) {
#pragma unroll
    for (int i = 0; i < kBytes / 4; ++i) {
      uint32_t u;
      T* e = reinterpret_cast<T*>(&u);
      e[0] = from_f<T>(v[2 * i]);
      e[1] = from_f<T>(v[2 * i + 1]);
      reinterpret_cast<uint32_t*>(r)[i] = u;
    }
  } else {
#pragma unroll
    for (int i = 0; i < C; ++i) r[i] = from_f<T>(v[i]);
  }
}

__device__ __forceinline__ void block_loss_reduce(float v, float* loss, bool enabled) {
  if (!enabled) return;
  __shared__ float red[4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(loss, red[0] + red[1] + red[2] + red[3]);
}

'''


def generate():
    seen = {}
    bodies, entries = [], []
    for dataset, fname in HIERARCHIES:
        pg = os.path.join(NBDT, "hierarchies", dataset, fname)
        pw = os.path.join(NBDT, "wnids", dataset + ".txt")
        wnids = [l.strip() for l in open(pw)] if os.path.exists(pw) else None
        C, children, root = flatten(pg, wnids)
        if C > MAX_LEAVES:
            continue
        h = fnv1a(C, children, root)
        if h in seen:
            continue
        name = f"tree_spec_{h:016x}"
        seen[h] = name
        bodies.append(emit_kernel(name, C, children, root))
        entries.append((h, C, name, f"{dataset}/{fname}"))
    out = [PRELUDE] + bodies + ["\n}  // namespace\n"]
    out.append("const SpecEntry kSpecTable[] = {")
    for h, C, name, src in entries:
        fns = ", ".join(
            "{" + ", ".join(f"(const void*){name}<{T}, {m}>" for T in ("float", "__nv_bfloat16", "__half")) + "}"
            for m in ("kSpecSoftLoss", "kSpecHardLoss", "kSpecSoftFwd"))
        out.append(f"    {{0x{h:016x}ull, {C}, {{{fns}}}}},  // {src}")
    out.append("};")
    out.append(f"const int kSpecCount = {len(entries)};\n")
    return "\n".join(out)


def main():
    path = os.path.join(HERE, "tree_spec_gen.cu")
    text = generate()
    if not os.path.exists(path) or open(path).read() != text:
        with open(path, "w") as f:
            f.write(text)
    return path


if __name__ == "__main__":
    print(main())
