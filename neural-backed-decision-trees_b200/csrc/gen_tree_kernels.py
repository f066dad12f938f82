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
    """Straight-line kernel for one hierarchy.

    Soft modes work in the PROBABILITY domain like the reference's traverse_tree (model.py:207-242: products of
    per-node softmax probabilities): a binary node costs two MUFU ops (ex2, rcp), leaves none, and
    exp(class_prob) for the second cross-entropy is a degree-6 polynomial on [0, 1] on the FMA pipe (the v1
    kernel was bound by the MUFU pipe at 70 %, profiles/r01_tree_spec_ncu.md).  The hard loss needs log path
    probabilities that must not underflow (per-node log-softmax, loss.py:212-257) and stays in the log2 domain.
    """
    N = len(children)
    NV = C + N
    parent = [-1] * NV
    for i, ch in enumerate(children):
        for c in ch:
            parent[c] = C + i
    cnt = [1] * C + [0] * N
    order = [C + root]
    for v in order:
        if v >= C:
            order.extend(children[v - C])
    for v in reversed(order):
        if parent[v] >= 0:
            cnt[parent[v]] += cnt[v]
    depth = [0] * NV
    for v in order:
        if parent[v] >= 0:
            depth[v] = depth[parent[v]] + 1
    inner_bottom_up = [v for v in reversed(order) if v >= C]
    inner_top_down = [v for v in order if v >= C]
    rootv = C + root

    L = []
    w = L.append

    def S(vv):
        return f"x{vv}" if vv < C else f"S{vv}"

    def G(vv):
        return f"g{vv}" if vv < C else f"G{vv}"

    def k2(c):  # log2(e) / #leaves: node logit of child c in log2 units = S(c) * k2(c)
        return f"{1.4426950408889634 / cnt[c]!r}f"

    def onehot(c):
        return f"oh{c}"

    w(f"// hierarchy {name}: {C} leaves, {N} inner nodes, max depth {max(depth[:C])}")
    w("template <typename T, int MODE>")
    # registers: ~8 per leaf live at the widest point; cap at 80 (6 CTAs/SM) for <= 12 leaves, 128 (4 CTAs/SM) above
    w(f"__global__ void __launch_bounds__(128, {6 if C <= 12 else 4}) {name}(const SpecParams p) {{")
    w("  const T* __restrict__ x = reinterpret_cast<const T*>(p.x);")
    w("  T* __restrict__ gout = reinterpret_cast<T*>(p.grad);")
    w("  float loss_acc = 0.f;")
    w("  const float kx = p.w_xent * p.inv_rows, kt = p.w_tree * p.inv_rows;")
    w("  const float kg = kt * p.grad_scale, kgx = kx * p.grad_scale;")
    w("  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < p.B; row += (long long)gridDim.x * blockDim.x) {")
    w("    const T* xr = x + row * p.ld;")
    w(f"    float v[{C}];")
    w(f"    load_row<T, {C}>(xr, v);")
    for c in range(C):
        w(f"    const float x{c} = v[{c}];")
    # plain cross-entropy statistics of the raw logits
    w("    float lse1 = 0.f, inv_s1 = 0.f, xy = 0.f; int y = 0; bool ign = false;  // ign: target out of range -> zero loss, zero gradient")
    for c in range(C):
        w(f"    float e{c} = 0.f;")
    w("    if (MODE != kSpecSoftFwd) {")
    w("      { const long long yt = p.targets[row]; ign = (unsigned long long)yt >= %dull; y = ign ? 0 : (int)yt; }" % C)
    w("      float m1 = x0;")
    for c in range(1, C):
        w(f"      m1 = fmaxf(m1, x{c});")
    w("      const float nm = -m1 * kLog2e;")
    for c in range(C):
        w(f"      e{c} = ex2(fmaf(x{c}, kLog2e, nm));")
    w("      const float s1 = " + " + ".join(f"e{c}" for c in range(C)) + ";")
    w("      lse1 = fmaf(lg2(s1), kLn2, m1); inv_s1 = rcp(s1);")
    w("      xy = to_f<T>(xr[y]);")
    w("    }")
    for c in range(C):
        w(f"    const float oh{c} = (y == {c}) ? 1.f : 0.f;")
    # up sweep: subtree sums
    for vv in inner_bottom_up:
        w(f"    const float S{vv} = " + " + ".join(S(c) for c in children[vv - C]) + ";")
    # down sweep.  pr{c}: probability of child c at its parent; soft: P{c} path probability; hard: LP{c} log2 path prob.
    for c in range(C):
        w(f"    float g{c};")
    for vv in range(NV):
        if vv != rootv:
            w(f"    float pr{vv};")
    w("    if (MODE != kSpecHardLoss) {")
    for vv in inner_top_down:
        ch = children[vv - C]
        if len(ch) == 2:
            c0, c1 = ch
            w(f"      {{ const float d = fmaf({S(c0)}, {k2(c0)}, -({S(c1)} * {k2(c1)}));")
            w("        const float t = ex2(-fabsf(d)); const float r = rcp(1.f + t); const float lo = t * r;")
            w(f"        pr{c0} = d >= 0.f ? r : lo; pr{c1} = d >= 0.f ? lo : r; }}")
        else:
            w("      {")
            for c in ch:
                w(f"        const float a{c} = {S(c)} * {k2(c)};")
            w(f"        float m = a{ch[0]};")
            for c in ch[1:]:
                w(f"        m = fmaxf(m, a{c});")
            for c in ch:
                w(f"        const float q{c} = ex2(a{c} - m);")
            w("        const float r = rcp(" + " + ".join(f"q{c}" for c in ch) + ");")
            for c in ch:
                w(f"        pr{c} = q{c} * r;")
            w("      }")
    for vv in inner_top_down:
        for c in children[vv - C]:
            if vv == rootv:
                w(f"      const float P{c} = pr{c};")
            else:
                w(f"      const float P{c} = P{vv} * pr{c};")
    w("      if (MODE == kSpecSoftFwd) {")
    w(f"        float o[{C}] = {{" + ", ".join(f"P{c}" for c in range(C)) + "};")
    w(f"        store_row<float, {C}>(p.class_probs + row * {C}, o);")
    w("        continue;")
    w("      }")
    for c in range(C):
        w(f"      const float q{c} = exp01(P{c});")
    w("      const float s2 = " + " + ".join(f"q{c}" for c in range(C)) + ";")
    w("      const float inv_s2 = rcp(s2);")
    w("      float py = oh0 * P0;")
    for c in range(1, C):
        w(f"      py = fmaf(oh{c}, P{c}, py);")
    w("      if (!ign) loss_acc += kx * (lse1 - xy) + kt * (kLn2 * lg2(s2) - py);")
    for c in range(C):
        w(f"      g{c} = kg * (q{c} * inv_s2 - {onehot(c)}) * P{c};")
    w("    } else {")
    for vv in inner_top_down:
        ch = children[vv - C]
        if len(ch) == 2:
            c0, c1 = ch
            w(f"      float lp{c0}, lp{c1};")
            w(f"      {{ const float d = fmaf({S(c0)}, {k2(c0)}, -({S(c1)} * {k2(c1)}));")
            w("        const float t = ex2(-fabsf(d)); const float l = -lg2(1.f + t); const float ll = l - fabsf(d);")
            w(f"        lp{c0} = d >= 0.f ? l : ll; lp{c1} = d >= 0.f ? ll : l;")
            w("        const float r = rcp(1.f + t); const float lo = t * r;")
            w(f"        pr{c0} = d >= 0.f ? r : lo; pr{c1} = d >= 0.f ? lo : r; }}")
        else:
            for c in ch:
                w(f"      float lp{c};")
            w("      {")
            for c in ch:
                w(f"        const float a{c} = {S(c)} * {k2(c)};")
            w(f"        float m = a{ch[0]};")
            for c in ch[1:]:
                w(f"        m = fmaxf(m, a{c});")
            for c in ch:
                w(f"        const float q{c} = ex2(a{c} - m);")
            w("        const float s = " + " + ".join(f"q{c}" for c in ch) + ";")
            w("        const float lse = m + lg2(s); const float r = rcp(s);")
            for c in ch:
                w(f"        lp{c} = a{c} - lse; pr{c} = q{c} * r;")
            w("      }")
    for vv in inner_top_down:
        for c in children[vv - C]:
            if vv == rootv:
                w(f"      const float LP{c} = lp{c};")
            else:
                w(f"      const float LP{c} = LP{vv} + lp{c};")
    w("      float LPy = oh0 * LP0;")
    for c in range(1, C):
        w(f"      LPy = fmaf(oh{c}, LP{c}, LPy);")
    w("      if (!ign) loss_acc += kx * (lse1 - xy) - kt * kLn2 * LPy;")
    for c in range(C):
        w(f"      g{c} = -kg * oh{c};")
    w("    }")
    w("    if (gout == nullptr) continue;")
    # up sweep of the leaf gradients, down sweep of the accumulated logit gradient
    for vv in inner_bottom_up:
        w(f"    const float G{vv} = " + " + ".join(G(c) for c in children[vv - C]) + ";")
    for vv in inner_top_down:
        for c in children[vv - C]:
            term = f"fmaf(-pr{c}, G{vv}, {G(c)})"
            if vv == rootv:
                w(f"    const float A{c} = {term}" + ("" if cnt[c] == 1 else f" * {1.0 / cnt[c]!r}f") + ";")
            elif cnt[c] == 1:
                w(f"    const float A{c} = A{vv} + {term};")
            else:
                w(f"    const float A{c} = fmaf({term}, {1.0 / cnt[c]!r}f, A{vv});")
    w(f"    float o[{C}];")
    for c in range(C):
        w(f"    o[{c}] = ign ? 0.f : fmaf(kgx, fmaf(e{c}, inv_s1, -{onehot(c)}), A{c});")
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
__device__ __forceinline__ float rcp(float v) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
// exp(t) for t in [0, 1] (a class probability): degree-6 Chebyshev-node interpolant, max rel. error 1.5e-7 --
// six FMA-pipe instructions instead of a MUFU op.
__device__ __forceinline__ float exp01(float t) {
  float r = 0.002307842718437314f;
  r = fmaf(r, t, 0.0069595566019415855f);
  r = fmaf(r, t, 0.04264117032289505f);
  r = fmaf(r, t, 0.16631804406642914f);
  r = fmaf(r, t, 0.5000588297843933f);
  r = fmaf(r, t, 0.9999963045120239f);
  return fmaf(r, t, 1.0f);
}

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
