// Fused tree-head / tree-supervision-loss kernel for ARBITRARY hierarchies, second generation
// (soft forward, soft loss fwd+bwd, hard loss fwd+bwd).  tree.cu keeps the first-generation kernel for the modes
// that are not bandwidth-critical (decision paths, per-node outputs, generic-criterion backward) and for
// hierarchies above kMaxLeaves leaves.
//
// Why a second kernel: the first one makes four level-synchronous sweeps over table-driven nodes (two bottom-up for
// the subtree sums S[v] and G[v]); profiles/r01_tree_sweep_v3.jsonl has it at 2-4 % of HBM bandwidth for 100..1000
// leaves (issue-bound, ~800 instructions per leaf).  Here:
//
//   * Leaves are renumbered in DFS order on the host, so the leaves under ANY node are a contiguous range
//     [beg, end) and every subtree sum is a difference of two prefix sums: S[v] = PS[end] - PS[beg].  The two
//     bottom-up sweeps disappear; a prefix scan over C leaves replaces each (fp64 for the logits: a difference of
//     fp32 prefix sums over 1000 logits would cost 1e-5 of relative accuracy in the node logits; fp32 for the
//     gradient sums, which only need accuracy relative to their own magnitude).
//   * Inner nodes are numbered in BFS order and their child edges consecutively, so one top-down pass per direction
//     does everything for a node: child softmax -> child probabilities pr[e] and path probabilities P[e]
//     (forward), A[e] = A[parent] + (G[e] - pr[e] G[node]) / #leaves(e) (backward).  Tables are read coalesced.
//   * Soft modes work in the probability domain like the reference (model.py:207-242) -- a binary node costs
//     ex2 + rcp -- and exp(class_prob) is a polynomial on the FMA pipe (see tree_spec_gen.cu).  The hard loss keeps
//     log2 path probabilities (per-node log-softmax, loss.py:212-257).
//   * G lanes cooperate on one row (G = 4..32 by leaf count; 32/G rows per warp run in lock step on the same
//     tree, so control flow is warp-uniform and only __syncwarp separates the levels).
//
// Per-row shared memory: PS[C+1] fp64 (later the fp32 gradient prefix sums), P[S] and pr[S] fp32 (S = C + N - 1
// edges): 4 KB per row at 100 leaves, 24 KB at 1000.  HBM traffic = algorithmic bytes: logits + targets read once,
// gradient (or class probabilities) written once.
//
// Reference semantics: nbdt/model.py:83-123,207-242; nbdt/loss.py:191-266 (same closed form as tree.cu).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <new>
#include <vector>

#include "common.h"
#include "tree_spec.h"
#include "tree_v2.h"

struct TreeV2 {
  int device;
  int C, N, S, D;          // leaves, inner nodes, edges, levels of inner nodes
  int sm_count, max_smem;
  void* arena;
  const int4* node;        // [N] BFS order: {first edge, parent edge (-1: root), #children, unused}
  const unsigned* be;      // [S] per edge: beg | end << 16  (DFS leaf positions of the child's subtree)
  const float* inv;        // [S] 1 / #leaves under the child
  const int* level_ptr;    // [D+1] into the BFS EDGE order (edges of depth d+1 are consecutive)
  const int* eparent;      // [S] edge that ends in the edge's parent node (-1: child of the root)
  const unsigned* pbe;     // [S] beg | end << 16 of the edge's PARENT node
  const int* class_pos;    // [C] DFS position of class c
  const int* class_edge;   // [C] edge that ends in the leaf of class c
  const int* pos_edge;     // [C] edge that ends in the leaf at DFS position i
};

namespace {

// Measured (profiles/r01_tree_sweep_v6.jsonl): 15-25 % faster than the first-generation kernel at 100-200 leaves, 20 % slower
// at 1000 (one row per warp, 24 KB of shared memory per row): used up to 512 leaves.
constexpr int kMaxLeaves = 512;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct Tab {
  const int4* node;
  const unsigned* be;
  const float* inv;
  const int* level_ptr;
  const int* eparent;
  const unsigned* pbe;
  const int* class_pos;
  const int* class_edge;
  const int* pos_edge;
  int C, N, S, D;
};

struct V2Params {
  Tab t;
  SpecParams p;
  int row_floats;  // shared-memory floats per row (odd, so that the rows of a warp start in different banks)
};

__device__ __forceinline__ float ex2(float v) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float lg2(float v) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float rcp(float v) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
// exp(t), t in [0, 1]: degree-6 interpolant, max rel. error 1.5e-7 (same polynomial as tree_spec_gen.cu)
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

template <int G> __device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, G);
  return v;
}
template <int G> __device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o, G));
  return v;
}
// exclusive scan over the G lanes of a group
template <int G, typename V> __device__ __forceinline__ V group_exclusive_scan(V v, int j) {
  V incl = v;
#pragma unroll
  for (int o = 1; o < G; o <<= 1) {
    const V up = __shfl_up_sync(0xffffffffu, incl, o, G);
    if (j >= o) incl += up;
  }
  return incl - v;
}

// G lanes per row, EPT = ceil(C / G) elements per lane (compile-time bound), MODE = kSpec{SoftLoss,HardLoss,SoftFwd}.
template <int G, int EPT, int MODE, typename T>
__global__ void __launch_bounds__(256, 3) tree_v2_kernel(const V2Params q) {
  // Rows are addressed as integer offsets into the __shared__ array (not through per-row generic pointers): the
  // compiler then emits plain LDS/STS with 32-bit offsets instead of re-deriving the shared window base for every
  // predicated access (profiles/r01_tree_v2_ncu.md).
  extern __shared__ __align__(16) float smf[];
  double* const smd = reinterpret_cast<double*>(smf);
  const Tab& t = q.t;
  const SpecParams& p = q.p;
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane / G, j = lane % G;
  const int C = t.C, S = t.S;
  // row layout: [PS: (C+1) doubles][P: S floats][pr: S floats]  (row_floats includes the doubles as 2 floats each)
  const int ro = (warp * RPW + r) * q.row_floats;  // even: the fp64 prefix sums stay 8-byte aligned
  const int oP = ro + 2 * (C + 1), oR = oP + S;
#define PS_(i) smd[(ro >> 1) + (i)]   /* fp64 prefix sums of the DFS-ordered logits */
#define PG_(i) smf[ro + (i)]          /* fp32 gradient prefix sums (reuse the PS storage) */
#define P_(i) smf[oP + (i)]           /* staging for the logits, then path (log-)probabilities, then A */
#define PR_(i) smf[oR + (i)]          /* child probabilities, then the per-edge gradient terms */
  const T* __restrict__ x = reinterpret_cast<const T*>(p.x);
  T* __restrict__ gout = reinterpret_cast<T*>(p.grad);
  const int per = (C + G - 1) / G;        // positions owned by one lane in the scans: [j*per, min(C, (j+1)*per))
  const float kx = p.w_xent * p.inv_rows, kt = p.w_tree * p.inv_rows;
  const float kg = kt * p.grad_scale, kgx = kx * p.grad_scale;

  const long long chunks = (p.B + RPW - 1) / RPW;
  const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
  float loss_acc = 0.f;
  for (long long chunk = (long long)blockIdx.x * (blockDim.x >> 5) + warp; chunk < chunks; chunk += warps_total) {
    const long long row = chunk * RPW + r;
    const bool valid = row < p.B;
    const T* xr = x + (valid ? row : 0) * p.ld;

    // ---- 1. load the row (class c = j + k*G), stage it in DFS order, plain cross-entropy statistics ----
    float ev[EPT];
    float m1 = -INFINITY;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
      const int c = j + k * G;
      float v = 0.f;
      if (c < C) {
        if (valid) v = to_f<T>(xr[c]);
        P_(__ldg(t.class_pos + c)) = v;
        m1 = fmaxf(m1, v);
      }
      ev[k] = v;
    }
    float lse1 = 0.f, inv_s1 = 0.f, xy = 0.f;
    int y = 0, pos_y = 0;
    bool ign = false;  // target outside [0, C): the row is ignored (zero loss, zero gradient), like nn.CrossEntropyLoss's ignore_index
    if (MODE != kSpecSoftFwd) {
      m1 = group_max<G>(m1);
      const float nm = -m1 * kLog2e;
      float s1 = 0.f;
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        const int c = j + k * G;
        ev[k] = c < C ? ex2(fmaf(ev[k], kLog2e, nm)) : 0.f;
        s1 += ev[k];
      }
      s1 = group_sum<G>(s1);
      lse1 = fmaf(lg2(s1), kLn2, m1);
      inv_s1 = rcp(s1);
      if (valid) {
        const long long yy = p.targets[row];
        ign = (unsigned long long)yy >= (unsigned long long)C;
        y = ign ? 0 : (int)yy;
        xy = to_f<T>(xr[y]);
      }
      pos_y = __ldg(t.class_pos + y);
    }
    __syncwarp();

    // ---- 2. fp64 prefix sums over the DFS-ordered logits: PS_(i) = sum of positions < i ----
    {
      // two passes over the lane's positions (sum, then running prefix) instead of EPT fp64 registers
      double run = 0.0;
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        const int i = j * per + k;
        if (k < per && i < C) run += (double)P_(i);
      }
      double acc = group_exclusive_scan<G, double>(run, j);
      if (j == 0) PS_(0) = 0.0;
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        const int i = j * per + k;
        if (k < per && i < C) {
          acc += (double)P_(i);
          PS_(i + 1) = acc;
        }
      }
    }
    __syncwarp();

    // ---- 3a. child softmax of every inner node (no dependency between nodes: no synchronisation, the loop
    //          iterations overlap).  PR_(e) = child probability; hard loss: P_(e) = log2 child probability ----
#pragma unroll 2
    for (int n = j; n < t.N; n += G) {
      const int4 nd = __ldg(t.node + n);
      const int e0 = nd.x, K = nd.z;
      if (K == 2) {
        const unsigned be0 = __ldg(t.be + e0), be1 = __ldg(t.be + e0 + 1);
        const double p0 = PS_(be0 & 0xffffu), p1 = PS_(be0 >> 16), p2 = PS_(be1 >> 16);
        const float a0 = (float)(p1 - p0) * (__ldg(t.inv + e0) * kLog2e);
        const float a1 = (float)(p2 - p1) * (__ldg(t.inv + e0 + 1) * kLog2e);
        const float dlt = a0 - a1;
        const float tt = ex2(-fabsf(dlt));
        const float rr = rcp(1.f + tt), lo = tt * rr;
        const bool first = dlt >= 0.f;
        PR_(e0) = first ? rr : lo;
        PR_(e0 + 1) = first ? lo : rr;
        if (MODE == kSpecHardLoss) {
          const float l = -lg2(1.f + tt), ll = l - fabsf(dlt);
          P_(e0) = first ? l : ll;
          P_(e0 + 1) = first ? ll : l;
        }
      } else {
        float mx = -INFINITY;
        for (int e = e0; e < e0 + K; ++e) {
          const unsigned b = __ldg(t.be + e);
          const float a = (float)(PS_(b >> 16) - PS_(b & 0xffffu)) * (__ldg(t.inv + e) * kLog2e);
          PR_(e) = a;
          mx = fmaxf(mx, a);
        }
        float sum = 0.f;
        for (int e = e0; e < e0 + K; ++e) {
          const float a = PR_(e) - mx;
          const float ee = ex2(a);
          if (MODE == kSpecHardLoss) P_(e) = a;  // completed with -lg2(sum) below
          PR_(e) = ee;
          sum += ee;
        }
        const float rs = rcp(sum), ls = lg2(sum);
        for (int e = e0; e < e0 + K; ++e) {
          PR_(e) *= rs;
          if (MODE == kSpecHardLoss) P_(e) -= ls;
        }
      }
    }
    __syncwarp();
    // ---- 3b. path (log-)probabilities, one level of edges at a time: P_(e) = P_(parent edge) (*|+) child value ----
    for (int d = 0; d < t.D; ++d) {
      const int e1 = __ldg(t.level_ptr + d + 1);
      for (int e = __ldg(t.level_ptr + d) + j; e < e1; e += G) {
        const int ep = __ldg(t.eparent + e);
        if (MODE == kSpecHardLoss) P_(e) = (ep < 0 ? 0.f : P_(ep)) + P_(e);
        else P_(e) = (ep < 0 ? 1.f : P_(ep)) * PR_(e);
      }
      __syncwarp();
    }

    if (MODE == kSpecSoftFwd) {
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        const int c = j + k * G;
        if (c < C && valid) p.class_probs[row * C + c] = P_(__ldg(t.class_edge + c));
      }
      __syncwarp();
      continue;
    }

    // ---- 4. leaf stage: loss terms, dL/dLP[leaf], fp32 prefix sums of those over the DFS positions ----
    float row_loss;
    if (MODE == kSpecSoftLoss) {
      float pl[EPT];
      float s2 = 0.f, py = 0.f;
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        const int i = j * per + k;
        float pv = 0.f, qv = 0.f;
        if (k < per && i < C) {
          pv = P_(__ldg(t.pos_edge + i));
          qv = exp01(pv);
          if (i == pos_y) py = pv;
        }
        pl[k] = pv;
        s2 += qv;
      }
      s2 = group_sum<G>(s2);
      py = group_sum<G>(py);
      const float inv_s2 = rcp(s2);
      row_loss = kx * (lse1 - xy) + kt * (kLn2 * lg2(s2) - py);
      float run = 0.f;
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        const int i = j * per + k;
        if (k < per && i < C) run += kg * (exp01(pl[k]) * inv_s2 - (i == pos_y ? 1.f : 0.f)) * pl[k];
        pl[k] = run;
      }
      const float off = group_exclusive_scan<G, float>(run, j);
      __syncwarp();  // PS (fp64) is dead from here on: its storage becomes PG
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        const int i = j * per + k;
        if (k < per && i < C) PG_(i + 1) = pl[k] + off;
      }
      if (j == 0) PG_(0) = 0.f;
    } else {  // hard loss: tree term = -log(path probability of the target's leaf)
      const float lpy = P_(__ldg(t.class_edge + y));
      row_loss = kx * (lse1 - xy) - kt * kLn2 * lpy;
    }
    if (j == 0 && valid && !ign) loss_acc += row_loss;
    if (gout == nullptr) {
      __syncwarp();
      continue;
    }
    __syncwarp();

    // ---- 5a. per-edge gradient term t[e] = (G[e] - PR_(e) * G[parent node]) / #leaves(e), G = range sums of the leaf
    //          gradients (prefix-sum differences); independent per edge, overwrites PR_(e) ----
#pragma unroll 2
    for (int e = j; e < S; e += G) {
      const unsigned b = __ldg(t.be + e), pb = __ldg(t.pbe + e);
      float Ge, Gn;
      if (MODE == kSpecSoftLoss) {
        Ge = PG_(b >> 16) - PG_(b & 0xffffu);
        Gn = PG_(pb >> 16) - PG_(pb & 0xffffu);
      } else {
        Ge = ((int)(b & 0xffffu) <= pos_y && pos_y < (int)(b >> 16)) ? -kg : 0.f;
        Gn = ((int)(pb & 0xffffu) <= pos_y && pos_y < (int)(pb >> 16)) ? -kg : 0.f;
      }
      PR_(e) = fmaf(-PR_(e), Gn, Ge) * __ldg(t.inv + e);
    }
    __syncwarp();
    // ---- 5b. A[e] = A[parent edge] + t[e], one level of edges at a time (A overwrites P) ----
    for (int d = 0; d < t.D; ++d) {
      const int e1 = __ldg(t.level_ptr + d + 1);
      for (int e = __ldg(t.level_ptr + d) + j; e < e1; e += G) {
        const int ep = __ldg(t.eparent + e);
        P_(e) = (ep < 0 ? 0.f : P_(ep)) + PR_(e);
      }
      __syncwarp();
    }

    // ---- 6. dL/dx[c] = A[leaf c] + w_x/B (softmax(x) - onehot(y)) ----
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
      const int c = j + k * G;
      if (c < C && valid)
        gout[row * p.ld + c] = from_f<T>(ign ? 0.f : fmaf(kgx, fmaf(ev[k], inv_s1, c == y ? -1.f : 0.f), P_(__ldg(t.class_edge + c))));
    }
    __syncwarp();
  }

  if (MODE != kSpecSoftFwd) {
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
    if (lane == 0) red[warp] = loss_acc;
    __syncthreads();
    if (warp == 0) {
      float v = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) atomicAdd(p.loss, v);
    }
  }
}
#undef PS_
#undef PG_
#undef P_
#undef PR_

template <int G, int EPT, int MODE, typename T>
int launch_cfg(const TreeV2* t, const V2Params& q, cudaStream_t stream) {
  constexpr int RPW = 32 / G;
  const size_t row_bytes = (size_t)q.row_floats * sizeof(float);
  const size_t per_warp = row_bytes * RPW;
  const size_t budget = (size_t)t->max_smem - 2048;
  if (per_warp > budget) return -1;
  int wpc = (int)std::min<size_t>(8, budget / per_warp);
  // prefer several CTAs per SM: shrink the CTA until at least two fit (or one warp is left)
  while (wpc > 2 && per_warp * wpc * 2 > budget) wpc >>= 1;
  const size_t smem = per_warp * wpc;
  auto kern = tree_v2_kernel<G, EPT, MODE, T>;
  NBDT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int ctas_per_sm = 1;
  NBDT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, wpc * 32, smem));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  const long long chunks = (q.p.B + RPW - 1) / RPW;
  long long grid = std::min<long long>((chunks + wpc - 1) / wpc, (long long)t->sm_count * ctas_per_sm);
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, wpc * 32, smem, stream>>>(q);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

template <int MODE, typename T>
int launch_mode(const TreeV2* t, const V2Params& q, cudaStream_t stream) {
  const int C = t->C;
  if (C <= 32) return launch_cfg<4, 8, MODE, T>(t, q, stream);
  if (C <= 64) return launch_cfg<4, 16, MODE, T>(t, q, stream);
  if (C <= 128) return launch_cfg<8, 16, MODE, T>(t, q, stream);
  if (C <= 256) return launch_cfg<16, 16, MODE, T>(t, q, stream);
  return launch_cfg<32, 16, MODE, T>(t, q, stream);
}

template <typename T>
int launch_dtype(const TreeV2* t, int mode, const V2Params& q, cudaStream_t stream) {
  switch (mode) {
    case kSpecSoftLoss: return launch_mode<kSpecSoftLoss, T>(t, q, stream);
    case kSpecHardLoss: return launch_mode<kSpecHardLoss, T>(t, q, stream);
    case kSpecSoftFwd: return launch_mode<kSpecSoftFwd, T>(t, q, stream);
    default: return -1;
  }
}

}  // namespace

int tree_v2_create(const nbdt_tree_desc* d, int device, int sm_count, int max_smem, TreeV2** out) {
  *out = nullptr;
  const int C = d->num_classes, N = d->num_inodes;
  if (C > kMaxLeaves || getenv("NBDT_B200_TREE_V1")) return NBDT_OK;  // first-generation kernel only
  const int S = C + N - 1;
  // DFS over the hierarchy (children in table order): leaf positions and [beg, end) of every node
  std::vector<int> beg(C + N, 0), end(C + N, 0), class_pos(C, 0);
  {
    int next = 0;
    std::vector<std::pair<int, int>> stack;  // (node, next child slot)
    stack.push_back({C + d->root, d->child_ptr[d->root]});
    beg[C + d->root] = 0;
    while (!stack.empty()) {
      auto& top = stack.back();
      const int v = top.first, i = v - C;
      if (top.second == d->child_ptr[i + 1]) {
        end[v] = next;
        stack.pop_back();
        continue;
      }
      const int ch = d->child_node[top.second++];
      beg[ch] = next;
      if (ch < C) {
        class_pos[ch] = next++;
        end[ch] = next;
      } else {
        stack.push_back({ch, d->child_ptr[ch - C]});
      }
    }
  }
  // BFS numbering of the inner nodes, consecutive edges per node; BFS order makes the edges of one depth consecutive
  std::vector<int> bfs, depth(N, 0), parent_edge(N, -1);
  bfs.push_back(d->root);
  std::vector<int4> node;
  std::vector<unsigned> be, pbe;
  std::vector<float> inv;
  std::vector<int> class_edge(C, 0), pos_edge(C, 0), eparent, edepth;
  for (size_t qi = 0; qi < bfs.size(); ++qi) {
    const int i = bfs[qi];
    const int e0 = (int)be.size();
    const int K = d->child_ptr[i + 1] - d->child_ptr[i];
    node.push_back(make_int4(e0, parent_edge[i], K, 0));
    for (int k = d->child_ptr[i]; k < d->child_ptr[i + 1]; ++k) {
      const int ch = d->child_node[k];
      const int e = (int)be.size();
      be.push_back((unsigned)beg[ch] | ((unsigned)end[ch] << 16));
      pbe.push_back((unsigned)beg[C + i] | ((unsigned)end[C + i] << 16));
      inv.push_back(1.0f / (float)(end[ch] - beg[ch]));
      eparent.push_back(parent_edge[i]);
      edepth.push_back(depth[i]);
      if (ch < C) {
        class_edge[ch] = e;
        pos_edge[class_pos[ch]] = e;
      } else {
        parent_edge[ch - C] = e;
        depth[ch - C] = depth[i] + 1;
        bfs.push_back(ch - C);
      }
    }
  }
  if ((int)bfs.size() != N || (int)be.size() != S) return NBDT_OK;  // malformed: tree.cu has already validated; be safe
  std::vector<int> level_ptr;  // edges grouped by the depth of their parent node (non-decreasing in BFS order)
  for (int e = 0; e < S; ++e)
    if (e == 0 || edepth[e] != edepth[e - 1]) level_ptr.push_back(e);
  const int D = (int)level_ptr.size();
  level_ptr.push_back(S);

  TreeV2* t = new (std::nothrow) TreeV2();
  NBDT_REQUIRE(t != nullptr, "out of host memory");
  t->device = device; t->C = C; t->N = N; t->S = S; t->D = D; t->sm_count = sm_count; t->max_smem = max_smem;
  t->arena = nullptr;
  // arena layout (16-byte aligned pieces)
  auto align16 = [](size_t v) { return (v + 15) & ~size_t(15); };
  const size_t o_node = 0;
  const size_t o_be = align16(o_node + node.size() * sizeof(int4));
  const size_t o_inv = align16(o_be + be.size() * 4);
  const size_t o_lvl = align16(o_inv + inv.size() * 4);
  const size_t o_epar = align16(o_lvl + level_ptr.size() * 4);
  const size_t o_pbe = align16(o_epar + eparent.size() * 4);
  const size_t o_cpos = align16(o_pbe + pbe.size() * 4);
  const size_t o_cedge = align16(o_cpos + class_pos.size() * 4);
  const size_t o_pedge = align16(o_cedge + class_edge.size() * 4);
  const size_t bytes = align16(o_pedge + pos_edge.size() * 4);
  std::vector<unsigned char> host(bytes, 0);
  memcpy(host.data() + o_node, node.data(), node.size() * sizeof(int4));
  memcpy(host.data() + o_be, be.data(), be.size() * 4);
  memcpy(host.data() + o_inv, inv.data(), inv.size() * 4);
  memcpy(host.data() + o_lvl, level_ptr.data(), level_ptr.size() * 4);
  memcpy(host.data() + o_epar, eparent.data(), eparent.size() * 4);
  memcpy(host.data() + o_pbe, pbe.data(), pbe.size() * 4);
  memcpy(host.data() + o_cpos, class_pos.data(), class_pos.size() * 4);
  memcpy(host.data() + o_cedge, class_edge.data(), class_edge.size() * 4);
  memcpy(host.data() + o_pedge, pos_edge.data(), pos_edge.size() * 4);
  int cur = -1;
  cudaGetDevice(&cur);
  cudaError_t ce = cudaSetDevice(device);
  if (ce == cudaSuccess) ce = cudaMalloc(&t->arena, bytes);
  if (ce == cudaSuccess) ce = cudaMemcpy(t->arena, host.data(), bytes, cudaMemcpyHostToDevice);
  if (cur >= 0) cudaSetDevice(cur);
  if (ce != cudaSuccess) {
    if (t->arena) cudaFree(t->arena);
    delete t;
    cudaGetLastError();
    return nbdt::set_error(NBDT_ECUDA, "uploading tree tables (v2) to device %d failed: %s", device, cudaGetErrorString(ce));
  }
  const unsigned char* base = reinterpret_cast<const unsigned char*>(t->arena);
  t->node = reinterpret_cast<const int4*>(base + o_node);
  t->be = reinterpret_cast<const unsigned*>(base + o_be);
  t->inv = reinterpret_cast<const float*>(base + o_inv);
  t->level_ptr = reinterpret_cast<const int*>(base + o_lvl);
  t->eparent = reinterpret_cast<const int*>(base + o_epar);
  t->pbe = reinterpret_cast<const unsigned*>(base + o_pbe);
  t->class_pos = reinterpret_cast<const int*>(base + o_cpos);
  t->class_edge = reinterpret_cast<const int*>(base + o_cedge);
  t->pos_edge = reinterpret_cast<const int*>(base + o_pedge);
  *out = t;
  return NBDT_OK;
}

void tree_v2_destroy(TreeV2* t) {
  if (!t) return;
  if (t->arena) {
    int cur = -1;
    cudaGetDevice(&cur);
    cudaSetDevice(t->device);
    cudaFree(t->arena);
    if (cur >= 0) cudaSetDevice(cur);
  }
  delete t;
}

int tree_v2_launch(const TreeV2* t, int mode, int dtype, const SpecParams& p, void* stream) {
  if (!t) return -1;
  V2Params q;
  q.t.node = t->node; q.t.be = t->be; q.t.inv = t->inv; q.t.level_ptr = t->level_ptr;
  q.t.eparent = t->eparent; q.t.pbe = t->pbe;
  q.t.class_pos = t->class_pos; q.t.class_edge = t->class_edge; q.t.pos_edge = t->pos_edge;
  q.t.C = t->C; q.t.N = t->N; q.t.S = t->S; q.t.D = t->D;
  q.p = p;
  q.row_floats = (2 * (t->C + 1) + 2 * t->S) | 1;
  q.row_floats += 1;  // keep rows 8-byte aligned for the fp64 prefix sums (even float count) ...
  if ((q.row_floats / 2) % 2 == 0) q.row_floats += 2;  // ... with an odd number of 8-byte words per row
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case NBDT_F32: return launch_dtype<float>(t, mode, q, st);
    case NBDT_BF16: return launch_dtype<__nv_bfloat16>(t, mode, q, st);
    case NBDT_F16: return launch_dtype<__half>(t, mode, q, st);
    default: return -1;
  }
}
