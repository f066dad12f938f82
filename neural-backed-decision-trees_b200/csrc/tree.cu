// Fused NBDT tree-head / tree-supervision-loss kernels for sm_100a.
//
// One warp owns 32/G consecutive rows of the logits matrix ("G lanes per row", G in {1..32} chosen
// per tree so that each lane has ~16-32 tree nodes of work).  For its rows the warp keeps two fp32
// arrays over all NV = C + N tree nodes in shared memory (leaf c at slot c, inner node i at C+i):
//
//   V1: subtree sums S[v]  ->  log child-probability lp[v]  ->  accumulated leaf gradient A[v]
//   V2: log path probability LP[v]                          ->  dL/dlp[v] (= subtree sums of leaf grads)
//
// and makes four sweeps over the hierarchy, each a precomputed schedule of "rounds" of <= G mutually
// independent inner nodes (by height for the upward sweeps, by depth for the downward ones):
//
//   up   : S[n] = sum_children S[c]                       (node logit of child c = S[c] / #leaves(c))
//   down : lp[c] = log_softmax over the children of n ;  LP[c] = LP[n] + lp[c]
//   leaf : class_probs = exp(LP[leaf]);  loss terms;  g[leaf] = dL/dLP[leaf]
//   up   : G[n] = sum_children G[c]
//   down : A[c] = A[n] + (G[c] - exp(lp[c]) * G[n]) / #leaves(c) ;  dL/dx[c] = A[leaf c] + CE term
//
// Everything between the coalesced load of the rows and the coalesced store of their gradient
// stays in shared memory / registers: algorithmic HBM traffic is one read of the logits (+ targets)
// and one write of the gradient (or class_probs).  No __syncthreads in the row loop: rows are
// warp-private, rounds are separated by __syncwarp.
//
// Reference semantics: nbdt/model.py:83-123 (node logits = mean of leaf logits, softmax, argmax,
// entropy), model.py:145-186 (hard walk), model.py:207-242 (soft path product), loss.py:191-266.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#include "common.h"
#include "tree_spec.h"
#include "tree_jit.h"
#include "tree_v2.h"

namespace {

constexpr int kNumG = 6;  // G = 1,2,4,8,16,32

struct Sched {
  const int* up;    // [up_rounds * G]   inode index or -1
  const int* down;  // [down_rounds * G]
  int up_rounds;
  int down_rounds;
};

struct TreeDev {
  int C, N, NV, NVp, S, root, Dmax, Kmax;
  unsigned magicC;          // floor(2^32 / C) + 1  (exact e / C for e < 2^32 / C)
  const int* child_ptr;     // [N+1]
  const int* child_node;    // [S]
  const float* inv_cnt;     // [NV]  1 / #leaves under node
  Sched sched[kNumG];
};

enum Mode { kSoftFwd = 0, kHardFwd = 1, kSoftLoss = 2, kHardLoss = 3, kNodes = 4, kSoftBwd = 5 };

struct Params {
  TreeDev t;
  const void* x;
  long long B;
  long long ld;
  const long long* targets;
  // outputs
  float* class_probs;     // kSoftFwd
  const float* grad_probs; // kSoftBwd: dL/dclass_probs, fp32 [B, C]
  long long* pred;        // kHardFwd
  int* path_child;        // kHardFwd (optional)
  int* path_inode;
  float* path_prob;
  float* path_entropy;
  float* node_logits;     // kNodes (optional each)
  float* node_probs;
  long long* node_preds;
  float* node_entropy;
  float* loss;            // loss modes (device scalar, pre-zeroed)
  void* grad;             // loss modes (optional)
  float w_xent;           // w_x
  float w_tree;           // w_t  (hard: w_t * tsw * 2 / N)
  float inv_rows;         // 1 / batch_rows
  float grad_scale;
  long long hw;           // > 0: class-major (N, C, H, W) logits / gradient / class probabilities, hw pixels per image (see elem_off)
};

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
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int G> __device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int G> __device__ __forceinline__ void round_sync() {
  if (G > 1) __syncwarp();
}

// S[n] = sum of children's S, bottom-up.
template <int G>
__device__ __forceinline__ void sweep_up(const TreeDev& t, float* __restrict__ v, int j) {
  const Sched& s = t.sched[31 - __clz(G)];
  for (int rd = 0; rd < s.up_rounds; ++rd) {
    const int i = __ldg(s.up + rd * G + j);
    if (i >= 0) {
      const int b = __ldg(t.child_ptr + i), e = __ldg(t.child_ptr + i + 1);
      float acc = 0.f;
      for (int k = b; k < e; ++k) acc += v[__ldg(t.child_node + k)];
      v[t.C + i] = acc;
    }
    round_sync<G>();
  }
}

// Element (row, class c) of the logits (and of the gradient / class-probability outputs, which share the layout): row-major with
// stride ld, or class-major (N, C, H, W) when hw > 0 -- the segmentation variants' tensors, addressed in place instead of through
// the reference's permute + reshape copy (nbdt/utils.py:338-347, model.py:376-399, loss.py:318-327).
__device__ __forceinline__ long long elem_off(long long row, int c, long long ld, long long hw, int C) {
  if (hw > 0) {
    const long long img = row / hw;
    return (img * C + c) * hw + (row - img * hw);
  }
  return row * ld + c;
}

template <int G, int MODE, typename T>
__global__ void __launch_bounds__(512) tree_kernel(const Params p) {
  extern __shared__ float smem[];
  const TreeDev& t = p.t;
  constexpr int RPW = 32 / G;  // rows per warp
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int r = lane / G;  // row within the warp's chunk
  const int j = lane % G;  // lane within the row group
  const int C = t.C, NVp = t.NVp;
  float* V1 = smem + (size_t)warp * (2 * RPW * NVp);
  float* V2 = V1 + RPW * NVp;
  float* v1 = V1 + r * NVp;
  float* v2 = V2 + r * NVp;
  const T* __restrict__ x = reinterpret_cast<const T*>(p.x);
  const Sched& sc = t.sched[31 - __clz(G)];

  const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
  const long long gwarp = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  const long long chunks = (p.B + RPW - 1) / RPW;
  float loss_acc = 0.f;

  for (long long chunk = gwarp; chunk < chunks; chunk += warps_total) {
    const long long row0 = chunk * RPW;
    const int nrows = (int)min((long long)RPW, p.B - row0);
    const bool valid = r < nrows;
    const long long myrow = row0 + r;

    // ---- load: the warp's rows are RPW*C consecutive logical elements; lane-contiguous reads ----
    for (int e = lane; e < RPW * C; e += 32) {
      const int rr = __umulhi((unsigned)e, t.magicC);
      const int c = e - rr * C;
      float val = 0.f;
      if (rr < nrows) val = to_f<T>(x[elem_off(row0 + rr, c, p.ld, p.hw, C)]);
      V1[rr * NVp + c] = val;
    }
    __syncwarp();

    // ---- plain cross-entropy statistics of the raw logits (loss modes) ----
    float lse1 = 0.f, xy = 0.f;
    int y = 0;
    if (MODE == kSoftLoss || MODE == kHardLoss) {
      float m1 = -INFINITY;
      for (int c = j; c < C; c += G) m1 = fmaxf(m1, v1[c]);
      m1 = group_max<G>(m1);
      float s1 = 0.f;
      for (int c = j; c < C; c += G) s1 += __expf(v1[c] - m1);
      s1 = group_sum<G>(s1);
      lse1 = m1 + __logf(s1);
      // target outside [0, C): the row is ignored (zero loss, zero gradient), like nn.CrossEntropyLoss's ignore_index; y = -1 marks it
      if (valid) {
        const long long yt = p.targets[myrow];
        y = (unsigned long long)yt < (unsigned long long)C ? (int)yt : -1;
      }
      xy = v1[y < 0 ? 0 : y];
      __syncwarp();
    }

    // ---- sweep 1 (up): subtree sums ----
    sweep_up<G>(t, v1, j);

    if (MODE == kHardFwd) {
      // greedy walk root -> leaf by one lane per row (depth <= Dmax, children <= Kmax)
      if (j == 0 && valid) {
        int node = C + t.root;
        int d = 0;
        while (node >= C) {
          const int i = node - C;
          const int b = __ldg(t.child_ptr + i), e = __ldg(t.child_ptr + i + 1);
          float m = -INFINITY;
          int best = b;
          for (int k = b; k < e; ++k) {
            const int ch = __ldg(t.child_node + k);
            const float nl = v1[ch] * __ldg(t.inv_cnt + ch);
            if (nl > m) { m = nl; best = k; }  // strict > keeps the first maximum (torch.max)
          }
          if (p.path_child) {
            float s = 0.f, dot = 0.f;
            for (int k = b; k < e; ++k) {
              const int ch = __ldg(t.child_node + k);
              const float z = v1[ch] * __ldg(t.inv_cnt + ch) - m;
              const float ez = __expf(z);
              s += ez;
              dot += ez * z;
            }
            // p_k = ez/s ; entropy = -sum p_k (z_k - log s) = log s - dot/s ; prob(best) = 1/s
            const long long o = myrow * t.Dmax + d;
            p.path_child[o] = best - b;
            if (p.path_inode) p.path_inode[o] = i;
            if (p.path_prob) p.path_prob[o] = 1.f / s;
            if (p.path_entropy) p.path_entropy[o] = __logf(s) - dot / s;
          }
          node = __ldg(t.child_node + best);
          ++d;
        }
        p.pred[myrow] = node;
        if (p.path_child) {
          for (; d < t.Dmax; ++d) {
            const long long o = myrow * t.Dmax + d;
            p.path_child[o] = -1;
            if (p.path_inode) p.path_inode[o] = -1;
            if (p.path_prob) p.path_prob[o] = 0.f;
            if (p.path_entropy) p.path_entropy[o] = 0.f;
          }
        }
      }
      __syncwarp();
      continue;
    }

    if (MODE == kNodes) {
      // per-node outputs, packed by child slot; node-parallel over the down schedule
      for (int rd = 0; rd < sc.down_rounds; ++rd) {
        const int i = __ldg(sc.down + rd * G + j);
        if (i >= 0 && valid) {
          const int b = __ldg(t.child_ptr + i), e = __ldg(t.child_ptr + i + 1);
          float m = -INFINITY;
          int best = b;
          for (int k = b; k < e; ++k) {
            const int ch = __ldg(t.child_node + k);
            const float nl = v1[ch] * __ldg(t.inv_cnt + ch);
            if (p.node_logits) p.node_logits[myrow * t.S + k] = nl;
            if (nl > m) { m = nl; best = k; }
          }
          float s = 0.f, dot = 0.f;
          for (int k = b; k < e; ++k) {
            const int ch = __ldg(t.child_node + k);
            const float z = v1[ch] * __ldg(t.inv_cnt + ch) - m;
            const float ez = __expf(z);
            s += ez;
            dot += ez * z;
          }
          const float inv = 1.f / s;
          if (p.node_probs)
            for (int k = b; k < e; ++k) {
              const int ch = __ldg(t.child_node + k);
              p.node_probs[myrow * t.S + k] = __expf(v1[ch] * __ldg(t.inv_cnt + ch) - m) * inv;
            }
          if (p.node_preds) p.node_preds[myrow * t.N + i] = best - b;
          if (p.node_entropy) p.node_entropy[myrow * t.N + i] = __logf(s) - dot * inv;
        }
      }
      __syncwarp();
      continue;
    }

    // ---- sweep 2 (down): lp[c] = log_softmax over siblings, LP[c] = LP[parent] + lp[c] ----
    if (j == 0) v2[C + t.root] = 0.f;
    round_sync<G>();
    for (int rd = 0; rd < sc.down_rounds; ++rd) {
      const int i = __ldg(sc.down + rd * G + j);
      if (i >= 0) {
        const int b = __ldg(t.child_ptr + i), e = __ldg(t.child_ptr + i + 1);
        const float base = v2[C + i];
        if (e - b == 2) {  // induced hierarchies are binary: log-sigmoid form
          const int c0 = __ldg(t.child_node + b), c1 = __ldg(t.child_node + b + 1);
          const float a0 = v1[c0] * __ldg(t.inv_cnt + c0);
          const float a1 = v1[c1] * __ldg(t.inv_cnt + c1);
          const float m = fmaxf(a0, a1);
          const float lse = m + __logf(1.f + __expf(-fabsf(a0 - a1)));
          const float l0 = a0 - lse, l1 = a1 - lse;
          v1[c0] = l0; v1[c1] = l1;
          v2[c0] = base + l0; v2[c1] = base + l1;
        } else {
          float m = -INFINITY;
          for (int k = b; k < e; ++k) {
            const int ch = __ldg(t.child_node + k);
            m = fmaxf(m, v1[ch] * __ldg(t.inv_cnt + ch));
          }
          float s = 0.f;
          for (int k = b; k < e; ++k) {
            const int ch = __ldg(t.child_node + k);
            s += __expf(v1[ch] * __ldg(t.inv_cnt + ch) - m);
          }
          const float lse = m + __logf(s);
          for (int k = b; k < e; ++k) {
            const int ch = __ldg(t.child_node + k);
            const float lp = v1[ch] * __ldg(t.inv_cnt + ch) - lse;
            v1[ch] = lp;
            v2[ch] = base + lp;
          }
        }
      }
      round_sync<G>();
    }

    if (MODE == kSoftFwd) {
      // class_probs = exp(LP[leaf]); coalesced fp32 store of the warp's rows
      if (G == 1) __syncwarp();
      for (int e = lane; e < RPW * C; e += 32) {
        const int rr = __umulhi((unsigned)e, t.magicC);
        const int c = e - rr * C;
        if (rr < nrows) p.class_probs[elem_off(row0 + rr, c, C, p.hw, C)] = __expf(V2[rr * NVp + c]);
      }
      __syncwarp();
      continue;
    }

    // ---- leaf stage: loss terms and dL/dLP[leaf] ----
    const float kx = p.w_xent * p.inv_rows;  // weight of one row's plain CE in the mean
    const float kt = p.w_tree * p.inv_rows;
    float row_loss = 0.f;
    if (MODE == kSoftBwd) {
      // generic criterion on top of the soft rules: dL/dLP[c] = dL/dp[c] * p[c]
      for (int c = j; c < C; c += G)
        v2[c] = valid ? p.grad_scale * __ldg(p.grad_probs + myrow * C + c) * __expf(v2[c]) : 0.f;
    } else if (MODE == kSoftLoss) {
      float s2 = 0.f;
      for (int c = j; c < C; c += G) s2 += __expf(__expf(v2[c]));  // probs in [0,1]: no max shift needed
      s2 = group_sum<G>(s2);
      const float lse2 = __logf(s2);
      const float py = __expf(v2[y < 0 ? 0 : y]);
      row_loss = kx * (lse1 - xy) + kt * (lse2 - py);
      if (p.grad) {
        round_sync<G>();
        const float kg = kt * p.grad_scale;
        for (int c = j; c < C; c += G) {
          const float pc = __expf(v2[c]);
          const float sm = __expf(pc - lse2);
          v2[c] = y < 0 ? 0.f : kg * (sm - (c == y ? 1.f : 0.f)) * pc;
        }
      }
    } else {  // kHardLoss:  tree term = -LP[y] (sum of -lp along the target's path)
      row_loss = kx * (lse1 - xy) - kt * v2[y < 0 ? 0 : y];
      if (p.grad) {
        round_sync<G>();
        const float kg = kt * p.grad_scale;
        for (int c = j; c < C; c += G) v2[c] = (c == y) ? -kg : 0.f;
      }
    }
    if (j == 0 && valid && y >= 0) loss_acc += row_loss;
    if (!p.grad) {
      __syncwarp();
      continue;
    }
    round_sync<G>();

    // ---- sweep 3 (up): G[n] = sum_children G[c] ----
    sweep_up<G>(t, v2, j);

    // ---- sweep 4 (down): A[c] = A[n] + (G[c] - exp(lp[c]) G[n]) / cnt[c] ----
    for (int rd = 0; rd < sc.down_rounds; ++rd) {
      const int i = __ldg(sc.down + rd * G + j);
      if (i >= 0) {
        const int b = __ldg(t.child_ptr + i), e = __ldg(t.child_ptr + i + 1);
        const float Gn = v2[C + i];
        const float An = (i == t.root) ? 0.f : v1[C + i];
        for (int k = b; k < e; ++k) {
          const int ch = __ldg(t.child_node + k);
          const float pk = __expf(v1[ch]);
          v1[ch] = An + (v2[ch] - pk * Gn) * __ldg(t.inv_cnt + ch);
        }
      }
      round_sync<G>();
    }
    if (G == 1) __syncwarp();

    // ---- store: dL/dx = A[leaf] + w_x/B (softmax(x) - onehot(y)); x re-read (L2-resident) ----
    {
      T* __restrict__ g = reinterpret_cast<T*>(p.grad);
      const float kgx = (MODE == kSoftBwd) ? 0.f : kx * p.grad_scale;
      // warp-uniform trip count: every lane must reach the shuffles (full-mask shfl in a loop that
      // some lanes have left is undefined -- it faulted as "illegal instruction" on sm_100a)
      for (int e0 = 0; e0 < RPW * C; e0 += 32) {
        const int e = e0 + lane;
        const int rr = min((int)__umulhi((unsigned)e, t.magicC), RPW - 1);
        const int c = e - rr * C;
        const float l1 = __shfl_sync(0xffffffffu, lse1, rr * G);
        const int yy = __shfl_sync(0xffffffffu, y, rr * G);
        if (e < RPW * C && rr < nrows) {
          const long long o = elem_off(row0 + rr, c, p.ld, p.hw, C);
          float gx = V1[rr * NVp + c];
          if (MODE != kSoftBwd) gx = yy < 0 ? 0.f : gx + kgx * (__expf(to_f<T>(x[o]) - l1) - (c == yy ? 1.f : 0.f));
          g[o] = from_f<T>(gx);
        }
      }
    }
    __syncwarp();
  }

  if (MODE == kSoftLoss || MODE == kHardLoss) {
    // block reduction of the per-row losses, one atomic per CTA
    __shared__ float red[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
    if (lane == 0) red[warp] = loss_acc;
    __syncthreads();
    if (warp == 0) {
      float v = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) atomicAdd(p.loss, v);
    }
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side: table construction and launch
// ---------------------------------------------------------------------------------------------
struct nbdt_tree {
  int device;
  TreeDev dev;
  void* arena;  // single device allocation holding every table
  int sm_count;
  int max_smem;
  int force_g;  // env NBDT_B200_TREE_G (0 = heuristic)
  const SpecEntry* spec;  // hierarchy-specialised straight-line kernels, when generated for this tree
  TreeV2* v2;             // second-generation generic kernel (tree_v2.cu) for the loss / soft-forward modes
  TreeJit* jit;           // run-time compiled straight-line kernels (tree_jit.cu), any hierarchy up to 1536 leaves
  mutable char last_kernel[96];  // kernel family of the last soft-forward / loss launch (bench lines, tests)
};

namespace {

unsigned long long tree_hash(const nbdt_tree_desc* d) {
  unsigned long long h = 0xCBF29CE484222325ull;
  auto mix = [&](int v) {
    for (int b = 0; b < 4; ++b) {
      h ^= (unsigned long long)((v >> (8 * b)) & 0xFF);
      h *= 0x100000001B3ull;
    }
  };
  mix(d->num_classes); mix(d->num_inodes); mix(d->root);
  for (int i = 0; i <= d->num_inodes; ++i) mix(d->child_ptr[i]);
  for (int i = 0; i < d->child_ptr[d->num_inodes]; ++i) mix(d->child_node[i]);
  return h;
}

// Launch the specialised kernel if one exists for this tree/mode/dtype and the rows are aligned for its
// vector width.  Returns -1 when the generic kernel must be used.
int try_launch_spec(const nbdt_tree* t, int spec_mode, int dtype, const Params& p, cudaStream_t stream) {
  if (!t->spec || t->force_g || dtype < 0 || dtype > 2 || p.hw > 0) return -1;
  const int esz = dtype == NBDT_F32 ? 4 : 2;
  const int row_bytes = t->dev.C * esz;
  const int vec = row_bytes % 16 == 0 ? 16 : (row_bytes % 8 == 0 ? 8 : (row_bytes % 4 == 0 ? 4 : esz));
  auto aligned = [&](const void* ptr) { return ptr == nullptr || (reinterpret_cast<uintptr_t>(ptr) % vec) == 0; };
  if (!aligned(p.x) || !aligned(p.grad) || (p.ld * esz) % vec != 0) return -1;
  if (spec_mode == kSpecSoftFwd && (reinterpret_cast<uintptr_t>(p.class_probs) % 16) != 0) return -1;
  SpecParams sp;
  sp.x = p.x; sp.B = p.B; sp.ld = p.ld; sp.targets = p.targets; sp.class_probs = p.class_probs; sp.loss = p.loss;
  sp.grad = p.grad; sp.w_xent = p.w_xent; sp.w_tree = p.w_tree; sp.inv_rows = p.inv_rows; sp.grad_scale = p.grad_scale;
  sp.hw = 0;
  const void* fn = t->spec->fn[spec_mode][dtype];
  long long grid = (p.B + 127) / 128;
  grid = std::min<long long>(grid, (long long)t->sm_count * 12);
  void* args[1] = {&sp};
  NBDT_CUDA_OK(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(128), args, 0, stream));
  return NBDT_OK;
}

int pick_group(const nbdt_tree* t) {
  if (t->force_g) return t->force_g;
  const int nv = t->dev.NV;
  int g = 1;
  while (g < 32 && nv > 24 * g) g <<= 1;
  return g;
}

template <int G, int MODE, typename T>
int launch_gt(const nbdt_tree* t, const Params& p, cudaStream_t stream) {
  const int rpw = 32 / G;
  const size_t smem_per_warp = (size_t)2 * rpw * t->dev.NVp * sizeof(float);
  const size_t budget = (size_t)t->max_smem - 1024;
  if (smem_per_warp > budget)
    return nbdt::set_error(NBDT_ETOOLARGE, "hierarchy with %d nodes needs %zu B of shared memory per row",
                           t->dev.NV, smem_per_warp);
  // warps per CTA: as many as fit, at most 8 (more CTAs per SM beat bigger CTAs for latency hiding)
  int wpc = (int)std::min<size_t>(8, budget / smem_per_warp);
  if (smem_per_warp * 8 > budget) wpc = (int)std::min<size_t>(16, budget / smem_per_warp);
  const size_t smem = smem_per_warp * wpc;
  auto kern = tree_kernel<G, MODE, T>;
  NBDT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int ctas_per_sm = 1;
  NBDT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, wpc * 32, smem));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  const long long chunks = (p.B + rpw - 1) / rpw;
  long long grid = (chunks + wpc - 1) / wpc;
  grid = std::min<long long>(grid, (long long)t->sm_count * ctas_per_sm);
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, wpc * 32, smem, stream>>>(p);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

template <int MODE, typename T>
int launch_t(const nbdt_tree* t, const Params& p, cudaStream_t stream) {
  switch (pick_group(t)) {
    case 1: return launch_gt<1, MODE, T>(t, p, stream);
    case 2: return launch_gt<2, MODE, T>(t, p, stream);
    case 4: return launch_gt<4, MODE, T>(t, p, stream);
    case 8: return launch_gt<8, MODE, T>(t, p, stream);
    case 16: return launch_gt<16, MODE, T>(t, p, stream);
    default: return launch_gt<32, MODE, T>(t, p, stream);
  }
}

template <int MODE>
int launch(const nbdt_tree* t, int dtype, const Params& p, cudaStream_t stream) {
  int cur = -1;
  NBDT_CUDA_OK(cudaGetDevice(&cur));
  if (cur != t->device) NBDT_CUDA_OK(cudaSetDevice(t->device));
  int rc = -1;
  if (MODE == kSoftLoss) rc = try_launch_spec(t, kSpecSoftLoss, dtype, p, stream);
  if (MODE == kHardLoss) rc = try_launch_spec(t, kSpecHardLoss, dtype, p, stream);
  if (MODE == kSoftFwd) rc = try_launch_spec(t, kSpecSoftFwd, dtype, p, stream);
  int which = rc >= 0 ? 1 : 0;  // 1 build-time specialised, 2 run-time specialised, 3 tree_v2, 0 generic
  if (rc < 0 && t->jit && !t->force_g && (MODE == kSoftLoss || MODE == kHardLoss || MODE == kSoftFwd)) {
    SpecParams sp;
    sp.x = p.x; sp.B = p.B; sp.ld = p.ld; sp.targets = p.targets; sp.class_probs = p.class_probs; sp.loss = p.loss;
    sp.grad = p.grad; sp.w_xent = p.w_xent; sp.w_tree = p.w_tree; sp.inv_rows = p.inv_rows; sp.grad_scale = p.grad_scale;
    sp.hw = p.hw;
    rc = tree_jit_launch(t->jit, MODE == kSoftLoss ? kSpecSoftLoss : (MODE == kHardLoss ? kSpecHardLoss : kSpecSoftFwd),
                         dtype, sp, stream);
    if (rc >= 0) which = 2;
  }
  if (rc < 0 && t->v2 && !t->force_g && p.hw == 0 && (MODE == kSoftLoss || MODE == kHardLoss || MODE == kSoftFwd)) {
    SpecParams sp;
    sp.x = p.x; sp.B = p.B; sp.ld = p.ld; sp.targets = p.targets; sp.class_probs = p.class_probs; sp.loss = p.loss;
    sp.grad = p.grad; sp.w_xent = p.w_xent; sp.w_tree = p.w_tree; sp.inv_rows = p.inv_rows; sp.grad_scale = p.grad_scale;
    sp.hw = 0;
    rc = tree_v2_launch(t->v2, MODE == kSoftLoss ? kSpecSoftLoss : (MODE == kHardLoss ? kSpecHardLoss : kSpecSoftFwd),
                        dtype, sp, stream);
  }
  if (rc >= 0) {
    if (which == 1) snprintf(t->last_kernel, sizeof t->last_kernel, "tree_spec_%016llx (build-time generated, <= 24 leaves)", t->spec ? t->spec->hash : 0ull);
    else if (which == 2) snprintf(t->last_kernel, sizeof t->last_kernel, "%s", tree_jit_last_variant(t->jit));
    else snprintf(t->last_kernel, sizeof t->last_kernel, "tree_v2_kernel (table-driven)");
    if (cur != t->device) cudaSetDevice(cur);
    return rc;
  }
  snprintf(t->last_kernel, sizeof t->last_kernel, "tree_kernel (table-driven)");
  switch (dtype) {
    case NBDT_F32: rc = launch_t<MODE, float>(t, p, stream); break;
    case NBDT_BF16: rc = launch_t<MODE, __nv_bfloat16>(t, p, stream); break;
    case NBDT_F16: rc = launch_t<MODE, __half>(t, p, stream); break;
    default: rc = nbdt::set_error(NBDT_EINVAL, "unknown dtype %d", dtype);
  }
  if (cur != t->device) cudaSetDevice(cur);
  return rc;
}

int check_common(const nbdt_tree* tree, const void* logits, int64_t B, int64_t ld) {
  NBDT_REQUIRE(tree != nullptr, "tree is NULL");
  NBDT_REQUIRE(B >= 0, "negative row count");
  NBDT_REQUIRE(B == 0 || logits != nullptr, "logits is NULL");
  NBDT_REQUIRE(ld >= tree->dev.C, "ld (%lld) smaller than the number of classes (%d)", (long long)ld, tree->dev.C);
  return NBDT_OK;
}

Params base_params(const nbdt_tree* tree, const void* logits, int64_t B, int64_t ld) {
  Params p = {};
  p.t = tree->dev;
  p.x = logits;
  p.B = B;
  p.ld = ld;
  p.inv_rows = 1.f;
  p.grad_scale = 1.f;
  return p;
}

}  // namespace

extern "C" {

int nbdt_tree_create(const nbdt_tree_desc* d, int device, nbdt_tree** out) {
  NBDT_REQUIRE(d && out, "NULL argument");
  *out = nullptr;
  const int C = d->num_classes, N = d->num_inodes;
  NBDT_REQUIRE(C >= 1 && N >= 1, "need at least one leaf and one inner node (C=%d, N=%d)", C, N);
  NBDT_REQUIRE(C <= 8192, "at most 8192 classes are supported (got %d)", C);
  NBDT_REQUIRE(d->root >= 0 && d->root < N, "root index %d out of range", d->root);
  NBDT_REQUIRE(d->child_ptr && d->child_node, "NULL table");
  NBDT_REQUIRE(d->child_ptr[0] == 0, "child_ptr[0] must be 0");
  const int NV = C + N;
  const int S = d->child_ptr[N];
  if (S != NV - 1) return nbdt::set_error(NBDT_ETREE, "not a tree: %d child slots for %d nodes", S, NV);

  std::vector<int> parent(NV, -1), depth(NV, -1), height(NV, 0), cnt(NV, 0);
  int kmax = 0;
  for (int i = 0; i < N; ++i) {
    const int b = d->child_ptr[i], e = d->child_ptr[i + 1];
    if (e <= b) return nbdt::set_error(NBDT_ETREE, "inner node %d has no children", i);
    kmax = std::max(kmax, e - b);
    for (int k = b; k < e; ++k) {
      const int ch = d->child_node[k];
      if (ch < 0 || ch >= NV) return nbdt::set_error(NBDT_ETREE, "child id %d out of range", ch);
      if (parent[ch] != -1 || ch == C + d->root)
        return nbdt::set_error(NBDT_ETREE, "node %d has more than one parent (hierarchy must be a tree)", ch);
      parent[ch] = C + i;
    }
  }
  // BFS from the root: depth, order; every node must be reached exactly once
  std::vector<int> order;
  order.reserve(NV);
  order.push_back(C + d->root);
  depth[C + d->root] = 0;
  for (size_t q = 0; q < order.size(); ++q) {
    const int v = order[q];
    if (v < C) continue;
    const int i = v - C;
    for (int k = d->child_ptr[i]; k < d->child_ptr[i + 1]; ++k) {
      const int ch = d->child_node[k];
      depth[ch] = depth[v] + 1;
      order.push_back(ch);
    }
  }
  if ((int)order.size() != NV) return nbdt::set_error(NBDT_ETREE, "%d of %d nodes unreachable from the root", NV - (int)order.size(), NV);
  int dmax = 0;
  for (int q = NV - 1; q >= 0; --q) {
    const int v = order[q];
    if (v < C) { cnt[v] = 1; dmax = std::max(dmax, depth[v]); }
    if (parent[v] >= 0) {
      cnt[parent[v]] += cnt[v];
      height[parent[v]] = std::max(height[parent[v]], height[v] + 1);
    }
  }
  if (cnt[C + d->root] != C) return nbdt::set_error(NBDT_ETREE, "root covers %d of %d leaves", cnt[C + d->root], C);

  // schedules: inner nodes grouped by height (up) / depth (down), chunked into rounds of G
  std::vector<std::vector<int>> by_h(height[C + d->root] + 1), by_d(dmax + 1);
  for (int i = 0; i < N; ++i) {
    by_h[height[C + i]].push_back(i);
    by_d[depth[C + i]].push_back(i);
  }
  std::vector<int> host_int;  // all int tables, concatenated
  auto push = [&](const std::vector<int>& v) { size_t o = host_int.size(); host_int.insert(host_int.end(), v.begin(), v.end()); return o; };
  const size_t off_ptr = push(std::vector<int>(d->child_ptr, d->child_ptr + N + 1));
  const size_t off_child = push(std::vector<int>(d->child_node, d->child_node + S));
  size_t off_up[kNumG], off_down[kNumG];
  int up_rounds[kNumG], down_rounds[kNumG];
  for (int gi = 0; gi < kNumG; ++gi) {
    const int G = 1 << gi;
    auto build = [&](const std::vector<std::vector<int>>& levels, int* rounds) {
      std::vector<int> s;
      for (const auto& lv : levels)
        for (size_t a = 0; a < lv.size(); a += G)
          for (int q = 0; q < G; ++q) s.push_back(a + q < lv.size() ? lv[a + q] : -1);
      *rounds = (int)s.size() / G;
      return s;
    };
    off_up[gi] = push(build(by_h, &up_rounds[gi]));
    off_down[gi] = push(build(by_d, &down_rounds[gi]));
  }
  std::vector<float> inv_cnt(NV);
  for (int v = 0; v < NV; ++v) inv_cnt[v] = 1.0f / (float)cnt[v];

  nbdt_tree* t = new (std::nothrow) nbdt_tree();
  NBDT_REQUIRE(t != nullptr, "out of host memory");
  t->device = device;
  t->arena = nullptr;
  nbdt::DeviceInfo info;
  int rc = nbdt::device_info(device, &info);
  if (rc) { delete t; return rc; }
  t->sm_count = info.sm_count;
  t->max_smem = info.max_smem_optin;
  t->force_g = 0;
  t->spec = nullptr;
  if (!getenv("NBDT_B200_TREE_GENERIC")) {
    const unsigned long long h = tree_hash(d);
    for (int i = 0; i < kSpecCount; ++i)
      if (kSpecTable[i].hash == h && kSpecTable[i].C == C) t->spec = &kSpecTable[i];
  }
  if (const char* env = getenv("NBDT_B200_TREE_G")) {
    const int g = atoi(env);
    if (g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32) t->force_g = g;
  }
  int cur = -1;
  cudaGetDevice(&cur);
  const size_t int_bytes = host_int.size() * sizeof(int);
  const size_t bytes = int_bytes + inv_cnt.size() * sizeof(float);
  cudaError_t ce = cudaSetDevice(device);
  if (ce == cudaSuccess) ce = cudaMalloc(&t->arena, bytes);
  if (ce == cudaSuccess) ce = cudaMemcpy(t->arena, host_int.data(), int_bytes, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy((char*)t->arena + int_bytes, inv_cnt.data(), inv_cnt.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (cur >= 0) cudaSetDevice(cur);
  if (ce != cudaSuccess) {
    if (t->arena) cudaFree(t->arena);
    delete t;
    cudaGetLastError();
    return nbdt::set_error(NBDT_ECUDA, "uploading tree tables to device %d failed: %s", device, cudaGetErrorString(ce));
  }
  const int* base = reinterpret_cast<const int*>(t->arena);
  TreeDev& dv = t->dev;
  dv.C = C; dv.N = N; dv.NV = NV; dv.NVp = NV | 1; dv.S = S; dv.root = d->root; dv.Dmax = dmax; dv.Kmax = kmax;
  dv.magicC = (unsigned)((1ull << 32) / (unsigned)C) + 1u;
  dv.child_ptr = base + off_ptr;
  dv.child_node = base + off_child;
  dv.inv_cnt = reinterpret_cast<const float*>((const char*)t->arena + int_bytes);
  for (int gi = 0; gi < kNumG; ++gi) {
    dv.sched[gi].up = base + off_up[gi];
    dv.sched[gi].down = base + off_down[gi];
    dv.sched[gi].up_rounds = up_rounds[gi];
    dv.sched[gi].down_rounds = down_rounds[gi];
  }
  t->v2 = nullptr;
  t->jit = nullptr;
  t->last_kernel[0] = 0;
  if (!getenv("NBDT_B200_TREE_GENERIC"))
    if (int rc3 = tree_jit_create(d, device, t->sm_count, t->max_smem, &t->jit)) {
      nbdt_tree_destroy(t);
      return rc3;
    }
  if (int rc2 = tree_v2_create(d, device, t->sm_count, t->max_smem, &t->v2)) {
    nbdt_tree_destroy(t);
    return rc2;
  }
  *out = t;
  return NBDT_OK;
}

void nbdt_tree_destroy(nbdt_tree* t) {
  if (!t) return;
  tree_v2_destroy(t->v2);
  tree_jit_destroy(t->jit);
  if (t->arena) {
    int cur = -1;
    cudaGetDevice(&cur);
    cudaSetDevice(t->device);
    cudaFree(t->arena);
    if (cur >= 0) cudaSetDevice(cur);
  }
  delete t;
}

int nbdt_tree_max_depth(const nbdt_tree* t) { return t ? t->dev.Dmax : -1; }
int nbdt_tree_num_slots(const nbdt_tree* t) { return t ? t->dev.S : -1; }
const char* nbdt_tree_kernel_variant(const nbdt_tree* t) { return t ? t->last_kernel : ""; }

int nbdt_soft_forward(const nbdt_tree* tree, const void* logits, int dtype, int64_t B, int64_t ld,
                      float* class_probs, void* stream) {
  if (int rc = check_common(tree, logits, B, ld)) return rc;
  if (B == 0) return NBDT_OK;
  NBDT_REQUIRE(class_probs != nullptr, "class_probs is NULL");
  Params p = base_params(tree, logits, B, ld);
  p.class_probs = class_probs;
  return launch<kSoftFwd>(tree, dtype, p, (cudaStream_t)stream);
}

int nbdt_hard_forward(const nbdt_tree* tree, const void* logits, int dtype, int64_t B, int64_t ld,
                      int64_t* pred, int32_t* path_child, int32_t* path_inode, float* path_prob,
                      float* path_entropy, void* stream) {
  if (int rc = check_common(tree, logits, B, ld)) return rc;
  if (B == 0) return NBDT_OK;
  NBDT_REQUIRE(pred != nullptr, "pred is NULL");
  NBDT_REQUIRE(path_child != nullptr || (!path_inode && !path_prob && !path_entropy),
               "path_child is required when any other path output is requested");
  Params p = base_params(tree, logits, B, ld);
  p.pred = reinterpret_cast<long long*>(pred);
  p.path_child = path_child;
  p.path_inode = path_inode;
  p.path_prob = path_prob;
  p.path_entropy = path_entropy;
  return launch<kHardFwd>(tree, dtype, p, (cudaStream_t)stream);
}

int nbdt_node_outputs(const nbdt_tree* tree, const void* logits, int dtype, int64_t B, int64_t ld,
                      float* node_logits, float* node_probs, int64_t* preds, float* entropy, void* stream) {
  if (int rc = check_common(tree, logits, B, ld)) return rc;
  if (B == 0) return NBDT_OK;
  Params p = base_params(tree, logits, B, ld);
  p.node_logits = node_logits;
  p.node_probs = node_probs;
  p.node_preds = reinterpret_cast<long long*>(preds);
  p.node_entropy = entropy;
  return launch<kNodes>(tree, dtype, p, (cudaStream_t)stream);
}

int nbdt_soft_backward(const nbdt_tree* tree, const void* logits, int dtype, int64_t B, int64_t ld,
                       const float* grad_probs, float grad_scale, void* grad_logits, void* stream) {
  if (int rc = check_common(tree, logits, B, ld)) return rc;
  if (B == 0) return NBDT_OK;
  NBDT_REQUIRE(grad_probs != nullptr && grad_logits != nullptr, "NULL gradient buffer");
  Params p = base_params(tree, logits, B, ld);
  p.grad_probs = grad_probs;
  p.grad = grad_logits;
  p.grad_scale = grad_scale;
  return launch<kSoftBwd>(tree, dtype, p, (cudaStream_t)stream);
}

static int loss_common(int mode, const nbdt_tree* tree, const void* logits, int dtype, const int64_t* targets,
                       int64_t B, int64_t ld, int64_t batch_rows, float w_xent, float w_tree,
                       float grad_scale, float* loss, void* grad_logits, void* stream) {
  if (int rc = check_common(tree, logits, B, ld)) return rc;
  NBDT_REQUIRE(loss != nullptr, "loss is NULL");
  NBDT_CUDA_OK(cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream));
  if (B == 0) return NBDT_OK;
  NBDT_REQUIRE(targets != nullptr, "targets is NULL");
  Params p = base_params(tree, logits, B, ld);
  p.targets = reinterpret_cast<const long long*>(targets);
  p.loss = loss;
  p.grad = grad_logits;
  p.w_xent = w_xent;
  p.w_tree = w_tree;
  p.inv_rows = 1.0f / (float)(batch_rows > 0 ? batch_rows : B);
  p.grad_scale = grad_scale;
  return mode == kSoftLoss ? launch<kSoftLoss>(tree, dtype, p, (cudaStream_t)stream)
                           : launch<kHardLoss>(tree, dtype, p, (cudaStream_t)stream);
}

int nbdt_soft_loss(const nbdt_tree* tree, const void* logits, int dtype, const int64_t* targets, int64_t B,
                   int64_t ld, int64_t batch_rows, float w_xent, float w_tree, float grad_scale, float* loss,
                   void* grad_logits, void* stream) {
  return loss_common(kSoftLoss, tree, logits, dtype, targets, B, ld, batch_rows, w_xent, w_tree, grad_scale,
                     loss, grad_logits, stream);
}

int nbdt_hard_loss(const nbdt_tree* tree, const void* logits, int dtype, const int64_t* targets, int64_t B,
                   int64_t ld, int64_t batch_rows, float w_xent, float w_tree, float tsw, float grad_scale,
                   float* loss, void* grad_logits, void* stream) {
  NBDT_REQUIRE(tree != nullptr, "tree is NULL");
  const float wt = w_tree * tsw * 2.0f / (float)tree->dev.N;
  return loss_common(kHardLoss, tree, logits, dtype, targets, B, ld, batch_rows, w_xent, wt, grad_scale, loss,
                     grad_logits, stream);
}

// ---- class-major (N, C, H, W) logits: the segmentation variants without the permute + reshape copy ----
// Row r of the hierarchy kernels = pixel r % hw of image r / hw (hw = H*W); logits, gradient and class probabilities are all
// (N, C, H, W) contiguous; targets and predictions are (N, H, W) = one entry per row.  Replace coerce_tensor / uncoerce_tensor
// around the rules in SegNBDT.forward (nbdt/model.py:376-399) and SoftSegTreeSupLoss.forward (nbdt/loss.py:318-327).
// Run-time specialised kernels when hw is a multiple of 32 (a block of 32 rows is 32 consecutive pixels of one image: every
// class is one coalesced 128-byte line), the table-driven kernel otherwise.
int nbdt_soft_forward_nchw(const nbdt_tree* tree, const void* logits, int dtype, int64_t N, int64_t hw, float* class_probs,
                           void* stream) {
  NBDT_REQUIRE(tree != nullptr && N >= 0 && hw > 0, "bad arguments");
  if (int rc = check_common(tree, logits, N * hw, tree->dev.C)) return rc;
  if (N == 0) return NBDT_OK;
  NBDT_REQUIRE(class_probs != nullptr, "class_probs is NULL");
  Params p = base_params(tree, logits, N * hw, tree->dev.C);
  p.hw = hw;
  p.class_probs = class_probs;
  return launch<kSoftFwd>(tree, dtype, p, (cudaStream_t)stream);
}

int nbdt_hard_forward_nchw(const nbdt_tree* tree, const void* logits, int dtype, int64_t N, int64_t hw, int64_t* pred,
                           void* stream) {
  NBDT_REQUIRE(tree != nullptr && N >= 0 && hw > 0, "bad arguments");
  if (int rc = check_common(tree, logits, N * hw, tree->dev.C)) return rc;
  if (N == 0) return NBDT_OK;
  NBDT_REQUIRE(pred != nullptr, "pred is NULL");
  Params p = base_params(tree, logits, N * hw, tree->dev.C);
  p.hw = hw;
  p.pred = reinterpret_cast<long long*>(pred);
  return launch<kHardFwd>(tree, dtype, p, (cudaStream_t)stream);
}

int nbdt_soft_loss_nchw(const nbdt_tree* tree, const void* logits, int dtype, const int64_t* targets, int64_t N, int64_t hw,
                        int64_t batch_rows, float w_xent, float w_tree, float grad_scale, float* loss, void* grad_logits,
                        void* stream) {
  NBDT_REQUIRE(tree != nullptr && N >= 0 && hw > 0, "bad arguments");
  if (int rc = check_common(tree, logits, N * hw, tree->dev.C)) return rc;
  NBDT_REQUIRE(loss != nullptr, "loss is NULL");
  NBDT_CUDA_OK(cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream));
  if (N == 0) return NBDT_OK;
  NBDT_REQUIRE(targets != nullptr, "targets is NULL");
  Params p = base_params(tree, logits, N * hw, tree->dev.C);
  p.hw = hw;
  p.targets = reinterpret_cast<const long long*>(targets);
  p.loss = loss;
  p.grad = grad_logits;
  p.w_xent = w_xent;
  p.w_tree = w_tree;
  p.inv_rows = 1.0f / (float)(batch_rows > 0 ? batch_rows : N * hw);
  p.grad_scale = grad_scale;
  return launch<kSoftLoss>(tree, dtype, p, (cudaStream_t)stream);
}

}  // extern "C"
