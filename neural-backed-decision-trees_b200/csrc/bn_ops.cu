// Memory-bound backbone kernels around the tensor-core convolutions (NHWC bf16 activations, fp32
// parameters and statistics): training-mode BatchNorm statistics / apply(+ReLU, +residual) / backward,
// global average pooling, the 3-channel stem convolution, fused SGD.
//
// Every kernel streams its tensors once with 16-byte vector accesses (8 bf16 channels per thread); the
// reductions keep per-channel partial sums in registers, combine them across the CTA in shared memory and
// write one partial per CTA that a sliced finalize kernel sums in a fixed order (deterministic).
// Roofline: HBM; algorithmic bytes are listed per kernel below (M = N*H*W rows, C channels); measured
// fractions: profiles/r01_bn_kernels.md.
//
// Replaces, in the reference's backbones, nn.BatchNorm2d + F.relu (+ residual add)
// (nbdt/models/resnet.py:50-54,70-73 and the pytorchcv pre-activation blocks), F.avg_pool2d
// (resnet.py:142) and optim.SGD (main.py:207).
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "common.h"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
    f[2 * i] = __bfloat162float(b.x);
    f[2 * i + 1] = __bfloat162float(b.y);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&b);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// Thread mapping shared by every BatchNorm kernel: a CTA of 256 threads covers R = 256/octets consecutive rows x
// `octets` channel octets (thread -> r = tid / octets, o = tid % octets; active iff r < R), i.e. R*C*2 contiguous
// bytes per pass, and advances by R*gridDim.x rows.  A thread owns ONE channel octet for its whole life, so the
// per-channel constants (scale/shift, mean, gradient coefficients) live in registers: no shared-memory reads and no
// index arithmetic in the streaming loops (the v1 kernels spent their time in a 64-bit `i % octets` per 16-byte
// access, profiles/r01_bn_kernels.md).  f(o, first_row, row_stride) runs the thread's row loop.
template <typename F>
__device__ __forceinline__ void for_octet_rows(int octets, F f) {
  if (octets <= kThreads) {
    const int R = kThreads / octets;
    const int r = threadIdx.x / octets, o = threadIdx.x - r * octets;
    if (r < R) f(o, (long long)blockIdx.x * R + r, (long long)gridDim.x * R);
  } else {
    for (int o = threadIdx.x; o < octets; o += kThreads) f(o, (long long)blockIdx.x, (long long)gridDim.x);
  }
}

__device__ __forceinline__ void load8(const float* __restrict__ p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// Per-channel reductions on that mapping: NACC float accumulators per channel; `Op` provides
//   void init(o)                 -- load the thread's per-channel constants
//   Pack load(row, o)            -- the global loads of one row (issued UNROLL rows at a time for memory-level parallelism)
//   void accum(pack, acc)        -- the arithmetic
// Each CTA writes its partial sums to out[blockIdx.x][NACC][C]; reduce_partials2() adds them in a fixed order
// (deterministic; same-address atomics serialise in L2).
// acc64 (optional, only for C <= 8 * kThreads): instead of a partial row per CTA, every CTA adds its sums to ONE fp64
// accumulator [NACC][C] (zeroed by the caller) and the CONSUMER kernel derives its coefficients from it in its prologue: the
// separate finalize launch (five 1024-thread CTAs chasing ~600 partial rows: 10-15 us of pure latency per BatchNorm, 50 per
// train step) disappears.  fp64 atomics: the order of the ~600 additions per channel changes the sum in the 16th digit only.
template <int NACC, int UNROLL, typename Op>
__device__ __forceinline__ void reduce_rows(long long M, int C, float* __restrict__ out, Op op, double* __restrict__ acc64 = nullptr) {
  extern __shared__ float red[];  // [R][NACC][C]  (only when R > 1)
  const int octets = C / 8;
  float* part = out + (long long)blockIdx.x * NACC * C;
  const int R = octets <= kThreads ? kThreads / octets : 1;
  for_octet_rows(octets, [&](int o, long long row, long long stride) {
    float acc[NACC][8];
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[a][e] = 0.f;
    op.init(o);
    for (; row + (UNROLL - 1) * stride < M; row += UNROLL * stride) {
      decltype(op.load(row, o)) pk[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) pk[u] = op.load(row + u * stride, o);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) op.accum(pk[u], acc);
    }
    for (; row < M; row += stride) op.accum(op.load(row, o), acc);
    if (R == 1 && acc64) {  // 128 <= octets <= 256: one thread per octet, nothing to combine inside the CTA
#pragma unroll
      for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(acc64 + (size_t)a * C + o * 8 + e, (double)acc[a][e]);
      return;
    }
    float* dst = R > 1 ? red + (size_t)(threadIdx.x / octets) * NACC * C : part;
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
      float4* d4 = reinterpret_cast<float4*>(dst + (size_t)a * C + o * 8);
      d4[0] = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
      d4[1] = make_float4(acc[a][4], acc[a][5], acc[a][6], acc[a][7]);
    }
  });
  if (R > 1) {
    __syncthreads();
    for (int i = threadIdx.x; i < NACC * C; i += kThreads) {
      float s = 0.f;
      for (int q = 0; q < R; ++q) s += red[(size_t)q * NACC * C + i];
      if (acc64) atomicAdd(acc64 + i, (double)s);
      else part[i] = s;
    }
  }
}

// Totals of the per-CTA partials [nparts][2][C] for channel c (fixed order, fp64).  CTA = 1024 threads = 32 channels
// (lanes) x 32 slices of the partial list (warps): a thread's chain of dependent L2 reads is nparts/256 iterations
// with sixteen loads in flight (the v1 finalize kernels read all partials from one thread per channel: ~20 us of pure
// latency per BatchNorm; round 1's four loads in flight still took ten round trips).  The totals are returned to warp 0.
constexpr int kFinalizeThreads = 1024;
__device__ __forceinline__ void reduce_partials2(const float* __restrict__ partials, int nparts, int C, int c,
                                                 double& t0, double& t1) {
  __shared__ double slices[2][32][33];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  // eight rows of the partial list per thread and iteration = sixteen independent L2 reads in flight: the 592 rows of a
  // 4-CTAs-per-SM reduction are three dependent round trips instead of ten (the finalize kernels are pure latency)
  constexpr int U = 8;
  double a[U], bq[U];
#pragma unroll
  for (int u = 0; u < U; ++u) { a[u] = 0.0; bq[u] = 0.0; }
  if (c < C) {
    const long long n = 2LL * C;
    int b = slice;
    for (; b + 32 * (U - 1) < nparts; b += 32 * U) {
      float va[U], vb[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        va[u] = __ldg(partials + (long long)(b + 32 * u) * n + c);
        vb[u] = __ldg(partials + (long long)(b + 32 * u) * n + C + c);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) { a[u] += (double)va[u]; bq[u] += (double)vb[u]; }
    }
    for (; b < nparts; b += 32) {
      a[0] += (double)__ldg(partials + (long long)b * n + c);
      bq[0] += (double)__ldg(partials + (long long)b * n + C + c);
    }
  }
  double sa = 0.0, sb = 0.0;
#pragma unroll
  for (int u = 0; u < U; ++u) { sa += a[u]; sb += bq[u]; }
  slices[0][slice][lane] = sa;
  slices[1][slice][lane] = sb;
  __syncthreads();
  t0 = 0.0;
  t1 = 0.0;
  if (slice == 0) {
#pragma unroll 8
    for (int q = 0; q < 32; ++q) {
      t0 += slices[0][q][lane];
      t1 += slices[1][q][lane];
    }
  }
}

// `last` (both reduction ops): M - 1 to walk the rows from the END of the tensor, -1 for the natural order.  The producer of the
// tensor (convolution / data-gradient epilogue) wrote it front to back, so its tail is what the 126 MB L2 still holds when the
// reduction starts; reading back to front turns that part into L2 hits, and leaves the FRONT hot for the apply pass that follows
// (which walks front to back).  NBDT_B200_BN_REVERSE=0 restores the natural order (A/B).
struct StatsOp {
  const uint4* x;
  int octets;
  long long last;
  __device__ __forceinline__ void init(int) {}
  __device__ __forceinline__ uint4 load(long long row, int o) const { return __ldg(x + (last >= 0 ? last - row : row) * octets + o); }
  __device__ __forceinline__ void accum(const uint4& v, float (&acc)[2][8]) const {
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[0][e] += f[e];
      acc[1][e] = fmaf(f[e], f[e], acc[1][e]);
    }
  }
};

// partials[cta][0][c] = sum_rows x, [1][c] = sum_rows x^2.          bytes: M*C*2 read
__global__ void __launch_bounds__(kThreads, 4) bn_stats_kernel(const uint4* __restrict__ x, long long M, int C,
                                                            float* __restrict__ partials, double* __restrict__ acc64, long long last) {
  reduce_rows<2, 8>(M, C, partials, StatsOp{x, C / 8, last}, acc64);
}

// What the apply kernel needs to turn the accumulated sums into this BatchNorm's coefficients (and to publish them).
struct BnFromSums {
  const double* sums;  // [2][C]: sum x, sum x^2 over M rows; nullptr = coefficients are given
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  float* mean_invstd;  // out [2][C] (saved for the backward pass)
  float* scale_shift;  // out [2][C]
  long long M;
  int c_real;
  float momentum, eps;
};

// mean/invstd, affine scale/shift, running-stat update (momentum, unbiased variance).  grid = ceil(C/32) x 1024 threads.
__global__ void __launch_bounds__(kFinalizeThreads) bn_finalize_kernel(const float* __restrict__ partials, int nparts, long long M,
                                                               int C, int c_real, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta,
                                                               float* __restrict__ running_mean,
                                                               float* __restrict__ running_var, float momentum, float eps,
                                                               float* __restrict__ mean_invstd /*[2][C]*/,
                                                               float* __restrict__ scale_shift /*[2][C]*/) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  double s0, s1;
  reduce_partials2(partials, nparts, C, c, s0, s1);
  if (threadIdx.x >= 32 || c >= C) return;
  if (c >= c_real) {  // zero-padded channel: stays exactly zero
    mean_invstd[c] = 0.f; mean_invstd[C + c] = 0.f; scale_shift[c] = 0.f; scale_shift[C + c] = 0.f;
    return;
  }
  const double inv_m = 1.0 / (double)M;
  const double mean = s0 * inv_m;
  double var = s1 * inv_m - mean * mean;
  if (var < 0) var = 0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  mean_invstd[c] = (float)mean;
  mean_invstd[C + c] = invstd;
  const float sc = gamma[c] * invstd;
  scale_shift[c] = sc;
  scale_shift[C + c] = beta[c] - (float)mean * sc;
  if (running_mean) {
    const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// eval mode: scale/shift from the running statistics
__global__ void bn_eval_scale_kernel(int C, int c_real, const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                     float eps, float* __restrict__ scale_shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (c >= c_real) { scale_shift[c] = 0.f; scale_shift[C + c] = 0.f; return; }
  const float sc = gamma[c] * rsqrtf(running_var[c] + eps);
  scale_shift[c] = sc;
  scale_shift[C + c] = beta[c] - running_mean[c] * sc;
}

// y = act(x * scale[c] + shift[c] (+ residual)).            bytes: M*C*2 read (+M*C*2) + M*C*2 written
template <bool RES>
__global__ void __launch_bounds__(kThreads) bn_apply_kernel(const uint4* __restrict__ x, const uint4* __restrict__ res,
                                                            uint4* __restrict__ y, long long M, int C,
                                                            const float* __restrict__ scale_shift, int relu, const BnFromSums fs) {
  const int octets = C / 8;
  for_octet_rows(octets, [&](int o, long long row, long long stride) {
    float sc[8], sh[8];
    if (fs.sums) {
      // every thread finalizes its own eight channels (same arithmetic as bn_finalize_kernel); the threads that start at row 0
      // publish mean / invstd / scale / shift and move the running statistics
      const bool publish = row == 0;
      const double inv_m = 1.0 / (double)fs.M;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = o * 8 + e;
        if (c >= fs.c_real) {  // zero-padded channel: stays exactly zero
          sc[e] = 0.f; sh[e] = 0.f;
          if (publish) { fs.mean_invstd[c] = 0.f; fs.mean_invstd[C + c] = 0.f; fs.scale_shift[c] = 0.f; fs.scale_shift[C + c] = 0.f; }
          continue;
        }
        const double mean = fs.sums[c] * inv_m;
        double var = fs.sums[C + c] * inv_m - mean * mean;
        if (var < 0) var = 0;
        const float invstd = (float)(1.0 / sqrt(var + (double)fs.eps));
        sc[e] = fs.gamma[c] * invstd;
        sh[e] = fs.beta[c] - (float)mean * sc[e];
        if (publish) {
          fs.mean_invstd[c] = (float)mean;
          fs.mean_invstd[C + c] = invstd;
          fs.scale_shift[c] = sc[e];
          fs.scale_shift[C + c] = sh[e];
          if (fs.running_mean) {
            const double unbiased = fs.M > 1 ? var * (double)fs.M / (double)(fs.M - 1) : var;
            fs.running_mean[c] = (1.f - fs.momentum) * fs.running_mean[c] + fs.momentum * (float)mean;
            fs.running_var[c] = (1.f - fs.momentum) * fs.running_var[c] + fs.momentum * (float)unbiased;
          }
        }
      }
    } else {
      load8(scale_shift + o * 8, sc);
      load8(scale_shift + C + o * 8, sh);
    }
    auto one = [&](long long rw, const uint4& vx, const uint4& vr) {
      float f[8], r[8];
      unpack8(vx, f);
      if (RES) unpack8(vr, r);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v = fmaf(f[e], sc[e], sh[e]);
        if (RES) v += r[e];
        // activation code: 0 none, 1 ReLU, 2 SiLU (inference path of the MBConv backbones)
        f[e] = relu == 1 ? fmaxf(v, 0.f) : (relu == 2 ? __fdividef(v, 1.f + __expf(-v)) : v);
      }
      y[rw * octets + o] = pack8(f);
    };
    constexpr int U = 4;
    for (; row + (U - 1) * stride < M; row += U * stride) {
      uint4 vx[U], vr[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        vx[u] = __ldg(x + (row + u * stride) * octets + o);
        if (RES) vr[u] = __ldg(res + (row + u * stride) * octets + o);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) one(row + u * stride, vx[u], vr[u]);
    }
    for (; row < M; row += stride)
      one(row, __ldg(x + row * octets + o), RES ? __ldg(res + row * octets + o) : make_uint4(0, 0, 0, 0));
  });
}

// ReLU mask of the forward pass: kMaskNone (no activation), kMaskAct (read the saved forward output: needed when a
// residual was added before the ReLU), kMaskRecompute (x*scale + shift > 0, bit-identical to the forward's fmaf:
// saves reading one tensor in each backward kernel).
// kMaskSiLU: the forward activation was SiLU(z), z = x*scale + shift: g = dy * SiLU'(z), SiLU'(z) = s (1 + z (1 - s)), s = sigmoid(z)
enum { kMaskNone = 0, kMaskAct = 1, kMaskRecompute = 2, kMaskSiLU = 3 };
__device__ __forceinline__ float silu_grad(float z) {
  const float sg = __fdividef(1.f, 1.f + __expf(-z));
  return sg * fmaf(z, 1.f - sg, 1.f);
}

struct BwdPack {
  uint4 x, g, a;
};
template <int MASK>
struct BwdReduceOp {
  const uint4* x;
  const uint4* act;
  const uint4* dy;
  const float* mean_invstd;
  const float* scale_shift;
  int octets, C;
  long long last;
  float mean[8], sc[8], sh[8];
  __device__ __forceinline__ void init(int o) {
    load8(mean_invstd + o * 8, mean);
    if (MASK == kMaskRecompute || MASK == kMaskSiLU) {
      load8(scale_shift + o * 8, sc);
      load8(scale_shift + C + o * 8, sh);
    }
  }
  __device__ __forceinline__ BwdPack load(long long row, int o) const {
    BwdPack p;
    const long long rr = last >= 0 ? last - row : row;
    p.x = __ldg(x + rr * octets + o);
    p.g = __ldg(dy + rr * octets + o);
    if (MASK == kMaskAct) p.a = __ldg(act + rr * octets + o);
    return p;
  }
  __device__ __forceinline__ void accum(const BwdPack& p, float (&acc)[2][8]) const {
    float fx[8], fg[8], fa[8];
    unpack8(p.x, fx);
    unpack8(p.g, fg);
    if (MASK == kMaskAct) unpack8(p.a, fa);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      bool on = true;
      if (MASK == kMaskAct) on = fa[e] > 0.f;
      if (MASK == kMaskRecompute) on = fmaf(fx[e], sc[e], sh[e]) > 0.f;
      float g = on ? fg[e] : 0.f;
      if (MASK == kMaskSiLU) g = fg[e] * silu_grad(fmaf(fx[e], sc[e], sh[e]));
      acc[0][e] += g;
      acc[1][e] = fmaf(g, fx[e] - mean[e], acc[1][e]);
    }
  }
};

// partials[cta][0][c] = sum g, [1][c] = sum g * (x - mean), with g = dy masked by the forward ReLU.
// bytes: M*C*2 * (2 or 3) read
template <int MASK>
__global__ void __launch_bounds__(kThreads, 4) bn_bwd_reduce_kernel(const uint4* __restrict__ x, const uint4* __restrict__ act,
                                                                 const uint4* __restrict__ dy, long long M, int C,
                                                                 const float* __restrict__ mean_invstd,
                                                                 const float* __restrict__ scale_shift,
                                                                 float* __restrict__ partials, double* __restrict__ acc64, long long last) {
  BwdReduceOp<MASK> op;
  op.x = x; op.act = act; op.dy = dy; op.mean_invstd = mean_invstd; op.scale_shift = scale_shift;
  op.octets = C / 8; op.C = C; op.last = last;
  reduce_rows<2, 4>(M, C, partials, op, acc64);
}

struct BnBwdFromSums {
  const double* sums;  // [2][C]: sum g, sum g (x - mean); nullptr = coef is given
  const float* gamma;
  const float* mean_invstd;
  float* dsums;        // out [2][C]: dbeta, dgamma
  int c_real;
};

// dsums[0][c] = dbeta = sum g, dsums[1][c] = dgamma = invstd * sum g (x - mean); coef[3][C] = (k1, a, b) with
// dx = k1*g + a*x + b:  k1 = gamma*invstd, a = -k1*invstd*dgamma/M, b = -k1*dbeta/M - a*mean.
__global__ void __launch_bounds__(kFinalizeThreads) bn_bwd_finalize_kernel(const float* __restrict__ partials, int nparts, long long M,
                                                                   int C, int c_real, const float* __restrict__ gamma,
                                                                   const float* __restrict__ mean_invstd,
                                                                   float* __restrict__ dsums, float* __restrict__ coef) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  double s0, s1;
  reduce_partials2(partials, nparts, C, c, s0, s1);
  if (threadIdx.x >= 32 || c >= C) return;
  const float mean = mean_invstd[c], invstd = mean_invstd[C + c];
  const float dbeta = (float)s0, dgamma = (float)(s1 * (double)invstd);
  dsums[c] = dbeta;
  dsums[C + c] = dgamma;
  const float inv_m = 1.f / (float)M;
  const float k1 = c < c_real ? gamma[c] * invstd : 0.f;
  const float a = -k1 * invstd * dgamma * inv_m;
  coef[c] = k1;
  coef[C + c] = a;
  coef[2 * C + c] = -k1 * dbeta * inv_m - a * mean;
}

// dx = k1*g + a*x + b (+ extra);  optionally dres = g (masked dy).
// bytes: M*C*2 * (2..4) read + M*C*2 * (1..2) written
template <int MASK, bool EXTRA, bool DRES>
__global__ void __launch_bounds__(kThreads, 3) bn_bwd_apply_kernel(const uint4* __restrict__ x, const uint4* __restrict__ act,
                                                                const uint4* __restrict__ dy, const uint4* __restrict__ extra,
                                                                uint4* __restrict__ dx, uint4* __restrict__ dres,
                                                                long long M, int C, const float* __restrict__ coef,
                                                                const float* __restrict__ scale_shift, const BnBwdFromSums fs) {
  const int octets = C / 8;
  for_octet_rows(octets, [&](int o, long long row, long long stride) {
    float k1[8], ka[8], kb[8], sc[8], sh[8];
    if (fs.sums) {  // the arithmetic of bn_bwd_finalize_kernel, per thread for its own eight channels
      const bool publish = row == 0;
      const float inv_m = 1.f / (float)M;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = o * 8 + e;
        const float mean = fs.mean_invstd[c], invstd = fs.mean_invstd[C + c];
        const float dbeta = (float)fs.sums[c], dgamma = (float)(fs.sums[C + c] * (double)invstd);
        if (publish) { fs.dsums[c] = dbeta; fs.dsums[C + c] = dgamma; }
        k1[e] = c < fs.c_real ? fs.gamma[c] * invstd : 0.f;
        ka[e] = -k1[e] * invstd * dgamma * inv_m;
        kb[e] = -k1[e] * dbeta * inv_m - ka[e] * mean;
      }
    } else {
      load8(coef + o * 8, k1);
      load8(coef + C + o * 8, ka);
      load8(coef + 2 * C + o * 8, kb);
    }
    if (MASK == kMaskRecompute || MASK == kMaskSiLU) {
      load8(scale_shift + o * 8, sc);
      load8(scale_shift + C + o * 8, sh);
    }
    auto one = [&](long long rw, const uint4& vx, const uint4& vg, const uint4& va, const uint4& ve) {
      float fx[8], fg[8], fa[8], fe[8], out[8];
      unpack8(vx, fx);
      unpack8(vg, fg);
      if (MASK == kMaskAct) unpack8(va, fa);
      if (EXTRA) unpack8(ve, fe);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        bool on = true;
        if (MASK == kMaskAct) on = fa[e] > 0.f;
        if (MASK == kMaskRecompute) on = fmaf(fx[e], sc[e], sh[e]) > 0.f;
        float g = on ? fg[e] : 0.f;
        if (MASK == kMaskSiLU) g = fg[e] * silu_grad(fmaf(fx[e], sc[e], sh[e]));
        float v = fmaf(k1[e], g, fmaf(ka[e], fx[e], kb[e]));
        if (EXTRA) v += fe[e];
        out[e] = v;
        fg[e] = g;
      }
      dx[rw * octets + o] = pack8(out);
      if (DRES) dres[rw * octets + o] = pack8(fg);
    };
    // rows in flight per thread: ~8 x 16-byte loads whatever the number of tensors read
    constexpr int U = (MASK == kMaskAct || EXTRA) ? ((MASK == kMaskAct && EXTRA) ? 2 : 3) : 4;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (; row + (U - 1) * stride < M; row += U * stride) {
      uint4 vx[U], vg[U], va[U], ve[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long k = (row + u * stride) * octets + o;
        vx[u] = __ldg(x + k);
        vg[u] = __ldg(dy + k);
        va[u] = MASK == kMaskAct ? __ldg(act + k) : zero;
        ve[u] = EXTRA ? __ldg(extra + k) : zero;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) one(row + u * stride, vx[u], vg[u], va[u], ve[u]);
    }
    for (; row < M; row += stride) {
      const long long k = row * octets + o;
      one(row, __ldg(x + k), __ldg(dy + k), MASK == kMaskAct ? __ldg(act + k) : zero, EXTRA ? __ldg(extra + k) : zero);
    }
  });
}

// feat[n][c] = mean over HW rows of a[n, :, c]  (fp32 out).      bytes: N*HW*C*2 read
// CTA (slice s, image n) reduces pixels [s*pix_per_cta, ...) on the row-octet mapping; with one slice per image the
// result is written directly (deterministic: the 8x8 maps of the CIFAR backbones), with several slices (the 112x112 ...
// 7x7 maps of EfficientNet's squeeze-excitation, where one CTA per image was 42 % of the forward time) the slices are
// combined with fp32 atomics into a zeroed output.
__global__ void __launch_bounds__(kThreads) avgpool_fwd_kernel(const uint4* __restrict__ a, float* __restrict__ feat, int HW,
                                                               int C, int pix_per_cta, int use_atomics) {
  __shared__ float red[kThreads * 8];
  const int octets = C / 8;
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
  const float inv = 1.f / (float)HW;
  const uint4* img = a + (size_t)n * HW * octets;
  if (octets <= kThreads) {
    const int R = kThreads / octets;
    const int r = threadIdx.x / octets, o = threadIdx.x - r * octets;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (r < R) {
      for (int p = p0 + r; p < p1; p += R) {
        float f[8];
        unpack8(__ldg(img + (size_t)p * octets + o), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] += f[e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) red[(r * octets + o) * 8 + e] = s[e];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kThreads) {
      float t = 0.f;
      for (int q = 0; q < R; ++q) t += red[q * C + c];
      if (use_atomics) atomicAdd(feat + (size_t)n * C + c, t * inv);
      else feat[(size_t)n * C + c] = t * inv;
    }
  } else {
    for (int o = threadIdx.x; o < octets; o += kThreads) {
      float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int p = p0; p < p1; ++p) {
        float f[8];
        unpack8(__ldg(img + (size_t)p * octets + o), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] += f[e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (use_atomics) atomicAdd(feat + (size_t)n * C + o * 8 + e, s[e] * inv);
        else feat[(size_t)n * C + o * 8 + e] = s[e] * inv;
      }
    }
  }
}
// da[n, p, c] = dfeat[n][c] / HW
__global__ void avgpool_bwd_kernel(const float* __restrict__ dfeat, uint4* __restrict__ da, long long total_octets,
                                   int HW, int C) {
  const int octets = C / 8;
  const float inv = 1.f / (float)HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_octets; i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i % octets);
    const long long n = i / ((long long)octets * HW);
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = dfeat[n * C + o * 8 + e] * inv;
    da[i] = pack8(f);
  }
}

// Stem: y[n,h,w,co] = sum_{ci<3, r, s} w[co][ci][r][s] * x[n,ci,h+r-1,w+s-1]   (x NCHW fp32, y NHWC bf16,
// c_phys >= cout physical channels, padding channels zero).  One thread per output pixel.
__global__ void __launch_bounds__(128) stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       __nv_bfloat16* __restrict__ y, int N, int H, int W, int cout,
                                                       int c_phys, int stride) {
  extern __shared__ float ws[];  // [27][cout]
  for (int i = threadIdx.x; i < 27 * cout; i += blockDim.x) {
    const int co = i % cout, k = i / cout;  // k = ci*9 + r*3 + s
    ws[i] = w[co * 27 + k];
  }
  __syncthreads();
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)N * Ho * Wo) return;
  const int wq = (int)(pix % Wo) * stride, hq = (int)((pix / Wo) % Ho) * stride;
  const long long n = pix / ((long long)Wo * Ho);
  float in[27];
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int hh = hq + r - 1, ww = wq + s - 1;
        in[ci * 9 + r * 3 + s] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(x + ((n * 3 + ci) * H + hh) * W + ww) : 0.f;
      }
  uint4* dst = reinterpret_cast<uint4*>(y + pix * c_phys);
  for (int c0 = 0; c0 < c_phys; c0 += 8) {
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float acc = 0.f;
      if (c0 + e < cout) {
#pragma unroll
        for (int k = 0; k < 27; ++k) acc = fmaf(in[k], ws[k * cout + c0 + e], acc);
      }
      o[e] = acc;
    }
    dst[c0 / 8] = pack8(o);
  }
}

// Stem weight gradient: dw[co][k] = sum_pixels dy[pix][co] * patch[pix][k]  (k = ci*9 + r*3 + s; a 16 x 27 x #pixels GEMM
// with a tiny output).  Thread (lane, g) walks a stream of pixels and keeps the 4 x 27 outputs of channels 4g..4g+3
// in registers: 27 patch loads (L1-resident NCHW gathers) + one 8-byte dy load feed 108 FMAs.  A warp has one g, so
// the end-of-kernel reduction is a shuffle tree + 108 atomics per warp.  (The first version staged 64-pixel tiles in
// shared memory between two __syncthreads with 8 warps per SM: latency-bound, 0.4 ms per step.)
__global__ void __launch_bounds__(128) stem_wgrad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                         float* __restrict__ dw, int N, int H, int W, int cout,
                                                         int c_phys, int stride) {
  const int lane = threadIdx.x;                         // pixel lane
  const int co0 = (blockIdx.y * 4 + threadIdx.y) * 4;   // first of this thread's 4 output channels
  if (co0 >= cout) return;                              // whole warp (one threadIdx.y per warp)
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;  // dy is [N, Ho, Wo, c_phys]
  const unsigned total = (unsigned)N * Ho * Wo;
  float acc[4][27];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[c][k] = 0.f;
  for (unsigned pix = blockIdx.x * 32 + lane; pix < total; pix += gridDim.x * 32) {
    const int wq = (int)(pix % (unsigned)Wo) * stride, hq = (int)((pix / (unsigned)Wo) % (unsigned)Ho) * stride;
    const unsigned n = pix / ((unsigned)Wo * (unsigned)Ho);
    const uint2 gv = __ldg(reinterpret_cast<const uint2*>(dy + (size_t)pix * c_phys + co0));
    const float g[4] = {__uint_as_float(gv.x << 16), __uint_as_float(gv.x & 0xffff0000u),
                        __uint_as_float(gv.y << 16), __uint_as_float(gv.y & 0xffff0000u)};
    const float* xb = x + (size_t)n * 3 * H * W;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int hh = hq + r - 1, ww = wq + s - 1;
          const float v = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(xb + ((size_t)ci * H + hh) * W + ww) : 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[c][ci * 9 + r * 3 + s] = fmaf(g[c], v, acc[c][ci * 9 + r * 3 + s]);
        }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      float v = acc[c][k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && co0 + c < cout) atomicAdd(dw + (co0 + c) * 27 + k, v);
    }
}

// p -= lr * (m = momentum*m + g + wd*p)   (torch.optim.SGD, dampening 0, no nesterov; first step m = g + wd*p)
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, long long n,
                           float lr, float momentum, float wd, float grad_scale, int first_step) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = fmaf(wd, p[i], g[i] * grad_scale);
    const float buf = first_step ? d : fmaf(momentum, m[i], d);
    m[i] = buf;
    p[i] -= lr * buf;
  }
}

int grid_for(long long work_items, int sm_count, int per_sm = 8) {
  const long long blocks = (work_items + kThreads - 1) / kThreads;
  return (int)std::max<long long>(1, std::min<long long>(blocks, (long long)sm_count * per_sm));
}

int sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  nbdt::DeviceInfo info;
  if (nbdt::device_info(dev, &info)) return 148;
  return info.sm_count;
}


// Launch geometry of the row-octet mapping: CTAs of 256 threads, `per_sm` per SM, never more than one pass of rows.
int rows_grid(long long M, int C, int per_sm) {
  const int octets = C / 8;
  const int R = octets <= kThreads ? kThreads / octets : 1;
  return (int)std::max<long long>(1, std::min<long long>((M + R - 1) / R, (long long)sm_count() * per_sm));
}
size_t reduce_smem(int C) {
  const int octets = C / 8;
  const int R = octets <= kThreads ? kThreads / octets : 1;
  return R > 1 ? (size_t)R * 2 * C * sizeof(float) : 0;
}
constexpr int kReduceCtasPerSm = 4;
long long reverse_last(long long M) {
  static const bool off = getenv("NBDT_B200_BN_REVERSE") && atoi(getenv("NBDT_B200_BN_REVERSE")) == 0;
  return off ? -1 : M - 1;
}

}  // namespace

extern "C" {

// Training-mode statistics of x[M][C] (bf16): fills mean_invstd[2][C], scale_shift[2][C], updates the running
// statistics (may be NULL).  `sums`: fp32 scratch of nbdt_bn_scratch_floats(C).  c_real <= C: channels beyond are padding.
int nbdt_bn_stats(const void* x, long long M, int C, int c_real, const float* gamma, const float* beta,
                  float* running_mean, float* running_var, float momentum, float eps, float* sums,
                  float* mean_invstd, float* scale_shift, void* stream) {
  NBDT_REQUIRE(x && gamma && beta && sums && mean_invstd && scale_shift, "NULL tensor");
  NBDT_REQUIRE(C % 8 == 0 && C <= 8192 && M > 0, "bad shape M=%lld C=%d", M, C);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = rows_grid(M, C, kReduceCtasPerSm);
  bn_stats_kernel<<<grid, kThreads, reduce_smem(C), st>>>(reinterpret_cast<const uint4*>(x), M, C, sums, nullptr, reverse_last(M));
  bn_finalize_kernel<<<(C + 31) / 32, kFinalizeThreads, 0, st>>>(sums, grid, M, C, c_real, gamma, beta, running_mean, running_var,
                                                         momentum, eps, mean_invstd, scale_shift);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

// Training-mode BatchNorm coefficients from per-CTA partial sums somebody else produced: partials[part][0][c] = sum x,
// [1][c] = sum x^2 over M rows in total -- the rows nbdt_conv2d_fwd_stats (conv_gemm.cu) writes from the convolution's
// epilogue.  Same finalize kernel as nbdt_bn_stats (fp64 reduction in a fixed order, running-statistics update).
int nbdt_bn_finalize_partials(const float* partials, int nparts, long long M, int C, int c_real, const float* gamma,
                              const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                              float* mean_invstd, float* scale_shift, void* stream) {
  NBDT_REQUIRE(partials && gamma && beta && mean_invstd && scale_shift, "NULL tensor");
  NBDT_REQUIRE(C % 8 == 0 && C <= 8192 && M > 0 && nparts > 0, "bad shape M=%lld C=%d parts=%d", M, C, nparts);
  bn_finalize_kernel<<<(C + 31) / 32, kFinalizeThreads, 0, (cudaStream_t)stream>>>(partials, nparts, M, C, c_real, gamma, beta,
                                                                                running_mean, running_var, momentum, eps,
                                                                                mean_invstd, scale_shift);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

// Floats the `sums` / `dsums` scratch buffers of nbdt_bn_stats / nbdt_bn_backward must hold for C channels:
// reduced result [2][C] + gradient coefficients [3][C] + per-CTA partial sums.
long long nbdt_bn_scratch_floats(int C) { return (long long)(kReduceCtasPerSm * sm_count()) * 2 * C + 5LL * C; }

int nbdt_bn_eval_scale(int C, int c_real, const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, float eps, float* scale_shift, void* stream) {
  NBDT_REQUIRE(gamma && beta && running_mean && running_var && scale_shift, "NULL tensor");
  bn_eval_scale_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(C, c_real, gamma, beta, running_mean,
                                                                          running_var, eps, scale_shift);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

// y = act(x * scale + shift (+ residual)); x, residual, y: bf16 [M][C].
int nbdt_bn_apply(const void* x, const void* residual, void* y, long long M, int C, const float* scale_shift, int relu,
                  void* stream) {
  NBDT_REQUIRE(x && y && scale_shift, "NULL tensor");
  NBDT_REQUIRE(C % 8 == 0 && C <= 8192, "bad channel count %d", C);
  if (M == 0) return NBDT_OK;
  const int grid = rows_grid(M, C, 6);
  const uint4* xs = reinterpret_cast<const uint4*>(x);
  const uint4* rs = reinterpret_cast<const uint4*>(residual);
  uint4* ys = reinterpret_cast<uint4*>(y);
  const BnFromSums none = {};
  if (residual)
    bn_apply_kernel<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(xs, rs, ys, M, C, scale_shift, relu, none);
  else
    bn_apply_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(xs, rs, ys, M, C, scale_shift, relu, none);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

// (Experiment, see bn_backward_impl: not faster than nbdt_bn_stats + nbdt_bn_apply; nbdt/_b200/nn_ops.py uses it only with
// NBDT_B200_BN_FOLD_FINALIZE=1.)
// Training-mode BatchNorm forward in two launches instead of three: the statistics kernel adds every CTA's sums to one fp64
// accumulator (the first 4*C floats of `sums`, zeroed here), and the apply kernel derives mean / invstd / scale / shift from it in
// its prologue (every thread for its own eight channels), publishes them for the backward pass and moves the running statistics.
// Arguments: nbdt_bn_stats + nbdt_bn_apply.  (nbdt_bn_stats + nbdt_bn_apply remain for callers that need the coefficients first.)
int nbdt_bn_stats_apply(const void* x, const void* residual, void* y, long long M, int C, int c_real, const float* gamma,
                        const float* beta, float* running_mean, float* running_var, float momentum, float eps, float* sums,
                        float* mean_invstd, float* scale_shift, int relu, void* stream) {
  NBDT_REQUIRE(x && y && gamma && beta && sums && mean_invstd && scale_shift, "NULL tensor");
  NBDT_REQUIRE(C % 8 == 0 && C <= 8192 && M > 0, "bad shape M=%lld C=%d", M, C);
  NBDT_REQUIRE((reinterpret_cast<uintptr_t>(sums) & 7) == 0, "sums must be 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  double* acc = reinterpret_cast<double*>(sums);
  NBDT_CUDA_OK(cudaMemsetAsync(acc, 0, (size_t)2 * C * sizeof(double), st));
  const int grid = rows_grid(M, C, kReduceCtasPerSm);
  bn_stats_kernel<<<grid, kThreads, reduce_smem(C), st>>>(reinterpret_cast<const uint4*>(x), M, C, nullptr, acc, reverse_last(M));
  BnFromSums fs;
  fs.sums = acc; fs.gamma = gamma; fs.beta = beta; fs.running_mean = running_mean; fs.running_var = running_var;
  fs.mean_invstd = mean_invstd; fs.scale_shift = scale_shift; fs.M = M; fs.c_real = c_real; fs.momentum = momentum; fs.eps = eps;
  const int agrid = rows_grid(M, C, 6);
  const uint4* xs = reinterpret_cast<const uint4*>(x);
  const uint4* rs = reinterpret_cast<const uint4*>(residual);
  uint4* ys = reinterpret_cast<uint4*>(y);
  if (residual) bn_apply_kernel<true><<<agrid, kThreads, 0, st>>>(xs, rs, ys, M, C, nullptr, relu, fs);
  else bn_apply_kernel<false><<<agrid, kThreads, 0, st>>>(xs, rs, ys, M, C, nullptr, relu, fs);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

// Backward of y = act(bn(x) (+ residual)).  ReLU mask: `act` (the forward OUTPUT y) when given; else recomputed from
// x and `scale_shift` when given (only valid when no residual was added before the ReLU); else no activation.
// dx = d/dx (+ extra); dres (may be NULL) = gradient w.r.t. the residual; dsums[0..2C) = (dbeta, dgamma).
static int bn_backward_impl(int mask, const void* x, const void* act, const float* scale_shift, const void* dy,
                            const void* extra, void* dx, void* dres, long long M, int C, int c_real, const float* gamma,
                            const float* mean_invstd, float* dsums, void* stream) {
  NBDT_REQUIRE(x && dy && dx && gamma && mean_invstd && dsums, "NULL tensor");
  NBDT_REQUIRE(C % 8 == 0 && C <= 8192 && M > 0, "bad shape M=%lld C=%d", M, C);
  cudaStream_t st = (cudaStream_t)stream;
  float* coef = dsums + 2 * C;      // [3][C]
  float* partials = dsums + 5 * C;  // [grid][2][C]
  const uint4* xs = reinterpret_cast<const uint4*>(x);
  const uint4* as = reinterpret_cast<const uint4*>(act);
  const uint4* gs = reinterpret_cast<const uint4*>(dy);
  const uint4* es = reinterpret_cast<const uint4*>(extra);
  uint4* dxs = reinterpret_cast<uint4*>(dx);
  uint4* drs = reinterpret_cast<uint4*>(dres);
  const int grid = rows_grid(M, C, kReduceCtasPerSm);
  const size_t smem = reduce_smem(C);
  // NBDT_B200_BN_FOLD_FINALIZE=1: two launches -- the reduce kernel adds into one fp64 accumulator (first 4*C floats of the partials
  // area, zeroed here) and the apply kernel derives (k1, a, b) from it in its prologue and publishes dbeta / dgamma.  Measured
  // SLOWER than the partial rows + finalize launch (WRN-28-10 step 19.2 vs 18.6 ms in one run, profiles/r02_bn_fold_finalize_ab.md:
  // ~600 same-address fp64 atomics per channel serialise in L2 for as long as the finalize kernel takes), so it is off by default.
  static const bool finalize_kernel = getenv("NBDT_B200_BN_FOLD_FINALIZE") == nullptr;
  double* acc = finalize_kernel ? nullptr : reinterpret_cast<double*>(partials);
  NBDT_REQUIRE(finalize_kernel || (reinterpret_cast<uintptr_t>(partials) & 7) == 0, "dsums must be 8-byte aligned");
  if (acc) NBDT_CUDA_OK(cudaMemsetAsync(acc, 0, (size_t)2 * C * sizeof(double), st));
  if (mask == kMaskAct)
    bn_bwd_reduce_kernel<kMaskAct><<<grid, kThreads, smem, st>>>(xs, as, gs, M, C, mean_invstd, scale_shift, partials, acc, reverse_last(M));
  else if (mask == kMaskRecompute)
    bn_bwd_reduce_kernel<kMaskRecompute><<<grid, kThreads, smem, st>>>(xs, as, gs, M, C, mean_invstd, scale_shift, partials, acc, reverse_last(M));
  else if (mask == kMaskSiLU)
    bn_bwd_reduce_kernel<kMaskSiLU><<<grid, kThreads, smem, st>>>(xs, as, gs, M, C, mean_invstd, scale_shift, partials, acc, reverse_last(M));
  else
    bn_bwd_reduce_kernel<kMaskNone><<<grid, kThreads, smem, st>>>(xs, as, gs, M, C, mean_invstd, scale_shift, partials, acc, reverse_last(M));
  BnBwdFromSums fs = {};
  if (acc) { fs.sums = acc; fs.gamma = gamma; fs.mean_invstd = mean_invstd; fs.dsums = dsums; fs.c_real = c_real; }
  else bn_bwd_finalize_kernel<<<(C + 31) / 32, kFinalizeThreads, 0, st>>>(partials, grid, M, C, c_real, gamma, mean_invstd, dsums, coef);
  const int agrid = rows_grid(M, C, 4);
#define NBDT_BWD_APPLY(MASK, EXTRA, DRES) \
  bn_bwd_apply_kernel<MASK, EXTRA, DRES><<<agrid, kThreads, 0, st>>>(xs, as, gs, es, dxs, drs, M, C, coef, scale_shift, fs)
#define NBDT_BWD_APPLY_M(MASK)                       \
  do {                                               \
    if (extra && dres) NBDT_BWD_APPLY(MASK, true, true);   \
    else if (extra) NBDT_BWD_APPLY(MASK, true, false);     \
    else if (dres) NBDT_BWD_APPLY(MASK, false, true);      \
    else NBDT_BWD_APPLY(MASK, false, false);               \
  } while (0)
  if (mask == kMaskAct) NBDT_BWD_APPLY_M(kMaskAct);
  else if (mask == kMaskRecompute) NBDT_BWD_APPLY_M(kMaskRecompute);
  else if (mask == kMaskSiLU) NBDT_BWD_APPLY_M(kMaskSiLU);
  else NBDT_BWD_APPLY_M(kMaskNone);
#undef NBDT_BWD_APPLY_M
#undef NBDT_BWD_APPLY
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

int nbdt_bn_backward(const void* x, const void* act, const float* scale_shift, const void* dy, const void* extra,
                     void* dx, void* dres, long long M, int C, int c_real, const float* gamma,
                     const float* mean_invstd, float* dsums, void* stream) {
  const int mask = act ? kMaskAct : (scale_shift ? kMaskRecompute : kMaskNone);
  return bn_backward_impl(mask, x, act, scale_shift, dy, extra, dx, dres, M, C, c_real, gamma, mean_invstd, dsums, stream);
}

// Same for y = SiLU(bn(x)) (no residual before the activation): the activation gradient is recomputed from x and scale_shift.
int nbdt_bn_backward_silu(const void* x, const float* scale_shift, const void* dy, const void* extra, void* dx,
                          long long M, int C, int c_real, const float* gamma, const float* mean_invstd, float* dsums,
                          void* stream) {
  NBDT_REQUIRE(scale_shift != nullptr, "scale_shift is required for the SiLU backward");
  return bn_backward_impl(kMaskSiLU, x, nullptr, scale_shift, dy, extra, dx, nullptr, M, C, c_real, gamma, mean_invstd, dsums,
                          stream);
}

int nbdt_avgpool_fwd(const void* a, float* feat, int N, int HW, int C, void* stream) {
  NBDT_REQUIRE(a && feat && C % 8 == 0 && N <= 65535, "bad arguments");
  if (N == 0) return NBDT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // slices per image: enough CTAs to fill the GPU (~4 per SM), at least 64 pixels each
  const int want = std::max(1, (4 * sm_count() + N - 1) / N);
  const int slices = std::max(1, std::min(want, (HW + 63) / 64));
  const int pix_per_cta = (HW + slices - 1) / slices;
  const int gx = (HW + pix_per_cta - 1) / pix_per_cta;
  if (gx > 1) NBDT_CUDA_OK(cudaMemsetAsync(feat, 0, (size_t)N * C * sizeof(float), st));
  avgpool_fwd_kernel<<<dim3(gx, N), kThreads, 0, st>>>(reinterpret_cast<const uint4*>(a), feat, HW, C, pix_per_cta, gx > 1);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

int nbdt_avgpool_bwd(const float* dfeat, void* da, int N, int HW, int C, void* stream) {
  NBDT_REQUIRE(dfeat && da && C % 8 == 0, "bad arguments");
  const long long total = (long long)N * HW * (C / 8);
  if (total == 0) return NBDT_OK;
  avgpool_bwd_kernel<<<grid_for(total, sm_count()), kThreads, 0, (cudaStream_t)stream>>>(dfeat, reinterpret_cast<uint4*>(da),
                                                                                         total, HW, C);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

// 3-channel 3x3 stem (padding 1): x NCHW fp32 [N,3,H,W], w fp32 [cout][3][3][3], y NHWC bf16 [N,Ho,Wo,c_phys].
int nbdt_stem_fwd_strided(const float* x, const float* w, void* y, int N, int H, int W, int cout, int c_phys, int stride,
                          void* stream) {
  NBDT_REQUIRE(x && w && y, "NULL tensor");
  NBDT_REQUIRE(c_phys % 8 == 0 && c_phys >= cout && cout <= 256, "bad stem channels %d/%d", cout, c_phys);
  NBDT_REQUIRE(stride == 1 || stride == 2, "stem stride %d is not supported", stride);
  const long long pix = (long long)N * ((H + stride - 1) / stride) * ((W + stride - 1) / stride);
  if (pix == 0) return NBDT_OK;
  stem_fwd_kernel<<<(unsigned)((pix + 127) / 128), 128, 27 * cout * sizeof(float), (cudaStream_t)stream>>>(
      x, w, reinterpret_cast<__nv_bfloat16*>(y), N, H, W, cout, c_phys, stride);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}
int nbdt_stem_fwd(const float* x, const float* w, void* y, int N, int H, int W, int cout, int c_phys, void* stream) {
  return nbdt_stem_fwd_strided(x, w, y, N, H, W, cout, c_phys, 1, stream);
}

// dw fp32 [cout][3][3][3] (overwritten) from x NCHW fp32 and dy NHWC bf16 [N,H,W,c_phys].
int nbdt_stem_wgrad_strided(const float* x, const void* dy, float* dw, int N, int H, int W, int cout, int c_phys,
                            int stride, void* stream) {
  NBDT_REQUIRE(x && dy && dw, "NULL tensor");
  NBDT_REQUIRE(stride == 1 || stride == 2, "stem stride %d is not supported", stride);
  NBDT_REQUIRE((long long)N * H * W < (1LL << 31) - 64 * 4096, "stem: too many pixels (%lld)", (long long)N * H * W);
  cudaStream_t st = (cudaStream_t)stream;
  NBDT_CUDA_OK(cudaMemsetAsync(dw, 0, (size_t)cout * 27 * sizeof(float), st));
  const long long pix = (long long)N * ((H + stride - 1) / stride) * ((W + stride - 1) / stride);
  if (pix == 0) return NBDT_OK;
  NBDT_REQUIRE(c_phys % 4 == 0 && (reinterpret_cast<uintptr_t>(dy) & 7) == 0, "stem wgrad: dy must be 8-byte aligned");
  const int gx = (int)std::min<long long>((pix + 31) / 32, (long long)sm_count() * 4);
  stem_wgrad_kernel<<<dim3(gx, (cout + 15) / 16), dim3(32, 4), 0, st>>>(x, reinterpret_cast<const __nv_bfloat16*>(dy), dw, N,
                                                                         H, W, cout, c_phys, stride);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}
int nbdt_stem_wgrad(const float* x, const void* dy, float* dw, int N, int H, int W, int cout, int c_phys, void* stream) {
  return nbdt_stem_wgrad_strided(x, dy, dw, N, H, W, cout, c_phys, 1, stream);
}

// One fused SGD update over a flat fp32 parameter buffer (momentum, weight decay; gradient pre-scaled by grad_scale).
int nbdt_sgd_step(float* params, const float* grads, float* momentum_buf, long long n, float lr, float momentum,
                  float weight_decay, float grad_scale, int first_step, void* stream) {
  NBDT_REQUIRE(params && grads && momentum_buf, "NULL tensor");
  if (n == 0) return NBDT_OK;
  sgd_kernel<<<grid_for(n, sm_count()), kThreads, 0, (cudaStream_t)stream>>>(params, grads, momentum_buf, n, lr, momentum,
                                                                             weight_decay, grad_scale, first_step);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

}  // extern "C"
