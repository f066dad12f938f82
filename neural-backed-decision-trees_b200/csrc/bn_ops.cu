// Memory-bound backbone kernels around the tensor-core convolutions (NHWC bf16 activations, fp32
// parameters and statistics): training-mode BatchNorm statistics / apply(+ReLU, +residual) / backward,
// global average pooling, the 3-channel stem convolution, fused SGD.
//
// Every kernel streams its tensors once with 16-byte vector accesses (8 bf16 channels per thread),
// keeps per-channel partial sums in registers, reduces them across the CTA in shared memory and
// finishes with one fp32 atomic per channel per CTA.  Roofline: HBM; algorithmic bytes are listed
// per kernel below (M = N*H*W rows, C channels).
//
// Replaces, in the reference's backbones, nn.BatchNorm2d + F.relu (+ residual add)
// (nbdt/models/resnet.py:50-54,70-73 and the pytorchcv pre-activation blocks), F.avg_pool2d
// (resnet.py:142) and optim.SGD (main.py:207).
#include <cuda_bf16.h>

#include <algorithm>

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

// Shared skeleton of the per-channel reductions: thread (r, o) owns channel octet o of rows
// r, r + R, ... of this CTA's grid-stride share; NACC float accumulators per channel.  `Op` provides
//   Pack load(row, o)            -- the global loads of one row (issued UNROLL rows at a time for memory-level parallelism)
//   void accum(pack, o, acc)     -- the arithmetic
// Each CTA writes its partial sums to out[blockIdx.x][NACC][C]; a tiny second kernel adds the partials in a fixed
// order (deterministic, and no same-address atomics: with one atomic per channel per CTA the reduction kernels
// took ~130 us regardless of tensor size, profiles/r01_bench_launches.md).
template <int NACC, int UNROLL, typename Op>
__device__ __forceinline__ void reduce_rows(long long M, int C, float* __restrict__ out, Op op) {
  extern __shared__ float red[];  // [R][NACC][C]  (only when R > 1)
  const int octets = C / 8;
  float acc[NACC][8];
  float* part = out + (long long)blockIdx.x * NACC * C;
  auto run = [&](long long row, long long stride, int o) {
    for (; row + (UNROLL - 1) * stride < M; row += UNROLL * stride) {
      decltype(op.load(row, o)) pk[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) pk[u] = op.load(row + u * stride, o);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) op.accum(pk[u], o, acc);
    }
    for (; row < M; row += stride) op.accum(op.load(row, o), o, acc);
  };
  if (octets <= kThreads) {
    const int R = kThreads / octets;
    const int r = threadIdx.x / octets, o = threadIdx.x % octets;
    const bool active = r < R;
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[a][e] = 0.f;
    if (active) run((long long)blockIdx.x * R + r, (long long)gridDim.x * R, o);
    if (R > 1) {
      if (active)
#pragma unroll
        for (int a = 0; a < NACC; ++a)
#pragma unroll
          for (int e = 0; e < 8; ++e) red[((r * NACC + a) * octets + o) * 8 + e] = acc[a][e];
      __syncthreads();
      if (r == 0) {
#pragma unroll
        for (int a = 0; a < NACC; ++a)
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float s = 0.f;
            for (int q = 0; q < R; ++q) s += red[((q * NACC + a) * octets + o) * 8 + e];
            part[(long long)a * C + o * 8 + e] = s;
          }
      }
    } else if (active) {
#pragma unroll
      for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int e = 0; e < 8; ++e) part[(long long)a * C + o * 8 + e] = acc[a][e];
    }
  } else {
    for (int o = threadIdx.x; o < octets; o += kThreads) {
#pragma unroll
      for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[a][e] = 0.f;
      run(blockIdx.x, gridDim.x, o);
#pragma unroll
      for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int e = 0; e < 8; ++e) part[(long long)a * C + o * 8 + e] = acc[a][e];
    }
  }
}

// out[i] = sum over the nparts per-CTA partials, i < n (fixed order).
__global__ void sum_partials_kernel(const float* __restrict__ partials, int nparts, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s[4] = {0.f, 0.f, 0.f, 0.f};  // independent chains: the loads are latency bound
  int b = 0;
  for (; b + 3 < nparts; b += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) s[u] += __ldg(partials + (long long)(b + u) * n + i);
  }
  for (; b < nparts; ++b) s[0] += __ldg(partials + (long long)b * n + i);
  out[i] = (s[0] + s[1]) + (s[2] + s[3]);
}

struct StatsOp {
  const uint4* x;
  int octets;
  __device__ __forceinline__ uint4 load(long long row, int o) const { return __ldg(x + row * octets + o); }
  __device__ __forceinline__ void accum(const uint4& v, int, float (&acc)[2][8]) const {
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[0][e] += f[e];
      acc[1][e] = fmaf(f[e], f[e], acc[1][e]);
    }
  }
};

// sums[0][c] = sum_rows x, sums[1][c] = sum_rows x^2.          bytes: M*C*2 read
__global__ void __launch_bounds__(kThreads) bn_stats_kernel(const uint4* __restrict__ x, long long M, int C,
                                                            float* __restrict__ sums) {
  reduce_rows<2, 8>(M, C, sums, StatsOp{x, C / 8});
}

// mean/invstd, affine scale/shift, running-stat update (momentum, unbiased variance) -- C threads of work.
__global__ void bn_finalize_kernel(const float* __restrict__ partials, int nparts, long long M, int C, int c_real,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                   float eps, float* __restrict__ mean_invstd /*[2][C]*/,
                                   float* __restrict__ scale_shift /*[2][C]*/) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (c >= c_real) {  // zero-padded channel: stays exactly zero
    mean_invstd[c] = 0.f; mean_invstd[C + c] = 0.f; scale_shift[c] = 0.f; scale_shift[C + c] = 0.f;
    return;
  }
  float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};  // independent chains (latency-bound loads)
  int b = 0;
  for (; b + 3 < nparts; b += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      a0[u] += __ldg(partials + (long long)(b + u) * 2 * C + c);
      a1[u] += __ldg(partials + (long long)(b + u) * 2 * C + C + c);
    }
  }
  for (; b < nparts; ++b) {
    a0[0] += __ldg(partials + (long long)b * 2 * C + c);
    a1[0] += __ldg(partials + (long long)b * 2 * C + C + c);
  }
  const float s0 = (a0[0] + a0[1]) + (a0[2] + a0[3]), s1 = (a1[0] + a1[1]) + (a1[2] + a1[3]);
  const double inv_m = 1.0 / (double)M;
  const double mean = (double)s0 * inv_m;
  double var = (double)s1 * inv_m - mean * mean;
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
__global__ void __launch_bounds__(kThreads) bn_apply_kernel(const uint4* __restrict__ x, const uint4* __restrict__ res,
                                                            uint4* __restrict__ y, long long total_octets, int C,
                                                            const float* __restrict__ scale_shift, int relu) {
  extern __shared__ float ss[];  // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) ss[i] = scale_shift[i];
  __syncthreads();
  const int octets = C / 8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  auto one = [&](long long i, const uint4& vx, const uint4& vr) {
    const int c0 = (int)(i % octets) * 8;
    float f[8], r[8];
    unpack8(vx, f);
    unpack8(vr, r);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = fmaf(f[e], ss[c0 + e], ss[C + c0 + e]);
      if (res) v += r[e];
      f[e] = relu ? fmaxf(v, 0.f) : v;
    }
    y[i] = pack8(f);
  };
  const uint4 zero = make_uint4(0, 0, 0, 0);
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < total_octets; i += 4 * stride) {
    uint4 vx[4], vr[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      vx[u] = __ldg(x + i + u * stride);
      vr[u] = res ? __ldg(res + i + u * stride) : zero;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) one(i + u * stride, vx[u], vr[u]);
  }
  for (; i < total_octets; i += stride) one(i, __ldg(x + i), res ? __ldg(res + i) : zero);
}

struct BwdPack {
  uint4 x, g, a;
};
struct BwdReduceOp {
  const uint4* x;
  const uint4* act;
  const uint4* dy;
  const float* mean_invstd;
  int octets, C;
  __device__ __forceinline__ BwdPack load(long long row, int o) const {
    BwdPack p;
    p.x = __ldg(x + row * octets + o);
    p.g = __ldg(dy + row * octets + o);
    p.a = act ? __ldg(act + row * octets + o) : make_uint4(0, 0, 0, 0);
    return p;
  }
  __device__ __forceinline__ void accum(const BwdPack& p, int o, float (&acc)[2][8]) const {
    float fx[8], fg[8], fa[8];
    unpack8(p.x, fx);
    unpack8(p.g, fg);
    unpack8(p.a, fa);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float g = (act && !(fa[e] > 0.f)) ? 0.f : fg[e];
      const float xhat = (fx[e] - __ldg(mean_invstd + o * 8 + e)) * __ldg(mean_invstd + C + o * 8 + e);
      acc[0][e] += g;
      acc[1][e] = fmaf(g, xhat, acc[1][e]);
    }
  }
};

// dsums[0][c] = sum g, dsums[1][c] = sum g * xhat, with g = dy * (act > 0) when the forward applied ReLU.
// bytes: M*C*2 * (2 or 3) read
__global__ void __launch_bounds__(kThreads) bn_bwd_reduce_kernel(const uint4* __restrict__ x, const uint4* __restrict__ act,
                                                                 const uint4* __restrict__ dy, long long M, int C,
                                                                 const float* __restrict__ mean_invstd,
                                                                 float* __restrict__ dsums) {
  reduce_rows<2, 4>(M, C, dsums, BwdReduceOp{x, act, dy, mean_invstd, C / 8, C});
}

// dx = gamma*invstd * (g - dbeta/M - xhat * dgamma/M) (+ extra);  optionally dres = g (masked dy).
// bytes: M*C*2 * (2..4) read + M*C*2 * (1..2) written
__global__ void __launch_bounds__(kThreads) bn_bwd_apply_kernel(const uint4* __restrict__ x, const uint4* __restrict__ act,
                                                                const uint4* __restrict__ dy, const uint4* __restrict__ extra,
                                                                uint4* __restrict__ dx, uint4* __restrict__ dres,
                                                                long long total_octets, long long M, int C, int c_real,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ mean_invstd,
                                                                const float* __restrict__ dsums) {
  extern __shared__ float sm[];  // [4][C]: mean, k1 = gamma*invstd, k2 = dbeta/M, k3 = invstd*dgamma/M ... see below
  const float inv_m = 1.f / (float)M;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float invstd = mean_invstd[C + c];
    const float g = c < c_real ? gamma[c] : 0.f;
    sm[c] = mean_invstd[c];
    sm[C + c] = invstd;
    sm[2 * C + c] = g * invstd;
    sm[3 * C + c] = dsums[c] * inv_m;
  }
  __syncthreads();
  float* dg_over_m = sm + 4 * C;  // [C]
  for (int c = threadIdx.x; c < C; c += blockDim.x) dg_over_m[c] = dsums[C + c] * inv_m;
  __syncthreads();
  const int octets = C / 8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  auto one = [&](long long i, const uint4& vx, const uint4& vg, const uint4& va, const uint4& ve) {
    const int c0 = (int)(i % octets) * 8;
    float fx[8], fg[8], fa[8], fe[8];
    unpack8(vx, fx);
    unpack8(vg, fg);
    unpack8(va, fa);
    unpack8(ve, fe);
    float o[8], gm[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = c0 + e;
      const float g = (act && !(fa[e] > 0.f)) ? 0.f : fg[e];
      const float xhat = (fx[e] - sm[c]) * sm[C + c];
      float v = sm[2 * C + c] * (g - sm[3 * C + c] - xhat * dg_over_m[c]);
      if (extra) v += fe[e];
      o[e] = v;
      gm[e] = g;
    }
    dx[i] = pack8(o);
    if (dres) dres[i] = pack8(gm);
  };
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < total_octets; i += 2 * stride) {
    uint4 vx[2], vg[2], va[2], ve[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long k = i + u * stride;
      vx[u] = __ldg(x + k);
      vg[u] = __ldg(dy + k);
      va[u] = act ? __ldg(act + k) : zero;
      ve[u] = extra ? __ldg(extra + k) : zero;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) one(i + u * stride, vx[u], vg[u], va[u], ve[u]);
  }
  for (; i < total_octets; i += stride)
    one(i, __ldg(x + i), __ldg(dy + i), act ? __ldg(act + i) : zero, extra ? __ldg(extra + i) : zero);
}

// feat[n][c] = mean over HW rows of a[n, :, c]  (fp32 out).      bytes: N*HW*C*2 read
__global__ void avgpool_fwd_kernel(const uint4* __restrict__ a, float* __restrict__ feat, int HW, int C) {
  const int octets = C / 8;
  const int n = blockIdx.x;
  for (int o = threadIdx.x; o < octets; o += blockDim.x) {
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = 0; p < HW; ++p) {
      float f[8];
      unpack8(__ldg(a + ((long long)n * HW + p) * octets + o), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) s[e] += f[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) feat[(long long)n * C + o * 8 + e] = s[e] / (float)HW;
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
                                                       int c_phys) {
  extern __shared__ float ws[];  // [27][cout]
  for (int i = threadIdx.x; i < 27 * cout; i += blockDim.x) {
    const int co = i % cout, k = i / cout;  // k = ci*9 + r*3 + s
    ws[i] = w[co * 27 + k];
  }
  __syncthreads();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)N * H * W) return;
  const int wq = (int)(pix % W), hq = (int)((pix / W) % H);
  const long long n = pix / ((long long)W * H);
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

// Stem weight gradient: dw[co][k] = sum_pixels dy[pix][co] * patch[pix][k];  CTA = tile of 64 pixels staged in
// shared memory, thread t owns outputs t, t + 256, ... of the cout*27; one atomic per output per CTA.
__global__ void __launch_bounds__(kThreads) stem_wgrad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                              float* __restrict__ dw, int N, int H, int W, int cout,
                                                              int c_phys) {
  constexpr int P = 64;
  extern __shared__ float sm[];  // patch[P][27] | g[P][cout]
  float* patch = sm;
  float* g = sm + P * 27;
  const long long total = (long long)N * H * W;
  const int nout = cout * 27;
  float acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0.f;
  for (long long p0 = (long long)blockIdx.x * P; p0 < total; p0 += (long long)gridDim.x * P) {
    __syncthreads();
    for (int i = threadIdx.x; i < P * 27; i += blockDim.x) {
      const int pp = i / 27, k = i % 27;
      const long long pix = p0 + pp;
      float v = 0.f;
      if (pix < total) {
        const int wq = (int)(pix % W), hq = (int)((pix / W) % H);
        const long long n = pix / ((long long)W * H);
        const int ci = k / 9, r = (k % 9) / 3, s = k % 3;
        const int hh = hq + r - 1, ww = wq + s - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = __ldg(x + ((n * 3 + ci) * H + hh) * W + ww);
      }
      patch[i] = v;
    }
    for (int i = threadIdx.x; i < P * cout; i += blockDim.x) {
      const int pp = i / cout, co = i % cout;
      const long long pix = p0 + pp;
      g[i] = pix < total ? __bfloat162float(dy[pix * c_phys + co]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int o = threadIdx.x + q * kThreads;
      if (o < nout) {
        const int co = o / 27, k = o % 27;
        float s = 0.f;
        for (int pp = 0; pp < P; ++pp) s = fmaf(g[pp * cout + co], patch[pp * 27 + k], s);
        acc[q] += s;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int o = threadIdx.x + q * kThreads;
    if (o < nout) atomicAdd(dw + o, acc[q]);
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

}  // namespace

extern "C" {

// Training-mode statistics of x[M][C] (bf16): fills mean_invstd[2][C], scale_shift[2][C], updates the running
// statistics (may be NULL).  `sums` is a [2][C] fp32 scratch buffer.  c_real <= C: channels beyond are padding.
int nbdt_bn_stats(const void* x, long long M, int C, int c_real, const float* gamma, const float* beta,
                  float* running_mean, float* running_var, float momentum, float eps, float* sums,
                  float* mean_invstd, float* scale_shift, void* stream) {
  NBDT_REQUIRE(x && gamma && beta && sums && mean_invstd && scale_shift, "NULL tensor");
  NBDT_REQUIRE(C % 8 == 0 && C <= 8192 && M > 0, "bad shape M=%lld C=%d", M, C);
  cudaStream_t st = (cudaStream_t)stream;
  const int octets = C / 8;
  const int R = std::max(1, kThreads / octets);
  const int grid = (int)std::max<long long>(1, std::min<long long>((M + R - 1) / R, (long long)sm_count() * 4));
  const size_t smem = R > 1 ? (size_t)R * 2 * C * sizeof(float) : 0;
  bn_stats_kernel<<<grid, kThreads, smem, st>>>(reinterpret_cast<const uint4*>(x), M, C, sums);
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, grid, M, C, c_real, gamma, beta, running_mean, running_var,
                                                      momentum, eps, mean_invstd, scale_shift);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

// Number of floats the `sums` / `dsums` scratch buffers of nbdt_bn_stats / nbdt_bn_backward must hold for C channels
// (per-CTA partial sums + the reduced result).
long long nbdt_bn_scratch_floats(int C) { return (long long)(4 * sm_count() + 1) * 2 * C; }

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
  NBDT_REQUIRE(C % 8 == 0 && C <= 4096, "bad channel count %d", C);
  const long long total = M * (C / 8);
  if (total == 0) return NBDT_OK;
  bn_apply_kernel<<<grid_for(total, sm_count()), kThreads, 2 * C * sizeof(float), (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(residual), reinterpret_cast<uint4*>(y), total, C,
      scale_shift, relu);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

// Backward of y = act(bn(x) (+ residual)).  act: the forward OUTPUT y when ReLU was applied, else NULL.
// dx = d/dx (+ extra); dres (may be NULL) = gradient w.r.t. the residual; dgamma_dbeta[2][C] = (dbeta, dgamma) sums.
int nbdt_bn_backward(const void* x, const void* act, const void* dy, const void* extra, void* dx, void* dres,
                     long long M, int C, int c_real, const float* gamma, const float* mean_invstd, float* dsums,
                     void* stream) {
  NBDT_REQUIRE(x && dy && dx && gamma && mean_invstd && dsums, "NULL tensor");
  NBDT_REQUIRE(C % 8 == 0 && C <= 2048 && M > 0, "bad shape M=%lld C=%d", M, C);
  cudaStream_t st = (cudaStream_t)stream;
  const int octets = C / 8;
  const int R = std::max(1, kThreads / octets);
  const int grid = (int)std::max<long long>(1, std::min<long long>((M + R - 1) / R, (long long)sm_count() * 2));
  const size_t smem = R > 1 ? (size_t)R * 2 * C * sizeof(float) : 0;
  float* partials = dsums + 2 * C;  // dsums: [2][C] result followed by the per-CTA partials
  bn_bwd_reduce_kernel<<<grid, kThreads, smem, st>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(act),
                                                     reinterpret_cast<const uint4*>(dy), M, C, mean_invstd, partials);
  sum_partials_kernel<<<(2 * C + 127) / 128, 128, 0, st>>>(partials, grid, 2 * C, dsums);
  const long long total = M * octets;
  bn_bwd_apply_kernel<<<grid_for(total, sm_count()), kThreads, 5 * C * sizeof(float), st>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(act), reinterpret_cast<const uint4*>(dy),
      reinterpret_cast<const uint4*>(extra), reinterpret_cast<uint4*>(dx), reinterpret_cast<uint4*>(dres), total, M, C,
      c_real, gamma, mean_invstd, dsums);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

int nbdt_avgpool_fwd(const void* a, float* feat, int N, int HW, int C, void* stream) {
  NBDT_REQUIRE(a && feat && C % 8 == 0, "bad arguments");
  if (N == 0) return NBDT_OK;
  avgpool_fwd_kernel<<<N, std::min(kThreads, std::max(32, C / 8)), 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(a), feat, HW, C);
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

// 3-channel 3x3 stride-1 stem: x NCHW fp32 [N,3,H,W], w fp32 [cout][3][3][3], y NHWC bf16 [N,H,W,c_phys].
int nbdt_stem_fwd(const float* x, const float* w, void* y, int N, int H, int W, int cout, int c_phys, void* stream) {
  NBDT_REQUIRE(x && w && y, "NULL tensor");
  NBDT_REQUIRE(c_phys % 8 == 0 && c_phys >= cout && cout <= 256, "bad stem channels %d/%d", cout, c_phys);
  const long long pix = (long long)N * H * W;
  if (pix == 0) return NBDT_OK;
  stem_fwd_kernel<<<(unsigned)((pix + 127) / 128), 128, 27 * cout * sizeof(float), (cudaStream_t)stream>>>(
      x, w, reinterpret_cast<__nv_bfloat16*>(y), N, H, W, cout, c_phys);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

// dw fp32 [cout][3][3][3] (overwritten) from x NCHW fp32 and dy NHWC bf16 [N,H,W,c_phys].
int nbdt_stem_wgrad(const float* x, const void* dy, float* dw, int N, int H, int W, int cout, int c_phys, void* stream) {
  NBDT_REQUIRE(x && dy && dw, "NULL tensor");
  NBDT_REQUIRE(cout * 27 <= 8 * kThreads, "stem with %d output channels is not supported", cout);
  cudaStream_t st = (cudaStream_t)stream;
  NBDT_CUDA_OK(cudaMemsetAsync(dw, 0, (size_t)cout * 27 * sizeof(float), st));
  const long long pix = (long long)N * H * W;
  if (pix == 0) return NBDT_OK;
  const int grid = (int)std::min<long long>((pix + 63) / 64, (long long)sm_count() * 4);
  stem_wgrad_kernel<<<grid, kThreads, (64 * 27 + 64 * cout) * sizeof(float), st>>>(
      x, reinterpret_cast<const __nv_bfloat16*>(dy), dw, N, H, W, cout, c_phys);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
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
