// MBConv (EfficientNet) kernels around the tensor-core 1x1 convolutions: depthwise k x k convolution with the
// folded BatchNorm + SiLU fused in its epilogue, squeeze-excitation channel scaling.  NHWC bf16 activations,
// fp32 weights / scales.  All HBM-bound: a depthwise conv reads its input once (the k*k re-use is served by L1/L2)
// and writes its output once -- algorithmic bytes N*H*W*C*2 + N*Ho*Wo*C*2.
//
// Replaces, for the reference's EfficientNet backbones (`from torchvision.models import *` /
// `from pytorchcv.models.efficientnet import *`, nbdt/models/__init__.py:3-4), the depthwise nn.Conv2d (groups = C)
// + BatchNorm2d + SiLU of torchvision.models.efficientnet.MBConv and the `scale * input` of SqueezeExcitation.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "common.h"

namespace {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
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
__device__ __forceinline__ float activate(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return __fdividef(v, 1.f + __expf(-v));  // SiLU
  return v;
}

// Depthwise K x K convolution, stride S, padding K/2.  A thread owns one channel octet (8 channels, one 16-byte
// vector) of a strip of TW consecutive output pixels of one output row: per kernel row it loads the
// (TW-1)*S + K input vectors of that row once and slides the K taps over them (weights [K*K][C] fp32 through L1).
// Work items are ordered channel-octet fastest, so a warp reads 512 contiguous bytes per input pixel.
// FLIP: the taps are read in reverse order -- the stride-1 DATA GRADIENT is this same convolution of dy with the flipped filter
// (dx[hi,wi] = sum w[K-1-r'][K-1-s'] * dy[hi+r'-PAD, wi+s'-PAD]), so nbdt_dwconv_dgrad runs this kernel for stride 1.
template <int K, int S, int TW, bool FLIP = false>
__global__ void __launch_bounds__(256) dwconv_fwd_kernel(const uint4* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ scale_shift, uint4* __restrict__ y,
                                                         int N, int H, int W, int C, int Ho, int Wo, int act) {
  constexpr int IN = (TW - 1) * S + K;
  constexpr int PAD = K / 2;
  const unsigned octets = C / 8;
  const unsigned strips = (Wo + TW - 1) / TW;
  const unsigned total = (unsigned)N * Ho * strips * octets;
  for (unsigned item = blockIdx.x * blockDim.x + threadIdx.x; item < total; item += gridDim.x * blockDim.x) {
    const unsigned o = item % octets;
    unsigned t = item / octets;
    const unsigned strip = t % strips;
    t /= strips;
    const unsigned ho = t % (unsigned)Ho;
    const unsigned n = t / (unsigned)Ho;
    const int wo0 = strip * TW;
    const int wi0 = wo0 * S - PAD;
    float acc[TW][8];
#pragma unroll
    for (int i = 0; i < TW; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[i][e] = 0.f;
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int hi = (int)ho * S + r - PAD;
      if (hi < 0 || hi >= H) continue;
      const uint4* row = x + ((size_t)(n * H + hi) * W) * octets + o;
      uint4 in[IN];
#pragma unroll
      for (int j = 0; j < IN; ++j) {
        const int wi = wi0 + j;
        in[j] = (wi >= 0 && wi < W) ? __ldg(row + (size_t)wi * octets) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int s = 0; s < K; ++s) {
        const float4* wp = reinterpret_cast<const float4*>(w + (size_t)(FLIP ? K * K - 1 - (r * K + s) : r * K + s) * C + o * 8);
        const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < TW; ++i) {
          float f[8];
          unpack8(in[i * S + s], f);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[i][e] = fmaf(f[e], wv[e], acc[i][e]);
        }
      }
    }
    float sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sc[e] = scale_shift ? __ldg(scale_shift + o * 8 + e) : 1.f;
      sh[e] = scale_shift ? __ldg(scale_shift + C + o * 8 + e) : 0.f;
    }
    uint4* dst = y + ((size_t)(n * Ho + ho) * Wo) * octets + o;
#pragma unroll
    for (int i = 0; i < TW; ++i) {
      if (wo0 + i < Wo) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = activate(fmaf(acc[i][e], sc[e], sh[e]), act);
        dst[(size_t)(wo0 + i) * octets] = pack8(v);
      }
    }
  }
}

// y[n, p, c] = x[n, p, c] * s[n][c]   (squeeze-excitation).  blockIdx.y = image; thread -> (pixel row r, octet o).
// bytes: N*HW*C*2 read + written
__global__ void __launch_bounds__(256) se_scale_kernel(const uint4* __restrict__ x, const float* __restrict__ s,
                                                       uint4* __restrict__ y, int HW, int C) {
  const int octets = C / 8;
  const int n = blockIdx.y;
  const uint4* xi = x + (size_t)n * HW * octets;
  uint4* yi = y + (size_t)n * HW * octets;
  auto run = [&](int o, int p0, int pstride) {
    float sv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) sv[e] = __ldg(s + (size_t)n * C + o * 8 + e);
    for (int p = p0; p < HW; p += pstride) {
      float f[8];
      unpack8(__ldg(xi + (size_t)p * octets + o), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] *= sv[e];
      yi[(size_t)p * octets + o] = pack8(f);
    }
  };
  if (octets <= 256) {
    const int R = 256 / octets;
    const int r = threadIdx.x / octets, o = threadIdx.x - r * octets;
    if (r < R) run(o, blockIdx.x * R + r, gridDim.x * R);
  } else {
    for (int o = threadIdx.x; o < octets; o += 256) run(o, blockIdx.x, gridDim.x);
  }
}

// Data gradient of the depthwise convolution: dx[n,hi,wi,c] = sum over taps (r,s) with (hi + PAD - r) and (wi + PAD - s)
// divisible by S of w[r][s][c] * dy[n, (hi+PAD-r)/S, (wi+PAD-s)/S, c].  One thread per (input pixel, channel octet).
// bytes: N*Ho*Wo*C*2 read (the K*K re-use is served by L1/L2) + N*H*W*C*2 written
template <int K, int S>
__global__ void __launch_bounds__(256) dwconv_dgrad_kernel(const uint4* __restrict__ dy, const float* __restrict__ w,
                                                           uint4* __restrict__ dx, int N, int H, int W, int C, int Ho,
                                                           int Wo) {
  constexpr int PAD = K / 2;
  const unsigned octets = C / 8;
  const unsigned total = (unsigned)N * H * W * octets;
  for (unsigned item = blockIdx.x * blockDim.x + threadIdx.x; item < total; item += gridDim.x * blockDim.x) {
    const unsigned o = item % octets;
    unsigned t = item / octets;
    const int wi = t % (unsigned)W;
    t /= (unsigned)W;
    const int hi = t % (unsigned)H;
    const unsigned n = t / (unsigned)H;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int hn = hi + PAD - r;
      if (hn < 0 || (S == 2 && (hn & 1))) continue;
      const int ho = hn / S;
      if (ho >= Ho) continue;
#pragma unroll
      for (int s = 0; s < K; ++s) {
        const int wn = wi + PAD - s;
        if (wn < 0 || (S == 2 && (wn & 1))) continue;
        const int wo = wn / S;
        if (wo >= Wo) continue;
        float f[8];
        unpack8(__ldg(dy + ((size_t)(n * Ho + ho) * Wo + wo) * octets + o), f);
        const float4* wp = reinterpret_cast<const float4*>(w + (size_t)(r * K + s) * C + o * 8);
        const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(f[e], wv[e], acc[e]);
      }
    }
    dx[item] = pack8(acc);
  }
}

// Weight gradient of the depthwise convolution: dw[r*K+s][c] = sum over output pixels of dy[n,ho,wo,c] * x[n, ho*S+r-PAD,
// wo*S+s-PAD, c].  Thread (pixel lane, channel octet) keeps the K x 8 partial sums of ONE kernel row r = blockIdx.y in
// registers over its share of the output pixels; lanes are combined in shared memory, CTAs with fp32 atomics.
// bytes: (N*H*W + N*Ho*Wo)*C*2 read per kernel row (K passes; the operands of neighbouring rows overlap in L2)
template <int K, int S>
__global__ void __launch_bounds__(256) dwconv_wgrad_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy,
                                                           float* __restrict__ dw, int N, int H, int W, int C, int Ho,
                                                           int Wo) {
  constexpr int PAD = K / 2;
  __shared__ float red[256 * 8];
  const int octets = C / 8;
  const int r = blockIdx.y;  // kernel row
  const int R = octets <= 256 ? 256 / octets : 1;
  const unsigned total = (unsigned)N * Ho * Wo;
  for (int o = threadIdx.x % octets, lane = threadIdx.x / octets; o < octets; o += 256) {
    float acc[K][8];
#pragma unroll
    for (int s = 0; s < K; ++s)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[s][e] = 0.f;
    if (lane < R || octets > 256) {
      const unsigned step = octets <= 256 ? gridDim.x * R : gridDim.x;
      for (unsigned pix = blockIdx.x * (octets <= 256 ? R : 1) + (octets <= 256 ? lane : 0); pix < total; pix += step) {
        const int wo = pix % (unsigned)Wo;
        const unsigned t = pix / (unsigned)Wo;
        const int ho = t % (unsigned)Ho;
        const unsigned n = t / (unsigned)Ho;
        const int hi = ho * S + r - PAD;
        if (hi < 0 || hi >= H) continue;
        float g[8];
        unpack8(__ldg(dy + (size_t)pix * octets + o), g);
        const uint4* row = x + ((size_t)(n * H + hi) * W) * octets + o;
#pragma unroll
        for (int s = 0; s < K; ++s) {
          const int wi = wo * S + s - PAD;
          if (wi < 0 || wi >= W) continue;
          float f[8];
          unpack8(__ldg(row + (size_t)wi * octets), f);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[s][e] = fmaf(g[e], f[e], acc[s][e]);
        }
      }
    }
    // combine the R pixel lanes of this CTA (one tap column at a time through shared memory), then one atomic per value
#pragma unroll
    for (int s = 0; s < K; ++s) {
      __syncthreads();
      if (octets <= 256) {
        if (lane < R)
#pragma unroll
          for (int e = 0; e < 8; ++e) red[(lane * octets + o) * 8 + e] = acc[s][e];
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += 256) {
          float v = 0.f;
          for (int q = 0; q < R; ++q) v += red[q * C + c];
          atomicAdd(dw + (size_t)(r * K + s) * C + c, v);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(dw + (size_t)(r * K + s) * C + o * 8 + e, acc[s][e]);
      }
    }
    if (octets <= 256) break;
  }
}

// Weight gradient, ONE pass over x and dy (round 2; the kernel above makes K passes, one per kernel row, and ran at 7 % of the
// HBM peak in the EfficientNet-B0 train step, profiles/r02_bench_c5_*.json).  A thread owns CG channels and ALL K x K taps of them
// in registers (K*K*CG accumulators: 50 at K = 5 / CG = 2, 36 at K = 3 / CG = 4) and walks over work items = strips of TW
// consecutive output pixels of one output row: per kernel row it loads the (TW-1)*S + K input vectors once and slides the K taps
// over them against the strip's TW dy vectors -- (TW-1)*S + K loads for TW*K*CG fused multiply-adds.  Consecutive threads are
// consecutive channel groups of the same pixel (a warp reads 128 or 256 contiguous bytes per pixel).  The threads of a CTA that
// share a channel group are combined in shared memory, CTAs with fp32 atomics (dw zeroed by the caller).
// bytes: (N*H*W + N*Ho*Wo)*C*2 read once (neighbouring rows / strips overlap in L1 / L2)
template <int CG> struct DwVec;
template <> struct DwVec<2> { typedef uint32_t type; };
template <> struct DwVec<4> { typedef uint2 type; };
__device__ __forceinline__ void dw_unpack(uint32_t v, float (&f)[2]) {
  f[0] = __uint_as_float(v << 16);
  f[1] = __uint_as_float(v & 0xffff0000u);
}
__device__ __forceinline__ void dw_unpack(uint2 v, float (&f)[4]) {
  f[0] = __uint_as_float(v.x << 16);
  f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16);
  f[3] = __uint_as_float(v.y & 0xffff0000u);
}
__device__ __forceinline__ uint32_t dw_zero(uint32_t) { return 0u; }
__device__ __forceinline__ uint2 dw_zero(uint2) { return make_uint2(0u, 0u); }

template <int K, int S, int CG>
__global__ void __launch_bounds__(256) dwconv_wgrad1_kernel(const void* __restrict__ xv, const void* __restrict__ dyv,
                                                            float* __restrict__ dw, int N, int H, int W, int C, int Ho,
                                                            int Wo) {
  typedef typename DwVec<CG>::type V;
  constexpr int PAD = K / 2, TW = 4, IN = (TW - 1) * S + K;
  __shared__ float red[256 * CG];
  const V* __restrict__ x = reinterpret_cast<const V*>(xv);
  const V* __restrict__ dy = reinterpret_cast<const V*>(dyv);
  const int groups = C / CG;                              // channel groups of the tensor
  const int g0 = blockIdx.y * 256;                        // first group of this CTA
  const int G = min(groups - g0, 256);                    // groups of this CTA
  const int slots = 256 / G;                              // threads per group (>= 1)
  const int g = threadIdx.x % G, slot = threadIdx.x / G;
  const bool active = slot < slots;
  const int cg = g0 + g;
  const unsigned strips = (Wo + TW - 1) / TW;
  const unsigned total = (unsigned)N * Ho * strips;
  float acc[K * K][CG];
#pragma unroll
  for (int t = 0; t < K * K; ++t)
#pragma unroll
    for (int e = 0; e < CG; ++e) acc[t][e] = 0.f;
  if (active) {
    for (unsigned item = blockIdx.x * slots + slot; item < total; item += gridDim.x * slots) {
      const unsigned strip = item % strips;
      unsigned t = item / strips;
      const int ho = t % (unsigned)Ho;
      const unsigned n = t / (unsigned)Ho;
      const int wo0 = strip * TW;
      const int wi0 = wo0 * S - PAD;
      float gy[TW][CG];
      const V* grow = dy + ((size_t)(n * Ho + ho) * Wo) * groups + cg;
#pragma unroll
      for (int i = 0; i < TW; ++i) {
        const V v = (wo0 + i < Wo) ? __ldg(grow + (size_t)(wo0 + i) * groups) : dw_zero(V());
        dw_unpack(v, gy[i]);
      }
#pragma unroll
      for (int r = 0; r < K; ++r) {
        const int hi = ho * S + r - PAD;
        if (hi < 0 || hi >= H) continue;
        const V* row = x + ((size_t)(n * H + hi) * W) * groups + cg;
        V in[IN];
#pragma unroll
        for (int j = 0; j < IN; ++j) {
          const int wi = wi0 + j;
          in[j] = (wi >= 0 && wi < W) ? __ldg(row + (size_t)wi * groups) : dw_zero(V());
        }
#pragma unroll
        for (int j = 0; j < IN; ++j) {
          float f[CG];
          dw_unpack(in[j], f);
          // input vector j meets output pixel i at tap s = j - i*S
#pragma unroll
          for (int i = 0; i < TW; ++i) {
            const int sft = j - i * S;
            if (sft >= 0 && sft < K) {
#pragma unroll
              for (int e = 0; e < CG; ++e) acc[r * K + sft][e] = fmaf(gy[i][e], f[e], acc[r * K + sft][e]);
            }
          }
        }
      }
    }
  }
  // combine the threads of this CTA that share a channel group (one tap at a time through shared memory), then one atomic per value
#pragma unroll
  for (int t = 0; t < K * K; ++t) {
    __syncthreads();
    if (active)
#pragma unroll
      for (int e = 0; e < CG; ++e) red[(slot * G + g) * CG + e] = acc[t][e];
    __syncthreads();
    for (int c = threadIdx.x; c < G * CG; c += 256) {
      float v = 0.f;
      for (int q = 0; q < slots; ++q) v += red[q * G * CG + c];
      atomicAdd(dw + (size_t)t * C + g0 * CG + c, v);
    }
  }
}

// ds[n][c] = sum over the HW pixels of image n of x[n,p,c] * dy[n,p,c]  (gradient of the squeeze-excitation scale).
// Same slicing as avgpool_fwd_kernel (bn_ops.cu): CTA (slice, image); slices combined with fp32 atomics into a zeroed output.
__global__ void __launch_bounds__(256) se_dscale_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy,
                                                        float* __restrict__ ds, int HW, int C, int pix_per_cta) {
  __shared__ float red[256 * 8];
  const int octets = C / 8;
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
  const uint4* xi = x + (size_t)n * HW * octets;
  const uint4* gi = dy + (size_t)n * HW * octets;
  if (octets <= 256) {
    const int R = 256 / octets;
    const int r = threadIdx.x / octets, o = threadIdx.x - r * octets;
    float sacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (r < R) {
      for (int p = p0 + r; p < p1; p += R) {
        float f[8], g[8];
        unpack8(__ldg(xi + (size_t)p * octets + o), f);
        unpack8(__ldg(gi + (size_t)p * octets + o), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) sacc[e] = fmaf(f[e], g[e], sacc[e]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) red[(r * octets + o) * 8 + e] = sacc[e];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
      float t = 0.f;
      for (int q = 0; q < R; ++q) t += red[q * C + c];
      atomicAdd(ds + (size_t)n * C + c, t);
    }
  } else {
    for (int o = threadIdx.x; o < octets; o += 256) {
      float sacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int p = p0; p < p1; ++p) {
        float f[8], g[8];
        unpack8(__ldg(xi + (size_t)p * octets + o), f);
        unpack8(__ldg(gi + (size_t)p * octets + o), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) sacc[e] = fmaf(f[e], g[e], sacc[e]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(ds + (size_t)n * C + o * 8 + e, sacc[e]);
    }
  }
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

// Depthwise conv + folded BatchNorm + activation: y = act(dwconv(x; w) * scale + shift).
// x bf16 [N,H,W,C]; w fp32 [ksize*ksize][C] (tap-major); scale_shift fp32 [2][C] or NULL; act 0 none / 1 ReLU / 2 SiLU;
// y bf16 [N,Ho,Wo,C] with Ho = (H + stride - 1) / stride (padding ksize/2).
int nbdt_dwconv_fwd(const void* x, const float* w, const float* scale_shift, void* y, int N, int H, int W, int C,
                    int ksize, int stride, int act, void* stream) {
  NBDT_REQUIRE(x && w && y, "NULL tensor");
  NBDT_REQUIRE((ksize == 3 || ksize == 5) && (stride == 1 || stride == 2), "unsupported depthwise conv %dx%d stride %d",
               ksize, ksize, stride);
  NBDT_REQUIRE(C % 8 == 0 && N > 0 && H > 0 && W > 0, "bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  constexpr int TW = 4;
  const long long items = (long long)N * Ho * ((Wo + TW - 1) / TW) * (C / 8);
  NBDT_REQUIRE(items < (1LL << 31) - (1 << 24), "depthwise conv too large (%lld work items)", items);
  const int grid = (int)std::min<long long>((items + 255) / 256, (long long)sm_count() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  const uint4* xs = reinterpret_cast<const uint4*>(x);
  uint4* ys = reinterpret_cast<uint4*>(y);
#define NBDT_DW(K, S) dwconv_fwd_kernel<K, S, TW><<<grid, 256, 0, st>>>(xs, w, scale_shift, ys, N, H, W, C, Ho, Wo, act)
  if (ksize == 3 && stride == 1) NBDT_DW(3, 1);
  else if (ksize == 3) NBDT_DW(3, 2);
  else if (stride == 1) NBDT_DW(5, 1);
  else NBDT_DW(5, 2);
#undef NBDT_DW
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

// Squeeze-excitation scaling: y[n,p,c] = x[n,p,c] * s[n][c]; x, y bf16 [N][HW][C] (may alias), s fp32 [N][C].
int nbdt_se_scale(const void* x, const float* s, void* y, int N, int HW, int C, void* stream) {
  NBDT_REQUIRE(x && s && y, "NULL tensor");
  NBDT_REQUIRE(C % 8 == 0 && C <= 8192 && N > 0 && N <= 65535 && HW > 0, "bad shape N=%d HW=%d C=%d", N, HW, C);
  const int octets = C / 8;
  const int R = octets <= 256 ? 256 / octets : 1;
  const int gx = std::max(1, std::min((HW + R - 1) / R, std::max(1, sm_count() * 8 / N)));
  se_scale_kernel<<<dim3(gx, N), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(x), s,
                                                                  reinterpret_cast<uint4*>(y), HW, C);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

// Data gradient of nbdt_dwconv_fwd without the fused affine/activation: dx bf16 [N,H,W,C] from dy bf16 [N,Ho,Wo,C].
int nbdt_dwconv_dgrad(const void* dy, const float* w, void* dx, int N, int H, int W, int C, int ksize, int stride,
                      void* stream) {
  NBDT_REQUIRE(dy && w && dx, "NULL tensor");
  NBDT_REQUIRE((ksize == 3 || ksize == 5) && (stride == 1 || stride == 2), "unsupported depthwise conv %dx%d stride %d",
               ksize, ksize, stride);
  NBDT_REQUIRE(C % 8 == 0 && N > 0 && H > 0 && W > 0, "bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  const long long items = (long long)N * H * W * (C / 8);
  NBDT_REQUIRE(items < (1LL << 31) - (1 << 24), "depthwise conv too large (%lld work items)", items);
  cudaStream_t st = (cudaStream_t)stream;
  const uint4* gs = reinterpret_cast<const uint4*>(dy);
  uint4* xs = reinterpret_cast<uint4*>(dx);
  if (stride == 1) {  // the forward kernel on dy with the flipped filter: strips of 4 pixels share their input loads
    constexpr int TW = 4;
    const long long fitems = (long long)N * H * ((W + TW - 1) / TW) * (C / 8);
    const int fgrid = (int)std::min<long long>((fitems + 255) / 256, (long long)sm_count() * 16);
    if (ksize == 3) dwconv_fwd_kernel<3, 1, TW, true><<<fgrid, 256, 0, st>>>(gs, w, nullptr, xs, N, H, W, C, H, W, 0);
    else dwconv_fwd_kernel<5, 1, TW, true><<<fgrid, 256, 0, st>>>(gs, w, nullptr, xs, N, H, W, C, H, W, 0);
    NBDT_CUDA_OK(cudaGetLastError());
    return NBDT_OK;
  }
  const int grid = (int)std::min<long long>((items + 255) / 256, (long long)sm_count() * 16);
#define NBDT_DWD(K, S) dwconv_dgrad_kernel<K, S><<<grid, 256, 0, st>>>(gs, w, xs, N, H, W, C, Ho, Wo)
  if (ksize == 3 && stride == 1) NBDT_DWD(3, 1);
  else if (ksize == 3) NBDT_DWD(3, 2);
  else if (stride == 1) NBDT_DWD(5, 1);
  else NBDT_DWD(5, 2);
#undef NBDT_DWD
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

// Weight gradient: dw fp32 [ksize*ksize][C] (tap-major, overwritten) from x bf16 [N,H,W,C] and dy bf16 [N,Ho,Wo,C].
int nbdt_dwconv_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int C, int ksize, int stride,
                      void* stream) {
  NBDT_REQUIRE(x && dy && dw, "NULL tensor");
  NBDT_REQUIRE((ksize == 3 || ksize == 5) && (stride == 1 || stride == 2), "unsupported depthwise conv %dx%d stride %d",
               ksize, ksize, stride);
  NBDT_REQUIRE(C % 8 == 0 && C <= 2048 && N > 0 && H > 0 && W > 0, "bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  NBDT_REQUIRE((long long)N * Ho * Wo < (1LL << 31) - (1 << 24), "depthwise conv too large");
  cudaStream_t st = (cudaStream_t)stream;
  NBDT_CUDA_OK(cudaMemsetAsync(dw, 0, (size_t)ksize * ksize * C * sizeof(float), st));
  static const bool k_passes = getenv("NBDT_B200_DWWGRAD_KPASS") != nullptr;  // the round-1 kernel, for A/B measurements
  if (!k_passes) {
    const int cg = ksize == 3 ? 4 : 2;
    const int groups = C / cg;
    const int slots = std::max(1, 256 / std::min(groups, 256));
    const long long items = (long long)N * Ho * ((Wo + 3) / 4);
    const int gx = (int)std::max<long long>(1, std::min<long long>((items + slots - 1) / slots, (long long)sm_count() * 2));
    const dim3 grid(gx, (groups + 255) / 256);
#define NBDT_DWW1(K, S, CG) dwconv_wgrad1_kernel<K, S, CG><<<grid, 256, 0, st>>>(x, dy, dw, N, H, W, C, Ho, Wo)
    if (ksize == 3 && stride == 1) NBDT_DWW1(3, 1, 4);
    else if (ksize == 3) NBDT_DWW1(3, 2, 4);
    else if (stride == 1) NBDT_DWW1(5, 1, 2);
    else NBDT_DWW1(5, 2, 2);
#undef NBDT_DWW1
    NBDT_CUDA_OK(cudaGetLastError());
    return NBDT_OK;
  }
  const int octets = C / 8;
  const int R = octets <= 256 ? 256 / octets : 1;
  const long long pix = (long long)N * Ho * Wo;
  const int gx = (int)std::max<long long>(1, std::min<long long>((pix + R - 1) / R, (long long)sm_count() * 4 / ksize + 1));
  const uint4* xs = reinterpret_cast<const uint4*>(x);
  const uint4* gs = reinterpret_cast<const uint4*>(dy);
#define NBDT_DWW(K, S) dwconv_wgrad_kernel<K, S><<<dim3(gx, K), 256, 0, st>>>(xs, gs, dw, N, H, W, C, Ho, Wo)
  if (ksize == 3 && stride == 1) NBDT_DWW(3, 1);
  else if (ksize == 3) NBDT_DWW(3, 2);
  else if (stride == 1) NBDT_DWW(5, 1);
  else NBDT_DWW(5, 2);
#undef NBDT_DWW
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

// ds[n][c] = sum_p x[n,p,c] * dy[n,p,c]: gradient of nbdt_se_scale with respect to s.  x, dy bf16 [N][HW][C]; ds fp32 [N][C].
int nbdt_se_dscale(const void* x, const void* dy, float* ds, int N, int HW, int C, void* stream) {
  NBDT_REQUIRE(x && dy && ds, "NULL tensor");
  NBDT_REQUIRE(C % 8 == 0 && C <= 8192 && N > 0 && N <= 65535 && HW > 0, "bad shape N=%d HW=%d C=%d", N, HW, C);
  cudaStream_t st = (cudaStream_t)stream;
  NBDT_CUDA_OK(cudaMemsetAsync(ds, 0, (size_t)N * C * sizeof(float), st));
  const int want = std::max(1, (4 * sm_count() + N - 1) / N);
  const int slices = std::max(1, std::min(want, (HW + 63) / 64));
  const int pix_per_cta = (HW + slices - 1) / slices;
  const int gx = (HW + pix_per_cta - 1) / pix_per_cta;
  se_dscale_kernel<<<dim3(gx, N), 256, 0, st>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(dy), ds, HW, C,
                                                pix_per_cta);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

}  // extern "C"
