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
template <int K, int S, int TW>
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
        const float4* wp = reinterpret_cast<const float4*>(w + (size_t)(r * K + s) * C + o * 8);
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

}  // extern "C"
