// The callers either side of the hot path, as small sm_100a kernels behind the C ABI:
//
//   nbdt_topk_count        top-k accuracy counters kept ON THE DEVICE (reference nbdt/metrics.py:16-20 runs torch.topk and a
//                          per-sample Python loop with a host sync every batch; nbdt/analysis.py:224-229 feeds it the rules'
//                          output): one warp per row ranks the target, two 64-bit counters accumulate, the host reads them
//                          only in report().
//   nbdt_ward_linkage      Ward agglomerative clustering of the C x D classifier rows -- the induced hierarchy of
//                          nbdt/graph.py:386-448 (sklearn AgglomerativeClustering(linkage="ward")) -- on the device: fp64 pairwise
//                          squared distances, then one CTA merges the globally closest pair C-1 times with Lance-Williams
//                          updates and a nearest-neighbour array (O(C) per merge).  SoftTreeLoss re-induces the hierarchy from the
//                          live classifier during training (nbdt/loss.py:303-315); this keeps that step off the host.
//   nbdt_augment_crop_flip the CIFAR / TinyImagenet training transform (RandomCrop(pad) + RandomHorizontalFlip + ToTensor +
//                          Normalize, nbdt/data/cifar.py:10-21) for a whole batch of uint8 images resident in HBM, instead of two
//                          DataLoader worker processes (main.py:136-141).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "common.h"

namespace {

// ------------------------------------------------------------------------------------------------------------------
// top-k counters
// ------------------------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<uint16_t>(const uint16_t* p) { return __uint_as_float((unsigned)*p << 16); }  // bf16
struct half_tag { uint16_t v; };
template <> __device__ __forceinline__ float ldf<half_tag>(const half_tag* p) {
  float r;
  asm("{ .reg .b16 h; mov.b16 h, %1; cvt.f32.f16 %0, h; }" : "=f"(r) : "h"(p->v));
  return r;
}

template <typename T>
__global__ void __launch_bounds__(256) topk_count_kernel(const T* __restrict__ x, const long long* __restrict__ y, long long B, int C,
                                                         long long ld, int k, unsigned long long* counters) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned hits = 0, rows = 0;
  for (long long row = (long long)blockIdx.x * 8 + warp; row < B; row += (long long)gridDim.x * 8) {
    const long long t = y[row];
    if (t < 0 || t >= C) continue;  // ignored rows (nn.CrossEntropyLoss's ignore_index) count for nothing
    const T* xr = x + row * ld;
    const float xt = ldf<T>(xr + t);
    int rank = 0;  // entries that torch.topk would list before the target: larger, or equal with a smaller index
    for (int c = lane; c < C; c += 32) {
      const float v = ldf<T>(xr + c);
      rank += (v > xt) || (v == xt && c < (int)t);
    }
    for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
    hits += rank < k;
    rows += 1;
  }
  __shared__ unsigned sh[8], sr[8];
  if (lane == 0) { sh[warp] = hits; sr[warp] = rows; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned h = 0, r = 0;
    for (int w = 0; w < 8; ++w) { h += sh[w]; r += sr[w]; }
    if (r) { atomicAdd(counters, (unsigned long long)h); atomicAdd(counters + 1, (unsigned long long)r); }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Ward linkage
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ward_dist_kernel(const float* __restrict__ x, int C, int D, double* __restrict__ d2) {
  const int i = blockIdx.x;
  for (int j = threadIdx.x; j < C; j += blockDim.x) {
    double acc = 0.0;
    for (int k = 0; k < D; ++k) {
      const double d = (double)x[(size_t)i * D + k] - (double)x[(size_t)j * D + k];
      acc += d * d;
    }
    d2[(size_t)i * C + j] = acc;
  }
}

struct MinPair { double d; int i; };
__device__ __forceinline__ MinPair min_pair(MinPair a, MinPair b) {  // smaller distance, then smaller index
  return (b.d < a.d || (b.d == a.d && b.i < a.i)) ? b : a;
}
__device__ MinPair block_min(MinPair v, MinPair* scratch) {
  for (int o = 16; o > 0; o >>= 1) {
    MinPair u;
    u.d = __shfl_xor_sync(0xffffffffu, v.d, o);
    u.i = __shfl_xor_sync(0xffffffffu, v.i, o);
    v = min_pair(v, u);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    MinPair u = lane < (int)(blockDim.x >> 5) ? scratch[lane] : MinPair{1e300, 0x7fffffff};
    for (int o = 16; o > 0; o >>= 1) {
      MinPair w;
      w.d = __shfl_xor_sync(0xffffffffu, u.d, o);
      w.i = __shfl_xor_sync(0xffffffffu, u.i, o);
      u = min_pair(u, w);
    }
    if (lane == 0) scratch[0] = u;
  }
  __syncthreads();
  return scratch[0];
}

// One CTA.  Slot i holds a live cluster while size[i] > 0; merging (i, j), i < j, keeps slot i.  nn[i] / nnd[i]: nearest live
// slot of slot i and its (squared, Lance-Williams) distance.  children[step] = ids of the merged clusters (smaller first),
// id = slot for the original points, C + step for the cluster created at `step`: scipy.cluster.hierarchy's numbering.
__global__ void __launch_bounds__(1024) ward_merge_kernel(double* __restrict__ d2, int C, int* __restrict__ children,
                                                           double* __restrict__ heights, int* __restrict__ work) {
  __shared__ MinPair scratch[32];
  __shared__ int rescans[1024];
  __shared__ int nrescan;
  int* size = work;            // [C]
  int* id = work + C;          // [C]
  int* nn = work + 2 * C;      // [C]
  double* nnd = reinterpret_cast<double*>(work + 4 * C);  // [C], 8-byte aligned when C is even; the launcher pads
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < C; i += nt) { size[i] = 1; id[i] = i; }
  __syncthreads();
  auto rescan_row = [&](int r) {  // all threads: nearest live neighbour of slot r
    MinPair best{1e300, 0x7fffffff};
    for (int k = tid; k < C; k += nt)
      if (k != r && size[k] > 0) best = min_pair(best, MinPair{d2[(size_t)r * C + k], k});
    best = block_min(best, scratch);
    if (tid == 0) { nn[r] = best.i; nnd[r] = best.d; }
    __syncthreads();
  };
  for (int r = 0; r < C; ++r) rescan_row(r);
  for (int step = 0; step < C - 1; ++step) {
    MinPair best{1e300, 0x7fffffff};
    for (int i = tid; i < C; i += nt)
      if (size[i] > 0) best = min_pair(best, MinPair{nnd[i], i});
    best = block_min(best, scratch);
    int i = best.i, j = nn[i];
    if (j < i) { const int t = i; i = j; j = t; }
    const double dij = d2[(size_t)i * C + j];
    const int ni = size[i], nj = size[j];
    __syncthreads();
    if (tid == 0) {
      const int a = id[i], b = id[j];
      children[2 * step] = a < b ? a : b;
      children[2 * step + 1] = a < b ? b : a;
      heights[step] = sqrt(dij);
      nrescan = 0;
    }
    __syncthreads();
    // Lance-Williams (Ward, squared distances) into row / column i; slot j dies
    for (int k = tid; k < C; k += nt) {
      if (k == i || k == j || size[k] <= 0) continue;
      const int nk = size[k];
      const double dn = ((double)(ni + nk) * d2[(size_t)k * C + i] + (double)(nj + nk) * d2[(size_t)k * C + j] - (double)nk * dij) /
                        (double)(ni + nj + nk);
      d2[(size_t)k * C + i] = dn;
      d2[(size_t)i * C + k] = dn;
      if (nn[k] == i || nn[k] == j) {
        rescans[atomicAdd(&nrescan, 1)] = k;          // its nearest neighbour changed: full rescan
      } else if (dn < nnd[k] || (dn == nnd[k] && i < nn[k])) {
        nn[k] = i;
        nnd[k] = dn;
      }
    }
    __syncthreads();
    if (tid == 0) { size[i] = ni + nj; size[j] = 0; id[i] = C + step; }
    __syncthreads();
    if (step == C - 2) break;
    const int nr = nrescan;
    for (int q = 0; q < nr; ++q) rescan_row(rescans[q]);
    rescan_row(i);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// batch augmentation
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) augment_kernel(const uint8_t* __restrict__ src, const int* __restrict__ params, float* __restrict__ dst,
                                                      int N, int H, int W, int pad, float m0, float m1, float m2, float s0, float s1, float s2) {
  const long long total = (long long)N * H * W;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % W), i = (int)((e / W) % H), n = (int)(e / ((long long)W * H));
    const int dx = params[3 * n], dy = params[3 * n + 1], flip = params[3 * n + 2];
    const int jj = flip ? W - 1 - j : j;           // RandomHorizontalFlip acts on the cropped image
    const int si = i + dy - pad, sj = jj + dx - pad;  // RandomCrop(H, padding=pad): window offset (dy, dx) in the zero-padded image
    float r = 0.f, g = 0.f, b = 0.f;
    if (si >= 0 && si < H && sj >= 0 && sj < W) {
      const uint8_t* px = src + (((long long)n * H + si) * W + sj) * 3;
      r = px[0] * (1.f / 255.f); g = px[1] * (1.f / 255.f); b = px[2] * (1.f / 255.f);
    }
    const long long plane = (long long)H * W, o = (long long)n * 3 * plane + (long long)i * W + j;
    dst[o] = (r - m0) / s0;
    dst[o + plane] = (g - m1) / s1;
    dst[o + 2 * plane] = (b - m2) / s2;
  }
}

}  // namespace

extern "C" {

int nbdt_topk_count(const void* scores, int dtype, const int64_t* targets, int64_t B, int64_t C, int64_t ld, int k,
                    unsigned long long* counters, void* stream) {
  NBDT_REQUIRE(B >= 0 && C >= 1 && ld >= C && k >= 1, "bad shape (B=%lld, C=%lld, ld=%lld, k=%d)", (long long)B, (long long)C, (long long)ld, k);
  NBDT_REQUIRE(counters != nullptr, "counters is NULL");
  if (B == 0) return NBDT_OK;
  NBDT_REQUIRE(scores && targets, "NULL argument");
  const unsigned grid = (unsigned)std::min<long long>((B + 7) / 8, 148 * 8);
  const long long* y = reinterpret_cast<const long long*>(targets);
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case NBDT_F32: topk_count_kernel<float><<<grid, 256, 0, s>>>((const float*)scores, y, B, (int)C, ld, k, counters); break;
    case NBDT_BF16: topk_count_kernel<uint16_t><<<grid, 256, 0, s>>>((const uint16_t*)scores, y, B, (int)C, ld, k, counters); break;
    case NBDT_F16: topk_count_kernel<half_tag><<<grid, 256, 0, s>>>((const half_tag*)scores, y, B, (int)C, ld, k, counters); break;
    default: return nbdt::set_error(NBDT_EINVAL, "unknown dtype %d", dtype);
  }
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

long long nbdt_ward_workspace_bytes(int C) {
  if (C < 2) return 0;
  return (long long)C * C * 8 + ((long long)4 * C + 2) * 4 + (long long)C * 8 + 64;
}

int nbdt_ward_linkage(const float* centers, int C, int D, int32_t* children, double* heights, void* workspace, void* stream) {
  NBDT_REQUIRE(C >= 2 && D >= 1, "need at least two points (C=%d, D=%d)", C, D);
  NBDT_REQUIRE(C <= 1024, "at most 1024 points (got %d): the merge kernel keeps its rescan list in shared memory", C);
  NBDT_REQUIRE(centers && children && heights && workspace, "NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  double* d2 = reinterpret_cast<double*>(workspace);
  // size / id / nn (3C ints, padded to 4C) then nnd (C doubles): keep the doubles 8-byte aligned
  int* work = reinterpret_cast<int*>(d2 + (size_t)C * C);
  ward_dist_kernel<<<C, 256, 0, s>>>(centers, C, D, d2);
  NBDT_CUDA_OK(cudaGetLastError());
  ward_merge_kernel<<<1, 1024, 0, s>>>(d2, C, children, heights, work);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

int nbdt_augment_crop_flip(const void* src_u8_nhwc, const int32_t* params, float* dst_nchw, int N, int H, int W, int pad,
                           const float* mean3, const float* std3, void* stream) {
  NBDT_REQUIRE(N >= 0 && H >= 1 && W >= 1 && pad >= 0, "bad shape");
  if (N == 0) return NBDT_OK;
  NBDT_REQUIRE(src_u8_nhwc && params && dst_nchw && mean3 && std3, "NULL argument");
  const long long total = (long long)N * H * W;
  const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 16);
  augment_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)src_u8_nhwc, params, dst_nchw, N, H, W, pad, mean3[0], mean3[1], mean3[2],
                                                          std3[0], std3[1], std3[2]);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

}  // extern "C"
