// Implicit-GEMM convolution on 5th-generation tensor cores (tcgen05 + TMEM), operands fed by TMA.
//
// Activations are NHWC bf16.  A convolution (forward, or data-gradient) is a sum over filter taps of
// plain GEMMs between a spatially SHIFTED view of the input and one [Cout x Cin] slice of the packed
// weights:
//
//     out[n, i, j, :] = sum_t  W_t * in_{map(t)}[n, i + dh_t, j + dw_t, :]          ("tap GEMM")
//
// The shifted view is never materialised (no im2col buffer): a 4-D TMA tensor map (C, W, H, N)
// loads the box {64 channels x box_w x box_h x box_n} = 128 pixels x 64 channels at the shifted
// coordinates, out-of-bounds pixels (the padding halo) are zero-filled by the TMA unit, and the box
// lands in shared memory exactly as the K-major, 128B-swizzled A tile the UMMA descriptor expects.
// Stride-2 convolutions use up to four tensor maps over the (row, column) parity classes of the
// input, which turns every tap into a unit-stride shifted view again; the stride-2 data-gradient
// is the same kernel run once per OUTPUT parity class with a strided output view.
//
// CTA layout (192 threads, persistent, one CTA per SM):
//   warp 0    TMA producer: per (tap, 64-channel k-block) one A box + one B (weights) box into a
//             kStages-deep shared-memory ring, completion signalled on `full` mbarriers
//   warp 1    MMA issuer: one elected thread issues the tcgen05.mma's of a stage (K = 16 each),
//             tcgen05.commit releases the stage (`empty`) and finally publishes the accumulator
//   warps 2-5 epilogue: tcgen05.ld the fp32 accumulator (TMEM, double buffered so the next tile's MMAs
//             overlap), optional residual add, convert to bf16, vectorised global stores
//
// Three forward / data-gradient kernels share that layout (launch_tap_gemm picks one):
//   tap_gemm_kernel    one CTA per 128-pixel tile, M = 128 (small problems)
//   tap_gemm2_kernel   a 2-CTA cluster per pair of tiles, tcgen05.mma.cta_group::2 with M = 256: each CTA stages
//                      half of the weight tile
//   tap_gemm2h_kernel  the same plus an H-halo activation tile that serves the three taps of a filter column
//                      (stride-1 3x3 on maps 16 / 32 pixels wide)
// and wgrad_kernel computes the weight gradient (pixel index = K, MN-major operands, halo tiles for the taps of a
// filter row / column, team scheduling for operands larger than L2).
//
// Roofline: tensor pipe by FLOPs (2 * N*Ho*Wo * Cout * Cin * taps per launch); in practice the L2 -> shared-memory
// fill rate (~12-14 TB/s chip-wide) bounds the 160-channel layers: every variant above exists to move fewer bytes per MMA
// (profiles/r01_conv_ncu.md, r01_conv_ncu_v2.md).
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "sm100.cuh"

namespace {

using namespace sm100;

constexpr int kBlockM = 128;  // output pixels per tile (UMMA M)
constexpr int kBlockK = 64;   // bf16 channels per k-block = one 128-byte swizzle row
constexpr int kMaxTaps = 9;

struct TapGemmParams {
  int num_taps, kblocks, last_ksteps;  // k-blocks of 64 channels; K=16 steps in the last (partial) one
  int n_tiles, tiles_w, tiles_h, tiles_n;
  int box_w, box_h, box_n;
  int N_img, H_out, W_out, Cout;
  long long out_sn, out_sh, out_sw;  // output strides in elements (channel stride 1)
  __nv_bfloat16* out;
  const __nv_bfloat16* residual;  // same geometry as out, or nullptr
  float* stats;                   // nullptr, or [gridDim.x][2][Cout]: every CTA's sum and sum of squares per output channel of the
                                  // bf16 values it stored (BatchNorm batch statistics of the consumer, bn_ops.cu finalizes them)
  int tap_map[kMaxTaps];
  int tap_dh[kMaxTaps];
  int tap_dw[kMaxTaps];
  int tap_wrow[kMaxTaps];  // first row of this tap's [Cout x Cin] slice in the packed weight matrix
  // H-halo pair kernel (tap_gemm2h_kernel): the 9 taps as 3 column groups (same dw) of 3 taps (dh = -1, 0, +1 in any order)
  int grp_dw[3];
  int grp_dh[3][3];
  int grp_wrow[3][3];
};

// BatchNorm statistics in the epilogue.  An epilogue warp holds 32 output rows x 32 channels of every chunk it stores: a
// butterfly transpose-reduce (5 rounds, 31 shuffles per statistic) leaves channel j's sum and sum of squares over the 32 rows
// in lane j, which adds them to the CTA's accumulators in shared memory; at the end of the kernel every CTA writes ONE row of
// partial sums (no global atomics, deterministic finalize in bn_ops.cu).  The statistics are those of the bf16 values that
// are stored, i.e. exactly what the separate pass over the tensor (bn_stats_kernel) would have read again.
constexpr int kStatsMaxC = 1024;                     // output channels the shared accumulators cover
constexpr int kStatsBytes = 2 * kStatsMaxC * 4;      // appended to every tap GEMM's dynamic shared memory

__device__ __forceinline__ void stats_accumulate(const uint32_t (&o)[16], int lane, float* s_sum, float* s_sq) {
  float a[32], b[32];
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    a[2 * e] = __uint_as_float(o[e] << 16);
    a[2 * e + 1] = __uint_as_float(o[e] & 0xffff0000u);
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) b[i] = a[i] * a[i];
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;  // this lane keeps the upper half of what is left, its partner the lower half
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float sa = up ? a[i] : a[i + off], ka = up ? a[i + off] : a[i];
      const float sb = up ? b[i] : b[i + off], kb = up ? b[i + off] : b[i];
      a[i] = ka + __shfl_xor_sync(0xffffffffu, sa, off);
      b[i] = kb + __shfl_xor_sync(0xffffffffu, sb, off);
    }
  }
  atomicAdd(s_sum + lane, a[0]);
  atomicAdd(s_sq + lane, b[0]);
}

// One pipeline stage holds kUnits consecutive (tap, 64-channel k-block) units: the MMA thread's per-stage
// handshake (barrier probe, fence, commit ~ 160 cycles + ~32 cycles per tcgen05.mma issue, measured with
// scripts/ubench_mma.cu) then sits under 8 MMAs = 640 tensor-pipe cycles instead of 4 = 320, which it exceeded.
template <int BLOCK_N>
struct TapGemmCfg {
  static constexpr int kUnits = 2;
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kUnitBytes = kABytes + kBBytes;
  static constexpr int kStageBytes = kUnits * kUnitBytes;
  static constexpr int kStages = (216 * 1024) / kStageBytes > 6 ? 6 : (216 * 1024) / kStageBytes;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/ + kStatsBytes;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(192, 1)
tap_gemm_kernel(const __grid_constant__ CUtensorMap tm_in0, const __grid_constant__ CUtensorMap tm_in1,
                const __grid_constant__ CUtensorMap tm_in2, const __grid_constant__ CUtensorMap tm_in3,
                const __grid_constant__ CUtensorMap tm_w, const TapGemmParams p) {
  using Cfg = TapGemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full = bars;                         // [kStages]
  uint64_t* empty = bars + Cfg::kStages;         // [kStages]
  uint64_t* tmem_full = bars + 2 * Cfg::kStages; // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_stats = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes + 256);  // [2][Cout] when p.stats
  if (p.stats)
    for (int i = threadIdx.x; i < 2 * p.Cout; i += blockDim.x) s_stats[i] = 0.f;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
  const int units = p.num_taps * p.kblocks;  // (tap, k-block) units per tile, tap-major

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_in0);
    prefetch_tmap(&tm_w);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full[a]), 1);
      mbar_init(smem_u32(&tmem_empty[a]), 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_base_slot), Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp runs the loop (warp-uniform control flow keeps descriptors / coordinates in uniform
    // registers); only the elected lane issues.  A loop under `if (lane == 0)` costs one R2UR per operand
    // per instruction and made TMA / MMA issue the bottleneck (profiles/r01_conv_issue_experiments.md).
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      int m = tile / p.n_tiles;
      const int w0 = (m % p.tiles_w) * p.box_w;
      m /= p.tiles_w;
      const int h0 = (m % p.tiles_h) * p.box_h;
      const int n0 = (m / p.tiles_h) * p.box_n;
      int t = 0, kb = 0;
      for (int u = 0; u < units; u += Cfg::kUnits) {
        const int nu = min(Cfg::kUnits, units - u);
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        const uint32_t bar = full0 + 8 * stage;
        const uint32_t dst = smem_base + stage * Cfg::kStageBytes;
        if (leader) mbar_expect_tx(bar, nu * Cfg::kUnitBytes);
#pragma unroll
        for (int j = 0; j < Cfg::kUnits; ++j) {
          if (j < nu) {
            const int map = p.tap_map[t];
            const CUtensorMap* tm = map == 0 ? &tm_in0 : (map == 1 ? &tm_in1 : (map == 2 ? &tm_in2 : &tm_in3));
            if (leader) {
              tma_load_4d(dst + j * Cfg::kUnitBytes, tm, bar, kb * kBlockK, w0 + p.tap_dw[t], h0 + p.tap_dh[t], n0);
              tma_load_2d(dst + j * Cfg::kUnitBytes + Cfg::kABytes, &tm_w, bar, kb * kBlockK,
                          p.tap_wrow[t] + n_tile * BLOCK_N);
            }
            if (++kb == p.kblocks) { kb = 0; ++t; }
          }
        }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, elected lane issues) =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, 0, 0);
    // descriptors differ between stages / units / k-steps only in the 14-bit start-address field (low word)
    const uint64_t desc0 = make_smem_desc(smem_base, 16, 1024);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
      int kb = 0;
      uint32_t accumulate = 0;
      for (int u = 0; u < units; u += Cfg::kUnits) {
        const int nu = min(Cfg::kUnits, units - u);
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        const uint64_t sdesc = desc0 + (uint64_t)((stage * Cfg::kStageBytes) >> 4);
#pragma unroll
        for (int j = 0; j < Cfg::kUnits; ++j) {
          if (j < nu) {
            // channels past Cin in a tap's last k-block are TMA zero fill; do not spend MMAs on them
            const int ksteps = (kb == p.kblocks - 1) ? p.last_ksteps : kBlockK / 16;
            const uint64_t adesc = sdesc + (uint64_t)((j * Cfg::kUnitBytes) >> 4);
            const uint64_t bdesc = adesc + (uint64_t)(Cfg::kABytes >> 4);
            if (leader) {
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                // +32 bytes (= 2 in descriptor units) per K=16 step inside the 128-byte swizzle row
                if (k < ksteps) umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, accumulate);
                accumulate = 1;
              }
            }
            accumulate = 1;
            if (++kb == p.kblocks) kb = 0;
          }
        }
        if (leader) umma_commit(empty0 + 8 * stage);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(smem_u32(&tmem_full[acc]));
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are this warp's
    const int row = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      int m = tile / p.n_tiles;
      const int w0 = (m % p.tiles_w) * p.box_w;
      m /= p.tiles_w;
      const int h0 = (m % p.tiles_h) * p.box_h;
      const int n0 = (m / p.tiles_h) * p.box_n;
      // row of the tile -> pixel (TMA box order: w fastest, then h, then n)
      const int pw = w0 + row % p.box_w;
      const int ph = h0 + (row / p.box_w) % p.box_h;
      const int pn = n0 + row / (p.box_w * p.box_h);
      const bool in_range = pw < p.W_out && ph < p.H_out && pn < p.N_img;
      const long long off = pn * p.out_sn + ph * p.out_sh + pw * p.out_sw + (long long)n_tile * BLOCK_N;

      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c0, v);
        tmem_ld_wait();
        uint32_t ow[16];  // the chunk as stored (bf16 pairs); zeros for rows outside the tensor
#pragma unroll
        for (int e = 0; e < 16; ++e) ow[e] = 0u;
        if (in_range) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + off + c0);
          const uint4* res = p.residual ? reinterpret_cast<const uint4*>(p.residual + off + c0) : nullptr;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]);
            if (res) {
              const uint4 r = res[q];  // plain load: residual may alias out (in-place accumulate)
              const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&rw[e]);
                f[2 * e] += __bfloat162float(b.x);
                f[2 * e + 1] += __bfloat162float(b.y);
              }
            }
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
              o[e] = *reinterpret_cast<const uint32_t*>(&b);
            }
            dst[q] = make_uint4(o[0], o[1], o[2], o[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) ow[q * 4 + e] = o[e];
          }
        }
        if (p.stats) {
          float* acc_c = s_stats + n_tile * BLOCK_N + c0;
          stats_accumulate(ow, lane, acc_c, acc_c + p.Cout);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.stats) {
    float* row = p.stats + (size_t)blockIdx.x * 2 * p.Cout;
    for (int i = threadIdx.x; i < 2 * p.Cout; i += blockDim.x) row[i] = s_stats[i];
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant: two CTAs (one cluster, one TPC) work on two neighbouring 128-pixel tiles with ONE
// tcgen05.mma.cta_group::2 of M = 256 per K step.  Each CTA stages its own A tile and only HALF of the weight tile
// (the tensor core reads B's N columns half from each CTA's shared memory), so a (tap, k-block) unit costs
// 16 KB + BLOCK_N/2 x 128 B of TMA fill per SM instead of 16 KB + BLOCK_N x 128 B: 26 KB instead of 36 KB at
// BLOCK_N = 160.  ncu had the single-CTA kernel at 59-71 % l1tex / 14-20 TB/s of L2->SM traffic with the tensor pipe
// only 41 % busy on the 160-channel layers (profiles/r01_conv_ncu.md): shared-memory fill was the limiter.
//
// Protocol: the leader (even) CTA's elected thread issues every MMA.  "Stage full" lives in the leader: both CTAs' TMA
// loads complete_tx on it (expect_tx = both halves).  "Stage empty" and "accumulator full" are multicast by
// tcgen05.commit to both CTAs; "accumulator empty" is counted in the leader (4 epilogue warps x 2 CTAs).
// ---------------------------------------------------------------------------------------------
template <int BLOCK_N>
struct TapGemm2Cfg {
  static constexpr int kUnits = 2;
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = (BLOCK_N / 2) * kBlockK * 2;  // this CTA's half of the weight tile
  static constexpr int kUnitBytes = kABytes + kBBytes;
  static constexpr int kStageBytes = kUnits * kUnitBytes;
  static constexpr int kStages = (216 * 1024) / kStageBytes > 6 ? 6 : (216 * 1024) / kStageBytes;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/ + kStatsBytes;
  static_assert(kUnitBytes % 1024 == 0 && kABytes % 1024 == 0, "swizzled tiles must stay 1024-byte aligned");
};

template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
tap_gemm2_kernel(const __grid_constant__ CUtensorMap tm_in0, const __grid_constant__ CUtensorMap tm_in1,
                 const __grid_constant__ CUtensorMap tm_in2, const __grid_constant__ CUtensorMap tm_in3,
                 const __grid_constant__ CUtensorMap tm_w, const TapGemmParams p) {
  using Cfg = TapGemm2Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full = bars;                         // [kStages]  (used in the leader CTA only)
  uint64_t* empty = bars + Cfg::kStages;         // [kStages]
  uint64_t* tmem_full = bars + 2 * Cfg::kStages; // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]        (used in the leader CTA only)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_stats = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes + 256);  // [2][Cout] when p.stats
  if (p.stats)
    for (int i = threadIdx.x; i < 2 * p.Cout; i += blockDim.x) s_stats[i] = 0.f;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();       // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int m_pairs = (m_tiles + 1) >> 1;
  const int total_pairs = m_pairs * p.n_tiles;   // pair tile -> (m pair, n tile), n fastest
  const int units = p.num_taps * p.kblocks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_in0);
    prefetch_tmap(&tm_w);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full[a]), 1);
      mbar_init(smem_u32(&tmem_empty[a]), 8);  // 4 epilogue warps of each CTA
    }
    fence_barrier_init();
  }
  cluster_sync();  // barriers of both CTAs are initialised before anyone signals across the pair
  if (warp == 1) {
    tmem2_alloc(smem_u32(tmem_base_slot), Cfg::kTmemCols);
    tmem2_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);

  // this CTA's 128-pixel tile of pair tile `pt`
  auto tile_origin = [&](int pt, int& n_tile, int& w0, int& h0, int& n0) {
    n_tile = pt % p.n_tiles;
    int m = (pt / p.n_tiles) * 2 + (int)rank;
    w0 = (m % p.tiles_w) * p.box_w;
    m /= p.tiles_w;
    h0 = (m % p.tiles_h) * p.box_h;
    n0 = (m / p.tiles_h) * p.box_n;  // >= N_img for the odd tile out: TMA zero fill, no stores
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    const bool leader_lane = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int pt = pair; pt < total_pairs; pt += num_pairs) {
      int n_tile, w0, h0, n0;
      tile_origin(pt, n_tile, w0, h0, n0);
      int t = 0, kb = 0;
      for (int u = 0; u < units; u += Cfg::kUnits) {
        const int nu = min(Cfg::kUnits, units - u);
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        const uint32_t bar = full0 + 8 * stage;  // same offset in the leader
        const uint32_t dst = smem_base + stage * Cfg::kStageBytes;
        if (leader_lane && rank == 0) mbar_expect_tx(bar, 2 * nu * Cfg::kUnitBytes);
#pragma unroll
        for (int j = 0; j < Cfg::kUnits; ++j) {
          if (j < nu) {
            const int map = p.tap_map[t];
            const CUtensorMap* tm = map == 0 ? &tm_in0 : (map == 1 ? &tm_in1 : (map == 2 ? &tm_in2 : &tm_in3));
            if (leader_lane) {
              tma2_load_4d(dst + j * Cfg::kUnitBytes, tm, bar, kb * kBlockK, w0 + p.tap_dw[t], h0 + p.tap_dh[t], n0);
              tma2_load_2d(dst + j * Cfg::kUnitBytes + Cfg::kABytes, &tm_w, bar, kb * kBlockK,
                           p.tap_wrow[t] + n_tile * BLOCK_N + (int)rank * (BLOCK_N / 2));
            }
            if (++kb == p.kblocks) { kb = 0; ++t; }
          }
        }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===================== MMA issuer (leader CTA only) =====================
      const bool leader_lane = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16(2 * kBlockM, BLOCK_N, 0, 0);
      const uint64_t desc0 = make_smem_desc(smem_base, 16, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int pt = pair; pt < total_pairs; pt += num_pairs) {
        mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        int kb = 0;
        uint32_t accumulate = 0;
        for (int u = 0; u < units; u += Cfg::kUnits) {
          const int nu = min(Cfg::kUnits, units - u);
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint64_t sdesc = desc0 + (uint64_t)((stage * Cfg::kStageBytes) >> 4);
#pragma unroll
          for (int j = 0; j < Cfg::kUnits; ++j) {
            if (j < nu) {
              const int ksteps = (kb == p.kblocks - 1) ? p.last_ksteps : kBlockK / 16;
              const uint64_t adesc = sdesc + (uint64_t)((j * Cfg::kUnitBytes) >> 4);
              const uint64_t bdesc = adesc + (uint64_t)(Cfg::kABytes >> 4);
              if (leader_lane) {
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                  if (k < ksteps) umma2_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, accumulate);
                  accumulate = 1;
                }
              }
              accumulate = 1;
              if (++kb == p.kblocks) kb = 0;
            }
          }
          if (leader_lane) umma2_commit(empty0 + 8 * stage);  // frees the stage in both CTAs
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if (leader_lane) umma2_commit(smem_u32(&tmem_full[acc]));  // accumulator ready in both CTAs
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 of both CTAs; each CTA holds its own 128 rows) =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int pt = pair; pt < total_pairs; pt += num_pairs) {
      int n_tile, w0, h0, n0;
      tile_origin(pt, n_tile, w0, h0, n0);
      const int pw = w0 + row % p.box_w;
      const int ph = h0 + (row / p.box_w) % p.box_h;
      const int pn = n0 + row / (p.box_w * p.box_h);
      const bool in_range = pw < p.W_out && ph < p.H_out && pn < p.N_img;
      const long long off = pn * p.out_sn + ph * p.out_sh + pw * p.out_sw + (long long)n_tile * BLOCK_N;

      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c0, v);
        tmem_ld_wait();
        uint32_t ow[16];  // the chunk as stored (bf16 pairs); zeros for rows outside the tensor
#pragma unroll
        for (int e = 0; e < 16; ++e) ow[e] = 0u;
        if (in_range) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + off + c0);
          const uint4* res = p.residual ? reinterpret_cast<const uint4*>(p.residual + off + c0) : nullptr;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]);
            if (res) {
              const uint4 r = res[q];  // plain load: residual may alias out (in-place accumulate)
              const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&rw[e]);
                f[2 * e] += __bfloat162float(b.x);
                f[2 * e + 1] += __bfloat162float(b.y);
              }
            }
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
              o[e] = *reinterpret_cast<const uint32_t*>(&b);
            }
            dst[q] = make_uint4(o[0], o[1], o[2], o[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) ow[q * 4 + e] = o[e];
          }
        }
        if (p.stats) {
          float* acc_c = s_stats + n_tile * BLOCK_N + c0;
          stats_accumulate(ow, lane, acc_c, acc_c + p.Cout);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(smem_u32(&tmem_empty[acc]));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  cluster_sync();  // no CTA may exit (or free TMEM) while its peer still reads its shared memory / signals its barriers
  if (p.stats) {
    float* row = p.stats + (size_t)blockIdx.x * 2 * p.Cout;
    for (int i = threadIdx.x; i < 2 * p.Cout; i += blockDim.x) row[i] = s_stats[i];
  }
  if (warp == 1) {
    tc_fence_after();
    tmem2_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair kernel with H-halo activation tiles (stride-1 3x3 convolutions on maps 16 or 32 pixels wide).
// The three taps of a filter COLUMN (same dw, dh = -1 / 0 / +1) read the same pixels shifted by whole image rows, and a
// 128-pixel tile is box_w x box_h pixels of ONE image: a tile with one halo row above and below
// ({64 ch x box_w x (box_h + 2)} box, zero-filled outside the image) holds all three views as row ranges starting
// (1 + dh) * box_w pixel rows into it -- for a K-major operand a pixel is one 128-byte row, and the tensor core
// un-swizzles by absolute address, so the A descriptor simply starts there.  One stage = one (column, k-block) unit:
// 24 KB of activations + 3 half weight tiles (30 KB at BLOCK_N = 160) for 12 MMAs = 57 B per tensor-pipe cycle per SM,
// against 83 B/cycle for tap_gemm2_kernel (which ran at the 12 TB/s L2->SM limit, not at the tensor pipe's).
// ---------------------------------------------------------------------------------------------
template <int BLOCK_N>
struct TapGemm2hCfg {
  static constexpr int kAHaloBytes = 192 * 128;                // 32 x (4+2) or 16 x (8+2) pixel rows of 64 channels
  static constexpr int kBBytes = (BLOCK_N / 2) * kBlockK * 2;  // this CTA's half of one tap's weight tile
  static constexpr int kStageBytes = kAHaloBytes + 3 * kBBytes;
  static constexpr int kStages = (216 * 1024) / kStageBytes > 6 ? 6 : (216 * 1024) / kStageBytes;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + kStatsBytes;
  static_assert(kStageBytes % 1024 == 0, "swizzled tiles must stay 1024-byte aligned");
};

template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
tap_gemm2h_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w,
                  const TapGemmParams p) {
  using Cfg = TapGemm2hCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full = bars;                         // [kStages]  (leader CTA only)
  uint64_t* empty = bars + Cfg::kStages;         // [kStages]
  uint64_t* tmem_full = bars + 2 * Cfg::kStages; // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]        (leader CTA only)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_stats = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes + 256);  // [2][Cout] when p.stats
  if (p.stats)
    for (int i = threadIdx.x; i < 2 * p.Cout; i += blockDim.x) s_stats[i] = 0.f;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int m_pairs = (m_tiles + 1) >> 1;
  const int total_pairs = m_pairs * p.n_tiles;
  const int units = 3 * p.kblocks;               // (column group, k-block), group-major
  const uint32_t a_tx = (uint32_t)(p.box_w * (p.box_h + 2) * 128);  // bytes of one halo box

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_in);
    prefetch_tmap(&tm_w);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full[a]), 1);
      mbar_init(smem_u32(&tmem_empty[a]), 8);
    }
    fence_barrier_init();
  }
  cluster_sync();
  if (warp == 1) {
    tmem2_alloc(smem_u32(tmem_base_slot), Cfg::kTmemCols);
    tmem2_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);

  auto tile_origin = [&](int pt, int& n_tile, int& w0, int& h0, int& n0) {
    n_tile = pt % p.n_tiles;
    int m = (pt / p.n_tiles) * 2 + (int)rank;
    w0 = (m % p.tiles_w) * p.box_w;
    m /= p.tiles_w;
    h0 = (m % p.tiles_h) * p.box_h;
    n0 = m / p.tiles_h;  // box_n == 1; >= N_img for the odd tile out: TMA zero fill, no stores
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    const bool leader_lane = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int pt = pair; pt < total_pairs; pt += num_pairs) {
      int n_tile, w0, h0, n0;
      tile_origin(pt, n_tile, w0, h0, n0);
      int g = 0, kb = 0;
      for (int u = 0; u < units; ++u) {
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        const uint32_t bar = full0 + 8 * stage;
        const uint32_t dst = smem_base + stage * Cfg::kStageBytes;
        if (leader_lane && rank == 0) mbar_expect_tx(bar, 2 * (a_tx + 3 * Cfg::kBBytes));
        if (leader_lane) {
          tma2_load_4d(dst, &tm_in, bar, kb * kBlockK, w0 + p.grp_dw[g], h0 - 1, n0);
#pragma unroll
          for (int r = 0; r < 3; ++r)
            tma2_load_2d(dst + Cfg::kAHaloBytes + r * Cfg::kBBytes, &tm_w, bar, kb * kBlockK,
                         p.grp_wrow[g][r] + n_tile * BLOCK_N + (int)rank * (BLOCK_N / 2));
        }
        if (++kb == p.kblocks) { kb = 0; ++g; }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===================== MMA issuer (leader CTA only) =====================
      const bool leader_lane = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16(2 * kBlockM, BLOCK_N, 0, 0);
      const uint64_t desc0 = make_smem_desc(smem_base, 16, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int pt = pair; pt < total_pairs; pt += num_pairs) {
        mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        int g = 0, kb = 0;
        uint32_t accumulate = 0;
        for (int u = 0; u < units; ++u) {
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint64_t sdesc = desc0 + (uint64_t)((stage * Cfg::kStageBytes) >> 4);
          const int ksteps = (kb == p.kblocks - 1) ? p.last_ksteps : kBlockK / 16;
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            // tap (dh, dw_g): its 128 pixel rows start (1 + dh) * box_w rows into the halo tile (128 B = 8 descriptor units per row)
            const uint64_t adesc = sdesc + (uint64_t)((1 + p.grp_dh[g][r]) * p.box_w * 8);
            const uint64_t bdesc = sdesc + (uint64_t)((Cfg::kAHaloBytes + r * Cfg::kBBytes) >> 4);
            if (leader_lane) {
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                if (k < ksteps) umma2_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, accumulate);
                accumulate = 1;
              }
            }
            accumulate = 1;
          }
          if (leader_lane) umma2_commit(empty0 + 8 * stage);
          if (++kb == p.kblocks) { kb = 0; ++g; }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if (leader_lane) umma2_commit(smem_u32(&tmem_full[acc]));
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 of both CTAs) =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int pt = pair; pt < total_pairs; pt += num_pairs) {
      int n_tile, w0, h0, n0;
      tile_origin(pt, n_tile, w0, h0, n0);
      const int pw = w0 + row % p.box_w;
      const int ph = h0 + row / p.box_w;
      const int pn = n0;
      const bool in_range = pw < p.W_out && ph < p.H_out && pn < p.N_img;
      const long long off = pn * p.out_sn + ph * p.out_sh + pw * p.out_sw + (long long)n_tile * BLOCK_N;

      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c0, v);
        tmem_ld_wait();
        uint32_t ow[16];  // the chunk as stored (bf16 pairs); zeros for rows outside the tensor
#pragma unroll
        for (int e = 0; e < 16; ++e) ow[e] = 0u;
        if (in_range) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + off + c0);
          const uint4* res = p.residual ? reinterpret_cast<const uint4*>(p.residual + off + c0) : nullptr;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]);
            if (res) {
              const uint4 rv = res[q];  // plain load: residual may alias out (in-place accumulate)
              const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&rw[e]);
                f[2 * e] += __bfloat162float(b.x);
                f[2 * e + 1] += __bfloat162float(b.y);
              }
            }
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
              o[e] = *reinterpret_cast<const uint32_t*>(&b);
            }
            dst[q] = make_uint4(o[0], o[1], o[2], o[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) ow[q * 4 + e] = o[e];
          }
        }
        if (p.stats) {
          float* acc_c = s_stats + n_tile * BLOCK_N + c0;
          stats_accumulate(ow, lane, acc_c, acc_c + p.Cout);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(smem_u32(&tmem_empty[acc]));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  cluster_sync();
  if (p.stats) {
    float* row = p.stats + (size_t)blockIdx.x * 2 * p.Cout;
    for (int i = threadIdx.x; i < 2 * p.Cout; i += blockDim.x) row[i] = s_stats[i];
  }
  if (warp == 1) {
    tc_fence_after();
    tmem2_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host: tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// bf16 tensor map with 128-byte swizzle and zero fill; dims/strides innermost first, strides in bytes
int make_tmap(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
              const cuuint32_t* box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return nbdt::set_error(NBDT_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint32_t elem_strides[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                  strides_bytes, box, elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return nbdt::set_error(NBDT_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu ...)",
                           (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1]);
  return NBDT_OK;
}

int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

struct View4 {  // an NHWC view: element (n,h,w,c) at base + n*sn + h*sh + w*sw + c
  const __nv_bfloat16* base;
  int N, H, W, C;
  long long sn, sh, sw;
};

int view_tmap(CUtensorMap* tm, const View4& v, int box_w, int box_h, int box_n) {
  cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
  cuuint64_t strides[3] = {(cuuint64_t)v.sw * 2, (cuuint64_t)v.sh * 2, (cuuint64_t)v.sn * 2};
  cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n};
  return make_tmap(tm, v.base, 4, dims, strides, box);
}

struct Tap {
  int map, dh, dw, wrow;
};

// Launch the tap GEMM: out view (strided) = sum over taps of W_t * in_view[tap.map] shifted.
// stats / stats_parts: optional BatchNorm partial sums of the output ([CTAs][2][Cout] floats) and the number of CTA rows written.
int launch_tap_gemm(const View4* in_views, int num_views, const Tap* taps, int num_taps, const __nv_bfloat16* wpack,
                    int w_rows, int Cin, const View4& out, const __nv_bfloat16* residual, cudaStream_t stream,
                    float* stats = nullptr, int* stats_parts = nullptr) {
  NBDT_REQUIRE(num_taps >= 1 && num_taps <= kMaxTaps, "bad tap count %d", num_taps);
  NBDT_REQUIRE(num_views >= 1 && num_views <= 4, "bad view count %d", num_views);
  NBDT_REQUIRE(Cin % 16 == 0, "input channels (%d) must be a multiple of 16 (pad the activation)", Cin);
  const int Cout = out.C;
  int block_n = 0;
  if (Cout % 160 == 0) block_n = 160;
  else if (Cout % 128 == 0) block_n = 128;
  else if (Cout % 64 == 0) block_n = 64;
  NBDT_REQUIRE(block_n != 0, "output channels (%d) must be a multiple of 64", Cout);
  NBDT_REQUIRE(out.sw % 8 == 0 && out.sh % 8 == 0 && out.sn % 8 == 0 && (reinterpret_cast<uintptr_t>(out.base) & 15) == 0,
               "output view must be 16-byte aligned");

  TapGemmParams p = {};
  p.box_w = std::min(pow2_ceil(out.W), kBlockM);
  p.box_h = std::min(pow2_ceil(out.H), kBlockM / p.box_w);
  p.box_n = kBlockM / (p.box_w * p.box_h);
  p.tiles_w = (out.W + p.box_w - 1) / p.box_w;
  p.tiles_h = (out.H + p.box_h - 1) / p.box_h;
  p.tiles_n = (out.N + p.box_n - 1) / p.box_n;
  p.n_tiles = Cout / block_n;
  p.num_taps = num_taps;
  p.kblocks = (Cin + kBlockK - 1) / kBlockK;
  p.last_ksteps = (Cin - (p.kblocks - 1) * kBlockK) / 16;
  p.N_img = out.N; p.H_out = out.H; p.W_out = out.W; p.Cout = Cout;
  p.out_sn = out.sn; p.out_sh = out.sh; p.out_sw = out.sw;
  p.out = const_cast<__nv_bfloat16*>(out.base);
  p.residual = residual;
  p.stats = stats;
  NBDT_REQUIRE(!stats || Cout <= kStatsMaxC, "fused BatchNorm statistics cover at most %d output channels (got %d)", kStatsMaxC, Cout);
  for (int t = 0; t < num_taps; ++t) {
    p.tap_map[t] = taps[t].map; p.tap_dh[t] = taps[t].dh; p.tap_dw[t] = taps[t].dw; p.tap_wrow[t] = taps[t].wrow;
  }
  CUtensorMap tm_in[4], tm_w;
  for (int v = 0; v < 4; ++v) {
    const View4& src = in_views[v < num_views ? v : 0];
    NBDT_REQUIRE(src.C == Cin, "view %d has %d channels, expected %d", v, src.C, Cin);
    if (int rc = view_tmap(&tm_in[v], src, p.box_w, p.box_h, p.box_n)) return rc;
  }
  int dev = 0;
  NBDT_CUDA_OK(cudaGetDevice(&dev));
  nbdt::DeviceInfo info;
  if (int rc = nbdt::device_info(dev, &info)) return rc;
  {
    cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)w_rows};
    cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)block_n};
    if (int rc = make_tmap(&tm_w, wpack, 2, dims, strides, box)) return rc;
  }
  const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
  // CTA-pair kernel (cta_group::2) when there is at least one pair tile per SM pair; NBDT_B200_CONV_PAIR=0 disables it
  static const bool pair_enabled = !(getenv("NBDT_B200_CONV_PAIR") && atoi(getenv("NBDT_B200_CONV_PAIR")) == 0);
  const int total_pairs = ((p.tiles_w * p.tiles_h * p.tiles_n + 1) / 2) * p.n_tiles;
  // pair kernels from this many pair tiles on (default: one per SM pair); NBDT_B200_CONV_PAIR_MIN lowers it for tests
  const int pair_min = getenv("NBDT_B200_CONV_PAIR_MIN") ? atoi(getenv("NBDT_B200_CONV_PAIR_MIN")) : info.sm_count / 2;
  // H-halo variant: 3x3 taps on ONE dense view, tiles = box_w x box_h pixels of one image with box_w in {16, 32}
  bool halo = pair_enabled && total_pairs >= pair_min && num_taps == 9 && num_views == 1 &&
              p.box_n == 1 && (p.box_w == 16 || p.box_w == 32) && p.box_w * (p.box_h + 2) <= 192 &&
              !getenv("NBDT_B200_CONV_NOHALO");
  if (halo) {  // group the taps by dw; every group must hold dh = -1, 0, +1
    int ng = 0, cnt[3] = {0, 0, 0};
    for (int t = 0; t < num_taps && halo; ++t) {
      int g = 0;
      while (g < ng && p.grp_dw[g] != taps[t].dw) ++g;
      if (g == ng) {
        if (ng == 3) { halo = false; break; }
        p.grp_dw[ng++] = taps[t].dw;
      }
      if (cnt[g] == 3 || taps[t].dh < -1 || taps[t].dh > 1 || taps[t].map != 0) { halo = false; break; }
      p.grp_dh[g][cnt[g]] = taps[t].dh;
      p.grp_wrow[g][cnt[g]] = taps[t].wrow;
      ++cnt[g];
    }
    halo = halo && ng == 3 && cnt[0] == 3 && cnt[1] == 3 && cnt[2] == 3;
  }
  if (halo) {
    CUtensorMap tm_a, tm_w2;
    if (int rc = view_tmap(&tm_a, in_views[0], p.box_w, p.box_h + 2, 1)) return rc;
    {
      cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)w_rows};
      cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
      cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)(block_n / 2)};
      if (int rc = make_tmap(&tm_w2, wpack, 2, dims, strides, box)) return rc;
    }
    const int grid2 = std::min(total_pairs, info.sm_count / 2) * 2;
    if (stats_parts) *stats_parts = grid2;
#define NBDT_LAUNCH_TAP2H(BN)                                                                                   \
  {                                                                                                             \
    auto kern = tap_gemm2h_kernel<BN>;                                                                          \
    NBDT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TapGemm2hCfg<BN>::kSmemBytes)); \
    kern<<<grid2, 192, TapGemm2hCfg<BN>::kSmemBytes, stream>>>(tm_a, tm_w2, p);                                  \
  }
    if (block_n == 160) NBDT_LAUNCH_TAP2H(160)
    else if (block_n == 128) NBDT_LAUNCH_TAP2H(128)
    else NBDT_LAUNCH_TAP2H(64)
#undef NBDT_LAUNCH_TAP2H
    NBDT_CUDA_OK(cudaGetLastError());
    return NBDT_OK;
  }
  if (pair_enabled && total_pairs >= pair_min) {
    CUtensorMap tm_w2;
    {
      cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)w_rows};
      cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
      cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)(block_n / 2)};
      if (int rc = make_tmap(&tm_w2, wpack, 2, dims, strides, box)) return rc;
    }
    const int grid2 = std::min(total_pairs, info.sm_count / 2) * 2;
    if (stats_parts) *stats_parts = grid2;
#define NBDT_LAUNCH_TAP2(BN)                                                                                    \
  {                                                                                                             \
    auto kern = tap_gemm2_kernel<BN>;                                                                           \
    NBDT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TapGemm2Cfg<BN>::kSmemBytes)); \
    kern<<<grid2, 192, TapGemm2Cfg<BN>::kSmemBytes, stream>>>(tm_in[0], tm_in[1], tm_in[2], tm_in[3], tm_w2, p); \
  }
    if (block_n == 160) NBDT_LAUNCH_TAP2(160)
    else if (block_n == 128) NBDT_LAUNCH_TAP2(128)
    else NBDT_LAUNCH_TAP2(64)
#undef NBDT_LAUNCH_TAP2
    NBDT_CUDA_OK(cudaGetLastError());
    return NBDT_OK;
  }
  const int grid = std::min(total_tiles, info.sm_count);
  if (stats_parts) *stats_parts = grid;
#define NBDT_LAUNCH_TAP(BN)                                                                                     \
  {                                                                                                             \
    auto kern = tap_gemm_kernel<BN>;                                                                            \
    NBDT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TapGemmCfg<BN>::kSmemBytes)); \
    kern<<<grid, 192, TapGemmCfg<BN>::kSmemBytes, stream>>>(tm_in[0], tm_in[1], tm_in[2], tm_in[3], tm_w, p);   \
  }
  if (block_n == 160) NBDT_LAUNCH_TAP(160)
  else if (block_n == 128) NBDT_LAUNCH_TAP(128)
  else NBDT_LAUNCH_TAP(64)
#undef NBDT_LAUNCH_TAP
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

View4 dense_view(const void* base, int N, int H, int W, int C) {
  View4 v;
  v.base = reinterpret_cast<const __nv_bfloat16*>(base);
  v.N = N; v.H = H; v.W = W; v.C = C;
  v.sw = C; v.sh = (long long)W * C; v.sn = (long long)H * W * C;
  return v;
}

// parity sub-view (ph, pw) of a dense tensor with even H, W: element (n, i, j) -> (n, 2i+ph, 2j+pw)
View4 parity_view(const void* base, int N, int H, int W, int C, int ph, int pw) {
  View4 v = dense_view(base, N, H, W, C);
  v.base += (long long)ph * v.sh + (long long)pw * v.sw;
  v.H = H / 2; v.W = W / 2;
  v.sh *= 2; v.sw *= 2;
  return v;
}


// =============================================================================================
// Weight gradient:  dW_t[co, ci] = sum over pixels  dy[n, i, j, co] * x_{map(t)}[n, i + dh_t, j + dw_t, ci]
//
// A GEMM with the PIXEL index as K: both operands are "MN-major" for the tensor core (the channel
// index -- M for dy, N for x -- is the contiguous one in NHWC).  TMA loads {64 channels x 64 pixels}
// boxes (128-byte rows, one per pixel, 128B swizzle); the UMMA descriptors address them as MN-major
// tiles (LBO = distance between 64-channel chunks, SBO = 1024 B between 8-pixel groups).
// Work item = (tap, 128-row Cout tile, Cin tile, K split); fp32 partial sums are reduced into
// dW[tap][Cout][Cin] with red.global.add.f32 (dW must be zeroed by the caller).
// =============================================================================================
struct WgradParams {
  int tap_groups, m_tiles, n_tiles;  // work item = (tap group, Cout tile, Cin tile)
  long long units_per_cta;           // stream-K: every CTA gets the same number of (item, pixel-block) units
  int pix_blocks;                    // total KPIX-pixel blocks = tiles_w * tiles_h * tiles_n
  int range_len;                     // pixel blocks per range (divides pix_blocks)
  int teams;                         // > 0: team mode with this many pixel ranges (see wgrad_kernel); 0: stream-K
  int stage_tx_bytes;                // halo mode: TMA bytes per stage (dy tile + the halo boxes, without the atom padding)
  int tiles_w, tiles_h;              // pixel-block grid inside one image (box_n images per block when the image is small)
  int box_w, box_h, box_n;
  int Cout, Cin;
  float* dw;                         // [taps][Cout][Cin] fp32
  int tap_map[kMaxTaps];
  int tap_dh[kMaxTaps];
  int tap_dw[kMaxTaps];
};

// TAPS taps that share the dy tile are accumulated side by side in TMEM (TAPS x BLOCK_N columns), so one
// stage = 1 dy tile + TAPS shifted x tiles carries TAPS * KPIX/16 MMAs per barrier handshake.
// HALO (stride-1 3x3 convs on maps at least 16 pixels wide): the three taps of a filter row are the same pixels
// shifted by -1 / 0 / +1 along W, so ONE x tile with a one-pixel halo on each side ({64 ch x (box_w+2) x box_h}
// boxes, zero-filled outside the image) serves all three -- the MMA descriptors of tap dw start (1 + dw) pixel rows
// into the tile (a pixel is one 128-byte row of an MN-major tile; the descriptor's base-offset field carries the
// start row's phase inside the 8-row swizzle atom).  TMA fill per 32-pixel k-block: 8 KB + 3 x 34 x 128 B = 21 KB
// instead of 8 + 3 x 12 = 44 KB.  The kernel ran at the L2->SM limit (11.7 TB/s, profiles/r01_conv_ncu.md).
// The tensor core un-swizzles by ABSOLUTE shared-memory address bits, so a descriptor may start at any 128-byte row of a
// TMA-written tile with base offset 0 (verified on B200; setting the base-offset field to the row phase is wrong here).
// HALO == 2 (maps narrower than 16 pixels: 8x8, 4x4): the taps of a filter COLUMN are shifts along H; a tile with a
// one-row halo above and below ({64 ch x box_w x (box_h+2)} boxes) serves the three of them at row offsets (1+dh)*box_w
// -- 8 + 3 x 48 x 128 B = 26 KB per k-block.
template <int BLOCK_N, int TAPS, int KPIX, int HALO = 0>
struct WgradCfg {
  static constexpr int kChunkBytes = KPIX * 128;                   // KPIX pixels x 64 channels bf16
  static constexpr int kAChunks = 2;                               // M = 128 channels of dy
  static constexpr int kBChunks = (BLOCK_N + 63) / 64;
  // mode 1: (32+2) x 1 or (16+2) x 2 pixel rows padded to 8-row atoms; mode 2: 8 x (4+2) or 2 x 4 x (4+2) pixel rows
  static constexpr int kHaloRows = HALO == 2 ? 48 : 40;
  static constexpr int kHaloChunkBytes = kHaloRows * 128;
  static constexpr int kTapBytes = kBChunks * kChunkBytes;
  static constexpr int kBBytes = HALO ? kBChunks * kHaloChunkBytes : TAPS * kTapBytes;
  static constexpr int kStageBytes = kAChunks * kChunkBytes + kBBytes;
  // at most ~190 KB of the SM's shared memory: the HBM-bound BatchNorm-backward CTAs that follow the data gradient on the
  // main stream must be able to co-reside with this kernel's CTAs (it runs on a side stream, nn_ops.ConvFunction)
  static constexpr int kMaxStages = (190 * 1024) / kStageBytes;
  static constexpr int kStages = kMaxStages > 8 ? 8 : kMaxStages;
  static constexpr int kAccCols = TAPS * BLOCK_N;
  static constexpr int kNumAcc = (2 * kAccCols <= 512) ? 2 : 1;    // double-buffer the accumulators when they fit
  static constexpr int kTmemCols = (kNumAcc * kAccCols <= 128) ? 128 : (kNumAcc * kAccCols <= 256 ? 256 : 512);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
  static_assert(kAccCols <= 512, "accumulators exceed TMEM");
  static_assert(kStages >= 2, "stage too large");
  static_assert(!HALO || (TAPS == 3 && KPIX == 32), "halo tiles are for the three taps of a 3x3 filter row");
  static_assert(kStageBytes % 1024 == 0, "stages must keep the 1024-byte swizzle alignment");
};

template <int BLOCK_N, int TAPS, int KPIX, int HALO>
__global__ void __launch_bounds__(224, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x0,
             const __grid_constant__ CUtensorMap tm_x1, const __grid_constant__ CUtensorMap tm_x2,
             const __grid_constant__ CUtensorMap tm_x3, const WgradParams p) {
  using Cfg = WgradCfg<BLOCK_N, TAPS, KPIX, HALO>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kStages;
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_items = p.tap_groups * p.m_tiles * p.n_tiles;
  // j-th tap of a group: filter row `group` (taps along W) except in H-halo mode, where a group is a filter column
  auto tap_of = [](int group, int j) { return HALO == 2 ? j * 3 + group : group * TAPS + j; };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_dy);
    prefetch_tmap(&tm_x0);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full[a]), 1);
      mbar_init(smem_u32(&tmem_empty[a]), 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_base_slot), Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);

  // Stream-K: the (item, pixel-block) space is cut into gridDim.x equal contiguous ranges, so every CTA
  // does the same number of k-blocks; a CTA emits one epilogue per item segment it touches.
  // item -> (n_tile, m_tile, tap group), n fastest.
  const long long total_units = (long long)total_items * p.pix_blocks;
  const long long u_begin = min((long long)blockIdx.x * p.units_per_cta, total_units);
  const long long u_end = min(u_begin + p.units_per_cta, total_units);
  // Unit order: (pixel range, item, k-block within the range).  CTAs with neighbouring unit ranges therefore work on
  // the SAME pixels for different items at the same time and share them through L2 (with item-major order every item
  // re-streamed both tensors from HBM: 1.7 GB of DRAM reads for 0.34 GB of operands, profiles/r01_conv_ncu.md).
  // Team mode (operands larger than L2): the pixel blocks are cut into p.teams ranges and piece q = (team, item)
  // covers ONE item over ONE range; CTA c runs pieces c, c + gridDim.x, ...  The total_items CTAs of a team start
  // together and walk the same pixel blocks at the same pace, so each block is fetched from HBM once and the other
  // items hit it in L2 (with one contiguous unit span per CTA the items of a range ran at different times:
  // 1.61 GB of DRAM reads for 0.34 GB of operands in the 160->160 layers, profiles/r01_conv_ncu.md).
  const bool team_mode = p.teams > 0;
  const long long cur_begin = team_mode ? (long long)blockIdx.x : u_begin;
  const long long cur_end = team_mode ? (long long)p.teams * total_items : u_end;
  // decodes the segment at cursor u and returns the cursor of the CTA's next segment
  auto decode = [&](long long u, int& group, int& m_tile, int& n_tile, int& kb0, int& kb1) -> long long {
    int item;
    long long next;
    if (team_mode) {
      const int team = (int)(u / total_items);
      item = (int)(u - (long long)team * total_items);
      kb0 = (int)((long long)team * p.pix_blocks / p.teams);
      kb1 = (int)((long long)(team + 1) * p.pix_blocks / p.teams);
      next = u + gridDim.x;
    } else {
      const long long seg = u / p.range_len;
      const int in_seg = (int)(u - seg * p.range_len);
      item = (int)(seg % total_items);
      const int range0 = (int)(seg / total_items) * p.range_len;
      kb0 = range0 + in_seg;
      kb1 = (int)min((long long)(range0 + p.range_len), kb0 + (u_end - u));
      next = u + (kb1 - kb0);
    }
    n_tile = item % p.n_tiles;
    item /= p.n_tiles;
    m_tile = item % p.m_tiles;
    group = item / p.m_tiles;
    return next;
  };

  if (warp == 0 || warp == 6) {
    // Two TMA-issuing warps (a stage is 2 + TAPS * kBChunks small boxes; one elected thread could not issue them
    // within the stage's MMA time): warp 0 arms the barrier and loads dy + tap 0, warp 6 loads the other taps.
    const bool leader = elect_one();
    const bool first = warp == 0;
    int stage = 0;
    uint32_t phase = 0;
    for (long long u = cur_begin; u < cur_end;) {
      int group, m_tile, n_tile, kb0, kb1;
      u = decode(u, group, m_tile, n_tile, kb0, kb1);
      if (!first && TAPS == 1) continue;
      for (int kb = kb0; kb < kb1; ++kb) {
        int b = kb;
        const int w0 = (b % p.tiles_w) * p.box_w;
        b /= p.tiles_w;
        const int h0 = (b % p.tiles_h) * p.box_h;
        const int n0 = (b / p.tiles_h) * p.box_n;
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        const uint32_t dst = smem_base + stage * Cfg::kStageBytes;
        const uint32_t bar = full0 + 8 * stage;
        if (leader && first) {
          mbar_expect_tx(bar, HALO ? p.stage_tx_bytes : Cfg::kStageBytes);
#pragma unroll
          for (int c = 0; c < Cfg::kAChunks; ++c)
            tma_load_4d(dst + c * Cfg::kChunkBytes, &tm_dy, bar, m_tile * kBlockM + c * 64, w0, h0, n0);
          if (HALO == 1) {  // one halo tile for the whole tap row: box {64, box_w + 2, box_h, box_n} from (w0 - 1, h0 + dh)
#pragma unroll
            for (int c = 0; c < Cfg::kBChunks; ++c)
              tma_load_4d(dst + Cfg::kAChunks * Cfg::kChunkBytes + c * Cfg::kHaloChunkBytes, &tm_x0, bar,
                          n_tile * BLOCK_N + c * 64, w0 - 1, h0 + p.tap_dh[tap_of(group, 0)], n0);
          }
          if (HALO == 2) {  // tap column: box {64, box_w, box_h + 2, box_n} from (w0 + dw, h0 - 1)
#pragma unroll
            for (int c = 0; c < Cfg::kBChunks; ++c)
              tma_load_4d(dst + Cfg::kAChunks * Cfg::kChunkBytes + c * Cfg::kHaloChunkBytes, &tm_x0, bar,
                          n_tile * BLOCK_N + c * 64, w0 + p.tap_dw[tap_of(group, 0)], h0 - 1, n0);
          }
        }
#pragma unroll
        for (int j = 0; j < TAPS; ++j) {
          if (HALO) continue;
          if ((j == 0) != first) continue;
          const int tap = group * TAPS + j;
          const int map = p.tap_map[tap];
          const CUtensorMap* tmx = map == 0 ? &tm_x0 : (map == 1 ? &tm_x1 : (map == 2 ? &tm_x2 : &tm_x3));
          if (leader) {
#pragma unroll
            for (int c = 0; c < Cfg::kBChunks; ++c)
              tma_load_4d(dst + Cfg::kAChunks * Cfg::kChunkBytes + j * Cfg::kTapBytes + c * Cfg::kChunkBytes, tmx, bar,
                          n_tile * BLOCK_N + c * 64, w0 + p.tap_dw[tap], h0 + p.tap_dh[tap], n0);
          }
        }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, 1, 1);  // both operands MN-major
    // MN-major tiles: LBO = distance between 64-channel chunks, SBO = 1024 B between 8-pixel groups
    const uint64_t desc0 = make_smem_desc(smem_base, Cfg::kChunkBytes, 1024);
    // halo mode: B descriptors (chunk stride = padded halo tile) + per (tap, K step) start row inside the tile
    const uint64_t descb0 = make_smem_desc(smem_base + Cfg::kAChunks * Cfg::kChunkBytes, Cfg::kHaloChunkBytes, 1024);
    uint64_t bdelta[TAPS][KPIX / 16];
#pragma unroll
    for (int j = 0; j < TAPS; ++j)
#pragma unroll
      for (int k = 0; k < KPIX / 16; ++k) {
        const int pix = k * 16;  // first pixel of the K step, in box order (w fastest, then h, then n)
        int row;
        if (HALO == 2) {
          const int per_img = p.box_w * p.box_h;
          row = (pix / per_img) * p.box_w * (p.box_h + 2) + ((pix % per_img) / p.box_w + j) * p.box_w + pix % p.box_w;
        } else {
          row = (pix / p.box_w) * (p.box_w + 2) + (pix % p.box_w) + 1 + (j - 1);
        }
        bdelta[j][k] = (uint64_t)(row * 8);  // + row * 128 bytes on the start address, base offset 0 (see above)
      }
    int stage = 0;
    uint32_t phase = 0;
    int item_idx = 0;
    for (long long u = cur_begin; u < cur_end; ++item_idx) {
      int group, m_tile, n_tile, kb0, kb1;
      u = decode(u, group, m_tile, n_tile, kb0, kb1);
      const int acc = item_idx % Cfg::kNumAcc;
      const uint32_t acc_phase = (item_idx / Cfg::kNumAcc) & 1;
      mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * Cfg::kAccCols;
      uint32_t accumulate = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        const uint64_t adesc = desc0 + (uint64_t)((stage * Cfg::kStageBytes) >> 4);
        if (leader) {
#pragma unroll
          for (int j = 0; j < TAPS; ++j) {
            const uint64_t bdesc = adesc + (uint64_t)((Cfg::kAChunks * Cfg::kChunkBytes + j * Cfg::kTapBytes) >> 4);
            const uint64_t hdesc = descb0 + (uint64_t)((stage * Cfg::kStageBytes) >> 4);
#pragma unroll
            for (int k = 0; k < KPIX / 16; ++k) {
              // K=16 pixels = two 8-pixel swizzle atoms = 2048 bytes (128 descriptor units)
              umma_bf16(tmem_d + j * BLOCK_N, adesc + 128 * k, HALO ? hdesc + bdelta[j][k] : bdesc + 128 * k, idesc,
                        (k > 0) ? 1u : accumulate);
            }
          }
          umma_commit(empty0 + 8 * stage);
        }
        accumulate = 1;
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(smem_u32(&tmem_full[acc]));
      __syncwarp();
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int item_idx = 0;
    for (long long u = cur_begin; u < cur_end; ++item_idx) {
      int group, m_tile, n_tile, kb0, kb1;
      u = decode(u, group, m_tile, n_tile, kb0, kb1);
      const int acc = item_idx % Cfg::kNumAcc;
      const uint32_t acc_phase = (item_idx / Cfg::kNumAcc) & 1;
      const int co = m_tile * kBlockM + row;
      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * Cfg::kAccCols;
#pragma unroll 1
      for (int j = 0; j < TAPS; ++j) {
        const int tap = tap_of(group, j);
        float* dst = p.dw + ((long long)tap * p.Cout + co) * p.Cin + n_tile * BLOCK_N;
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + j * BLOCK_N + c0, v);
          tmem_ld_wait();
          if (co < p.Cout) {
            // 16-byte vector reductions (red.global.add.v4.f32): a quarter of the L2 atomic operations of scalar
            // atomicAdd -- 148 CTAs x 128 x 480 scalar reductions per epilogue round cost ~30 us of L2 atomic throughput
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const int col = n_tile * BLOCK_N + c0 + e;
              if (col + 3 < p.Cin) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + e), "f"(__uint_as_float(v[e])),
                             "f"(__uint_as_float(v[e + 1])), "f"(__uint_as_float(v[e + 2])), "f"(__uint_as_float(v[e + 3]))
                             : "memory");
              } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  if (col + q < p.Cin) atomicAdd(dst + c0 + e + q, __uint_as_float(v[e + q]));
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// weight repacking: fp32 OIHW master weights <-> the bf16 tap-major matrices the GEMMs read
// ---------------------------------------------------------------------------------------------
// Tiled (32 output channels x 32 input channels x taps through shared memory) so that the OIHW side is read / written
// in runs of 32*taps contiguous floats and BOTH tap-major sides in runs of 32 contiguous elements -- the first version
// used one thread per packed element with three 64-bit divisions and stride-`taps` / stride-`cout` scattered accesses
// (0.76 ms per step for 2 x 27 launches, profiles/r01_bench_launches_v3.md).
constexpr int kPackTile = 32;

__device__ __forceinline__ void pack_weights_tile(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp,
                                                  __nv_bfloat16* __restrict__ wpt, int Cout, int Cin, int cin_pad, int cout_pad, int taps,
                                                  int bx, int by) {
  // w: [Cout][Cin][taps] (OIHW);  wp: [taps][cout_pad][cin_pad];  wpt: [taps][cin_pad][cout_pad]  (zero padded)
  extern __shared__ float tile[];  // [32 co][32 * taps + 1]
  const int row_len = kPackTile * taps + 1;
  const int ci0 = bx * kPackTile, co0 = by * kPackTile;
  const int per_row = kPackTile * taps;
  for (int i = threadIdx.x; i < kPackTile * per_row; i += blockDim.x) {
    const int co_l = i / per_row, rem = i - co_l * per_row;
    const int co = co0 + co_l, ci = ci0 + rem / taps;
    tile[co_l * row_len + rem] = (co < Cout && ci < Cin) ? __ldg(w + ((size_t)co * Cin + ci0) * taps + rem) : 0.f;
  }
  __syncthreads();
  const int n = kPackTile * kPackTile * taps;
  if (wp) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int ci_l = i % kPackTile, co_l = (i / kPackTile) % kPackTile, tp = i / (kPackTile * kPackTile);
      const int co = co0 + co_l, ci = ci0 + ci_l;
      if (co < cout_pad && ci < cin_pad)
        wp[((size_t)tp * cout_pad + co) * cin_pad + ci] = __float2bfloat16_rn(tile[co_l * row_len + ci_l * taps + tp]);
    }
  }
  if (wpt) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int co_l = i % kPackTile, ci_l = (i / kPackTile) % kPackTile, tp = i / (kPackTile * kPackTile);
      const int co = co0 + co_l, ci = ci0 + ci_l;
      if (co < cout_pad && ci < cin_pad)
        wpt[((size_t)tp * cin_pad + ci) * cout_pad + co] = __float2bfloat16_rn(tile[co_l * row_len + ci_l * taps + tp]);
    }
  }
}

__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp,
                                                           __nv_bfloat16* __restrict__ wpt, int Cout, int Cin,
                                                           int cin_pad, int cout_pad, int taps) {
  pack_weights_tile(w, wp, wpt, Cout, Cin, cin_pad, cout_pad, taps, blockIdx.x, blockIdx.y);
}

// One launch for every convolution layer of the network (instead of one per layer: 28 + 28 small launches per WRN-28-10 step were
// launch-latency bound, 1.3 ms).  job = 8 x int64: {fp32 OIHW tensor, packed tensor 0, packed tensor 1, Cout, Cin, cout_pad, cin_pad,
// taps | first tile << 32}; tiles of a job: (cin tiles) x (cout tiles), numbered x-major from `first tile`.
struct PackJob { long long a, b, c, Cout, Cin, cout_pad, cin_pad, taps_first; };
__device__ __forceinline__ const PackJob& find_job(const PackJob* __restrict__ jobs, int njobs, int tile_id) {
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {  // last job whose first tile <= tile_id
    const int mid = (lo + hi + 1) >> 1;
    if ((int)(jobs[mid].taps_first >> 32) <= tile_id) lo = mid; else hi = mid - 1;
  }
  return jobs[lo];
}

__device__ __forceinline__ void unpack_wgrad_tile(const float* __restrict__ dwp, float* __restrict__ dw, int Cout, int Cin, int cin_pad,
                                                  int cout_pad, int taps, int accumulate, int bx, int by) {
  // dwp: [taps][cout_pad][cin_pad] fp32  ->  dw: [Cout][Cin][taps] (OIHW)
  extern __shared__ float tile[];  // [32 co][32 * taps + 1]
  const int row_len = kPackTile * taps + 1;
  const int ci0 = bx * kPackTile, co0 = by * kPackTile;
  const int n = kPackTile * kPackTile * taps;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int ci_l = i % kPackTile, co_l = (i / kPackTile) % kPackTile, tp = i / (kPackTile * kPackTile);
    const int co = co0 + co_l, ci = ci0 + ci_l;
    tile[co_l * row_len + ci_l * taps + tp] =
        (co < Cout && ci < Cin) ? __ldg(dwp + ((size_t)tp * cout_pad + co) * cin_pad + ci) : 0.f;
  }
  __syncthreads();
  const int per_row = kPackTile * taps;
  for (int i = threadIdx.x; i < kPackTile * per_row; i += blockDim.x) {
    const int co_l = i / per_row, rem = i - co_l * per_row;
    const int co = co0 + co_l, ci = ci0 + rem / taps;
    if (co < Cout && ci < Cin) {
      float* dst = dw + ((size_t)co * Cin + ci0) * taps + rem;
      const float v = tile[co_l * row_len + rem];
      *dst = accumulate ? *dst + v : v;
    }
  }
}

__global__ void __launch_bounds__(256) unpack_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Cout,
                                                           int Cin, int cin_pad, int cout_pad, int taps, int accumulate) {
  unpack_wgrad_tile(dwp, dw, Cout, Cin, cin_pad, cout_pad, taps, accumulate, blockIdx.x, blockIdx.y);
}

__global__ void __launch_bounds__(256) pack_weights_batch_kernel(const PackJob* __restrict__ jobs, int njobs) {
  const PackJob& j = find_job(jobs, njobs, blockIdx.x);
  const int local = blockIdx.x - (int)(j.taps_first >> 32);
  const int tx = ((int)j.cin_pad + kPackTile - 1) / kPackTile;
  pack_weights_tile(reinterpret_cast<const float*>(j.a), reinterpret_cast<__nv_bfloat16*>(j.b), reinterpret_cast<__nv_bfloat16*>(j.c),
                    (int)j.Cout, (int)j.Cin, (int)j.cin_pad, (int)j.cout_pad, (int)(j.taps_first & 0xffffffff), local % tx, local / tx);
}

__global__ void __launch_bounds__(256) unpack_wgrad_batch_kernel(const PackJob* __restrict__ jobs, int njobs) {
  const PackJob& j = find_job(jobs, njobs, blockIdx.x);
  const int local = blockIdx.x - (int)(j.taps_first >> 32);
  const int tx = ((int)j.Cin + kPackTile - 1) / kPackTile;
  unpack_wgrad_tile(reinterpret_cast<const float*>(j.b), reinterpret_cast<float*>(j.a), (int)j.Cout, (int)j.Cin, (int)j.cin_pad,
                    (int)j.cout_pad, (int)(j.taps_first & 0xffffffff), 1, local % tx, local / tx);
}

}  // namespace

extern "C" {

// y[N,Ho,Wo,Cout] = conv(x[N,H,W,Cin], w) (+ residual), ksize in {1,3} (pad = ksize/2), stride in {1,2}.
// wpack: bf16 [ksize*ksize][Cout][Cin] (tap-major, K contiguous), from nbdt_conv_pack_weights.
static int conv2d_fwd_impl(const void* x, const void* wpack, void* y, const void* residual, int N, int H, int W, int Cin,
                           int Cout, int ksize, int stride, float* stats, int* stats_parts, void* stream) {
  NBDT_REQUIRE(x && wpack && y, "NULL tensor");
  NBDT_REQUIRE((ksize == 1 || ksize == 3) && (stride == 1 || stride == 2), "unsupported conv %dx%d stride %d", ksize, ksize, stride);
  NBDT_REQUIRE(stride == 1 || (H % 2 == 0 && W % 2 == 0), "stride-2 conv needs even H, W");
  const int Ho = H / stride, Wo = W / stride;
  View4 out = dense_view(y, N, Ho, Wo, Cout);
  View4 views[4];
  Tap taps[kMaxTaps];
  int nt = 0, nv = 1;
  const int pad = ksize / 2;
  if (stride == 1) {
    views[0] = dense_view(x, N, H, W, Cin);
    for (int r = 0; r < ksize; ++r)
      for (int s = 0; s < ksize; ++s) taps[nt++] = Tap{0, r - pad, s - pad, (r * ksize + s) * Cout};
  } else {
    nv = 4;
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) views[ph * 2 + pw] = parity_view(x, N, H, W, Cin, ph, pw);
    for (int r = 0; r < ksize; ++r)
      for (int s = 0; s < ksize; ++s) {
        // input row 2*ho + r - pad = 2*(ho + dh) + ph
        const int rr = r - pad, ss = s - pad;
        const int ph = ((rr % 2) + 2) % 2, pw = ((ss % 2) + 2) % 2;
        taps[nt++] = Tap{ph * 2 + pw, (rr - ph) / 2, (ss - pw) / 2, (r * ksize + s) * Cout};
      }
  }
  return launch_tap_gemm(views, nv, taps, nt, reinterpret_cast<const __nv_bfloat16*>(wpack), ksize * ksize * Cout, Cin,
                         out, reinterpret_cast<const __nv_bfloat16*>(residual), (cudaStream_t)stream, stats, stats_parts);
}

int nbdt_conv2d_fwd(const void* x, const void* wpack, void* y, const void* residual, int N, int H, int W, int Cin,
                    int Cout, int ksize, int stride, void* stream) {
  return conv2d_fwd_impl(x, wpack, y, residual, N, H, W, Cin, Cout, ksize, stride, nullptr, nullptr, stream);
}

// The same convolution, and in its epilogue the BatchNorm batch statistics of y: every CTA writes one row of per-channel partial
// sums, partials[cta][0][c] = sum of y[..., c], partials[cta][1][c] = sum of y[..., c]^2 over the pixels the CTA stored (the bf16
// values).  `partials` holds nbdt_conv2d_stats_floats(Cout) floats; *parts receives the number of rows written (host int, known
// at launch time).  nbdt_bn_finalize_partials (bn_ops.cu) turns them into the scale / shift of the BatchNorm that consumes y --
// the separate statistics pass over y (nbdt_bn_stats) disappears.  Replaces the F.batch_norm statistics of the reference's
// nn.BatchNorm2d modules that follow a convolution (nbdt/models/resnet.py:47-54,69-74 and the pytorchcv WRN units).
int nbdt_conv2d_fwd_stats(const void* x, const void* wpack, void* y, const void* residual, int N, int H, int W, int Cin,
                          int Cout, int ksize, int stride, float* partials, int* parts, void* stream) {
  NBDT_REQUIRE(partials && parts, "NULL statistics buffer");
  return conv2d_fwd_impl(x, wpack, y, residual, N, H, W, Cin, Cout, ksize, stride, partials, parts, stream);
}

// floats nbdt_conv2d_fwd_stats needs in `partials` for Cout output channels (0: too many channels for the fused statistics)
long long nbdt_conv2d_stats_floats(int Cout) {
  if (Cout > kStatsMaxC || Cout <= 0) return 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  nbdt::DeviceInfo info;
  if (nbdt::device_info(dev, &info)) return 0;
  return (long long)info.sm_count * 2 * Cout;
}

// dx[N,H,W,Cin] = conv_transpose(dy[N,Ho,Wo,Cout], w) (+ accumulate), the data gradient of nbdt_conv2d_fwd.
// wpack_t: bf16 [ksize*ksize][Cin][Cout] (tap-major, Cout contiguous), from nbdt_conv_pack_weights.
int nbdt_conv2d_dgrad(const void* dy, const void* wpack_t, void* dx, const void* accumulate, int N, int H, int W,
                      int Cin, int Cout, int ksize, int stride, void* stream) {
  NBDT_REQUIRE(dy && wpack_t && dx, "NULL tensor");
  NBDT_REQUIRE((ksize == 1 || ksize == 3) && (stride == 1 || stride == 2), "unsupported conv %dx%d stride %d", ksize, ksize, stride);
  NBDT_REQUIRE(stride == 1 || (H % 2 == 0 && W % 2 == 0), "stride-2 conv needs even H, W");
  const int Ho = H / stride, Wo = W / stride;
  const int pad = ksize / 2;
  const __nv_bfloat16* wp = reinterpret_cast<const __nv_bfloat16*>(wpack_t);
  View4 in = dense_view(dy, N, Ho, Wo, Cout);
  if (stride == 1) {
    // dx[h,w] = sum_{r,s} W[r,s]^T dy[h + pad - r, w + pad - s]
    Tap taps[kMaxTaps];
    int nt = 0;
    for (int r = 0; r < ksize; ++r)
      for (int s = 0; s < ksize; ++s) taps[nt++] = Tap{0, pad - r, pad - s, (r * ksize + s) * Cin};
    View4 out = dense_view(dx, N, H, W, Cin);
    return launch_tap_gemm(&in, 1, taps, nt, wp, ksize * ksize * Cin, Cout, out,
                           reinterpret_cast<const __nv_bfloat16*>(accumulate), (cudaStream_t)stream);
  }
  if (ksize == 1) {
    NBDT_REQUIRE(accumulate == nullptr, "1x1 stride-2 dgrad cannot accumulate; run it first and accumulate the 3x3 dgrad on top");
    NBDT_CUDA_OK(cudaMemsetAsync(dx, 0, (size_t)N * H * W * Cin * 2, (cudaStream_t)stream));
  }
  // stride 2: dx[2i+ph, 2j+pw] = sum over taps with r == ph + pad (mod 2), s likewise, of
  //           W[r,s]^T dy[i + (ph + pad - r)/2, j + (pw + pad - s)/2]
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      Tap taps[kMaxTaps];
      int nt = 0;
      for (int r = 0; r < ksize; ++r)
        for (int s = 0; s < ksize; ++s) {
          const int a = ph + pad - r, b = pw + pad - s;
          if ((a & 1) || (b & 1)) continue;
          taps[nt++] = Tap{0, a / 2, b / 2, (r * ksize + s) * Cin};
        }
      View4 out = parity_view(dx, N, H, W, Cin, ph, pw);
      const __nv_bfloat16* acc = accumulate ? parity_view(accumulate, N, H, W, Cin, ph, pw).base : nullptr;
      if (nt == 0) continue;  // (1x1 stride 2: odd positions get no gradient; dx was zero-filled above)
      if (int rc = launch_tap_gemm(&in, 1, taps, nt, wp, ksize * ksize * Cin, Cout, out, acc, (cudaStream_t)stream)) return rc;
    }
  return NBDT_OK;
}

// dwp[taps][Cout][Cin] (fp32, zeroed here) = weight gradient of nbdt_conv2d_fwd given x and dy.
int nbdt_conv2d_wgrad(const void* x, const void* dy, float* dwp, int N, int H, int W, int Cin, int Cout, int ksize,
                      int stride, void* stream) {
  NBDT_REQUIRE(x && dy && dwp, "NULL tensor");
  NBDT_REQUIRE((ksize == 1 || ksize == 3) && (stride == 1 || stride == 2), "unsupported conv %dx%d stride %d", ksize, ksize, stride);
  NBDT_REQUIRE(stride == 1 || (H % 2 == 0 && W % 2 == 0), "stride-2 conv needs even H, W");
  NBDT_REQUIRE(Cin % 8 == 0 && Cout % 8 == 0, "channel counts must be multiples of 8");
  const int Ho = H / stride, Wo = W / stride;
  const int pad = ksize / 2;
  cudaStream_t st = (cudaStream_t)stream;
  const int taps = ksize * ksize;
  NBDT_CUDA_OK(cudaMemsetAsync(dwp, 0, (size_t)taps * Cout * Cin * sizeof(float), st));

  WgradParams p = {};
  int nv = 1;
  View4 xviews[4];
  if (stride == 1) {
    xviews[0] = dense_view(x, N, H, W, Cin);
    int nt = 0;
    for (int r = 0; r < ksize; ++r)
      for (int s = 0; s < ksize; ++s) { p.tap_map[nt] = 0; p.tap_dh[nt] = r - pad; p.tap_dw[nt] = s - pad; ++nt; }
  } else {
    nv = 4;
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) xviews[ph * 2 + pw] = parity_view(x, N, H, W, Cin, ph, pw);
    int nt = 0;
    for (int r = 0; r < ksize; ++r)
      for (int s = 0; s < ksize; ++s) {
        const int rr = r - pad, ss = s - pad;
        const int ph = ((rr % 2) + 2) % 2, pw = ((ss % 2) + 2) % 2;
        p.tap_map[nt] = ph * 2 + pw; p.tap_dh[nt] = (rr - ph) / 2; p.tap_dw[nt] = (ss - pw) / 2; ++nt;
      }
  }
  // 3x3: the three taps of one filter row share the dy tile (TAPS = 3, 32-pixel k-blocks); 1x1: TAPS = 1, 64 pixels
  const int taps_per_group = taps == 9 ? 3 : 1;
  const int kpix = taps == 9 ? 32 : 64;
  int block_n;
  if (Cin % 160 == 0) block_n = 160;
  else if (Cin % 128 == 0) block_n = 128;
  else if (Cin % 64 == 0) block_n = 64;
  else return nbdt::set_error(NBDT_EINVAL, "wgrad: input channels (%d) must be a multiple of 64", Cin);
  p.tap_groups = taps / taps_per_group;
  p.m_tiles = (Cout + kBlockM - 1) / kBlockM;
  p.n_tiles = Cin / block_n;
  p.box_w = std::min(pow2_ceil(Wo), kpix);
  p.box_h = std::min(pow2_ceil(Ho), kpix / p.box_w);
  p.box_n = kpix / (p.box_w * p.box_h);
  p.tiles_w = (Wo + p.box_w - 1) / p.box_w;
  p.tiles_h = (Ho + p.box_h - 1) / p.box_h;
  const int tiles_n = (N + p.box_n - 1) / p.box_n;
  p.pix_blocks = p.tiles_w * p.tiles_h * tiles_n;
  p.Cout = Cout; p.Cin = Cin; p.dw = dwp;
  int dev = 0;
  NBDT_CUDA_OK(cudaGetDevice(&dev));
  nbdt::DeviceInfo info;
  if (int rc = nbdt::device_info(dev, &info)) return rc;
  const int base_items = p.tap_groups * p.m_tiles * p.n_tiles;
  // stream-K (operands L2-resident): one pixel range per item, so a CTA's span crosses as few (item) segments -- one
  // epilogue each -- as possible
  p.range_len = p.pix_blocks;
  const long long total_units = (long long)base_items * p.pix_blocks;
  int grid = (int)std::min<long long>(total_units, info.sm_count);
  p.units_per_cta = (total_units + grid - 1) / grid;
  // Operands that do not fit in L2 together: team scheduling (each pixel range read from HBM once, see the kernel).
  // teams = the smallest count whose pieces fill the SMs to >= 94 % (else the best), ranges of >= 64 k-blocks.
  p.teams = 0;
  const double operand_bytes = 2.0 * N * ((double)H * W * Cin + (double)Ho * Wo * Cout);
  if (operand_bytes > 96e6 && base_items <= info.sm_count && !getenv("NBDT_B200_WGRAD_STREAMK")) {
    double best_eff = 0.0;
    for (int t = 1; t <= 64 && p.pix_blocks / t >= 64; ++t) {
      const long long pieces = (long long)t * base_items;
      const long long rounds = (pieces + info.sm_count - 1) / info.sm_count;
      const double eff = (double)pieces / (double)(rounds * info.sm_count);
      if (eff > best_eff + 1e-9) { best_eff = eff; p.teams = t; }
      if (eff >= 0.94) break;
    }
    if (p.teams > 0) grid = (int)std::min<long long>((long long)p.teams * base_items, info.sm_count);
  }
  CUtensorMap tm_dy, tm_x[4];
  {
    View4 dyv = dense_view(dy, N, Ho, Wo, Cout);
    cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)dyv.sw * 2, (cuuint64_t)dyv.sh * 2, (cuuint64_t)dyv.sn * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)p.box_w, (cuuint32_t)p.box_h, (cuuint32_t)p.box_n};
    if (int rc = make_tmap(&tm_dy, dyv.base, 4, dims, strides, box)) return rc;
  }
  for (int v = 0; v < 4; ++v)
    if (int rc = view_tmap(&tm_x[v], xviews[v < nv ? v : 0], p.box_w, p.box_h, p.box_n)) return rc;
  // halo tiles for stride-1 3x3: mode 1 when the pixel boxes are whole 16-pixel runs of an image row, mode 2 (tap columns)
  // for narrower maps whose box holds 16 or 32 pixels of one image
  int halo = 0;
  if (taps_per_group == 3 && stride == 1 && !getenv("NBDT_B200_WGRAD_NOHALO")) {
    if (p.box_w == 16 || p.box_w == 32) halo = 1;
    else if (p.box_w * p.box_h * p.box_n == 32 && (p.box_w * p.box_h) % 16 == 0 && p.box_w * (p.box_h + 2) * p.box_n <= 48) halo = 2;
  }
  if (halo == 1) {
    if (int rc = view_tmap(&tm_x[0], xviews[0], p.box_w + 2, p.box_h, p.box_n)) return rc;
    p.stage_tx_bytes = 2 * kpix * 128 + ((block_n + 63) / 64) * (p.box_w + 2) * p.box_h * p.box_n * 128;
  } else if (halo == 2) {
    if (int rc = view_tmap(&tm_x[0], xviews[0], p.box_w, p.box_h + 2, p.box_n)) return rc;
    p.stage_tx_bytes = 2 * kpix * 128 + ((block_n + 63) / 64) * p.box_w * (p.box_h + 2) * p.box_n * 128;
  }
#define NBDT_LAUNCH_WG(BN, TP, KP, HL)                                                                          \
  {                                                                                                             \
    auto kern = wgrad_kernel<BN, TP, KP, HL>;                                                                   \
    NBDT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WgradCfg<BN, TP, KP, HL>::kSmemBytes)); \
    kern<<<grid, 224, WgradCfg<BN, TP, KP, HL>::kSmemBytes, st>>>(tm_dy, tm_x[0], tm_x[1], tm_x[2], tm_x[3], p); \
  }
  if (halo == 1) {
    if (block_n == 160) NBDT_LAUNCH_WG(160, 3, 32, 1)
    else if (block_n == 128) NBDT_LAUNCH_WG(128, 3, 32, 1)
    else NBDT_LAUNCH_WG(64, 3, 32, 1)
  } else if (halo == 2) {
    if (block_n == 160) NBDT_LAUNCH_WG(160, 3, 32, 2)
    else if (block_n == 128) NBDT_LAUNCH_WG(128, 3, 32, 2)
    else NBDT_LAUNCH_WG(64, 3, 32, 2)
  } else if (taps_per_group == 3) {
    if (block_n == 160) NBDT_LAUNCH_WG(160, 3, 32, 0)
    else if (block_n == 128) NBDT_LAUNCH_WG(128, 3, 32, 0)
    else NBDT_LAUNCH_WG(64, 3, 32, 0)
  } else {
    if (block_n == 160) NBDT_LAUNCH_WG(160, 1, 64, 0)
    else if (block_n == 128) NBDT_LAUNCH_WG(128, 1, 64, 0)
    else NBDT_LAUNCH_WG(64, 1, 64, 0)
  }
#undef NBDT_LAUNCH_WG
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

// fp32 OIHW weights -> bf16 wpack [taps][cout_pad][cin_pad] and/or wpack_t [taps][cin_pad][cout_pad] (zero padded).
int nbdt_conv_pack_weights(const float* w_oihw, void* wpack, void* wpack_t, int Cout, int Cin, int cout_pad,
                           int cin_pad, int ksize, void* stream) {
  NBDT_REQUIRE(w_oihw && (wpack || wpack_t), "NULL tensor");
  NBDT_REQUIRE(cout_pad >= Cout && cin_pad >= Cin, "padded sizes smaller than the real ones");
  const int taps = ksize * ksize;
  const dim3 grid((cin_pad + kPackTile - 1) / kPackTile, (cout_pad + kPackTile - 1) / kPackTile);
  pack_weights_kernel<<<grid, 256, kPackTile * (kPackTile * taps + 1) * sizeof(float), (cudaStream_t)stream>>>(
      w_oihw, reinterpret_cast<__nv_bfloat16*>(wpack), reinterpret_cast<__nv_bfloat16*>(wpack_t), Cout, Cin, cin_pad,
      cout_pad, taps);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

// fp32 dwp [taps][cout_pad][cin_pad] -> OIHW gradient [Cout][Cin][k][k] (optionally accumulating).
int nbdt_conv_unpack_wgrad(const float* dwp, float* dw_oihw, int Cout, int Cin, int cout_pad, int cin_pad, int ksize,
                           int accumulate, void* stream) {
  NBDT_REQUIRE(dwp && dw_oihw, "NULL tensor");
  const int taps = ksize * ksize;
  const dim3 grid((Cin + kPackTile - 1) / kPackTile, (Cout + kPackTile - 1) / kPackTile);
  unpack_wgrad_kernel<<<grid, 256, kPackTile * (kPackTile * taps + 1) * sizeof(float), (cudaStream_t)stream>>>(
      dwp, dw_oihw, Cout, Cin, cin_pad, cout_pad, taps, accumulate);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

// Batched forms over a device table of jobs (8 x int64 each, see PackJob): total_tiles = sum of the jobs' tile counts.
// pack: job = {w_oihw, wpack, wpack_t, Cout, Cin, cout_pad, cin_pad, taps | first_tile << 32}, tiles over (cin_pad, cout_pad);
// unpack (accumulating): job = {dw_oihw, dwp, 0, ...}, tiles over (Cin, Cout).
int nbdt_conv_pack_weights_batch(const void* jobs_device, int njobs, int total_tiles, void* stream) {
  NBDT_REQUIRE(njobs >= 0 && total_tiles >= 0, "negative count");
  if (njobs == 0 || total_tiles == 0) return NBDT_OK;
  NBDT_REQUIRE(jobs_device != nullptr, "NULL job table");
  pack_weights_batch_kernel<<<total_tiles, 256, kPackTile * (kPackTile * 9 + 1) * sizeof(float), (cudaStream_t)stream>>>(
      reinterpret_cast<const PackJob*>(jobs_device), njobs);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

int nbdt_conv_unpack_wgrad_batch(const void* jobs_device, int njobs, int total_tiles, void* stream) {
  NBDT_REQUIRE(njobs >= 0 && total_tiles >= 0, "negative count");
  if (njobs == 0 || total_tiles == 0) return NBDT_OK;
  NBDT_REQUIRE(jobs_device != nullptr, "NULL job table");
  unpack_wgrad_batch_kernel<<<total_tiles, 256, kPackTile * (kPackTile * 9 + 1) * sizeof(float), (cudaStream_t)stream>>>(
      reinterpret_cast<const PackJob*>(jobs_device), njobs);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

}  // extern "C"
