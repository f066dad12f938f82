// Microbenchmark: issue cost of tcgen05.mma (M=128, K=16, bf16) as a function of N, of the number of
// independent accumulators interleaved, and of the commit frequency.  One CTA per SM, garbage operands.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_mma scripts/ubench_mma.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../neural-backed-decision-trees_b200/csrc/sm100.cuh"
using namespace sm100;

template <int N, int NACC, int COMMIT_EVERY>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const int ncommits = (iters + COMMIT_EVERY - 1) / COMMIT_EVERY;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), ncommits); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t a_addr = smem_u32(smem);
    const uint64_t adesc = make_smem_desc(a_addr, 16, 1024);
    const uint64_t bdesc = make_smem_desc(a_addr + 16384, 16, 1024);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (leader) {
        umma_bf16(tmem + (i % NACC) * N, adesc + 2 * (i & 3), bdesc + 2 * (i & 3), idesc, i >= NACC ? 1u : 0u);
        if ((i + 1) % COMMIT_EVERY == 0 || i + 1 == iters) umma_commit(smem_u32(&bar));
      }
    }
    const long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    if (leader && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// serialized latency: issue NMMA MMAs, commit, wait for the barrier, repeat
template <int N, int NMMA>
__global__ void __launch_bounds__(128, 1) bench_latency(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t a_addr = smem_u32(smem);
    const uint64_t adesc = make_smem_desc(a_addr, 16, 1024);
    const uint64_t bdesc = make_smem_desc(a_addr + 16384, 16, 1024);
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (leader) {
        for (int k = 0; k < NMMA; ++k) umma_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
        umma_commit(smem_u32(&bar));
      }
      mbar_wait(smem_u32(&bar), phase);
      phase ^= 1;
      tc_fence_after();
    }
    const long long t1 = clock64();
    if (leader && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// mbarrier ping-pong between two warps (no tensor work): round trip of two arrive->wake hops
__global__ void __launch_bounds__(128, 1) bench_pingpong(long long* out, int iters) {
  __shared__ uint64_t bars[2];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1); fence_barrier_init(); }
  __syncthreads();
  if (warp == 0) {
    const bool leader = elect_one();
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (leader) mbar_arrive(smem_u32(&bars[0]));
      mbar_wait(smem_u32(&bars[1]), phase);
      phase ^= 1;
    }
    if (leader && blockIdx.x == 0) out[0] = clock64() - t0;
  } else if (warp == 1) {
    const bool leader = elect_one();
    uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(smem_u32(&bars[0]), phase);
      phase ^= 1;
      if (leader) mbar_arrive(smem_u32(&bars[1]));
    }
  }
}

// groups of 4 MMAs separated by a configurable "loop overhead": MODE bit0 commit to a side barrier, bit1 try_wait on an
// already-complete barrier, bit2 tcgen05.fence::after_thread_sync, and DELAY cycles of busy spinning.
template <int N, int MODE, int DELAY>
__global__ void __launch_bounds__(128, 1) bench_groups(long long* out, int groups) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, side, done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&side), groups); mbar_init(smem_u32(&done_bar), 1);
    fence_barrier_init();
    mbar_arrive(smem_u32(&done_bar));  // phase 0 of done_bar is complete: try_wait(parity 0) succeeds immediately
  }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t a_addr = smem_u32(smem);
    const uint64_t adesc = make_smem_desc(a_addr, 16, 1024);
    const uint64_t bdesc = make_smem_desc(a_addr + 16384, 16, 1024);
    unsigned sink = 0;
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (MODE & 2) sink += mbar_try_wait(smem_u32(&done_bar), 0);
      if (MODE & 4) tc_fence_after();
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
        if (MODE & 1) umma_commit(smem_u32(&side));
      }
      if (DELAY > 0) { const long long s0 = clock64(); while (clock64() - s0 < DELAY) {} }
    }
    if (leader) umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (leader && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = sink; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, int MODE, int DELAY>
void run_groups(long long* d_out, int groups) {
  auto k = bench_groups<N, MODE, DELAY>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int rep = 0; rep < 2; ++rep) k<<<148, 128, 100 * 1024>>>(d_out, groups);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("groups of 4 MMAs N=%d mode=%d (1 commit 2 trywait 4 fence) delay=%3d: %7.1f cyc/group (floor %d)  %s\n", N, MODE, DELAY,
         (double)h[0] / groups, 2 * N, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int N, int NMMA>
void run_latency(long long* d_out, int iters) {
  auto k = bench_latency<N, NMMA>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int rep = 0; rep < 2; ++rep) k<<<148, 128, 100 * 1024>>>(d_out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[1];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("serialized issue->commit->wait      N=%3d %d mma/round: %7.1f cyc/round  %s\n", N, NMMA, (double)h[0] / iters,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int N, int NACC, int CE>
void run(const char* name, long long* d_out, int iters) {
  auto k = bench<N, NACC, CE>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int rep = 0; rep < 2; ++rep) k<<<148, 128, 100 * 1024>>>(d_out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-34s N=%3d acc=%d commit/%-3d: issue %7.1f cyc/mma, complete %7.1f cyc/mma (pipe floor %d)  %s\n", name, N, NACC, CE,
         (double)h[0] / iters, (double)h[1] / iters, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  const int iters = 512;
  run<64, 1, 512>("single acc, one commit", d_out, iters);
  run<128, 1, 512>("single acc, one commit", d_out, iters);
  run<160, 1, 512>("single acc, one commit", d_out, iters);
  run<256, 1, 512>("single acc, one commit", d_out, iters);
  run<160, 2, 512>("two accs interleaved", d_out, iters);
  run<128, 2, 512>("two accs interleaved", d_out, iters);
  run<64, 4, 512>("four accs interleaved", d_out, iters);
  run<160, 1, 4>("single acc, commit every 4", d_out, iters);
  run<160, 1, 1>("single acc, commit every 1", d_out, iters);
  run<256, 1, 4>("single acc, commit every 4", d_out, iters);
  run<160, 2, 8>("two accs, commit every 8", d_out, iters);
  run_groups<160, 0, 0>(d_out, 256);
  run_groups<160, 1, 0>(d_out, 256);
  run_groups<160, 3, 0>(d_out, 256);
  run_groups<160, 7, 0>(d_out, 256);
  run_groups<160, 0, 100>(d_out, 256);
  run_groups<160, 0, 200>(d_out, 256);
  run_groups<160, 0, 300>(d_out, 256);
  run_groups<160, 0, 500>(d_out, 256);
  run_groups<160, 7, 100>(d_out, 256);
  run_groups<160, 7, 200>(d_out, 256);
  run_latency<160, 1>(d_out, 256);
  run_latency<160, 4>(d_out, 256);
  run_latency<64, 1>(d_out, 256);
  run_latency<256, 4>(d_out, 256);
  for (int rep = 0; rep < 2; ++rep) bench_pingpong<<<148, 128>>>(d_out, 1024);
  cudaDeviceSynchronize();
  long long h[1];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("mbarrier ping-pong between two warps: %7.1f cyc/round trip\n", (double)h[0] / 1024);
  return 0;
}
