// Development tool: issue rate of tcgen05.mma (cta_group::1, kind::f16, bf16 -> fp32) from shared-memory descriptors,
// one CTA per SM, operands = whatever is in shared memory.  Answers "what can one CTA's tensor core sustain for a
// 128 x N x 16 instruction stream" for K-major and MN-major SW128 operands, with 1 or 2 accumulators in flight.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I linkless_link_prediction_b200/csrc \
//        -o tools/_build/mmabench tools/mmabench.cu
#include <stdio.h>
#include <stdlib.h>

#include "tcgen05.cuh"

namespace llp { std::atomic<int64_t> g_launch_count{0}; int g_tuning[32] = {0}; }
using namespace llp::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// mode: 0 = K-major A and B (NT GEMM), 1 = MN-major A and B (weight gradient)
__global__ void __launch_bounds__(128, 1) mma_rate(int N, int mode, int iters, int k_per_commit, int accs, int per_mma, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[8];
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&holder), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = holder;
  if (warp == 1) {
    const uint32_t idesc = make_idesc(128, N, mode == 1);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32 * 1024);
    // descriptors precomputed once: the loop body is nothing but the MMA issue
    uint64_t ad[4], bd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (mode == 0) {
        ad[k] = make_smem_desc(sa + k * 32, 16, 1024);
        bd[k] = make_smem_desc(sb + k * 32, 16, 1024);
      } else {
        ad[k] = make_smem_desc(sa + (k & 1) * 2048 + (k >> 1) * 8192, 4096, 1024);
        bd[k] = make_smem_desc(sb + (k & 1) * 2048 + (k >> 1) * 16384, 4096, 1024);
      }
    }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int slot = it & 7;
      if (it >= 8) mbar_wait(smem_u32(&bars[slot]), (uint32_t)(((it >> 3) - 1) & 1));  // the slot's previous commit
      if (elect_one_sync()) {
        const uint32_t d = tmem + (uint32_t)((per_mma ? 0 : it % accs) * N);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(per_mma ? tmem + (uint32_t)((k % accs) * N) : d, ad[k], bd[k], idesc, 1u);
        umma_commit(smem_u32(&bars[slot]));
      }
      __syncwarp();
    }
    for (int it = iters - 8; it < iters; ++it) mbar_wait(smem_u32(&bars[it & 7]), (uint32_t)((it >> 3) & 1));
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tmem, 512); }
}

// cta_group::2: a CTA pair issues 256 x N x 16 MMAs (A: 128 rows from each CTA, B: N/2 rows from each CTA)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma_rate_pair(int N, int iters, int mask, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[8];
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&holder)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tcgen05_fence_before();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tcgen05_fence_after();
  const uint32_t tmem = holder;
  if (warp == 1 && rank == 0) {
    const uint32_t idesc = make_idesc(256, N, false);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32 * 1024);
    uint64_t ad[4], bd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { ad[k] = make_smem_desc(sa + k * 32, 16, 1024); bd[k] = make_smem_desc(sb + k * 32, 16, 1024); }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int slot = it & 7;
      if (it >= 8) mbar_wait(smem_u32(&bars[slot]), (uint32_t)(((it >> 3) - 1) & 1));
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem), "l"(ad[k]), "l"(bd[k]), "r"(idesc), "r"(1u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bars[slot])), "h"((uint16_t)mask) : "memory");
      }
      __syncwarp();
    }
    for (int it = iters - 8; it < iters; ++it) mbar_wait(smem_u32(&bars[it & 7]), (uint32_t)((it >> 3) & 1));
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x >> 1] = t1 - t0;
  }
  tcgen05_fence_before();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

int main() {
  long long* out;
  CK(cudaMalloc(&out, 148 * sizeof(long long)));
  CK(cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  long long h[148];
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {64, 128, 256})
      for (int cfg = 0; cfg < 5; ++cfg) {
        const int accs = cfg == 0 ? 1 : (cfg <= 2 ? 2 : 4);
        const int per_mma = cfg == 2 || cfg == 4;
        if (accs * N > 512) continue;
        const int iters = 4096, kpc = 4;
        mma_rate<<<148, 128, 100 * 1024>>>(N, mode, iters, kpc, accs, per_mma, out);
        CK(cudaDeviceSynchronize());
        mma_rate<<<148, 128, 100 * 1024>>>(N, mode, iters, kpc, accs, per_mma, out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        const double per_mma_clk = (double)mx / ((double)iters * kpc);
        const double flop_clk = 2.0 * 128 * N * 16 / per_mma_clk;
        printf("%s M=128 N=%3d K=16, %d accumulator(s)%s: %7.1f clk per MMA  -> %7.0f FLOP/clk/SM (%.0f%% of 8192)\n",
               mode == 0 ? "K-major " : "MN-major", N, accs, per_mma ? " alternating per MMA" : "", per_mma_clk, flop_clk, 100.0 * flop_clk / 8192.0);
      }
  CK(cudaFuncSetAttribute(mma_rate_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int cfg = 0; cfg < 4; ++cfg) {
    const int N = (cfg & 1) ? 256 : 128, mask = (cfg & 2) ? 3 : 1;
    const int iters = 4096;
    for (int rep = 0; rep < 2; ++rep) {
      mma_rate_pair<<<148, 128, 100 * 1024>>>(N, iters, mask, out);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h, out, 74 * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < 74; ++i) mx = h[i] > mx ? h[i] : mx;
    const double clk = (double)mx / ((double)iters * 4);
    printf("cta_group::2 K-major M=256 N=%3d K=16, commit mask %d: %7.1f clk per MMA -> %7.0f FLOP/clk per SM (%.0f%% of 8192)\n", N, mask, clk,
           2.0 * 256 * N * 16 / clk / 2, 100.0 * (2.0 * 256 * N * 16 / clk / 2) / 8192.0);
  }
  return 0;
}
