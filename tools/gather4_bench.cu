// Development tool: can the Blackwell TMA row-gather (cp.async.bulk.tensor.2d ... tile::gather4: four arbitrary rows of a
// 2-D tensor per single-thread instruction) feed a gather-reduce faster than per-lane 128-bit LDGs?  
// A CTA = one producer warp (elected lane issues gather4 into a ring of stages of R rows) + C consumer warps that sum the
// staged rows (fp32 accumulators, one 16-byte vector per lane and row) and release the stage.  Random row indices over
// an [N, F] bf16 matrix (F = 256: 512-byte rows, the C4 SpMM shape).  Reports correctness against an LDG gather of the
// same indices and GB/s of gathered bytes.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/gather4_bench tools/gather4_bench.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* map, int c0, int r0, int r1, int r2, int r3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

constexpr int kRowBytes = 512;   // F = 256 bf16
constexpr int kRows = 32;        // rows per stage (8 gather4 instructions)
constexpr int kStageBytes = kRows * kRowBytes;

// edges [e_begin, e_end) of this CTA; out[cta][lane vector] = sum of the gathered rows (checksum)
__global__ void __launch_bounds__(32 * 5) gather4_reduce(const __grid_constant__ CUtensorMap map, const int* __restrict__ idx,
                                                          int edges_per_cta, int stages, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * kStageBytes);
  uint64_t* empty = full + stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kConsumers = 4;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty[s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int* my = idx + (size_t)blockIdx.x * edges_per_cta;
  const int n_stage = edges_per_cta / kRows;
  if (warp == 0) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < n_stage; ++i) {
      mbar_wait(s32(&empty[s]), ph ^ 1);
      const int r = __ldg(my + i * kRows + lane);   // 32 row indices of this stage, one per lane
      // lanes 0, 4, 8, ... hold the first index of each group of four: pull the other three by shuffle
      const int r1 = __shfl_down_sync(0xffffffffu, r, 1), r2 = __shfl_down_sync(0xffffffffu, r, 2), r3 = __shfl_down_sync(0xffffffffu, r, 3);
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(kStageBytes) : "memory");
      __syncwarp();
      if ((lane & 3) == 0)
        gather4(s32(smem + (size_t)s * kStageBytes + (lane >> 2) * 4 * kRowBytes), &map, 0, r, r1, r2, r3, s32(&full[s]));
      __syncwarp();
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  } else {
    // consumer w takes stages w-1, w-1+4, ...: the ring position and phase follow from the stage's global number
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    for (int i = warp - 1; i < n_stage; i += kConsumers) {
      const int s = i % stages;
      const uint32_t ph = (uint32_t)(i / stages) & 1u;
      mbar_wait(s32(&full[s]), ph);
      const uint4* rows = reinterpret_cast<const uint4*>(smem + (size_t)s * kStageBytes);
#pragma unroll 8
      for (int r = 0; r < kRows; ++r) {
        const uint4 v = rows[r * (kRowBytes / 16) + lane];
        acc[0] += __uint_as_float(v.x << 16); acc[1] += __uint_as_float(v.x & 0xffff0000u);
        acc[2] += __uint_as_float(v.y << 16); acc[3] += __uint_as_float(v.y & 0xffff0000u);
        acc[4] += __uint_as_float(v.z << 16); acc[5] += __uint_as_float(v.z & 0xffff0000u);
        acc[6] += __uint_as_float(v.w << 16); acc[7] += __uint_as_float(v.w & 0xffff0000u);
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
    }
    float* o = out + ((size_t)blockIdx.x * kConsumers + (warp - 1)) * 256 + lane * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = acc[j];
  }
}

// reference: the round-1 SpMM inner loop shape — each warp gathers rows with 128-bit LDGs, 4 in flight per lane
__global__ void __launch_bounds__(128) ldg_reduce(const uint4* __restrict__ x, const int* __restrict__ idx, int edges_per_warp,
                                                  float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int* my = idx + (size_t)w * edges_per_warp;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
  for (int e = 0; e < edges_per_warp; e += 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(x + (size_t)__ldg(my + e + u) * 32 + lane);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      acc[0] += __uint_as_float(v[u].x << 16); acc[1] += __uint_as_float(v[u].x & 0xffff0000u);
      acc[2] += __uint_as_float(v[u].y << 16); acc[3] += __uint_as_float(v[u].y & 0xffff0000u);
      acc[4] += __uint_as_float(v[u].z << 16); acc[5] += __uint_as_float(v[u].z & 0xffff0000u);
      acc[6] += __uint_as_float(v[u].w << 16); acc[7] += __uint_as_float(v[u].w & 0xffff0000u);
    }
  }
  float* o = out + (size_t)w * 256 + lane * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = acc[j];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int N = 235868, F = 256;
  const long E = 2358104;
  const int stages = argc > 1 ? atoi(argv[1]) : 12;
  const int ctas = argc > 2 ? atoi(argv[2]) : 148;
  const int box_rows = argc > 3 ? atoi(argv[3]) : 1;
  std::vector<uint16_t> hx((size_t)N * F);
  srand(1);
  for (auto& v : hx) v = (uint16_t)(0x3f80 + (rand() & 0x3f));   // bf16 values in [1, 1.5)
  const int edges_per_cta = (int)(E / ctas / (kRows * 4) * (kRows * 4));
  std::vector<int> hidx((size_t)edges_per_cta * ctas);
  for (auto& v : hidx) v = (int)(((long)rand() * 32768 + rand()) % N);
  uint16_t* x; int* idx; float *out_a, *out_b;
  CK(cudaMalloc(&x, hx.size() * 2)); CK(cudaMalloc(&idx, hidx.size() * 4));
  CK(cudaMalloc(&out_a, (size_t)ctas * 4 * 256 * 4)); CK(cudaMalloc(&out_b, (size_t)ctas * 4 * 256 * 4));
  CK(cudaMemcpy(x, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(idx, hidx.data(), hidx.size() * 4, cudaMemcpyHostToDevice));
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  CUtensorMap map;
  cuuint64_t gdim[2] = {(cuuint64_t)F, (cuuint64_t)N};
  cuuint64_t gstride[1] = {(cuuint64_t)F * 2};
  cuuint32_t box[2] = {(cuuint32_t)F, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)fnp)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("tensor map (box {%d, %d}) encode: %d\n", F, box_rows, (int)r);
  if (r != CUDA_SUCCESS) return 1;
  const size_t smem = (size_t)stages * kStageBytes + 2 * stages * 8 + 64;
  CK(cudaFuncSetAttribute(gather4_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double gbytes = (double)edges_per_cta * ctas * kRowBytes;
  float ms;
  for (int it = 0; it < 3; ++it) {
    CK(cudaEventRecord(e0));
    gather4_reduce<<<ctas, 160, smem>>>(map, idx, edges_per_cta, stages, out_a);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("gather4 ring  stages %2d ctas %d: %8.1f us  %7.1f GB/s gathered\n", stages, ctas, ms * 1e3, gbytes / ms / 1e6);
  }
  // LDG reference on the same indices: 4 warps per "CTA slot" x the same number of edges per warp-slice
  const int warps = ctas * 4, edges_per_warp = edges_per_cta / 4;
  // NOTE: the consumer/stage assignment above is round-robin by stage, so per-warp sums differ; compare the TOTAL checksum
  for (int it = 0; it < 3; ++it) {
    CK(cudaEventRecord(e0));
    ldg_reduce<<<warps / 4, 128>>>((const uint4*)x, idx, edges_per_warp, out_b);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("ldg 4-in-flight, %d warps (1 per SM slot): %8.1f us  %7.1f GB/s gathered\n", warps, ms * 1e3, gbytes / ms / 1e6);
  }
  // full-occupancy LDG gather (the production kernel's shape: one warp per 64 edges)
  {
    const int epw = 64; const long nw = (long)edges_per_cta * ctas / epw;
    float* out_c; CK(cudaMalloc(&out_c, (size_t)nw * 256 * 4));
    for (int it = 0; it < 3; ++it) {
      CK(cudaEventRecord(e0));
      ldg_reduce<<<(unsigned)(nw / 4), 128>>>((const uint4*)x, idx, epw, out_c);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
      CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("ldg 4-in-flight, %ld warps of 64 edges (production shape): %8.1f us  %7.1f GB/s gathered\n", nw, ms * 1e3, gbytes / ms / 1e6);
    }
  }
  std::vector<float> ha((size_t)ctas * 4 * 256), hb((size_t)ctas * 4 * 256);
  CK(cudaMemcpy(ha.data(), out_a, ha.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hb.data(), out_b, hb.size() * 4, cudaMemcpyDeviceToHost));
  // per CTA and column: sum over its 4 consumers == sum over its 4 LDG warps (same edges, different grouping)
  double maxrel = 0;
  for (int c = 0; c < ctas; ++c)
    for (int j = 0; j < 256; ++j) {
      double a = 0, b = 0;
      for (int w = 0; w < 4; ++w) { a += ha[((size_t)c * 4 + w) * 256 + j]; b += hb[((size_t)c * 4 + w) * 256 + j]; }
      const double rel = fabs(a - b) / fabs(b);
      if (rel > maxrel) maxrel = rel;
    }
  printf("checksum: max relative difference gather4 vs LDG per (CTA, column) = %.3e (fp32 summation order differs)\n", maxrel);
  return 0;
}
