// Development tool: what HBM bandwidth can a READ-dominated kernel reach on this B200?  (MEASURED_PEAKS.json's figure is
// a copy: half reads, half writes.)  Three readers over a 1 GiB buffer: 128-bit LDG grid-stride, cp.async.bulk (1-D TMA)
// into a shared-memory ring with S stages of C KiB per CTA, and a plain copy for reference.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/membench tools/membench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void ldg_read(const uint4* __restrict__ p, size_t n, uint32_t* out) {
  uint32_t acc = 0;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    uint4 a = p[i], b = p[i + stride], c = p[i + 2 * stride], d = p[i + 3 * stride];
    acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
  }
  for (; i < n; i += stride) { uint4 a = p[i]; acc ^= a.x ^ a.y ^ a.z ^ a.w; }
  if (acc == 0x12345678u) out[0] = acc;
}

__global__ void copy_k(const uint4* __restrict__ p, uint4* __restrict__ q, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) q[i] = p[i];
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one elected thread streams this CTA's slice through a ring of `stages` buffers of `chunk` bytes; a consumer warp
// releases each buffer as soon as it has landed (no compute): pure load rate of the TMA path
__global__ void __launch_bounds__(64) bulk_read(const uint8_t* __restrict__ p, size_t bytes_per_cta, int chunk, int stages) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * chunk);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty[s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint8_t* src = p + (size_t)blockIdx.x * bytes_per_cta;
  const int n = (int)(bytes_per_cta / chunk);
  auto wait = [](uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
      asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                   : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
  };
  if (threadIdx.x == 0) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < n; ++i) {
      wait(s32(&empty[s]), ph ^ 1);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(s32(smem + (size_t)s * chunk)), "l"(src + (size_t)i * chunk), "r"(chunk), "r"(s32(&full[s])) : "memory");
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < n; ++i) {
      wait(s32(&full[s]), ph);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  }
}

template <typename F>
static float time_ms(F fn, int iters = 5) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  fn();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < iters; ++i) {
    CK(cudaEventRecord(a)); fn(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  const size_t bytes = (size_t)1 << 30;
  uint8_t *p, *q; uint32_t* out;
  CK(cudaMalloc(&p, bytes)); CK(cudaMalloc(&q, bytes)); CK(cudaMalloc(&out, 4));
  CK(cudaMemset(p, 1, bytes)); CK(cudaMemset(q, 2, bytes));
  const size_t n16 = bytes / 16;
  for (int blocks_per_sm : {4, 8, 16}) {
    float ms = time_ms([&] { ldg_read<<<148 * blocks_per_sm, 512>>>((const uint4*)p, n16, out); });
    printf("ldg.128 read-only   %2d CTAs/SM x 512 thr          : %7.1f GB/s\n", blocks_per_sm, bytes / ms / 1e6);
  }
  {
    float ms = time_ms([&] { copy_k<<<148 * 8, 512>>>((const uint4*)p, (uint4*)q, n16); });
    printf("ldg/stg copy (read+write bytes)                  : %7.1f GB/s\n", 2.0 * bytes / ms / 1e6);
  }
  CK(cudaFuncSetAttribute(bulk_read, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  const int cfgs[][3] = {{16, 12, 1}, {32, 6, 1}, {64, 3, 1}, {16, 6, 1}, {16, 6, 2}, {8, 12, 2}, {16, 3, 4}, {4, 12, 4}};  // chunk KiB, stages, CTAs/SM
  for (auto& c : cfgs) {
    const int chunk = c[0] * 1024, stages = c[1], per_sm = c[2];
    const int ctas = 148 * per_sm;
    size_t per_cta = bytes / ctas / chunk * chunk;
    size_t smem = (size_t)stages * chunk + 16 * stages + 64;
    float ms = time_ms([&] { bulk_read<<<ctas, 64, smem>>>(p, per_cta, chunk, stages); });
    printf("cp.async.bulk read  %d CTA/SM, %2d KiB x %2d stages (%3d KiB in flight/SM): %7.1f GB/s\n", per_sm, c[0], stages,
           c[0] * stages * per_sm, (double)per_cta * ctas / ms / 1e6);
  }
  return 0;
}
