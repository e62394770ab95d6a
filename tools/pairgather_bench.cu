// Development probe: is a two-pass, half-warp-per-row gather (16 lanes x 16 B = a 256-byte column slice of a 512-byte row,
// two edges per warp instruction, one pass per column half so that the pass's footprint is 60 MB) faster than the
// one-pass full-warp gather of the production SpMM?  Pure gather + fp32 accumulate over random rows of a [N, 256] bf16
// matrix, production launch shape (one warp per 64 edges), L2 flushed before every timed launch.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/pairgather_bench tools/pairgather_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void acc8(float (&acc)[8], const uint4& v) {
  acc[0] += __uint_as_float(v.x << 16); acc[1] += __uint_as_float(v.x & 0xffff0000u);
  acc[2] += __uint_as_float(v.y << 16); acc[3] += __uint_as_float(v.y & 0xffff0000u);
  acc[4] += __uint_as_float(v.z << 16); acc[5] += __uint_as_float(v.z & 0xffff0000u);
  acc[6] += __uint_as_float(v.w << 16); acc[7] += __uint_as_float(v.w & 0xffff0000u);
}

// one pass, full warp per row (production shape): 4 rows in flight per lane
__global__ void __launch_bounds__(128) full_rows(const uint4* __restrict__ x, const int* __restrict__ idx, int epw, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int* my = idx + w * epw;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int e = 0; e < epw; e += 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(x + (size_t)__ldg(my + e + u) * 32 + lane);
#pragma unroll
    for (int u = 0; u < 4; ++u) acc8(acc, v[u]);
  }
  float* o = out + w * 256 + lane * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = acc[j];
}

// one column half per launch (blockIdx.y = pass): half-warp per row, two edges per warp instruction, 4 instructions
// (= 8 edges) in flight per lane
__global__ void __launch_bounds__(128) half_rows(const uint4* __restrict__ x, const int* __restrict__ idx, int epw, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, h = lane >> 4, l16 = lane & 15;
  const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int pass = blockIdx.y;
  const int* my = idx + w * epw;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int e = 0; e < epw; e += 8) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(x + (size_t)__ldg(my + e + 2 * u + h) * 32 + pass * 16 + l16);
#pragma unroll
    for (int u = 0; u < 4; ++u) acc8(acc, v[u]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
  if (h == 0) {
    float* o = out + w * 256 + pass * 128 + l16 * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = acc[j];
  }
}

int main(int argc, char** argv) {
  const int N = 235868, F = 256;
  const long E = 2358104;
  const int epw = 64;
  const long nw = E / epw / 4 * 4;
  const bool powerlaw = argc > 1 && atoi(argv[1]) == 1;
  std::vector<uint16_t> hx((size_t)N * F);
  srand(1);
  for (auto& v : hx) v = (uint16_t)(0x3f80 + (rand() & 0x3f));
  std::vector<int> hidx((size_t)nw * epw);
  for (auto& v : hidx) {
    double u = (double)rand() / RAND_MAX;
    v = powerlaw ? (int)((double)(N - 1) * u * u * u) : (int)(((long)rand() * 32768 + rand()) % N);   // u^3: skewed towards low ids
  }
  uint16_t* x; int* idx; float *oa, *ob; char* flush;
  CK(cudaMalloc(&x, hx.size() * 2)); CK(cudaMalloc(&idx, hidx.size() * 4));
  CK(cudaMalloc(&oa, (size_t)nw * 256 * 4)); CK(cudaMalloc(&ob, (size_t)nw * 256 * 4)); CK(cudaMalloc(&flush, 512u << 20));
  CK(cudaMemcpy(x, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(idx, hidx.data(), hidx.size() * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double gb = (double)nw * epw * 512;
  float ms;
  for (int it = 0; it < 4; ++it) {
    CK(cudaMemset(flush, it, 512u << 20));
    CK(cudaEventRecord(e0));
    full_rows<<<(unsigned)(nw / 4), 128>>>((const uint4*)x, idx, epw, oa);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%s one pass, full warp per row      (cold L2): %7.1f us  %7.1f GB/s gathered\n", powerlaw ? "skewed " : "uniform", ms * 1e3, gb / ms / 1e6);
    CK(cudaMemset(flush, it, 512u << 20));
    CK(cudaEventRecord(e0));
    half_rows<<<dim3((unsigned)(nw / 4), 2), 128>>>((const uint4*)x, idx, epw, ob);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%s two passes, half warp per row    (cold L2): %7.1f us  %7.1f GB/s gathered\n", powerlaw ? "skewed " : "uniform", ms * 1e3, gb / ms / 1e6);
  }
  std::vector<float> ha(1 << 16), hb(1 << 16);
  CK(cudaMemcpy(ha.data(), oa, ha.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hb.data(), ob, hb.size() * 4, cudaMemcpyDeviceToHost));
  double md = 0;
  for (size_t i = 0; i < ha.size(); ++i) { double d = fabs(ha[i] - hb[i]) / fabs(ha[i]); if (d > md) md = d; }
  printf("max relative difference between the two gathers: %.3e\n", md);
  return 0;
}
