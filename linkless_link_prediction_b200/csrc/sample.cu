// Uniform random walks for LLP context sampling.  Replaces torch_cluster 1.6.0's
// uniform_sampling_kernel behind random_walk(row, col, start, walk_length, coalesced=False)
// (reference: src/main.py:37,43,45).  Bit-exact given the same `rand` tensor.
#include "common.cuh"

namespace llp {

__global__ void random_walk_kernel(const int64_t* __restrict__ rowptr, const int64_t* __restrict__ col,
                                   const int64_t* __restrict__ start, const float* __restrict__ rand, int64_t B, int64_t L,
                                   int64_t* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= B) return;
  int64_t cur = start[i];
  out[i * (L + 1)] = cur;
  for (int64_t l = 0; l < L; ++l) {
    int64_t rs = rowptr[cur], re = rowptr[cur + 1];
    int64_t deg = re - rs;
    if (deg > 0) {
      // fp32 product, truncated — exactly `int64_t(rand * (row_end - row_start))`
      int64_t off = (int64_t)(rand[i * L + l] * (float)deg);
      cur = col[rs + off];
    }
    out[i * (L + 1) + l + 1] = cur;
  }
}

}  // namespace llp

using namespace llp;

extern "C" int llp_random_walk(const int64_t* rowptr, const int64_t* col, const int64_t* start, const float* rand,
                               int64_t B, int64_t L, int64_t* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(rowptr && start && out && B >= 0 && L >= 0 && (L == 0 || rand));
  if (int rc = check_device()) return rc;
  if (B == 0) return 0;
  random_walk_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, stream>>>(rowptr, col, start, rand, B, L, out);
  LLP_LAUNCH_OK();
  return 0;
}
