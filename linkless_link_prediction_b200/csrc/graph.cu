// CSR construction of the message graph (stable radix sort by destination / source).
// Replaces the per-call gather/scatter bookkeeping PyG does for a dense [2,E] edge_index
// (reference: src/models.py:113,118 -> SAGEConv.propagate).  The SpMM is in spmm.cu.
#include <cub/cub.cuh>

#include "common.cuh"

namespace llp {

// ------------------------------------------------------------------------------------------------
// CSR build
// ------------------------------------------------------------------------------------------------
__global__ void csr_prepare_kernel(const int64_t* __restrict__ key, int64_t E, int32_t* __restrict__ key32,
                                   int32_t* __restrict__ idx) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < E) {
    key32[i] = (int32_t)key[i];
    idx[i] = (int32_t)i;
  }
}

__global__ void csr_finish_kernel(const int64_t* __restrict__ val, const int32_t* __restrict__ perm, int64_t E,
                                  int32_t* __restrict__ col) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < E) col[i] = (int32_t)val[perm[i]];
}

// rowptr[r] = number of sorted keys < r  (lower bound), r in [0, N]
__global__ void csr_rowptr_kernel(const int32_t* __restrict__ sorted_key, int64_t E, int64_t N,
                                  int32_t* __restrict__ rowptr, float* __restrict__ inv_deg) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r > N) return;
  auto lower = [&](int64_t target) {
    int64_t lo = 0, hi = E;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if ((int64_t)sorted_key[mid] < target) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  int64_t a = lower(r);
  rowptr[r] = (int32_t)a;
  if (inv_deg != nullptr && r < N) {
    int64_t b = lower(r + 1);
    float d = (float)(b - a);
    inv_deg[r] = 1.0f / fmaxf(d, 1.0f);
  }
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t cub_sort_temp_bytes(int64_t E, int end_bit) {
  size_t temp = 0;
  cub::DoubleBuffer<int32_t> k(nullptr, nullptr), v(nullptr, nullptr);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, temp, k, v, (int)E, 0, end_bit);
  if (e != cudaSuccess) {
    cudaGetLastError();
    temp = (size_t)16 << 20;  // conservative bound when no device is visible (onesweep needs O(E/tile) lookback words)
    temp += (size_t)E / 64;
  }
  return temp;
}

}  // namespace llp

using namespace llp;

extern "C" size_t llp_csr_build_workspace_bytes(int64_t N, int64_t E) {
  if (E < 0 || N < 0) return 0;
  return 4 * align256((size_t)E * 4 + 4) + align256(cub_sort_temp_bytes(E, 32)) + 1024;
}

extern "C" int llp_csr_build(const int64_t* edge_val, const int64_t* edge_key, int64_t E, int64_t N, int32_t* rowptr,
                             int32_t* col, int32_t* perm, float* inv_deg, void* workspace, size_t workspace_bytes,
                             void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(E >= 0 && N >= 0 && rowptr != nullptr);
  LLP_CHECK_ARG(E == 0 || (edge_val && edge_key && col && workspace));
  LLP_CHECK_ARG(E < (int64_t)INT32_MAX && N < (int64_t)INT32_MAX);
  if (int rc = check_device()) return rc;
  if (workspace_bytes < llp_csr_build_workspace_bytes(N, E)) return LLP_E_WORKSPACE;

  char* ws = reinterpret_cast<char*>(workspace);
  size_t slab = align256((size_t)E * 4 + 4);
  int32_t* key_a = reinterpret_cast<int32_t*>(ws);
  int32_t* key_b = reinterpret_cast<int32_t*>(ws + slab);
  int32_t* idx_a = reinterpret_cast<int32_t*>(ws + 2 * slab);
  int32_t* idx_b = reinterpret_cast<int32_t*>(ws + 3 * slab);
  void* cub_temp = ws + 4 * slab;
  int32_t* sorted_key = key_a;
  int32_t* sorted_idx = idx_a;

  if (E > 0) {
    int threads = 256;
    int64_t blocks = ceil_div(E, threads);
    csr_prepare_kernel<<<(unsigned)blocks, threads, 0, stream>>>(edge_key, E, key_a, idx_a);
    LLP_LAUNCH_OK();
    int end_bit = 1;
    while (end_bit < 31 && ((int64_t)1 << end_bit) < N) ++end_bit;
    size_t temp = 0;
    cub::DoubleBuffer<int32_t> k(key_a, key_b), v(idx_a, idx_b);
    LLP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp, k, v, (int)E, 0, end_bit, stream));
    if (4 * slab + temp > workspace_bytes) return LLP_E_WORKSPACE;
    LLP_CUDA(cub::DeviceRadixSort::SortPairs(cub_temp, temp, k, v, (int)E, 0, end_bit, stream));
    count_launch(4);
    sorted_key = k.Current();
    sorted_idx = v.Current();
    csr_finish_kernel<<<(unsigned)blocks, threads, 0, stream>>>(edge_val, sorted_idx, E, col);
    LLP_LAUNCH_OK();
    if (perm != nullptr) LLP_CUDA(cudaMemcpyAsync(perm, sorted_idx, (size_t)E * 4, cudaMemcpyDeviceToDevice, stream));
  }
  {
    int threads = 256;
    int64_t blocks = ceil_div(N + 1, threads);
    csr_rowptr_kernel<<<(unsigned)blocks, threads, 0, stream>>>(sorted_key, E, N, rowptr, inv_deg);
    LLP_LAUNCH_OK();
  }
  return 0;
}

