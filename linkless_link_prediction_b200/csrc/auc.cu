// ROC-AUC on device as exact integer pair counts.
// Replaces sklearn.metrics.roc_auc_score on host copies of the scores (train_teacher_gnn.py:147-153,251-266):
//     AUC = ( #{(p, n): s_n < s_p} + 0.5 * #{(p, n): s_n == s_p} ) / (n_pos * n_neg)
// which is what the trapezoid over the ROC curve evaluates to (Mann-Whitney U with ties at half weight).
// The negatives are radix-sorted once (order-preserving integer keys); every positive then takes a lower and an
// upper bound in the sorted keys (the list is L2-resident: 100,000 negatives = 400 KB).  Both counts are integers, so
// the result is bit-exact, independent of the order of the scores and of how positives are sharded across ranks.
#include <cub/cub.cuh>

#include "common.cuh"

namespace llp {
namespace auc {

// ascending order-preserving key; -0.0 and +0.0 compare equal (as the float comparison sklearn's argsort/diff uses)
__device__ __forceinline__ uint32_t score_key(float f) {
  uint32_t u = __float_as_uint(f + 0.0f);  // -0.0f + 0.0f == +0.0f
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void keys_kernel(const float* __restrict__ x, int64_t n, uint32_t* __restrict__ keys) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    keys[i] = score_key(x[i]);
}

// first index whose key is >= k (strict == false) or > k (strict == true)
__device__ __forceinline__ int64_t bound(const uint32_t* __restrict__ keys, int64_t n, uint32_t k, bool strict) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    uint32_t v = __ldg(keys + mid);
    bool go_right = strict ? (v <= k) : (v < k);
    if (go_right) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void pairs_kernel(const float* __restrict__ pos, int64_t n_pos, const uint32_t* __restrict__ neg_keys,
                             int64_t n_neg, unsigned long long* __restrict__ pairs) {
  unsigned long long less = 0, equal = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_pos; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t k = score_key(pos[i]);
    int64_t lb = bound(neg_keys, n_neg, k, false);
    int64_t ub = bound(neg_keys, n_neg, k, true);
    less += (unsigned long long)lb;
    equal += (unsigned long long)(ub - lb);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    less += __shfl_xor_sync(0xffffffffu, less, o);
    equal += __shfl_xor_sync(0xffffffffu, equal, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (less) atomicAdd(&pairs[0], less);  // integer atomics: order-independent, hence deterministic
    if (equal) atomicAdd(&pairs[1], equal);
  }
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t sort_temp_bytes(int64_t n) {
  size_t temp = 0;
  cub::DoubleBuffer<uint32_t> k(nullptr, nullptr);
  cudaError_t e = cub::DeviceRadixSort::SortKeys(nullptr, temp, k, (int)n, 0, 32);
  if (e != cudaSuccess) {
    cudaGetLastError();
    temp = ((size_t)16 << 20) + (size_t)n / 64;
  }
  return temp;
}

}  // namespace auc
}  // namespace llp

using namespace llp;

extern "C" size_t llp_auc_workspace_bytes(int64_t n_neg) {
  if (n_neg < 0) n_neg = 0;
  return 2 * auc::align256((size_t)n_neg * sizeof(uint32_t) + 4) + auc::align256(auc::sort_temp_bytes(n_neg)) + 256;
}

extern "C" int llp_auc_pairs(const float* pos, int64_t n_pos, const float* neg, int64_t n_neg, int64_t* pairs,
                             void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(pairs && n_pos >= 0 && n_neg >= 0 && (n_pos == 0 || pos) && (n_neg == 0 || neg));
  if (n_neg >= ((int64_t)1 << 31)) return LLP_E_SHAPE;
  if (int rc = check_device()) return rc;
  LLP_CUDA(cudaMemsetAsync(pairs, 0, 2 * sizeof(int64_t), stream));
  if (n_pos == 0 || n_neg == 0) return 0;
  LLP_CHECK_ARG(workspace);
  if (workspace_bytes < llp_auc_workspace_bytes(n_neg)) return LLP_E_WORKSPACE;
  char* base = reinterpret_cast<char*>(workspace);
  const size_t key_bytes = auc::align256((size_t)n_neg * sizeof(uint32_t) + 4);
  uint32_t* key_a = reinterpret_cast<uint32_t*>(base);
  uint32_t* key_b = reinterpret_cast<uint32_t*>(base + key_bytes);
  void* cub_temp = base + 2 * key_bytes;
  size_t temp = auc::sort_temp_bytes(n_neg);
  unsigned blocks = (unsigned)imin64(ceil_div(n_neg, 256), (int64_t)kNumSMs * 8);
  auc::keys_kernel<<<blocks, 256, 0, stream>>>(neg, n_neg, key_a);
  LLP_LAUNCH_OK();
  cub::DoubleBuffer<uint32_t> k(key_a, key_b);
  LLP_CUDA(cub::DeviceRadixSort::SortKeys(cub_temp, temp, k, (int)n_neg, 0, 32, stream));
  count_launch(3);
  blocks = (unsigned)imin64(ceil_div(n_pos, 128), (int64_t)kNumSMs * 16);
  auc::pairs_kernel<<<blocks, 128, 0, stream>>>(pos, n_pos, k.Current(), n_neg,
                                                reinterpret_cast<unsigned long long*>(pairs));
  LLP_LAUNCH_OK();
  return 0;
}
