// Hits@K on device: exact top-kmax of the negative scores by MSB-first radix select, then a
// strict-greater count of the positives against every K's threshold.
// Replaces ogb Evaluator._eval_hits (torch.topk on host tensors), train_teacher_gnn.py:120-145,226-249.
// Counts are integers => bit-exact and independent of the sharding.
#include <math_constants.h>

#include "common.cuh"

namespace llp {

// order-preserving float -> uint32 key (larger float <=> larger key); NaNs are not expected (sigmoid outputs)
__device__ __forceinline__ uint32_t float_key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

struct SelectState {
  uint32_t prefix;       // key bits fixed so far (high bits)
  uint32_t k_remaining;  // rank still to resolve inside the prefix bucket
  uint32_t n_greater;    // candidates collected so far (counter for the collect pass)
  uint32_t pad;
  uint32_t hist[4][256];
};

__global__ void select_init_kernel(SelectState* st, uint32_t k) {
  int i = threadIdx.x;
  if (i == 0) { st->prefix = 0; st->k_remaining = k; st->n_greater = 0; st->pad = 0; }
  for (int p = 0; p < 4; ++p) st->hist[p][i] = 0;
}

// histogram of digit `pass` (8 bits, MSB first) over elements matching the prefix on the higher bits
__global__ void select_hist_kernel(const float* __restrict__ x, int64_t n, int pass, SelectState* st) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int shift = 24 - 8 * pass;
  const uint32_t prefix = st->prefix;
  const uint32_t mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t k = float_key(x[i]);
    if ((k & mask) == prefix) atomicAdd(&h[(k >> shift) & 0xff], 1u);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&st->hist[pass][threadIdx.x], h[threadIdx.x]);
}

// pick the digit that contains the k-th largest; single thread (256 bins)
__global__ void select_pick_kernel(int pass, SelectState* st) {
  if (threadIdx.x != 0) return;
  const int shift = 24 - 8 * pass;
  uint32_t k = st->k_remaining, above = 0;
  int d = 255;
  for (; d > 0; --d) {
    uint32_t c = st->hist[pass][d];
    if (above + c >= k) break;
    above += c;
  }
  st->prefix |= ((uint32_t)d << shift);
  st->k_remaining = k - above;
}

// append every score strictly greater than the k-th value
__global__ void select_collect_kernel(const float* __restrict__ x, int64_t n, SelectState* st, float* __restrict__ cand,
                                      uint32_t cap) {
  const uint32_t kth = st->prefix;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = x[i];
    if (float_key(v) > kth) {
      uint32_t slot = atomicAdd(&st->n_greater, 1u);
      if (slot < cap) cand[slot] = v;
    }
  }
}

// single block: bitonic sort (descending) of the candidates padded with the k-th value (or -inf when n < kmax)
__global__ void select_sort_kernel(const float* __restrict__ cand, const float* __restrict__ all, int64_t n,
                                   const SelectState* st, int kmax, int pow2, float* __restrict__ out) {
  extern __shared__ float sv[];
  const bool small = n <= kmax;
  uint32_t ng = small ? (uint32_t)n : min(st->n_greater, (uint32_t)kmax);
  float fill = small ? -CUDART_INF_F : key_float(st->prefix);
  for (int i = threadIdx.x; i < pow2; i += blockDim.x) {
    float v;
    if (i < (int)ng) v = small ? all[i] : cand[i];
    else if (i < kmax) v = fill;
    else v = -CUDART_INF_F;
    sv[i] = v;
  }
  __syncthreads();
  for (int size = 2; size <= pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < pow2; i += blockDim.x) {
        int j = i ^ stride;
        if (j > i) {
          bool desc = (i & size) == 0;
          float a = sv[i], b = sv[j];
          // total order through the integer keys (handles -0.0 / +0.0 consistently)
          bool swap = desc ? (float_key(a) < float_key(b)) : (float_key(a) > float_key(b));
          if (swap) { sv[i] = b; sv[j] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < kmax; i += blockDim.x) out[i] = sv[i];
}

constexpr int kMaxThr = 16;
struct Thresholds { float v[kMaxThr]; };

__global__ void count_greater_kernel(const float* __restrict__ pos, int64_t n, const float* __restrict__ thr, int n_thr,
                                     unsigned long long* __restrict__ counts) {
  __shared__ float t[kMaxThr];
  if (threadIdx.x < n_thr) t[threadIdx.x] = thr[threadIdx.x];
  __syncthreads();
  uint32_t c[kMaxThr];
#pragma unroll
  for (int k = 0; k < kMaxThr; ++k) c[k] = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = pos[i];
#pragma unroll
    for (int k = 0; k < kMaxThr; ++k)
      if (k < n_thr && v > t[k]) ++c[k];
  }
#pragma unroll
  for (int k = 0; k < kMaxThr; ++k) {
    if (k < n_thr) {
      uint32_t s = c[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if ((threadIdx.x & 31) == 0 && s) atomicAdd(&counts[k], (unsigned long long)s);
    }
  }
}

}  // namespace llp

using namespace llp;

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }
constexpr int64_t kMaxTopK = 8192;

extern "C" size_t llp_topk_workspace_bytes(int64_t n, int64_t kmax) {
  (void)n;
  return sizeof(SelectState) + 256 + (size_t)(kmax > 0 ? kmax : 1) * sizeof(float);
}

extern "C" int llp_topk_desc(const float* scores, int64_t n, int64_t kmax, float* out, void* workspace,
                             size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(out && workspace && n >= 0 && kmax > 0 && (n == 0 || scores));
  if (kmax > kMaxTopK) return LLP_E_SHAPE;
  if (workspace_bytes < llp_topk_workspace_bytes(n, kmax)) return LLP_E_WORKSPACE;
  if (int rc = check_device()) return rc;
  SelectState* st = reinterpret_cast<SelectState*>(workspace);
  float* cand = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ((sizeof(SelectState) + 255) & ~(size_t)255));
  int pow2 = next_pow2((int)kmax);
  if (n > kmax) {
    select_init_kernel<<<1, 256, 0, stream>>>(st, (uint32_t)kmax);
    LLP_LAUNCH_OK();
    unsigned blocks = (unsigned)imin64(ceil_div(n, 256), (int64_t)kNumSMs * 8);
    for (int pass = 0; pass < 4; ++pass) {
      select_hist_kernel<<<blocks, 256, 0, stream>>>(scores, n, pass, st);
      LLP_LAUNCH_OK();
      select_pick_kernel<<<1, 32, 0, stream>>>(pass, st);
      LLP_LAUNCH_OK();
    }
    select_collect_kernel<<<blocks, 256, 0, stream>>>(scores, n, st, cand, (uint32_t)kmax);
    LLP_LAUNCH_OK();
  }
  select_sort_kernel<<<1, 1024, (size_t)pow2 * sizeof(float), stream>>>(cand, scores, n, st, (int)kmax, pow2, out);
  LLP_LAUNCH_OK();
  return 0;
}

extern "C" int llp_count_greater(const float* pos, int64_t n_pos, const float* thresholds, int64_t n_thr,
                                 int64_t* counts, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(thresholds && counts && n_pos >= 0 && n_thr > 0 && n_thr <= kMaxThr && (n_pos == 0 || pos));
  if (int rc = check_device()) return rc;
  LLP_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_thr * sizeof(int64_t), stream));
  if (n_pos == 0) return 0;
  unsigned blocks = (unsigned)imin64(ceil_div(n_pos, 256), (int64_t)kNumSMs * 8);
  count_greater_kernel<<<blocks, 256, 0, stream>>>(pos, n_pos, thresholds, (int)n_thr, reinterpret_cast<unsigned long long*>(counts));
  LLP_LAUNCH_OK();
  return 0;
}
