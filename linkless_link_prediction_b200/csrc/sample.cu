// Sampling kernels.
// (1) Uniform random walks for LLP context sampling.  Replaces torch_cluster 1.6.0's
// uniform_sampling_kernel behind random_walk(row, col, start, walk_length, coalesced=False)
// (reference: src/main.py:37,43,45).  Bit-exact given the same `rand` tensor.
// (2) The device half of PyG 2.2.0's dense negative_sampling (train_teacher_gnn.py:50-51, main.py:81-82,206-207): the
//     candidates come from CPython's random.sample on the host; PyG indexes an N*N - N boolean mask with them, keeps the
//     unmasked ones in order, cuts to the requested count and de-linearises.  Here: membership in the SORTED ids of the
//     existing edges (no mask), an ordered compaction and the (row, col) split in one call whose only output the host
//     waits for is the kept count.
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

// flag[i] = 1 if cand[i] is not an existing edge id (binary search in the sorted ids); flag[k] = 0
__global__ void negative_flags_kernel(const int64_t* __restrict__ cand, int64_t k, const int64_t* __restrict__ taken, int64_t n_taken,
                                      int32_t* __restrict__ flag) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > k) return;
  if (i == k) { flag[i] = 0; return; }
  const int64_t v = cand[i];
  int64_t lo = 0, hi = n_taken;   // first position with taken[pos] >= v
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(taken + mid) < v) lo = mid + 1; else hi = mid;
  }
  flag[i] = (lo < n_taken && __ldg(taken + lo) == v) ? 0 : 1;
}

// pos = exclusive scan of the flags: kept candidate i goes to slot pos[i] (the first max_out of them are written)
__global__ void negative_scatter_kernel(const int64_t* __restrict__ cand, int64_t k, const int32_t* __restrict__ pos, int64_t num_nodes,
                                        int64_t max_out, int64_t* __restrict__ kept, int64_t* __restrict__ edges,
                                        int32_t* __restrict__ count) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i == 0) *count = pos[k];
  if (i >= k) return;
  const int32_t p = pos[i];
  if (pos[i + 1] == p || p >= max_out) return;
  const int64_t v = cand[i];
  kept[p] = v;
  if (edges != nullptr) {   // vector_to_edge_index: r = id // (N-1); c = id % (N-1); c += (r <= c)
    const int64_t r = v / (num_nodes - 1);
    int64_t c = v - r * (num_nodes - 1);
    if (r <= c) c += 1;
    edges[p] = r;
    edges[max_out + p] = c;
  }
}

}  // namespace llp

using namespace llp;

static size_t neg_align(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t llp_negative_filter_workspace_bytes(int64_t k) {
  if (k < 0) return 256;
  return neg_align((size_t)(k + 1) * 4) + neg_align((size_t)scan_i32_tiles(k + 1) * 4 + 4) + 256;
}

extern "C" int llp_negative_filter(const int64_t* cand, int64_t k, const int64_t* taken_sorted, int64_t n_taken,
                                   int64_t num_nodes, int64_t max_out, int64_t* kept, int64_t* edges, int32_t* count,
                                   void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(k >= 0 && n_taken >= 0 && num_nodes >= 2 && max_out >= 0 && count && workspace);
  LLP_CHECK_ARG(k == 0 || (cand && (max_out == 0 || kept)) );
  LLP_CHECK_ARG(n_taken == 0 || taken_sorted);
  LLP_CHECK_ARG(k + 1 < (int64_t)INT32_MAX);
  if (int rc = check_device()) return rc;
  if (workspace_bytes < llp_negative_filter_workspace_bytes(k)) return LLP_E_WORKSPACE;
  int32_t* flag = reinterpret_cast<int32_t*>(workspace);
  int32_t* tile_sum = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(workspace) + neg_align((size_t)(k + 1) * 4));
  negative_flags_kernel<<<(unsigned)ceil_div(k + 1, 256), 256, 0, stream>>>(cand, k, taken_sorted, n_taken, flag);
  LLP_LAUNCH_OK();
  if (int rc = exclusive_scan_i32(flag, k + 1, tile_sum, stream)) return rc;
  negative_scatter_kernel<<<(unsigned)ceil_div(k > 0 ? k : 1, 256), 256, 0, stream>>>(cand, k, flag, num_nodes, max_out, kept, edges, count);
  LLP_LAUNCH_OK();
  return 0;
}

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
