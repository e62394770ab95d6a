// Backward of the edge gather-Hadamard z[m] = h[u[m]] * h[v[m]] (autograd of the two fancy-index gathers + mul in front
// of LinkPredictor: train_teacher_gnn.py:58, main.py:186,214, models.py:140; PyTorch uses index_put_ with atomics):
//
//     gh[n,:] = sum_{m: u[m]==n} dz[m,:]*h[v[m],:]  +  sum_{m: v[m]==n} dz[m,:]*h[u[m],:]
//
// as a gather-REDUCE instead of a scatter.  `llp_edge_plan` stably radix-sorts the 2M (node, edge) incidences of the
// batch by node (it depends only on u and v, so the host runs it on a side stream while the encoder works); the backward
// kernel then walks node rows in order, adds each row's incidences in sorted order and writes the row once in the
// activation dtype (rows without incidences are written as zeros).  No atomics, no fp32 staging buffer, no separate
// zero-fill / cast passes; the result is bit-reproducible.  Rows with more than kHubThreshold incidences (hub nodes of
// a power-law graph) are deferred to a block-per-row kernel (8 warps take every 8th incidence, fixed order combine) so
// that no single warp serialises on a long row.
#include <cub/cub.cuh>

#include "common.cuh"

namespace llp {
namespace eb {

constexpr int kHubThreshold = 48;
constexpr int kHubWarpsMax = 32;  // warps of a hub block (fewer when F is wide: [warps][F] floats of shared memory)
constexpr int kRowsPerWarp = 16;

__global__ void incidence_keys_kernel(const int64_t* __restrict__ u, const int64_t* __restrict__ v, int64_t M,
                                      int32_t* __restrict__ key, int32_t* __restrict__ idx) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < 2 * M) {
    key[i] = (int32_t)(i < M ? u[i] : v[i - M]);
    idx[i] = (int32_t)i;
  }
}

// From the sorted incidences: rowptr[r] = first sorted position whose node is >= r, and per sorted position the pair
// meta[p] = (edge m, the OTHER endpoint of that edge), so that the backward kernel has no index chains to chase.
__global__ void incidence_finish_kernel(const int32_t* __restrict__ sorted_key, const int32_t* __restrict__ sorted_idx,
                                        const int64_t* __restrict__ u, const int64_t* __restrict__ v, int64_t M, int64_t N,
                                        int32_t* __restrict__ rowptr, int2* __restrict__ meta) {
  const int64_t E = 2 * M;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > E) return;
  const int64_t prev = i == 0 ? -1 : (int64_t)sorted_key[i - 1];
  const int64_t cur = i == E ? N : (int64_t)sorted_key[i];
  for (int64_t r = prev + 1; r <= cur; ++r) rowptr[r] = (int32_t)i;
  if (i < E) {
    const int e = sorted_idx[i];
    const int64_t m = e < M ? e : e - M;
    meta[i] = make_int2((int)m, (int)(e < M ? v[m] : u[m]));
  }
}

template <typename T>
struct RowArgs {
  const T* h; int64_t ldh;
  const T* dz; int64_t lddz;
  const int32_t* rowptr; const int2* meta;
  T* gh; int64_t ldgh;
  int64_t N; int F;
  int32_t* hub_list; int32_t* hub_count;
};

// acc[:] += dz[m, c:c+VE] * h[other, c:c+VE] for sorted positions [p0, p1) stepping by `step`, four gathers in flight
template <typename T, int VE>
__device__ __forceinline__ void accumulate(const RowArgs<T>& a, int p0, int p1, int step, int c, float (&acc)[VE]) {
  int p = p0;
  for (; p + 3 * step < p1; p += 4 * step) {
    uint4 dv[4], hv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int2 e = __ldg(a.meta + p + q * step);
      dv[q] = ldg_nc_v4(a.dz + (int64_t)e.x * a.lddz + c);
      hv[q] = ldg_gather_v4(a.h + (int64_t)e.y * a.ldh + c);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float d[VE], x[VE];
      unpack16(dv[q], d, T());
      unpack16(hv[q], x, T());
#pragma unroll
      for (int i = 0; i < VE; ++i) acc[i] = fmaf(d[i], x[i], acc[i]);
    }
  }
  for (; p < p1; p += step) {
    const int2 e = __ldg(a.meta + p);
    float d[VE], x[VE];
    unpack16(ldg_nc_v4(a.dz + (int64_t)e.x * a.lddz + c), d, T());
    unpack16(ldg_gather_v4(a.h + (int64_t)e.y * a.ldh + c), x, T());
#pragma unroll
    for (int i = 0; i < VE; ++i) acc[i] = fmaf(d[i], x[i], acc[i]);
  }
}

// One warp per kRowsPerWarp consecutive node rows.  The row pointers of the group are one coalesced load, the
// (edge, other endpoint) pairs of all its incidences another (32 per round, one per lane, broadcast by shuffle), and
// the row gathers are issued four incidences at a time; rows are flushed (written once, zeros when empty) as the walk
// crosses their end.  A group that contains a hub row takes the row-by-row path and defers the hub to the hub kernel.
template <typename T>
__global__ void __launch_bounds__(256) hadamard_bwd_rows_kernel(const RowArgs<T> a) {
  constexpr int VE = Vec16<T>::n;
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t r0 = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) * kRowsPerWarp;
  if (r0 >= a.N) return;
  const int nrows = (int)(a.N - r0 < kRowsPerWarp ? a.N - r0 : kRowsPerWarp);
  const int rp = __ldg(a.rowptr + r0 + (lane < nrows ? lane : nrows));
  const int rp_next = __shfl_down_sync(kFull, rp, 1);
  const unsigned hubs = __ballot_sync(kFull, lane < nrows && rp_next - rp > kHubThreshold);
  const int p_begin = __shfl_sync(kFull, rp, 0), p_end = __shfl_sync(kFull, rp, nrows);

  if (hubs != 0) {  // rare: row by row, hub rows are only recorded
    for (int row = 0; row < nrows; ++row) {
      const int s = __shfl_sync(kFull, rp, row), e = __shfl_sync(kFull, rp, row + 1);
      if ((hubs >> row) & 1u) {
        if (lane == 0) a.hub_list[atomicAdd(a.hub_count, 1)] = (int32_t)(r0 + row);  // slot order does not matter
        continue;
      }
      for (int c = lane * VE; c < a.F; c += 32 * VE) {
        float acc[VE];
#pragma unroll
        for (int i = 0; i < VE; ++i) acc[i] = 0.0f;
        accumulate<T, VE>(a, s, e, 1, c, acc);
        stg_v4(a.gh + (r0 + row) * a.ldgh + c, pack16(acc, T()));
      }
    }
    return;
  }

  for (int c0 = 0; c0 < a.F; c0 += 32 * VE) {
    const int c = c0 + lane * VE;
    const bool active = c < a.F;
    float acc[VE];
#pragma unroll
    for (int i = 0; i < VE; ++i) acc[i] = 0.0f;
    int row = 0;
    int row_end = __shfl_sync(kFull, rp, 1);
    auto flush = [&]() {
      if (active) stg_v4(a.gh + (r0 + row) * a.ldgh + c, pack16(acc, T()));
#pragma unroll
      for (int i = 0; i < VE; ++i) acc[i] = 0.0f;
      ++row;
      row_end = __shfl_sync(kFull, rp, row + 1 <= nrows ? row + 1 : nrows);
    };
    for (int base = p_begin; base < p_end; base += 32) {
      const int cnt = p_end - base < 32 ? p_end - base : 32;
      int2 mine = make_int2(0, 0);
      if (lane < cnt) mine = __ldg(a.meta + base + lane);
      for (int j0 = 0; j0 < cnt; j0 += 4) {
        uint4 dv[4], hv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int jj = j0 + q < cnt ? j0 + q : cnt - 1;   // clamped duplicates are loaded but never added
          const int m = __shfl_sync(kFull, mine.x, jj), o = __shfl_sync(kFull, mine.y, jj);
          if (active) {
            dv[q] = ldg_nc_v4(a.dz + (int64_t)m * a.lddz + c);
            hv[q] = ldg_gather_v4(a.h + (int64_t)o * a.ldh + c);
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (j0 + q < cnt) {
            const int p = base + j0 + q;
            while (p >= row_end) flush();     // warp-uniform
            if (active) {
              float d[VE], x[VE];
              unpack16(dv[q], d, T());
              unpack16(hv[q], x, T());
#pragma unroll
              for (int i = 0; i < VE; ++i) acc[i] = fmaf(d[i], x[i], acc[i]);
            }
          }
        }
      }
    }
    while (row < nrows) flush();
  }
}

// one block per hub row: warp w of nw adds sorted positions p0+w, p0+w+nw, ...; the nw partial rows are combined in warp
// order (deterministic for a given F, which fixes nw)
template <typename T>
__global__ void __launch_bounds__(32 * kHubWarpsMax) hadamard_bwd_hubs_kernel(const RowArgs<T> a) {
  constexpr int VE = Vec16<T>::n;
  extern __shared__ float part[];  // [nw][F]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int n_hubs = *a.hub_count;
  for (int i = blockIdx.x; i < n_hubs; i += gridDim.x) {
    const int64_t r = a.hub_list[i];
    const int p0 = __ldg(a.rowptr + r), p1 = __ldg(a.rowptr + r + 1);
    for (int c = lane * VE; c < a.F; c += 32 * VE) {
      float acc[VE];
#pragma unroll
      for (int k = 0; k < VE; ++k) acc[k] = 0.0f;
      accumulate<T, VE>(a, p0 + w, p1, nw, c, acc);
#pragma unroll
      for (int k = 0; k < VE; ++k) part[w * a.F + c + k] = acc[k];
    }
    __syncthreads();
    for (int c = threadIdx.x * VE; c < a.F; c += blockDim.x * VE) {
      float s[VE];
#pragma unroll
      for (int k = 0; k < VE; ++k) {
        float t = 0.0f;
        for (int q = 0; q < nw; ++q) t += part[q * a.F + c + k];
        s[k] = t;
      }
      stg_v4(a.gh + r * a.ldgh + c, pack16(s, T()));
    }
    __syncthreads();
  }
}

// scalar fallback (rows not 16-byte multiples): one warp per row, any length, lanes stride over columns
template <typename T>
__global__ void __launch_bounds__(256) hadamard_bwd_rows_scalar_kernel(const RowArgs<T> a) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= a.N) return;
  const int p0 = a.rowptr[r], p1 = a.rowptr[r + 1];
  for (int c = lane; c < a.F; c += 32) {
    float acc = 0.0f;
    for (int p = p0; p < p1; ++p) {
      const int2 e = a.meta[p];
      acc = fmaf(to_f32(a.dz[(int64_t)e.x * a.lddz + c]), to_f32(a.h[(int64_t)e.y * a.ldh + c]), acc);
    }
    a.gh[r * a.ldgh + c] = from_f32<T>(acc);
  }
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t sort_temp_bytes(int64_t E) {
  size_t temp = 0;
  cub::DoubleBuffer<int32_t> k(nullptr, nullptr), v(nullptr, nullptr);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, temp, k, v, (int)E, 0, 32);
  if (e != cudaSuccess) {
    cudaGetLastError();
    temp = ((size_t)16 << 20) + (size_t)E / 64;
  }
  return temp;
}

struct Workspace {
  int32_t *key_a, *key_b, *idx_a, *idx_b;
  void* cub_temp;
  size_t cub_bytes, total;
};

static Workspace carve(char* base, int64_t M) {
  const int64_t E = 2 * M;
  Workspace w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align256(bytes); return p; };
  w.key_a = (int32_t*)take((size_t)E * 4 + 4);
  w.key_b = (int32_t*)take((size_t)E * 4 + 4);
  w.idx_a = (int32_t*)take((size_t)E * 4 + 4);
  w.idx_b = (int32_t*)take((size_t)E * 4 + 4);
  w.cub_bytes = sort_temp_bytes(E);
  w.cub_temp = take(w.cub_bytes);
  w.total = off + 256;
  return w;
}

template <typename T>
static int run(const void* h, int64_t ldh, int64_t F, int64_t M, const void* dz, int64_t lddz, int64_t N,
               const int32_t* rowptr, const int32_t* meta, void* gh, int64_t ldgh, int32_t* hub_ws, cudaStream_t stream) {
  LLP_CUDA(cudaMemsetAsync(hub_ws, 0, sizeof(int32_t), stream));
  RowArgs<T> a{(const T*)h, ldh, (const T*)dz, lddz, rowptr, reinterpret_cast<const int2*>(meta), (T*)gh, ldgh, N, (int)F,
               hub_ws + 64, hub_ws};
  const bool vec = aligned(h, 16) && aligned(dz, 16) && aligned(gh, 16) && (ldh * sizeof(T)) % 16 == 0 &&
                   (lddz * sizeof(T)) % 16 == 0 && (ldgh * sizeof(T)) % 16 == 0 && F % Vec16<T>::n == 0;
  if (!vec) {
    hadamard_bwd_rows_scalar_kernel<T><<<(unsigned)ceil_div(N * 32, 256), 256, 0, stream>>>(a);
    LLP_LAUNCH_OK();
    return 0;
  }
  hadamard_bwd_rows_kernel<T><<<(unsigned)ceil_div(ceil_div(N, kRowsPerWarp) * 32, 256), 256, 0, stream>>>(a);
  LLP_LAUNCH_OK();
  int hub_warps = (int)((48 * 1024) / (F * sizeof(float)));
  hub_warps = hub_warps > kHubWarpsMax ? kHubWarpsMax : hub_warps;
  if (hub_warps < 1) return LLP_E_SHAPE;
  const size_t smem = (size_t)hub_warps * F * sizeof(float);
  hadamard_bwd_hubs_kernel<T><<<kNumSMs * 2, 32 * hub_warps, smem, stream>>>(a);
  LLP_LAUNCH_OK();
  return 0;
}

}  // namespace eb
}  // namespace llp

using namespace llp;

extern "C" size_t llp_edge_plan_workspace_bytes(int64_t M) { return M < 0 ? 256 : eb::carve(nullptr, M).total; }

extern "C" int llp_edge_plan(const int64_t* u, const int64_t* v, int64_t M, int64_t N, int32_t* rowptr, int32_t* meta,
                             void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(M >= 0 && N >= 0 && rowptr && workspace && (M == 0 || (u && v && meta)));
  LLP_CHECK_ARG(2 * M < (int64_t)INT32_MAX && N < (int64_t)INT32_MAX);
  if (int rc = check_device()) return rc;
  if (workspace_bytes < llp_edge_plan_workspace_bytes(M)) return LLP_E_WORKSPACE;
  const int64_t E = 2 * M;
  eb::Workspace w = eb::carve(reinterpret_cast<char*>(workspace), M);
  cub::DoubleBuffer<int32_t> k(w.key_a, w.key_b), x(w.idx_a, w.idx_b);
  if (E > 0) {
    eb::incidence_keys_kernel<<<(unsigned)ceil_div(E, 256), 256, 0, stream>>>(u, v, M, w.key_a, w.idx_a);
    LLP_LAUNCH_OK();
    int end_bit = 1;
    while (end_bit < 31 && ((int64_t)1 << end_bit) < N) ++end_bit;
    size_t temp = w.cub_bytes;
    LLP_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, temp, k, x, (int)E, 0, end_bit, stream));
    count_launch(3);
  }
  eb::incidence_finish_kernel<<<(unsigned)ceil_div(E + 1, 256), 256, 0, stream>>>(
      k.Current(), x.Current(), u, v, M, N, rowptr, reinterpret_cast<int2*>(meta));
  LLP_LAUNCH_OK();
  return 0;
}

extern "C" size_t llp_edge_hadamard_bwd_workspace_bytes(int64_t M) {
  return (size_t)(64 + (M > 0 ? 2 * M : 0) / eb::kHubThreshold + 2) * sizeof(int32_t);
}

extern "C" int llp_edge_hadamard_bwd(int dtype, const void* h, int64_t ldh, int64_t F, int64_t M, const void* dz,
                                     int64_t lddz, int64_t N, const int32_t* rowptr, const int32_t* meta, void* gh,
                                     int64_t ldgh, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(F > 0 && M >= 0 && N >= 0 && F <= 1536);
  if (int rc = check_device()) return rc;
  if (N == 0) return 0;
  LLP_CHECK_ARG(gh && ldgh >= F && workspace && rowptr);
  LLP_CHECK_ARG(M == 0 || (h && dz && meta && ldh >= F && lddz >= F));
  if (workspace_bytes < llp_edge_hadamard_bwd_workspace_bytes(M)) return LLP_E_WORKSPACE;
  int32_t* hub_ws = reinterpret_cast<int32_t*>(workspace);
  if (dtype == LLP_F32) return eb::run<float>(h, ldh, F, M, dz, lddz, N, rowptr, meta, gh, ldgh, hub_ws, stream);
  if (dtype == LLP_BF16) return eb::run<__nv_bfloat16>(h, ldh, F, M, dz, lddz, N, rowptr, meta, gh, ldgh, hub_ws, stream);
  return LLP_E_BADARG;
}
