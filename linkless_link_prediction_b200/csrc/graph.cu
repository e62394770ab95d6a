// CSR construction + CSR gather-reduce SpMM (SAGE mean aggregation, forward and transpose-backward).
//
// Replaces the gather -> scatter-mean path PyG takes for a dense [2,E] edge_index
// (reference: src/models.py:113,118 -> SAGEConv.propagate).
//
// SpMM work decomposition (HBM-bound; no tensor cores on purpose):
//   * edges are cut into nominal chunks of kEPW; chunk c OWNS the rows whose first edge falls in
//     [c*kEPW, (c+1)*kEPW)  ->  one warp streams ~kEPW neighbour rows with 128-bit loads per lane,
//     kUnroll independent gathers in flight, fp32 accumulation in CSR order, segmented flush at
//     row boundaries (warp-uniform branches only);
//   * rows longer than kHub are split along the chunk grid: every chunk writes an fp32 partial for
//     its slice and a fix-up pass adds them in chunk order  ->  power-law hubs cannot stall a warp
//     and the result stays deterministic (no atomics);
//   * rows without edges are zero-filled by the fix-up pass.
#include <cub/cub.cuh>

#include "common.cuh"

namespace llp {

constexpr int kEPW = 64;    // nominal edges per warp-chunk
constexpr int kHub = 512;   // rows with more edges than this are split (must be >= kEPW)
constexpr int kSpmmThreads = 256;

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

// ------------------------------------------------------------------------------------------------
// SpMM plan
// ------------------------------------------------------------------------------------------------
namespace llp {

__global__ void spmm_plan_kernel(const int32_t* __restrict__ rowptr, int64_t N, int64_t n_chunks,
                                 int32_t* __restrict__ first_row) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c > n_chunks) return;
  if (c == n_chunks) { first_row[c] = (int32_t)N; return; }
  int64_t target = c * kEPW;
  int64_t lo = 0, hi = N;  // first r in [0,N] with rowptr[r] >= target (rowptr[N] = E >= target)
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((int64_t)rowptr[mid] < target) lo = mid + 1; else hi = mid;
  }
  first_row[c] = (int32_t)lo;
}

// ------------------------------------------------------------------------------------------------
// SpMM main kernel: one warp per chunk
// ------------------------------------------------------------------------------------------------
template <typename T, int VE, int NV>
struct RowAcc {
  float a[NV][VE];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int i = 0; i < VE; ++i) a[k][i] = 0.0f;
  }
};

// load the lane's slice of one source row (NV vectors of VE elements, 32*VE*NV columns per pass)
template <typename T, int VE, int NV>
__device__ __forceinline__ void load_row(const T* __restrict__ x, int64_t ldx, int64_t F, int src, int col0, int lane,
                                         uint4 (&v)[NV]) {
  const T* row = x + (int64_t)src * ldx;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    int c = col0 + (k * 32 + lane) * VE;
    if constexpr (VE * sizeof(T) == 16) {
      v[k] = (c < F) ? ldg_nc_v4(row + c) : make_uint4(0, 0, 0, 0);
    } else {  // scalar path: one element per lane
      float f = (c < F) ? to_f32(row[c]) : 0.0f;
      v[k] = make_uint4(__float_as_uint(f), 0, 0, 0);
    }
  }
}

template <typename T, int VE, int NV>
__device__ __forceinline__ void add_row(RowAcc<T, VE, NV>& acc, const uint4 (&v)[NV], float scale) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    if constexpr (VE * sizeof(T) == 16) {
      float f[VE];
      unpack16(v[k], f, T());
#pragma unroll
      for (int i = 0; i < VE; ++i) acc.a[k][i] = fmaf(f[i], scale, acc.a[k][i]);
    } else {
      acc.a[k][0] = fmaf(__uint_as_float(v[k].x), scale, acc.a[k][0]);
    }
  }
}

template <typename T, int VE, int NV>
__device__ __forceinline__ void store_row(T* __restrict__ out, int64_t ldo, int64_t F, int r, int col0, int lane,
                                          const RowAcc<T, VE, NV>& acc, float divisor) {
  T* row = out + (int64_t)r * ldo;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    int c = col0 + (k * 32 + lane) * VE;
    if (c < F) {
      if constexpr (VE * sizeof(T) == 16) {
        float f[VE];
#pragma unroll
        for (int i = 0; i < VE; ++i) f[i] = __fdiv_rn(acc.a[k][i], divisor);
        stg_v4(row + c, pack16(f, T()));
      } else {
        row[c] = from_f32<T>(__fdiv_rn(acc.a[k][0], divisor));
      }
    }
  }
}

template <typename T, int VE, int NV>
__device__ __forceinline__ void store_partial(float* __restrict__ p, int64_t F, int col0, int lane,
                                              const RowAcc<T, VE, NV>& acc) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    int c = col0 + (k * 32 + lane) * VE;
#pragma unroll
    for (int i = 0; i < VE; ++i)
      if (c + i < F) p[c + i] = acc.a[k][i];
  }
}

// Accumulate edges [e0, e1) (all belonging to ONE row) into acc.
template <typename T, int VE, int NV, int U>
__device__ __forceinline__ void accumulate_range(const T* __restrict__ x, int64_t ldx, int64_t F,
                                                 const int32_t* __restrict__ col, const float* __restrict__ src_scale,
                                                 int64_t e0, int64_t e1, int col0, int lane, RowAcc<T, VE, NV>& acc) {
  for (int64_t base = e0; base < e1; base += 32) {
    int cnt = (int)min((int64_t)32, e1 - base);
    int my = (lane < cnt) ? __ldg(col + base + lane) : 0;
    float mys = (src_scale != nullptr && lane < cnt) ? __ldg(src_scale + my) : 1.0f;
    for (int j0 = 0; j0 < cnt; j0 += U) {
      uint4 v[U][NV];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        int src = __shfl_sync(0xffffffffu, my, (j0 + u) & 31);
        if (j0 + u < cnt) load_row<T, VE, NV>(x, ldx, F, src, col0, lane, v[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float sc = __shfl_sync(0xffffffffu, mys, (j0 + u) & 31);
        if (j0 + u < cnt) add_row<T, VE, NV>(acc, v[u], sc);
      }
    }
  }
}

template <typename T, int VE, int NV, int U>
__global__ void __launch_bounds__(kSpmmThreads)
spmm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ first_row,
            int64_t n_chunks, const T* __restrict__ x, int64_t ldx, int64_t F, const float* __restrict__ src_scale,
            int mean, T* __restrict__ out, int64_t ldo, float* __restrict__ partial) {
  const int lane = threadIdx.x & 31;
  const int64_t c = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (c >= n_chunks) return;
  const int64_t cb = c * kEPW, ce = cb + kEPW;  // nominal edge range of this chunk
  const int r_begin = first_row[c], r_end = first_row[c + 1];
  constexpr int kColsPerPass = 32 * VE * NV;

  for (int col0 = 0; col0 < F; col0 += kColsPerPass) {
    RowAcc<T, VE, NV> acc;
    // (1) continuation slice of a hub row that started in an earlier chunk
    if (r_begin > 0) {
      int rp = r_begin - 1;
      int64_t ps = rowptr[rp], pe = rowptr[rp + 1];
      if (pe > cb && pe - ps > kHub) {
        acc.zero();
        accumulate_range<T, VE, NV, U>(x, ldx, F, col, src_scale, cb, min(pe, ce), col0, lane, acc);
        store_partial<T, VE, NV>(partial + (c * 2 + 0) * F, F, col0, lane, acc);
      }
    }
    if (r_begin >= r_end) continue;
    // (2) owned rows.  A hub row can only be the last owned row (it runs past the chunk end).
    int r_last = r_end - 1;
    int64_t ls = rowptr[r_last], le = rowptr[r_last + 1];
    const bool last_is_hub = (le - ls) > kHub;
    const int r_stream_end = last_is_hub ? r_last : r_end;  // rows [r_begin, r_stream_end) are streamed whole
    if (r_begin < r_stream_end) {
      int64_t e0 = rowptr[r_begin], e1 = rowptr[r_stream_end];
      int r = r_begin;
      int64_t row_start = e0, row_end = rowptr[r + 1];
      acc.zero();
      for (int64_t base = e0; base < e1; base += 32) {
        int cnt = (int)min((int64_t)32, e1 - base);
        int my = (lane < cnt) ? __ldg(col + base + lane) : 0;
        float mys = (src_scale != nullptr && lane < cnt) ? __ldg(src_scale + my) : 1.0f;
        for (int j0 = 0; j0 < cnt; j0 += U) {
          uint4 v[U][NV];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            int src = __shfl_sync(0xffffffffu, my, (j0 + u) & 31);
            if (j0 + u < cnt) load_row<T, VE, NV>(x, ldx, F, src, col0, lane, v[u]);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            float sc = __shfl_sync(0xffffffffu, mys, (j0 + u) & 31);
            if (j0 + u < cnt) {
              int64_t e = base + j0 + u;
              while (e == row_end) {  // warp-uniform: close finished rows (empty rows are skipped here)
                if (row_end > row_start) {
                  store_row<T, VE, NV>(out, ldo, F, r, col0, lane, acc, mean ? (float)(row_end - row_start) : 1.0f);
                  acc.zero();
                }
                ++r;
                row_start = row_end;
                row_end = rowptr[r + 1];
              }
              add_row<T, VE, NV>(acc, v[u], sc);
            }
          }
        }
      }
      // close the last streamed row(s)
      while (r < r_stream_end) {
        if (row_end > row_start) {
          store_row<T, VE, NV>(out, ldo, F, r, col0, lane, acc, mean ? (float)(row_end - row_start) : 1.0f);
          acc.zero();
        }
        ++r;
        if (r < r_stream_end) { row_start = row_end; row_end = rowptr[r + 1]; }
      }
    }
    if (last_is_hub) {
      acc.zero();
      accumulate_range<T, VE, NV, U>(x, ldx, F, col, src_scale, ls, min(le, ce), col0, lane, acc);
      store_partial<T, VE, NV>(partial + (c * 2 + 1) * F, F, col0, lane, acc);
    }
  }
}

// Fix-up: (a) combine hub-row partials in chunk order, (b) zero-fill rows without edges.
template <typename T>
__global__ void __launch_bounds__(kSpmmThreads)
spmm_fixup_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ first_row, int64_t n_chunks,
                  int64_t N, int64_t F, int mean, T* __restrict__ out, int64_t ldo, const float* __restrict__ partial) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // (a) chunk c is the FIRST continuation chunk of hub row rp  <=>  rp started in chunk c-1
  for (int64_t c = w + 1; c < n_chunks; c += n_warps) {
    int r0 = first_row[c];
    if (r0 == 0) continue;
    int rp = r0 - 1;
    int64_t ps = rowptr[rp], pe = rowptr[rp + 1];
    if (!(pe > c * kEPW && pe - ps > kHub && ps / kEPW == c - 1)) continue;
    int64_t c_last = (pe - 1) / kEPW;
    float divisor = mean ? (float)(pe - ps) : 1.0f;
    for (int64_t f = lane; f < F; f += 32) {
      float acc = partial[((c - 1) * 2 + 1) * F + f];
      for (int64_t cc = c; cc <= c_last; ++cc) acc += partial[(cc * 2 + 0) * F + f];
      out[(int64_t)rp * ldo + f] = from_f32<T>(__fdiv_rn(acc, divisor));
    }
  }
  // (b) zero rows: each warp inspects 32 rows at a time
  for (int64_t rb = w * 32; rb < N; rb += n_warps * 32) {
    int64_t r = rb + lane;
    bool empty = (r < N) && (rowptr[r + 1] == rowptr[r]);
    unsigned m = __ballot_sync(0xffffffffu, empty);
    while (m) {
      int j = __ffs(m) - 1;
      m &= m - 1;
      T* row = out + (rb + j) * ldo;
      for (int64_t f = lane; f < F; f += 32) row[f] = from_f32<T>(0.0f);
    }
  }
}

template <typename T>
static int spmm_launch(const int32_t* rowptr, const int32_t* col, const int32_t* first_row, int64_t N, int64_t E,
                       const void* x_, int64_t ldx, int64_t F, const float* src_scale, int mean, void* out_,
                       int64_t ldo, void* ws, cudaStream_t stream) {
  const T* x = reinterpret_cast<const T*>(x_);
  T* out = reinterpret_cast<T*>(out_);
  float* partial = reinterpret_cast<float*>(ws);
  int64_t n_chunks = llp_spmm_num_chunks(E);
  constexpr int VE = Vec16<T>::n;
  bool vec = aligned(x, 16) && aligned(out, 16) && (ldx * sizeof(T)) % 16 == 0 && (ldo * sizeof(T)) % 16 == 0 &&
             F % VE == 0;
  int64_t blocks = ceil_div(n_chunks * 32, kSpmmThreads);
  if (E > 0) {
    if (vec) {
      if (F * (int64_t)sizeof(T) <= 512)
        spmm_kernel<T, VE, 1, 8><<<(unsigned)blocks, kSpmmThreads, 0, stream>>>(rowptr, col, first_row, n_chunks, x, ldx, F, src_scale, mean, out, ldo, partial);
      else
        spmm_kernel<T, VE, 2, 4><<<(unsigned)blocks, kSpmmThreads, 0, stream>>>(rowptr, col, first_row, n_chunks, x, ldx, F, src_scale, mean, out, ldo, partial);
    } else {
      spmm_kernel<T, 1, 4, 4><<<(unsigned)blocks, kSpmmThreads, 0, stream>>>(rowptr, col, first_row, n_chunks, x, ldx, F, src_scale, mean, out, ldo, partial);
    }
    LLP_LAUNCH_OK();
  }
  int64_t fix_warps = max(n_chunks, ceil_div(N, 32));
  int64_t fix_blocks = imin64(ceil_div(fix_warps * 32, kSpmmThreads), (int64_t)kNumSMs * 16);
  spmm_fixup_kernel<T><<<(unsigned)imax64(fix_blocks, 1), kSpmmThreads, 0, stream>>>(rowptr, first_row, n_chunks, N, F, mean, out, ldo, partial);
  LLP_LAUNCH_OK();
  return 0;
}

}  // namespace llp

extern "C" int64_t llp_spmm_num_chunks(int64_t E) { return E <= 0 ? 1 : ceil_div(E, kEPW); }

extern "C" int llp_spmm_plan(const int32_t* rowptr, int64_t N, int64_t E, int32_t* chunk_first_row, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(rowptr && chunk_first_row && N >= 0 && E >= 0);
  if (int rc = check_device()) return rc;
  int64_t n_chunks = llp_spmm_num_chunks(E);
  spmm_plan_kernel<<<(unsigned)ceil_div(n_chunks + 1, 256), 256, 0, stream>>>(rowptr, N, n_chunks, chunk_first_row);
  LLP_LAUNCH_OK();
  return 0;
}

extern "C" size_t llp_spmm_workspace_bytes(int64_t E, int64_t F) {
  return (size_t)llp_spmm_num_chunks(E) * 2 * (size_t)(F > 0 ? F : 1) * sizeof(float);
}

extern "C" int llp_spmm(int dtype, const int32_t* rowptr, const int32_t* col, const int32_t* chunk_first_row,
                        int64_t N, int64_t E, const void* x, int64_t ldx, int64_t F, const float* src_scale, int mean,
                        void* out, int64_t ldo, void* workspace, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(rowptr && chunk_first_row && N >= 0 && E >= 0 && F > 0 && ldx >= F && ldo >= F);
  LLP_CHECK_ARG((E == 0 || (col && x && workspace)) && (N == 0 || out));
  if (int rc = check_device()) return rc;
  if (N == 0) return 0;
  if (dtype == LLP_F32) return spmm_launch<float>(rowptr, col, chunk_first_row, N, E, x, ldx, F, src_scale, mean, out, ldo, workspace, stream);
  if (dtype == LLP_BF16) return spmm_launch<__nv_bfloat16>(rowptr, col, chunk_first_row, N, E, x, ldx, F, src_scale, mean, out, ldo, workspace, stream);
  return LLP_E_BADARG;
}
