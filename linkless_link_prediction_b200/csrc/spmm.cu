// CSR gather-reduce SpMM: SAGE mean aggregation, forward and transpose-backward.
//
// Replaces the gather -> scatter-mean path PyG takes for a dense [2,E] edge_index
// (reference: src/models.py:113,118 -> SAGEConv.propagate -> torch_scatter.scatter(reduce='mean'),
// and its autograd index_add_).
//
// Work decomposition (HBM-bound byte shuffling; no tensor cores on purpose):
//   * edges are cut into nominal chunks of kEPW; chunk c OWNS the rows whose first edge falls in
//     [c*kEPW, (c+1)*kEPW).  Because rows are contiguous in edge space, everything a chunk must touch
//     is ONE contiguous edge range, streamed by one warp: 32 column indices per coalesced load, shuffled
//     to all lanes, U independent 128-bit row gathers in flight per lane, fp32 accumulation in CSR
//     order, warp-uniform flush at row boundaries (a group of U edges inside one row takes a predicate-free path);
//   * rows longer than kHub are split along the chunk grid: every chunk writes an fp32 partial for its
//     slice and the fix-up pass adds the partials in chunk order -> a power-law hub cannot serialise a
//     warp and the result has no atomics (deterministic);
//   * rows without edges are zero-filled by the fix-up pass.
// Design notes from the round-1 measurements (profiles/r01_spmm_*): the first version spent 76 instructions per edge
// (inlined IEEE divisions + store code in the streaming loop, I-cache misses) and ran at 8 warps/SM because of 158
// registers; staging rows through shared memory with cp.async.bulk (one TMA op per row) or cp.async/LDGSTS (16 B per
// lane) was measured too and was no faster (12 warps/SM, issue-bound), so the kernel keeps the gathered rows in
// registers, holds the budget at <= 64 registers (32 warps/SM) and keeps the per-row epilogue out of line.
#include "common.cuh"

namespace llp {

constexpr int kEPW = 128;   // nominal edges per warp-chunk
constexpr int kHub = 256;   // rows with more edges than this are split along the chunk grid (must be >= kEPW)
constexpr int kSpmmThreads = 128;

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

// ---- per-lane accumulator over NV vectors of VE elements --------------------------------------------
template <int VE, int NV>
struct RowAcc {
  float a[NV][VE];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int i = 0; i < VE; ++i) a[k][i] = 0.0f;
  }
};

template <typename T, int VE, int NV>
__device__ __forceinline__ void load_row(const T* __restrict__ x, int64_t ldx, int F, int src, int col0, int lane,
                                         uint4 (&v)[NV]) {
  const T* row = x + (int64_t)src * ldx + col0 + lane * VE;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = col0 + (k * 32 + lane) * VE;
    if constexpr (VE * sizeof(T) == 16) {
      v[k] = (c < F) ? ldg_nc_v4(row + k * 32 * VE) : make_uint4(0, 0, 0, 0);
    } else {  // scalar path: one element per lane
      float f = (c < F) ? to_f32(row[k * 32]) : 0.0f;
      v[k] = make_uint4(__float_as_uint(f), 0, 0, 0);
    }
  }
}

template <typename T, int VE, int NV, bool kScale>
__device__ __forceinline__ void add_row(RowAcc<VE, NV>& acc, const uint4 (&v)[NV], float scale) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    if constexpr (VE * sizeof(T) == 16) {
      float f[VE];
      unpack16(v[k], f, T());
#pragma unroll
      for (int i = 0; i < VE; ++i) acc.a[k][i] = kScale ? fmaf(f[i], scale, acc.a[k][i]) : acc.a[k][i] + f[i];
    } else {
      const float f = __uint_as_float(v[k].x);
      acc.a[k][0] = kScale ? fmaf(f, scale, acc.a[k][0]) : acc.a[k][0] + f;
    }
  }
}

// A row [rs, re) is done for this chunk.  Rows of up to kHub edges were streamed whole and go to `out` (mean: divided
// by the degree — an exact IEEE division in fp32 so the result is bit-identical to scatter-mean's true_divide, a
// reciprocal multiply for bf16 outputs).  The slice of a longer (hub) row goes to this chunk's fp32 partial slot
// (0: the row started in an earlier chunk, 1: it starts here).  Not inlined on purpose: it runs once per row, and
// keeping it out of the streaming loop keeps that loop small enough for the instruction cache.
template <typename T, int VE, int NV>
__device__ __noinline__ void close_row(const RowAcc<VE, NV>& acc, T* __restrict__ out, int64_t ldo, int F, int r, int col0,
                                       int lane, int mean, int rs, int re, int cb, int c, float* __restrict__ partial) {
  const int deg = re - rs;
  if (deg > kHub) {
    float* part = partial + ((int64_t)c * 2 + (rs < cb ? 0 : 1)) * F;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int col = col0 + (k * 32 + lane) * VE;
      if (col < F) {
#pragma unroll
        for (int i = 0; i < VE; ++i) part[col + i] = acc.a[k][i];
      }
    }
    return;
  }
  T* row = out + (int64_t)r * ldo;
  const float d = mean ? (float)deg : 1.0f;
  const float inv = 1.0f / d;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int col = col0 + (k * 32 + lane) * VE;
    if (col >= F) continue;
    float f[VE];
#pragma unroll
    for (int i = 0; i < VE; ++i) f[i] = sizeof(T) == 4 ? __fdiv_rn(acc.a[k][i], d) : acc.a[k][i] * inv;
    if constexpr (VE * sizeof(T) == 16) {
      stg_v4(row + col, pack16(f, T()));
    } else {
      row[col] = from_f32<T>(f[0]);
    }
  }
}

// The edge range [eb, ee) a chunk streams, the row r0 containing eb and whether that row is a hub continuation.
struct ChunkRange {
  int r0, eb, ee, row_start;
};
__device__ __forceinline__ ChunkRange chunk_range(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ first_row,
                                                  int c) {
  const int cb = c * kEPW, ce = cb + kEPW;
  const int r_begin = first_row[c], r_end = first_row[c + 1];
  ChunkRange cr;
  const int own_b = rowptr[r_begin];  // == E when the chunk owns no rows
  cr.r0 = r_begin;
  cr.eb = own_b;
  cr.row_start = own_b;
  bool cont = false;
  if (r_begin > 0) {
    const int ps = rowptr[r_begin - 1];
    if (own_b > cb && own_b - ps > kHub) {  // a hub row that started in an earlier chunk runs into this one
      cr.r0 = r_begin - 1;
      cr.eb = cb;
      cr.row_start = ps;
      cont = true;
    }
  }
  if (r_end > r_begin) {
    const int ls = rowptr[r_end - 1], le = rowptr[r_end];
    cr.ee = (le - ls > kHub) ? min(le, ce) : le;  // an owned hub row is streamed only up to the chunk end
  } else {
    cr.ee = cont ? min(own_b, ce) : cr.eb;
  }
  return cr;
}

template <typename T, int VE, int NV, int U, bool kScale, int kMinBlocks>
__global__ void __launch_bounds__(kSpmmThreads, kMinBlocks)
spmm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ first_row,
            int n_chunks, const T* __restrict__ x, int64_t ldx, int F, const float* __restrict__ src_scale, int mean,
            T* __restrict__ out, int64_t ldo, float* __restrict__ partial, int fake_seq_n) {
  const int lane = threadIdx.x & 31;
  const int c = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (c >= n_chunks) return;
  const ChunkRange cr = chunk_range(rowptr, first_row, c);
  if (cr.ee <= cr.eb) return;
  const int cb = c * kEPW;
  constexpr int kColsPerPass = 32 * VE * NV;

  for (int col0 = 0; col0 < F; col0 += kColsPerPass) {
    RowAcc<VE, NV> acc;
    acc.zero();
    int r = cr.r0;
    int row_start = cr.row_start;
    int row_end = rowptr[r + 1];
    for (int base = cr.eb; base < cr.ee; base += 32) {
      const int cnt = min(32, cr.ee - base);
      int my = (lane < cnt) ? __ldg(col + base + lane) : 0;
      if (fake_seq_n > 0) my = (base + lane) % fake_seq_n;
      float mys = 1.0f;
      if constexpr (kScale) mys = (lane < cnt) ? __ldg(src_scale + my) : 0.0f;
      int j0 = 0;
      // full groups of U edges: U independent row gathers in flight, no per-edge predicates
      for (; j0 + U <= cnt; j0 += U) {
        uint4 v[U][NV];
#pragma unroll
        for (int u = 0; u < U; ++u) load_row<T, VE, NV>(x, ldx, F, __shfl_sync(0xffffffffu, my, j0 + u), col0, lane, v[u]);
        const int e0 = base + j0;
        if (e0 + U <= row_end) {  // warp-uniform fast path: the whole group lies inside the open row
#pragma unroll
          for (int u = 0; u < U; ++u)
            add_row<T, VE, NV, kScale>(acc, v[u], kScale ? __shfl_sync(0xffffffffu, mys, j0 + u) : 1.0f);
        } else {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            while (e0 + u == row_end) {  // close finished rows, skip rows without edges
              if (row_end > row_start) {
                close_row<T, VE, NV>(acc, out, ldo, F, r, col0, lane, mean, row_start, row_end, cb, c, partial);
                acc.zero();
              }
              ++r;
              row_start = row_end;
              row_end = rowptr[r + 1];
            }
            add_row<T, VE, NV, kScale>(acc, v[u], kScale ? __shfl_sync(0xffffffffu, mys, j0 + u) : 1.0f);
          }
        }
      }
      // tail of the batch, one edge at a time
      for (; j0 < cnt; ++j0) {
        uint4 v[NV];
        load_row<T, VE, NV>(x, ldx, F, __shfl_sync(0xffffffffu, my, j0), col0, lane, v);
        const int e = base + j0;
        while (e == row_end) {
          if (row_end > row_start) {
            close_row<T, VE, NV>(acc, out, ldo, F, r, col0, lane, mean, row_start, row_end, cb, c, partial);
            acc.zero();
          }
          ++r;
          row_start = row_end;
          row_end = rowptr[r + 1];
        }
        add_row<T, VE, NV, kScale>(acc, v, kScale ? __shfl_sync(0xffffffffu, mys, j0) : 1.0f);
      }
    }
    // the last open row: complete (it ends exactly at ee) or a hub slice
    close_row<T, VE, NV>(acc, out, ldo, F, r, col0, lane, mean, row_start, row_end, cb, c, partial);
  }
}

// Fix-up: (a) combine hub-row partials in chunk order, (b) zero-fill rows without edges.
template <typename T>
__global__ void __launch_bounds__(256)
spmm_fixup_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ first_row, int n_chunks, int N, int F,
                  int mean, T* __restrict__ out, int64_t ldo, const float* __restrict__ partial) {
  const int lane = threadIdx.x & 31;
  const int w = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int n_warps = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
  // (a) chunk c is the FIRST continuation chunk of hub row rp  <=>  rp started in chunk c-1
  for (int c = w + 1; c < n_chunks; c += n_warps) {
    const int r0 = first_row[c];
    if (r0 == 0) continue;
    const int rp = r0 - 1;
    const int ps = rowptr[rp], pe = rowptr[rp + 1];
    if (!(pe > c * kEPW && pe - ps > kHub && ps / kEPW == c - 1)) continue;
    const int c_last = (pe - 1) / kEPW;
    const float divisor = mean ? (float)(pe - ps) : 1.0f;
    const float* first = partial + ((int64_t)(c - 1) * 2 + 1) * F;  // slot 1 of the owner chunk
    if (F % 4 == 0) {  // 128-bit loads, 8 partials in flight per lane, added in chunk order
      for (int f = lane * 4; f < F; f += 128) {
        float4 acc = *reinterpret_cast<const float4*>(first + f);
        int cc = c;
        for (; cc + 8 <= c_last + 1; cc += 8) {
          float4 t[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = __ldg(reinterpret_cast<const float4*>(partial + ((int64_t)(cc + i) * 2) * F + f));
#pragma unroll
          for (int i = 0; i < 8; ++i) { acc.x += t[i].x; acc.y += t[i].y; acc.z += t[i].z; acc.w += t[i].w; }
        }
        for (; cc <= c_last; ++cc) {
          float4 t = __ldg(reinterpret_cast<const float4*>(partial + ((int64_t)cc * 2) * F + f));
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        T* o = out + (int64_t)rp * ldo + f;
        o[0] = from_f32<T>(__fdiv_rn(acc.x, divisor)); o[1] = from_f32<T>(__fdiv_rn(acc.y, divisor));
        o[2] = from_f32<T>(__fdiv_rn(acc.z, divisor)); o[3] = from_f32<T>(__fdiv_rn(acc.w, divisor));
      }
    } else {
      for (int f = lane; f < F; f += 32) {
        float acc = first[f];
        for (int cc = c; cc <= c_last; ++cc) acc += partial[((int64_t)cc * 2) * F + f];
        out[(int64_t)rp * ldo + f] = from_f32<T>(__fdiv_rn(acc, divisor));
      }
    }
  }
  // (b) zero rows: each warp inspects 32 rows at a time
  for (int rb = w * 32; rb < N; rb += n_warps * 32) {
    const int r = rb + lane;
    const bool empty = (r < N) && (rowptr[r + 1] == rowptr[r]);
    unsigned m = __ballot_sync(0xffffffffu, empty);
    while (m) {
      const int j = __ffs(m) - 1;
      m &= m - 1;
      T* row = out + (int64_t)(rb + j) * ldo;
      for (int f = lane; f < F; f += 32) row[f] = from_f32<T>(0.0f);
    }
  }
}

int g_spmm_variant = 0;
int g_spmm_chunk_div = 1;  // experiment: process only the first n_chunks/div chunks
int g_spmm_fake_seq = 0;   // experiment: gather row (edge id mod N) instead of col[e]

template <typename T, bool kScale>
static int spmm_launch(const int32_t* rowptr, const int32_t* col, const int32_t* first_row, int64_t N, int64_t E,
                       const void* x_, int64_t ldx, int64_t F, const float* src_scale, int mean, void* out_,
                       int64_t ldo, void* ws, cudaStream_t stream) {
  const T* x = reinterpret_cast<const T*>(x_);
  T* out = reinterpret_cast<T*>(out_);
  float* partial = reinterpret_cast<float*>(ws);
  const int n_chunks = (int)llp_spmm_num_chunks(E);
  constexpr int VE = Vec16<T>::n;
  const bool vec = aligned(x, 16) && aligned(out, 16) && (ldx * sizeof(T)) % 16 == 0 && (ldo * sizeof(T)) % 16 == 0 &&
                   F % VE == 0;
  const unsigned blocks = (unsigned)ceil_div((int64_t)n_chunks * 32, kSpmmThreads);
  if (E > 0) {
#define LLP_SPMM_LAUNCH(VE_, NV_, U_, MB_) \
  spmm_kernel<T, VE_, NV_, U_, kScale, MB_><<<blocks, kSpmmThreads, 0, stream>>>(rowptr, col, first_row, n_chunks / g_spmm_chunk_div, x, ldx, (int)F, src_scale, mean, out, ldo, partial, g_spmm_fake_seq ? (int)N : 0)
    const int variant = g_spmm_variant;  // occupancy/register trade-off (llp_set_tuning(0, v)): 0 = 8 blocks/SM (<=64 regs)
    if (vec) {
      if (F * (int64_t)sizeof(T) <= 512) {
        if (variant == 1) LLP_SPMM_LAUNCH(VE, 1, 8, 6); else if (variant == 2) LLP_SPMM_LAUNCH(VE, 1, 8, 5); else if (variant == 3) LLP_SPMM_LAUNCH(VE, 1, 4, 8); else LLP_SPMM_LAUNCH(VE, 1, 8, 8);
      } else {
        if (variant == 1) LLP_SPMM_LAUNCH(VE, 2, 4, 6); else if (variant == 2) LLP_SPMM_LAUNCH(VE, 2, 4, 5); else if (variant == 3) LLP_SPMM_LAUNCH(VE, 2, 2, 8); else LLP_SPMM_LAUNCH(VE, 2, 4, 8);
      }
    } else {
      LLP_SPMM_LAUNCH(1, 4, 4, 8);
    }
#undef LLP_SPMM_LAUNCH
    LLP_LAUNCH_OK();
  }
  const int64_t fix_warps = imax64(n_chunks, ceil_div(N, 32));
  const int64_t fix_blocks = imin64(ceil_div(fix_warps * 32, 256), (int64_t)kNumSMs * 8);
  spmm_fixup_kernel<T><<<(unsigned)imax64(fix_blocks, 1), 256, 0, stream>>>(rowptr, first_row, n_chunks, (int)N, (int)F, mean, out, ldo, partial);
  LLP_LAUNCH_OK();
  return 0;
}

}  // namespace llp

using namespace llp;

extern "C" void llp_set_tuning(int key, int value) {
  if (key == 0) g_spmm_variant = value;
  if (key == 1) g_spmm_chunk_div = value < 1 ? 1 : value;
  if (key == 2) g_spmm_fake_seq = value;
}

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
  LLP_CHECK_ARG(E < (int64_t)INT32_MAX - kEPW && N < (int64_t)INT32_MAX && F < (1 << 24));
  if (int rc = check_device()) return rc;
  if (N == 0) return 0;
#define LLP_SPMM(T)                                                                                                     \
  return src_scale != nullptr                                                                                           \
             ? spmm_launch<T, true>(rowptr, col, chunk_first_row, N, E, x, ldx, F, src_scale, mean, out, ldo, workspace, stream)  \
             : spmm_launch<T, false>(rowptr, col, chunk_first_row, N, E, x, ldx, F, src_scale, mean, out, ldo, workspace, stream)
  if (dtype == LLP_F32) { LLP_SPMM(float); }
  if (dtype == LLP_BF16) { LLP_SPMM(__nv_bfloat16); }
#undef LLP_SPMM
  return LLP_E_BADARG;
}
