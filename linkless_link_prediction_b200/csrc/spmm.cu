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
//     to all lanes, kU independent 128-bit row gathers in flight per lane, fp32 accumulation in CSR
//     order, warp-uniform flush at row boundaries;
//   * rows longer than kHub are split along the chunk grid: every chunk writes an fp32 partial for its
//     slice and the fix-up pass adds the partials in chunk order -> a power-law hub cannot serialise a
//     warp and the result has no atomics (deterministic);
//   * rows without edges are zero-filled by the fix-up pass.
// Register budget is kept <= 64/thread (one streaming loop, 32-bit indices) so 32 warps/SM stay resident:
// the kernel lives on memory-level parallelism.
#include "common.cuh"

namespace llp {

constexpr int kEPW = 128;   // nominal edges per warp-chunk
constexpr int kHub = 512;   // rows with more edges than this are split (must be >= kEPW)
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
__device__ __forceinline__ void add_row(RowAcc<VE, NV>& acc, const uint4 (&v)[NV], float scale) {
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

// close a row segment: whole rows go to `out` (divided by the degree for the mean), hub slices to `partial`
template <typename T, int VE, int NV>
__device__ __forceinline__ void flush(const RowAcc<VE, NV>& acc, T* __restrict__ out, int64_t ldo, int F, int r, int col0,
                                      int lane, float divisor, float* __restrict__ part) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    int c = col0 + (k * 32 + lane) * VE;
    if (c >= F) continue;
    if (part != nullptr) {
#pragma unroll
      for (int i = 0; i < VE; ++i) part[c + i] = acc.a[k][i];
    } else {
      T* row = out + (int64_t)r * ldo;
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

// A row [rs, re) is done for this chunk: rows up to kHub edges were streamed whole and go to `out`; the slice of a
// longer (hub) row goes to this chunk's partial slot (0: the row started in an earlier chunk, 1: it starts here).
template <typename T, int VE, int NV>
__device__ __forceinline__ void close_row(const RowAcc<VE, NV>& acc, T* __restrict__ out, int64_t ldo, int F, int r,
                                          int col0, int lane, int mean, int rs, int re, int cb, int c,
                                          float* __restrict__ partial) {
  const int deg = re - rs;
  if (deg > kHub) {
    flush<T, VE, NV>(acc, out, ldo, F, r, col0, lane, 1.0f, partial + ((int64_t)c * 2 + (rs < cb ? 0 : 1)) * F);
  } else {
    flush<T, VE, NV>(acc, out, ldo, F, r, col0, lane, mean ? (float)deg : 1.0f, nullptr);
  }
}

template <typename T, int VE, int NV, int U>
__global__ void __launch_bounds__(kSpmmThreads, 2048 / kSpmmThreads / 2)
spmm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ first_row,
            int n_chunks, const T* __restrict__ x, int64_t ldx, int F, const float* __restrict__ src_scale, int mean,
            T* __restrict__ out, int64_t ldo, float* __restrict__ partial) {
  const int lane = threadIdx.x & 31;
  const int c = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (c >= n_chunks) return;
  const int cb = c * kEPW, ce = cb + kEPW;  // nominal edge range of this chunk (E < 2^31)
  const int r_begin = first_row[c], r_end = first_row[c + 1];

  // The edge range [eb, ee) this warp streams, and the row r0 containing eb.
  int r0 = r_begin, eb, ee;
  bool head_is_continuation = false;
  {
    int own_b = rowptr[r_begin];  // == E when the chunk owns no rows
    eb = own_b;
    if (r_begin > 0) {
      int ps = rowptr[r_begin - 1];
      if (own_b > cb && own_b - ps > kHub) {  // a hub row that started in an earlier chunk runs into this one
        r0 = r_begin - 1;
        eb = cb;
        head_is_continuation = true;
      }
    }
    if (r_end > r_begin) {
      int ls = rowptr[r_end - 1], le = rowptr[r_end];
      ee = (le - ls > kHub) ? min(le, ce) : le;  // an owned hub row is streamed only up to the chunk end
    } else {
      ee = head_is_continuation ? min(own_b, ce) : eb;
    }
  }
  if (ee <= eb) return;
  constexpr int kColsPerPass = 32 * VE * NV;

  for (int col0 = 0; col0 < F; col0 += kColsPerPass) {
    RowAcc<VE, NV> acc;
    acc.zero();
    int r = r0;
    int row_start = head_is_continuation ? rowptr[r0] : eb;
    int row_end = rowptr[r + 1];
    for (int base = eb; base < ee; base += 32) {
      const int cnt = min(32, ee - base);
      const int my = (lane < cnt) ? __ldg(col + base + lane) : 0;
      const float mys = (src_scale != nullptr && lane < cnt) ? __ldg(src_scale + my) : 1.0f;
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
            const int e = base + j0 + u;
            while (e == row_end) {  // warp-uniform: close finished rows, skip rows without edges
              if (row_end > row_start) {
                close_row<T, VE, NV>(acc, out, ldo, F, r, col0, lane, mean, row_start, row_end, cb, c, partial);
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
    // the last open row: complete (it ends exactly at ee) or a hub slice
    close_row<T, VE, NV>(acc, out, ldo, F, r, col0, lane, mean, row_start, row_end, cb, c, partial);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Bulk-copy variant (the default for 16-byte-aligned rows of up to 1 KB): every lane issues ONE cp.async.bulk
// (global -> shared, mbarrier complete_tx) for the neighbour row of "its" edge, so a single warp instruction puts up
// to 16 whole feature rows in flight without holding them in registers; two such batches are double-buffered per
// warp (16 KB of smem per warp, ~190 KB of gathers in flight per SM).  The warp then streams the rows out of shared
// memory (one conflict-free 128-bit LDS per lane per row) with the same segmented fp32 accumulation as above.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kBulkWarps = 4;
constexpr int kBulkBufBytes = 8192;  // per buffer; 2 buffers per warp

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

template <typename T, int NV>
__global__ void __launch_bounds__(kBulkWarps * 32)
spmm_bulk_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ first_row,
                 int n_chunks, const T* __restrict__ x, int64_t ldx, int F, const float* __restrict__ src_scale, int mean,
                 T* __restrict__ out, int64_t ldo, float* __restrict__ partial) {
  constexpr int VE = Vec16<T>::n;
  constexpr int kSlot = 512 * NV;                // bytes reserved per staged row
  constexpr int kBatch = kBulkBufBytes / kSlot;  // rows per buffer: 16 (NV=1) or 8 (NV=2)
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[kBulkWarps][2];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int c = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (c >= n_chunks) return;
  const int cb = c * kEPW, ce = cb + kEPW;
  const int r_begin = first_row[c], r_end = first_row[c + 1];
  int r0 = r_begin, eb, ee;
  bool head_is_continuation = false;
  {
    int own_b = rowptr[r_begin];
    eb = own_b;
    if (r_begin > 0) {
      int ps = rowptr[r_begin - 1];
      if (own_b > cb && own_b - ps > kHub) { r0 = r_begin - 1; eb = cb; head_is_continuation = true; }
    }
    if (r_end > r_begin) {
      int ls = rowptr[r_end - 1], le = rowptr[r_end];
      ee = (le - ls > kHub) ? min(le, ce) : le;
    } else {
      ee = head_is_continuation ? min(own_b, ce) : eb;
    }
  }
  if (ee <= eb) return;

  const uint32_t row_bytes = (uint32_t)F * sizeof(T);
  const uint32_t buf0 = smem_u32(smem) + (uint32_t)wib * 2 * kBulkBufBytes;
  const uint32_t bar0 = smem_u32(&bars[wib][0]), bar1 = smem_u32(&bars[wib][1]);
  if (lane == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const int n_batches = (ee - eb + kBatch - 1) / kBatch;
  float sc_even = 1.0f, sc_odd = 1.0f;  // per-lane src_scale of the edge this lane fetched, per buffer
  auto issue = [&](int b) {
    const int base = eb + b * kBatch;
    const int cnt = min(kBatch, ee - base);
    const uint32_t bar = (b & 1) ? bar1 : bar0;
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)cnt * row_bytes);
    float sc = 1.0f;
    if (lane < cnt) {
      const int src = __ldg(col + base + lane);
      if (src_scale != nullptr) sc = __ldg(src_scale + src);
      bulk_g2s(buf0 + (uint32_t)(b & 1) * kBulkBufBytes + (uint32_t)lane * kSlot, x + (int64_t)src * ldx, row_bytes, bar);
    }
    if (b & 1) sc_odd = sc; else sc_even = sc;
  };
  issue(0);
  if (n_batches > 1) issue(1);

  RowAcc<VE, NV> acc;
  acc.zero();
  int r = r0;
  int row_start = head_is_continuation ? rowptr[r0] : eb;
  int row_end = rowptr[r + 1];
  uint32_t phase0 = 0, phase1 = 0;
  for (int b = 0; b < n_batches; ++b) {
    const int base = eb + b * kBatch;
    const int cnt = min(kBatch, ee - base);
    if (b & 1) { mbar_wait(bar1, phase1); phase1 ^= 1; } else { mbar_wait(bar0, phase0); phase0 ^= 1; }
    const uint32_t buf = buf0 + (uint32_t)(b & 1) * kBulkBufBytes + (uint32_t)lane * 16;
    const float my_sc = (b & 1) ? sc_odd : sc_even;
    constexpr int U = 4;
    for (int j0 = 0; j0 < cnt; j0 += U) {
      uint4 v[U][NV];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < NV; ++k)
          v[u][k] = (j0 + u < cnt && (uint32_t)(lane * 16 + k * 512) < row_bytes) ? lds_v4(buf + (uint32_t)(j0 + u) * kSlot + k * 512)
                                                                                : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float sc = __shfl_sync(0xffffffffu, my_sc, (j0 + u) & 31);
        if (j0 + u < cnt) {
          const int e = base + j0 + u;
          while (e == row_end) {
            if (row_end > row_start) {
              close_row<T, VE, NV>(acc, out, ldo, F, r, 0, lane, mean, row_start, row_end, cb, c, partial);
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
    __syncwarp();                        // every lane is done reading this buffer
    if (b + 2 < n_batches) issue(b + 2);  // refill it
  }
  close_row<T, VE, NV>(acc, out, ldo, F, r, 0, lane, mean, row_start, row_end, cb, c, partial);
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
    for (int f = lane; f < F; f += 32) {
      float acc = partial[((int64_t)(c - 1) * 2 + 1) * F + f];
      int cc = c;
      for (; cc + 8 <= c_last + 1; cc += 8) {  // 8 independent loads in flight, added in chunk order
        float t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = __ldg(partial + ((int64_t)(cc + i) * 2) * F + f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += t[i];
      }
      for (; cc <= c_last; ++cc) acc += partial[((int64_t)cc * 2) * F + f];
      out[(int64_t)rp * ldo + f] = from_f32<T>(__fdiv_rn(acc, divisor));
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

bool g_spmm_force_register_path = false;  // llp_spmm_set_path(): tests exercise both implementations

template <typename T>
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
  if (E > 0 && vec && F * (int64_t)sizeof(T) <= 1024 && !g_spmm_force_register_path) {
    const unsigned bblocks = (unsigned)ceil_div((int64_t)n_chunks, kBulkWarps);
    const size_t smem = (size_t)kBulkWarps * 2 * kBulkBufBytes;
    if (F * (int64_t)sizeof(T) <= 512) {
      auto kern = spmm_bulk_kernel<T, 1>;
      static bool once = false;
      if (!once) { LLP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); once = true; }
      kern<<<bblocks, kBulkWarps * 32, smem, stream>>>(rowptr, col, first_row, n_chunks, x, ldx, (int)F, src_scale, mean, out, ldo, partial);
    } else {
      auto kern = spmm_bulk_kernel<T, 2>;
      static bool once = false;
      if (!once) { LLP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); once = true; }
      kern<<<bblocks, kBulkWarps * 32, smem, stream>>>(rowptr, col, first_row, n_chunks, x, ldx, (int)F, src_scale, mean, out, ldo, partial);
    }
    LLP_LAUNCH_OK();
  } else if (E > 0) {
    if (vec) {
      if (F * (int64_t)sizeof(T) <= 512)
        spmm_kernel<T, VE, 1, 8><<<blocks, kSpmmThreads, 0, stream>>>(rowptr, col, first_row, n_chunks, x, ldx, (int)F, src_scale, mean, out, ldo, partial);
      else
        spmm_kernel<T, VE, 2, 4><<<blocks, kSpmmThreads, 0, stream>>>(rowptr, col, first_row, n_chunks, x, ldx, (int)F, src_scale, mean, out, ldo, partial);
    } else {
      spmm_kernel<T, 1, 4, 4><<<blocks, kSpmmThreads, 0, stream>>>(rowptr, col, first_row, n_chunks, x, ldx, (int)F, src_scale, mean, out, ldo, partial);
    }
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

extern "C" void llp_spmm_set_path(int force_register_path) { g_spmm_force_register_path = force_register_path != 0; }

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
  if (dtype == LLP_F32) return spmm_launch<float>(rowptr, col, chunk_first_row, N, E, x, ldx, F, src_scale, mean, out, ldo, workspace, stream);
  if (dtype == LLP_BF16) return spmm_launch<__nv_bfloat16>(rowptr, col, chunk_first_row, N, E, x, ldx, F, src_scale, mean, out, ldo, workspace, stream);
  return LLP_E_BADARG;
}
