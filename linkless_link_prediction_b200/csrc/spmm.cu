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
//   * rows without edges are zero-filled by the fix-up pass (extra blocks of the same launch).
// Design notes from the round-1 measurements (profiles/r01_spmm_*): the first version spent 76 instructions per edge
// (inlined IEEE divisions + store code in the streaming loop, I-cache misses) and ran at 8 warps/SM because of 158
// registers; staging rows through shared memory with cp.async.bulk (one TMA op per row) or cp.async/LDGSTS (16 B per
// lane) was measured too and was no faster (12 warps/SM, issue-bound), so the kernel keeps the gathered rows in
// registers, holds the budget at <= 64 registers (32 warps/SM) and keeps the per-row epilogue out of line.
#include <type_traits>

#include "common.cuh"

namespace llp {

constexpr int kEPW = 64;    // nominal edges per warp-chunk (32 and 128 were measured: slower on the C4 graph)
constexpr int kHub = 64;    // rows with more edges than this are split along the chunk grid (>= kEPW; 128 and 256 measured: longer warp tails)
constexpr int kSpmmThreads = 128;

// Hub table: one record {c, row, row begin, row end} per row longer than kHub, appended by the FIRST continuation chunk
// c of that row (the row started in chunk c-1).  Record 0 is the header {n_small, n_big, index of big hub 0, -}: hubs of
// at most kFixWarpParts partials ("small": one warp of the fix-up kernel combines them) fill the records from 1 upwards,
// longer ones ("big": one block each) from the last record downwards.  The order inside either group does not matter
// (every hub is combined independently of the others).
constexpr int kFixWarpParts = 8;
__global__ void spmm_hub_list_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ first_row,
                                     int64_t n_chunks, int4* __restrict__ hub_tbl, int32_t* __restrict__ num_hubs) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x + 1;
  if (c >= n_chunks) return;
  const int r0 = first_row[c];
  if (r0 == 0) return;
  const int ps = rowptr[r0 - 1], pe = rowptr[r0];
  if (pe > c * kEPW && pe - ps > kHub && ps / kEPW == c - 1) {
    const int n_part = (pe - 1) / kEPW - (int)c + 2;  // slot 1 of chunk c-1, then slot 0 of chunks c .. c_last
    int32_t* hdr = reinterpret_cast<int32_t*>(hub_tbl);
    const int4 rec = make_int4((int)c, r0 - 1, ps, pe);
    if (n_part <= kFixWarpParts) hub_tbl[1 + atomicAdd(hdr + 0, 1)] = rec;
    else hub_tbl[n_chunks - atomicAdd(hdr + 1, 1)] = rec;
    atomicAdd(num_hubs, 1);
  }
}

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

// ---- lane vectors: 16 bytes (VE = 4 fp32 / 8 bf16), 8 bytes (2 fp32 / 4 bf16: rows of <= 256 bytes keep all 32 lanes
// busy that way) or one element (unaligned fallback) ------------------------------------------------------------------
template <typename T, int VE>
__device__ __forceinline__ uint4 load_vec(const T* p) {
  if constexpr (VE * sizeof(T) == 16) {
    return ldg_gather_v4(p);
  } else if constexpr (VE * sizeof(T) == 8) {
    uint2 r;
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return make_uint4(r.x, r.y, 0, 0);
  } else {
    return make_uint4(__float_as_uint(to_f32(*p)), 0, 0, 0);
  }
}
template <typename T, int VE>
__device__ __forceinline__ void unpack_vec(const uint4& v, float (&f)[VE]) {
  if constexpr (VE == 1) {
    f[0] = __uint_as_float(v.x);
  } else if constexpr (sizeof(T) == 4) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
    if constexpr (VE == 4) { f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w); }
  } else {
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
    if constexpr (VE == 8) {
      f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
      f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
    }
  }
}
template <typename T, int VE>
__device__ __forceinline__ void store_vec(T* p, const float (&f)[VE]) {
  if constexpr (VE == 1) {
    *p = from_f32<T>(f[0]);
  } else if constexpr (sizeof(T) == 4) {
    if constexpr (VE == 4) *reinterpret_cast<uint4*>(p) = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    else *reinterpret_cast<uint2*>(p) = make_uint2(__float_as_uint(f[0]), __float_as_uint(f[1]));
  } else {
    if constexpr (VE == 8) *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    else *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]));
  }
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
    v[k] = (c < F) ? load_vec<T, VE>(row + k * 32 * VE) : make_uint4(0, 0, 0, 0);
  }
}

// Packed fp32 arithmetic (sm_100: FADD2 / FFMA2, two IEEE round-to-nearest operations per issue slot).  The kernel is
// issue-bound (profiles/r01_spmm_ncu_full_summary.json: 71 % issue-active, ~25 warp instructions per edge), and the eight
// additions per 16-byte bf16 vector are its largest instruction group; each lane's result is bit-identical to the scalar
// FADD / FFMA it replaces.
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 ra, ra, rb;\n\t"
      "mov.b64 {%0, %1}, ra;\n\t}"
      : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fma2(float& a0, float& a1, float b0, float b1, float s) {
  asm("{\n\t.reg .b64 ra, rb, rs;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tmov.b64 rs, {%4, %4};\n\t"
      "fma.rn.f32x2 ra, rb, rs, ra;\n\tmov.b64 {%0, %1}, ra;\n\t}"
      : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1), "f"(s));
}

template <typename T, int VE, int NV, bool kScale>
__device__ __forceinline__ void add_row(RowAcc<VE, NV>& acc, const uint4 (&v)[NV], float scale) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float f[VE];
    unpack_vec<T, VE>(v[k], f);
    if constexpr (VE % 2 == 0) {
#pragma unroll
      for (int i = 0; i < VE; i += 2) {
        if constexpr (kScale) fma2(acc.a[k][i], acc.a[k][i + 1], f[i], f[i + 1], scale);
        else add2(acc.a[k][i], acc.a[k][i + 1], f[i], f[i + 1]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < VE; ++i) acc.a[k][i] = kScale ? fmaf(f[i], scale, acc.a[k][i]) : acc.a[k][i] + f[i];
    }
  }
}

// add_row for the streaming kernel: the bf16 -> fp32 unpack is an `asm volatile`, which keeps it next to its additions.
// Left to the compiler, the unpack of all U gathered rows is hoisted in front of the accumulation (32 extra live
// registers for U = 4: spills under the 64-register cap the kernel's occupancy needs).
template <typename T, int VE, int NV, bool kScale>
__device__ __forceinline__ void add_row_pinned(RowAcc<VE, NV>& acc, const uint4 (&v)[NV], float scale) {
  if constexpr (sizeof(T) == 2 && VE % 4 == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
      for (int i = 0; i < VE / 2; ++i) {
        float lo, hi;
        asm volatile("shl.b32 %0, %2, 16;\n\tand.b32 %1, %2, 0xffff0000;" : "=f"(lo), "=f"(hi) : "r"(w[i]));
        if constexpr (kScale) fma2(acc.a[k][2 * i], acc.a[k][2 * i + 1], lo, hi, scale);
        else add2(acc.a[k][2 * i], acc.a[k][2 * i + 1], lo, hi);
      }
    }
  } else {
    add_row<T, VE, NV, kScale>(acc, v, scale);
  }
}

// A row [rs, re) is done for this chunk.  Rows of up to kHub edges were streamed whole and go to `out` (mean: divided
// by the degree — an exact IEEE division in fp32 so the result is bit-identical to scatter-mean's true_divide, a
// reciprocal multiply for bf16 outputs).  The slice of a longer (hub) row goes to this chunk's fp32 partial slot
// (0: the row started in an earlier chunk, 1: it starts here).
template <typename T, int VE, int NV>
__device__ __forceinline__ void close_row(const RowAcc<VE, NV>& acc, T* __restrict__ out, int64_t ldo, int F, int r, int col0,
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
  float inv = 1.0f;   // bf16 outputs: reciprocal multiply (rcp.approx: 1 ulp of fp32, far below the bf16 rounding that follows)
  if (sizeof(T) == 2) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(d));
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int col = col0 + (k * 32 + lane) * VE;
    if (col >= F) continue;
    float f[VE];
#pragma unroll
    for (int i = 0; i < VE; ++i) f[i] = sizeof(T) == 4 ? __fdiv_rn(acc.a[k][i], d) : acc.a[k][i] * inv;
    store_vec<T, VE>(row + col, f);
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

// Chunk descriptors for the streaming kernel, written once per graph by llp_spmm_plan: what chunk_range() derives
// through a chain of three dependent loads (first_row -> rowptr -> rowptr) plus the first two row ends, as two 16-byte
// vectors per chunk {r0, eb, ee, row_start} {row_end, row_end_next, -, -}.  A warp lives for ~16 gather round trips; the
// descriptor takes four dependent round trips out of the front of every one of them.
__global__ void spmm_desc_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ first_row, int64_t N,
                                 int64_t n_chunks, int4* __restrict__ desc, int4* __restrict__ hub_tbl) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  if (c == 0) hub_tbl[0] = make_int4(0, 0, (int)n_chunks, 0);   // header of the hub table (spmm_hub_list_kernel)
  const ChunkRange cr = chunk_range(rowptr, first_row, (int)c);
  int4 d0 = make_int4(cr.r0, cr.eb, cr.ee, cr.row_start), d1 = make_int4(0, 0, 0, 0);
  if (cr.ee > cr.eb) {
    d1.x = rowptr[cr.r0 + 1];
    d1.y = rowptr[min((int64_t)cr.r0 + 2, N)];
  }
  desc[2 * c] = d0;
  desc[2 * c + 1] = d1;
}
// descriptors start at the first 32-byte boundary behind the [n_chunks + 1] first-row table
__host__ __device__ inline int64_t spmm_desc_offset_ints(int64_t n_chunks) { return (n_chunks + 1 + 7) / 8 * 8; }

// Row-run traversal: the warp walks its edge range row by row; inside a row, edges are consumed in groups of up to U
// whose gathers are all issued before the first one is used.  Every branch is warp-uniform and there is no per-edge
// boundary test: a group never crosses a row end (or the 32-wide index batch held in registers), so the inner body is
// load / unpack / add only.  Column indices are fetched one batch ahead.
template <typename T, int VE, int NV, int U, bool kScale, int kMinBlocks>
__global__ void __launch_bounds__(kSpmmThreads, kMinBlocks)
spmm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ first_row,
            int n_chunks, const T* __restrict__ x, int ldx, int F, const float* __restrict__ src_scale, int mean,
            T* __restrict__ out, int64_t ldo, float* __restrict__ partial, int fake_seq_n, int n_rows, int n_edges) {
  const int lane = threadIdx.x & 31;
  const int c = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (c >= n_chunks) return;
  const ChunkRange cr = chunk_range(rowptr, first_row, c);
  if (cr.ee <= cr.eb) return;
  const int cb = c * kEPW;
  constexpr int kColsPerPass = 32 * VE * NV;
  const uint32_t ld_bytes = (uint32_t)ldx * (uint32_t)sizeof(T);

  for (int col0 = 0; col0 < F; col0 += kColsPerPass) {
    const char* xlane = reinterpret_cast<const char*>(x + col0 + lane * VE);
    bool lane_ok[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) lane_ok[k] = col0 + (k * 32 + lane) * VE < F;
    const bool all_lanes = col0 + kColsPerPass <= F;  // warp-uniform: every lane vector of this pass lies inside the row
    RowAcc<VE, NV> acc;
    acc.zero();
    int r = cr.r0;
    int row_start = cr.row_start;
    int row_end = rowptr[r + 1];
    int e = cr.eb;
    // index batches are 32-aligned in edge space: lane i of batch `base` holds col[base + i]
    int base = e & ~31;
    auto fetch_idx = [&](int b) {
      const int i = b + lane;
      int v = (i < n_edges) ? __ldg(col + i) : 0;
#ifdef LLP_EXPERIMENT
      if (fake_seq_n > 0) v = i % fake_seq_n;
#endif
      return v;
    };
    int my = fetch_idx(base), my_next = fetch_idx(base + 32);
    float mys = 1.0f, mys_next = 1.0f;
    if constexpr (kScale) { mys = __ldg(src_scale + my); mys_next = __ldg(src_scale + my_next); }

    while (e < cr.ee) {
      while (row_end == e) {  // close finished rows, skip rows without edges
        if (row_end > row_start) {
          close_row<T, VE, NV>(acc, out, ldo, F, r, col0, lane, mean, row_start, row_end, cb, c, partial);
          acc.zero();
        }
        ++r;
        row_start = row_end;
        row_end = rowptr[r + 1];
      }
      const int run_end = min(row_end, cr.ee);
      while (e < run_end) {
        if (e >= base + 32) {  // next index batch (already in registers), prefetch the one after
          base += 32;
          my = my_next;
          my_next = fetch_idx(base + 32);
          if constexpr (kScale) { mys = mys_next; mys_next = __ldg(src_scale + my_next); }
        }
        const int o = e - base;                             // position inside the batch
        const int cnt = min(min(U, run_end - e), 32 - o);  // group: same row, same batch
        uint4 v[U][NV];
        if (sizeof(T) == 2 && cnt == U && all_lanes) {   // (fp32 rows: measured 6 % slower with this path, left as it was)
          // Full group and every lane inside the row (F == 32 * VE * NV: the widths this path is built for): no per-edge
          // or per-lane predicate at all — shuffle, one address multiply-add, load; then unpack and packed adds: ~16 warp
          // instructions per edge instead of ~58 through the guarded code below.  (A cascade of predicate-free groups of
          // 8 / 4 / 2 / 1 edges for the partial groups was measured too: slower, 234 vs 200 us — code size.)
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int src = __shfl_sync(0xffffffffu, my, (o + u) & 31);
            const char* row = xlane + (uint64_t)((uint32_t)src) * ld_bytes;
#pragma unroll
            for (int k = 0; k < NV; ++k) v[u][k] = load_vec<T, VE>(reinterpret_cast<const T*>(row) + k * 32 * VE);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            float sc = 1.0f;
            if constexpr (kScale) sc = __shfl_sync(0xffffffffu, mys, (o + u) & 31);
            add_row<T, VE, NV, kScale>(acc, v[u], sc);
          }
          e += U;
          continue;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int src = __shfl_sync(0xffffffffu, my, (o + u) & 31);
          if (u < cnt) {
            const char* row = xlane + (uint64_t)((uint32_t)src) * ld_bytes;
#pragma unroll
            for (int k = 0; k < NV; ++k)
              v[u][k] = lane_ok[k] ? load_vec<T, VE>(reinterpret_cast<const T*>(row) + k * 32 * VE) : make_uint4(0, 0, 0, 0);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float sc = 1.0f;
          if constexpr (kScale) sc = __shfl_sync(0xffffffffu, mys, (o + u) & 31);
          if (u < cnt) add_row<T, VE, NV, kScale>(acc, v[u], sc);
        }
        e += cnt;
      }
    }
    // the last open row: complete (it ends exactly at ee) or a hub slice
    close_row<T, VE, NV>(acc, out, ldo, F, r, col0, lane, mean, row_start, row_end, cb, c, partial);
  }
}

// Streaming variant for rows that are exactly one pass of the warp (F == 32 * VE * NV: 128 / 256 / 512 bf16, 64 / 128 /
// 256 fp32 — every width the C3-C5 encoders aggregate).  Same chunk plan, same per-row summation order, same split-row
// partials as spmm_kernel (bit-identical results); what changes is that the LOAD groups are decoupled from the rows.
// spmm_kernel never lets a group of U gathers cross a row end, so at an average degree of ~10 a third of the groups are
// partial: fewer loads in flight and the ~58-instruction guarded path.  Here the warp walks its edge range in U-aligned
// groups of exactly U gathers whatever rows they belong to (the 32-wide index batches are 32-aligned and U divides 32, so
// a group never straddles a batch either): shuffle, one multiply-add for the address, load — no predicate.  The first
// and last group of a chunk may hold edges of the neighbouring chunks (real edges, so the loads are safe; they are not
// accumulated).  A group that lies inside one row and inside the chunk's range — the common case — is then U unguarded
// adds; any other group is walked in row pieces: each finished row is closed at ONE out-of-line site, each piece is a
// fall-through chain of adds entered at its first edge and left at its last (one warp-uniform test per edge).  The next
// row end is fetched one row ahead so that closing a row never waits on memory.
// (ncu, C4 F = 256 bf16: 41 -> STREAM_INSTR warp instructions per edge.)
#define LLP_STREAM_ADD(k_)                                                    \
  {                                                                           \
    float sc = 1.0f;                                                          \
    if constexpr (kScale) sc = __shfl_sync(0xffffffffu, mys, o + (k_));       \
    add_row_pinned<T, VE, NV, kScale>(acc, v[k_], sc);                        \
  }
// kPeer (node-partitioned encoder): the source rows live in the peer-mapped buffers of W ranks (`x` is
// then a device table of W base pointers) and a column index is (owner rank << peer_shift) | row inside the owner's
// block.  Each lane turns ITS entry of an index batch into a full 64-bit row address once per batch (one cached table
// load); the gathers of remote rows are ordinary 128-bit loads that travel over NVLink, issued next to the local ones.
template <typename T, int VE, int NV, int U, bool kScale, int kMinBlocks, bool kPeer = false>
__global__ void __launch_bounds__(kSpmmThreads, kMinBlocks)
spmm_stream_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int4* __restrict__ desc,
                   int n_chunks, const T* __restrict__ x, int ldx, int F, const float* __restrict__ src_scale, int mean,
                   T* __restrict__ out, int64_t ldo, float* __restrict__ partial, int n_rows, int n_edges,
                   int peer_shift = 0, int peer_nloc = 0) {
  static_assert(U == 2 || U == 4 || U == 8, "load groups of 2, 4 or 8 (a group must not straddle a 32-wide index batch)");
  const int lane = threadIdx.x & 31;
  const int c = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (c >= n_chunks) return;
  const int4 d0 = __ldg(desc + 2 * c), d1 = __ldg(desc + 2 * c + 1);
  struct { int r0, eb, ee, row_start; } cr = {d0.x, d0.y, d0.z, d0.w};
  if (cr.ee <= cr.eb) return;
  const uint32_t ld_bytes = (uint32_t)ldx * (uint32_t)sizeof(T);
  const char* xlane = reinterpret_cast<const char*>(x + lane * VE);
  RowAcc<VE, NV> acc;
  acc.zero();
  int r = cr.r0;
  int row_start = cr.row_start;
  int row_end = d1.x;
  int row_end_next = d1.y;
  int e = cr.eb & ~(U - 1);
  int u = cr.eb - e;                         // first edge of the group to accumulate: > 0 only in the chunk's first group
  // the 32-wide index batch holding edge e is the 32-aligned one (e is U-aligned, U divides 32): lane i has col[(e & ~31) + i]
  auto fetch_idx = [&](int b) {
    const int i = b + lane;
    return (i < n_edges) ? __ldg(col + i) : 0;
  };
  int my = fetch_idx(e & ~31), my_next = fetch_idx((e & ~31) + 32);
  // kPeer: byte address of row `idx` for THIS lane's columns; scale index = the row's global (rank-major) number
  auto peer_addr = [&](int idx) -> uint64_t {
    const uint64_t base = (uint64_t)__ldg(reinterpret_cast<const unsigned long long*>(x) + (idx >> peer_shift));
    return base + (uint64_t)(uint32_t)(idx & ((1 << peer_shift) - 1)) * ld_bytes;
  };
  auto scale_index = [&](int idx) -> int {
    if constexpr (kPeer) return (idx >> peer_shift) * peer_nloc + (idx & ((1 << peer_shift) - 1));
    else return idx;
  };
  uint64_t myaddr = 0;
  if constexpr (kPeer) myaddr = peer_addr(my);
  // the scale of the CURRENT batch is requested when the batch becomes current (its indices are in registers by then)
  // and first read behind the batch's first gathers; requesting it for the NEXT batch would wait on that batch's indices
  float mys = 1.0f;
  if constexpr (kScale) mys = __ldg(src_scale + scale_index(my));
  const int ee = cr.ee;
  while (e < ee) {
    const int o = e & 31;
    uint4 v[U][NV];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const char* row;
      if constexpr (kPeer) {
        const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)myaddr, o + k);
        const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(myaddr >> 32), o + k);
        row = reinterpret_cast<const char*>(((uint64_t)hi << 32 | lo) + (uint64_t)(lane * VE * (int)sizeof(T)));
      } else {
        const int src = __shfl_sync(0xffffffffu, my, o + k);
        row = xlane + (uint64_t)((uint32_t)src) * ld_bytes;
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) v[k][j] = load_vec<T, VE>(reinterpret_cast<const T*>(row) + j * 32 * VE);
    }
    if (u == 0 && min(row_end, ee) - e >= U) {   // the whole group continues the open row: no test per edge
#pragma unroll
      for (int k = 0; k < U; ++k) LLP_STREAM_ADD(k)
    } else {
      const int u_end = min(U, ee - e);      // < U only in the chunk's last group
      while (u < u_end) {                    // one trip per row piece inside the group; every test is warp-uniform
        while (e + u == row_end) {           // edge e + u opens a new row: close the finished one, skip empty rows
          if (row_end > row_start) {
            close_row<T, VE, NV>(acc, out, ldo, F, r, 0, lane, mean, row_start, row_end, c * kEPW, c, partial);
            acc.zero();
          }
          ++r;
          row_start = row_end;
          row_end = row_end_next;
          row_end_next = rowptr[min(r + 2, n_rows)];
        }
        const int stop = min(u_end, row_end - e);   // > u
        switch (u) {                         // enter at the piece's first edge, leave after its last
          case 0: LLP_STREAM_ADD(0) if (stop == 1) break;  // fall through
          case 1: LLP_STREAM_ADD(1) if (U == 2 || stop == 2) break;
          case 2: if constexpr (U > 2) { LLP_STREAM_ADD(2) } if (stop == 3) break;
          case 3: if constexpr (U > 2) { LLP_STREAM_ADD(3) } if (U == 4 || stop == 4) break;
          case 4: if constexpr (U > 4) { LLP_STREAM_ADD(4) } if (stop == 5) break;
          case 5: if constexpr (U > 4) { LLP_STREAM_ADD(5) } if (stop == 6) break;
          case 6: if constexpr (U > 4) { LLP_STREAM_ADD(6) } if (stop == 7) break;
          case 7: if constexpr (U > 4) { LLP_STREAM_ADD(7) } break;
        }
        u = stop;
      }
      u = 0;
    }
    e += U;
    if ((e & 31) == 0) {   // next index batch (already in registers), prefetch the one after
      my = my_next;
      if constexpr (kPeer) myaddr = peer_addr(my);
      if constexpr (kScale) mys = __ldg(src_scale + scale_index(my));
      my_next = fetch_idx(e + 32);
    }
  }
  close_row<T, VE, NV>(acc, out, ldo, F, r, 0, lane, mean, row_start, row_end, c * kEPW, c, partial);
}
#undef LLP_STREAM_ADD

// Fix-up: adds the fp32 partials of every hub row and zero-fills the rows without edges.  Most hubs are barely longer
// than a chunk (C4: 2,870 rows above 64 edges, 1,850 of them below 128): a hub of at most kFixWarpParts partials is
// combined by ONE warp — eight independent 16-byte loads per lane and pass, added in chunk order — eight such hubs per
// block.  A longer hub is cut into slices of 128 columns and each (hub, slice) is taken by a whole block, before the small
// hubs (the longest hub is the critical path of the launch: every other SM idles behind it): warp w sums the partials
// w, w+8, ... (8 x 16 bytes in flight per lane) and the eight warp sums are combined in warp order through shared memory — for up to eight partials the same association as the
// single-warp path, so a result never depends on the path; deterministic, and a 12k-edge hub (200 partials) costs a few
// microseconds instead of one long serial chain.  The hub records carry the row and its bounds (no table chasing).
constexpr int kFixHubsPerBlock = 8;
constexpr int kFixZeroRows = 2048;   // rows scanned per zero-fill block
template <typename T>
__global__ void __launch_bounds__(256)
spmm_fixup_kernel(const int32_t* __restrict__ rowptr, const int4* __restrict__ hub_tbl, int hub_blocks, int n_rows, int F,
                  int mean, T* __restrict__ out, int64_t ldo, const float* __restrict__ partial) {
  __shared__ float red[8][128];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if ((int)blockIdx.x >= hub_blocks) {
    // the blocks behind the hub blocks zero-fill the rows without edges (kFixZeroRows rows per block, one row per lane
    // and pass; the warp then writes each empty row together): the main kernels only ever write rows that own edges
    const int base = ((int)blockIdx.x - hub_blocks) * kFixZeroRows + w * (kFixZeroRows / 8);
    const int lim = min(n_rows, base + kFixZeroRows / 8);
    bool empty[kFixZeroRows / 8 / 32];
#pragma unroll
    for (int i = 0; i < kFixZeroRows / 8 / 32; ++i) {   // all loads first: one round trip for the warp's 256 rows
      const int rr = base + i * 32 + lane;
      empty[i] = rr < lim && rowptr[rr + 1] == rowptr[rr];
    }
#pragma unroll
    for (int i = 0; i < kFixZeroRows / 8 / 32; ++i) {
      unsigned m = __ballot_sync(0xffffffffu, empty[i]);
      while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        T* row = out + (int64_t)(base + i * 32 + j) * ldo;
        for (int f = lane; f < F; f += 32) row[f] = from_f32<T>(0.0f);
      }
    }
    return;
  }
  const int4 hdr = hub_tbl[0];
  const int n_small = hdr.x, n_big = hdr.y, big0 = hdr.z;
  // Big hubs first (they are the critical path of the launch): one work item = (hub, slice of 128 columns); block b takes
  // the items b, b + hub_blocks, ...  Lane l owns columns 4l..4l+3 of the slice (one 16-byte vector when F % 4 == 0,
  // else the four columns l, l+32, l+64, l+96).
  const bool vec = F % 4 == 0;
  const int n_slices = (F + 127) / 128;
  for (int item = blockIdx.x; item < n_big * n_slices; item += hub_blocks) {   // block-uniform
    const int4 rec = hub_tbl[big0 - item / n_slices];
    const int f0 = (item % n_slices) * 128;
    const int c = rec.x, rp = rec.y, ps = rec.z, pe = rec.w;
    const int n_part = (pe - 1) / kEPW - c + 2;
    const float divisor = mean ? (float)(pe - ps) : 1.0f;
    auto part_ptr = [&](int k) { return partial + (k == 0 ? ((int64_t)(c - 1) * 2 + 1) : ((int64_t)(c + k - 1) * 2)) * F + f0; };
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = w; k < n_part; k += 64) {   // eight loads in flight (zeros behind the last partial)
      float4 t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k + 8 * i < n_part) {
          const float* p = part_ptr(k + 8 * i);
          if (vec) {
            if (f0 + lane * 4 < F) t[i] = __ldg(reinterpret_cast<const float4*>(p + lane * 4));
          } else {
            if (f0 + lane < F) t[i].x = __ldg(p + lane);
            if (f0 + lane + 32 < F) t[i].y = __ldg(p + lane + 32);
            if (f0 + lane + 64 < F) t[i].z = __ldg(p + lane + 64);
            if (f0 + lane + 96 < F) t[i].w = __ldg(p + lane + 96);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) { acc.x += t[i].x; acc.y += t[i].y; acc.z += t[i].z; acc.w += t[i].w; }
    }
    // red[w][j]: column f0 + j of warp w's sum
    if (vec) {
      *reinterpret_cast<float4*>(&red[w][lane * 4]) = acc;
    } else {
      red[w][lane] = acc.x; red[w][lane + 32] = acc.y; red[w][lane + 64] = acc.z; red[w][lane + 96] = acc.w;
    }
    __syncthreads();
    if (threadIdx.x < 128 && f0 + (int)threadIdx.x < F) {
      float sum = red[0][threadIdx.x];
#pragma unroll
      for (int i = 1; i < 8; ++i) sum += red[i][threadIdx.x];
      out[(int64_t)rp * ldo + f0 + threadIdx.x] = from_f32<T>(__fdiv_rn(sum, divisor));
    }
    __syncthreads();
  }
  const int h = blockIdx.x * kFixHubsPerBlock + w;
  if (h < n_small) {
    const int4 rec = hub_tbl[1 + h];
    const int c = rec.x, rp = rec.y, ps = rec.z, pe = rec.w;
    const int n_part = (pe - 1) / kEPW - c + 2;  // slot 1 of chunk c-1, then slot 0 of chunks c .. c_last
    const float divisor = mean ? (float)(pe - ps) : 1.0f;
    const float* p0 = partial + ((int64_t)(c - 1) * 2 + 1) * F;   // partial k >= 1 lies at p0 + (2k - 1) * F
    if (vec) {
      for (int f = lane * 4; f < F; f += 128) {
        float4 t[kFixWarpParts];
#pragma unroll
        for (int k = 0; k < kFixWarpParts; ++k)
          t[k] = k < n_part ? __ldg(reinterpret_cast<const float4*>(p0 + (k == 0 ? 0 : (int64_t)(2 * k - 1) * F) + f))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 acc = t[0];
#pragma unroll
        for (int k = 1; k < kFixWarpParts; ++k) { acc.x += t[k].x; acc.y += t[k].y; acc.z += t[k].z; acc.w += t[k].w; }
        float r[4] = {__fdiv_rn(acc.x, divisor), __fdiv_rn(acc.y, divisor), __fdiv_rn(acc.z, divisor), __fdiv_rn(acc.w, divisor)};
        T* dst = out + (int64_t)rp * ldo + f;
        if constexpr (sizeof(T) == 4) *reinterpret_cast<float4*>(dst) = make_float4(r[0], r[1], r[2], r[3]);
        else *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]));
      }
    } else {
      for (int f = lane; f < F; f += 32) {
        float acc = __ldg(p0 + f);
        for (int k = 1; k < n_part; ++k) acc += __ldg(p0 + (int64_t)(2 * k - 1) * F + f);
        out[(int64_t)rp * ldo + f] = from_f32<T>(__fdiv_rn(acc, divisor));
      }
    }
  }
}

int g_spmm_variant = 0;  // 0 = 8 blocks/SM (<= 64 registers): fastest for the row-run kernel (tools/kbench.py spmmsweep)
// Result-changing experiment knobs exist only in -DLLP_EXPERIMENT builds (make EXTRA=-DLLP_EXPERIMENT); the shipped
// library cannot skip work: llp_set_tuning ignores keys 1, 2, 11 and 13.
#ifdef LLP_EXPERIMENT
int g_spmm_chunk_div = 1;  // experiment: process only the first n_chunks/div chunks
int g_spmm_fake_seq = 0;   // experiment: gather row (edge id mod N) instead of col[e]
#define LLP_SPMM_CHUNKS(n) ((n) / g_spmm_chunk_div)
#define LLP_SPMM_FAKE(N) (g_spmm_fake_seq ? (int)(N) : 0)
#else
#define LLP_SPMM_CHUNKS(n) (n)
#define LLP_SPMM_FAKE(N) 0
#endif

template <typename T, bool kScale>
static int spmm_launch(const int32_t* rowptr, const int32_t* col, const int32_t* first_row, int64_t N, int64_t E,
                       const void* x_, int64_t ldx, int64_t F, const float* src_scale, int mean, void* out_,
                       int64_t ldo, void* ws, const int32_t* hub_list, int num_hubs, cudaStream_t stream) {
  const T* x = reinterpret_cast<const T*>(x_);
  T* out = reinterpret_cast<T*>(out_);
  float* partial = reinterpret_cast<float*>(ws);
  const int n_chunks = (int)llp_spmm_num_chunks(E);
  const int4* desc = reinterpret_cast<const int4*>(first_row + spmm_desc_offset_ints(n_chunks));
  constexpr int VE = Vec16<T>::n;
  // Odd widths (Cora's 1,433 input features): when both matrices are row-padded to the next 16-byte multiple — the host
  // side's buffers always are — the kernels run over the PADDED width with 128-bit accesses (the padding columns of the
  // output then hold sums of the input's padding columns; nobody reads them: every consumer takes F) instead of the scalar
  // path, which walks the edge list once per 128 columns: 303 us -> ~15 us for the layer-1 aggregation of the Cora teacher.
  const bool base_ok = aligned(x, 16) && aligned(out, 16) && (ldx * sizeof(T)) % 16 == 0 && (ldo * sizeof(T)) % 16 == 0;
  const int64_t F_padded = (F + VE - 1) / VE * VE;
  if (base_ok && F % VE != 0 && ldx >= F_padded && ldo >= F_padded) F = F_padded;
  const bool vec = base_ok && F % VE == 0;
  const unsigned blocks = (unsigned)ceil_div((int64_t)n_chunks * 32, kSpmmThreads);
  if (E > 0) {
    // rows of exactly one warp pass go to the streaming kernel (llp_set_tuning(3, 1): always the row-run kernel; A/B knob,
    // the two are bit-identical)
    const int stream_mode = LLP_SPMM_FAKE(N) != 0 ? 1 : g_tuning[3];   // 0: rows of one 16-byte (or 8-byte) vector per lane; 1: never; 2: every full-width row
#define LLP_SPMM_LAUNCH(VE_, NV_, U_, MB_) \
  do { \
    if (stream_mode != 1 && F == 32 * (VE_) * (NV_) && ((NV_) == 1 || stream_mode == 2)) \
      spmm_stream_kernel<T, VE_, NV_, U_, kScale, MB_><<<blocks, kSpmmThreads, 0, stream>>>(rowptr, col, desc, LLP_SPMM_CHUNKS(n_chunks), x, (int)ldx, (int)F, src_scale, mean, out, ldo, partial, (int)N, (int)E); \
    else \
      spmm_kernel<T, VE_, NV_, U_, kScale, MB_><<<blocks, kSpmmThreads, 0, stream>>>(rowptr, col, first_row, LLP_SPMM_CHUNKS(n_chunks), x, (int)ldx, (int)F, src_scale, mean, out, ldo, partial, LLP_SPMM_FAKE(N), (int)N, (int)E); \
  } while (0)
    const int variant = g_spmm_variant;  // occupancy/register trade-off (llp_set_tuning(0, v)): 0 = 8 blocks/SM (<=64 regs)
    const bool vec8 = aligned(x, 8) && aligned(out, 8) && (ldx * sizeof(T)) % 8 == 0 && (ldo * sizeof(T)) % 8 == 0 &&
                      F % (VE / 2) == 0 && F * (int64_t)sizeof(T) <= 256;
    if (vec8) {
      if (variant == 0) LLP_SPMM_LAUNCH(VE / 2, 1, 8, 8); else if (variant == 3) LLP_SPMM_LAUNCH(VE / 2, 1, 4, 8); else LLP_SPMM_LAUNCH(VE / 2, 1, 8, 6);
    } else if (vec) {
      if (F * (int64_t)sizeof(T) <= 512) {
        // default: groups of 4 gathers (with an average degree of ~10 far more groups are full, i.e. take the predicate-free
        // path, than with groups of 8: 189 / 208 us vs 200 / 226 us forward / transpose at F = 256 bf16 on the C4 graph)
        if (variant == 1) LLP_SPMM_LAUNCH(VE, 1, 8, 6); else if (variant == 2) LLP_SPMM_LAUNCH(VE, 1, 8, 5); else if (variant == 3) LLP_SPMM_LAUNCH(VE, 1, 8, 8);
        else if (variant == 4) LLP_SPMM_LAUNCH(VE, 1, 2, 10); else if (variant == 5) LLP_SPMM_LAUNCH(VE, 1, 4, 10); else if (variant == 6) LLP_SPMM_LAUNCH(VE, 1, 2, 12); else LLP_SPMM_LAUNCH(VE, 1, 4, 8);
      } else {
        if (variant == 1) LLP_SPMM_LAUNCH(VE, 2, 4, 6); else if (variant == 2) LLP_SPMM_LAUNCH(VE, 2, 4, 5); else if (variant == 3) LLP_SPMM_LAUNCH(VE, 2, 2, 8); else LLP_SPMM_LAUNCH(VE, 2, 4, 8);
      }
    } else {
      LLP_SPMM_LAUNCH(1, 4, 4, 8);
    }
#undef LLP_SPMM_LAUNCH
    LLP_LAUNCH_OK();
  }
  if (E > 0) {   // hub rows: add the partials; rows without edges: zeros (the main kernel writes neither)
    const int hub_blocks = (int)ceil_div(num_hubs, kFixHubsPerBlock);   // >= ceil(n_small / 8); big hubs stride over them
    const unsigned fix_blocks = (unsigned)(hub_blocks + ceil_div(N, kFixZeroRows));
    spmm_fixup_kernel<T><<<fix_blocks, 256, 0, stream>>>(rowptr, reinterpret_cast<const int4*>(hub_list), hub_blocks, (int)N, (int)F,
                                                        mean, out, ldo, partial);
    LLP_LAUNCH_OK();
  } else {
    LLP_CUDA(cudaMemset2DAsync(out, (size_t)ldo * sizeof(T), 0, (size_t)F * sizeof(T), (size_t)N, stream));
  }
  return 0;
}

// Peer variant: only the streaming kernel's full-width single-vector rows (the caller falls back to a staged all-gather
// for anything else).
template <typename T, bool kScale>
static int spmm_peer_launch(const int32_t* rowptr, const int32_t* col, const int32_t* first_row, int64_t N, int64_t E,
                            const void* const* peer_x, int peer_shift, int64_t peer_nloc, int64_t ldx, int64_t F,
                            const float* src_scale, int mean, void* out_, int64_t ldo, void* ws, const int32_t* hub_list,
                            int num_hubs, cudaStream_t stream) {
  T* out = reinterpret_cast<T*>(out_);
  float* partial = reinterpret_cast<float*>(ws);
  const int n_chunks = (int)llp_spmm_num_chunks(E);
  const int4* desc = reinterpret_cast<const int4*>(first_row + spmm_desc_offset_ints(n_chunks));
  constexpr int VE = Vec16<T>::n;
  const unsigned blocks = (unsigned)ceil_div((int64_t)n_chunks * 32, kSpmmThreads);
  const T* table = reinterpret_cast<const T*>(peer_x);
  if (E > 0) {
    if (F == 32 * VE)
      spmm_stream_kernel<T, VE, 1, 4, kScale, 8, true><<<blocks, kSpmmThreads, 0, stream>>>(
          rowptr, col, desc, n_chunks, table, (int)ldx, (int)F, src_scale, mean, out, ldo, partial, (int)N, (int)E, peer_shift, (int)peer_nloc);
    else if (F == 32 * (VE / 2))
      spmm_stream_kernel<T, VE / 2, 1, 8, kScale, 8, true><<<blocks, kSpmmThreads, 0, stream>>>(
          rowptr, col, desc, n_chunks, table, (int)ldx, (int)F, src_scale, mean, out, ldo, partial, (int)N, (int)E, peer_shift, (int)peer_nloc);
    else
      return LLP_E_SHAPE;
    LLP_LAUNCH_OK();
    const int hub_blocks = (int)ceil_div(num_hubs, kFixHubsPerBlock);
    const unsigned fix_blocks = (unsigned)(hub_blocks + ceil_div(N, kFixZeroRows));
    spmm_fixup_kernel<T><<<fix_blocks, 256, 0, stream>>>(rowptr, reinterpret_cast<const int4*>(hub_list), hub_blocks, (int)N, (int)F,
                                                        mean, out, ldo, partial);
    LLP_LAUNCH_OK();
  } else {
    LLP_CUDA(cudaMemset2DAsync(out, (size_t)ldo * sizeof(T), 0, (size_t)F * sizeof(T), (size_t)N, stream));
  }
  return 0;
}

}  // namespace llp

using namespace llp;

extern "C" void llp_set_tuning(int key, int value) {
#ifndef LLP_EXPERIMENT
  if (key == 1 || key == 2 || key == 11 || key == 13) return;  // work-skipping experiments: not compiled into this build
#endif
  if (key >= 0 && key < 32) g_tuning[key] = value;
  if (key == 0) g_spmm_variant = value;
#ifdef LLP_EXPERIMENT
  if (key == 1) g_spmm_chunk_div = value < 1 ? 1 : value;
  if (key == 2) g_spmm_fake_seq = value;
#endif
}

extern "C" int64_t llp_spmm_num_chunks(int64_t E) { return E <= 0 ? 1 : ceil_div(E, kEPW); }

extern "C" int64_t llp_spmm_plan_ints(int64_t E) {
  const int64_t n_chunks = llp_spmm_num_chunks(E);
  return spmm_desc_offset_ints(n_chunks) + 8 * n_chunks;
}

extern "C" int llp_spmm_plan(const int32_t* rowptr, int64_t N, int64_t E, int32_t* chunk_first_row, int32_t* hub_list,
                             int32_t* num_hubs, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(rowptr && chunk_first_row && hub_list && num_hubs && N >= 0 && E >= 0);
  if (int rc = check_device()) return rc;
  int64_t n_chunks = llp_spmm_num_chunks(E);
  spmm_plan_kernel<<<(unsigned)ceil_div(n_chunks + 1, 256), 256, 0, stream>>>(rowptr, N, n_chunks, chunk_first_row);
  LLP_LAUNCH_OK();
  spmm_desc_kernel<<<(unsigned)ceil_div(n_chunks, 256), 256, 0, stream>>>(
      rowptr, chunk_first_row, N, n_chunks, reinterpret_cast<int4*>(chunk_first_row + spmm_desc_offset_ints(n_chunks)),
      reinterpret_cast<int4*>(hub_list));
  LLP_LAUNCH_OK();
  LLP_CUDA(cudaMemsetAsync(num_hubs, 0, sizeof(int32_t), stream));
  if (n_chunks > 1) {
    spmm_hub_list_kernel<<<(unsigned)ceil_div(n_chunks, 256), 256, 0, stream>>>(rowptr, chunk_first_row, n_chunks,
                                                                                reinterpret_cast<int4*>(hub_list), num_hubs);
    LLP_LAUNCH_OK();
  }
  return 0;
}

extern "C" size_t llp_spmm_workspace_bytes(int64_t E, int64_t F) {
  const int64_t Fp = ((F > 0 ? F : 1) + 7) / 8 * 8;   // odd widths run over the padded width (spmm_launch)
  return (size_t)llp_spmm_num_chunks(E) * 2 * (size_t)Fp * sizeof(float);
}

extern "C" int llp_spmm(int dtype, const int32_t* rowptr, const int32_t* col, const int32_t* chunk_first_row,
                        int64_t N, int64_t E, const void* x, int64_t ldx, int64_t F, const float* src_scale, int mean,
                        void* out, int64_t ldo, void* workspace, const int32_t* hub_list, int64_t num_hubs, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(rowptr && chunk_first_row && N >= 0 && E >= 0 && F > 0 && ldx >= F && ldo >= F);
  LLP_CHECK_ARG((E == 0 || (col && x && workspace)) && (N == 0 || out) && num_hubs >= 0 && (E == 0 || hub_list));
  LLP_CHECK_ARG(E < (int64_t)INT32_MAX - kEPW && N < (int64_t)INT32_MAX && F < (1 << 24) && ldx * 4 < (int64_t)UINT32_MAX);
  if (int rc = check_device()) return rc;
  if (N == 0) return 0;
#define LLP_SPMM(T)                                                                                                     \
  return src_scale != nullptr                                                                                           \
             ? spmm_launch<T, true>(rowptr, col, chunk_first_row, N, E, x, ldx, F, src_scale, mean, out, ldo, workspace, hub_list, (int)num_hubs, stream)  \
             : spmm_launch<T, false>(rowptr, col, chunk_first_row, N, E, x, ldx, F, src_scale, mean, out, ldo, workspace, hub_list, (int)num_hubs, stream)
  if (dtype == LLP_F32) { LLP_SPMM(float); }
  if (dtype == LLP_BF16) { LLP_SPMM(__nv_bfloat16); }
#undef LLP_SPMM
  return LLP_E_BADARG;
}

// llp_spmm over peer-mapped source blocks (node-partitioned encoder): `peer_x` is a DEVICE array of `world` base
// pointers (rank r's [peer_nloc, ldx] block as mapped into this process, llp_ipc_open), `col` holds
// (owner rank << peer_shift) | local row, `src_scale` (optional) is indexed by owner * peer_nloc + local row.
// Supported widths: feat * sizeof(elt) in {256, 512} bytes (one 8- or 16-byte vector per lane); LLP_E_SHAPE otherwise.
extern "C" int llp_spmm_peer(int dtype, const int32_t* rowptr, const int32_t* col, const int32_t* chunk_first_row,
                             int64_t N, int64_t E, const void* const* peer_x, int world, int peer_shift, int64_t peer_nloc,
                             int64_t ldx, int64_t F, const float* src_scale, int mean, void* out, int64_t ldo,
                             void* workspace, const int32_t* hub_list, int64_t num_hubs, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(rowptr && chunk_first_row && N >= 0 && E >= 0 && F > 0 && ldx >= F && ldo >= F && peer_x && world >= 1);
  LLP_CHECK_ARG(peer_shift >= 0 && peer_shift < 31 && peer_nloc > 0 && peer_nloc <= ((int64_t)1 << peer_shift) &&
                ((int64_t)world << peer_shift) <= (int64_t)INT32_MAX);
  LLP_CHECK_ARG((E == 0 || (col && workspace && hub_list)) && (N == 0 || out) && num_hubs >= 0);
  LLP_CHECK_ARG(E < (int64_t)INT32_MAX - kEPW && N < (int64_t)INT32_MAX && ldx * 4 < (int64_t)UINT32_MAX);
  if (int rc = check_device()) return rc;
  if (N == 0) return 0;
  const int64_t elt = dtype == LLP_F32 ? 4 : 2;
  if ((ldx * elt) % 16 != 0 || (ldo * elt) % 16 != 0 || !aligned(out, 16)) return LLP_E_ALIGN;
#define LLP_SPMM_PEER(T)                                                                                                \
  return src_scale != nullptr                                                                                           \
             ? spmm_peer_launch<T, true>(rowptr, col, chunk_first_row, N, E, peer_x, peer_shift, peer_nloc, ldx, F, src_scale, mean, out, ldo, workspace, hub_list, (int)num_hubs, stream)  \
             : spmm_peer_launch<T, false>(rowptr, col, chunk_first_row, N, E, peer_x, peer_shift, peer_nloc, ldx, F, src_scale, mean, out, ldo, workspace, hub_list, (int)num_hubs, stream)
  if (dtype == LLP_F32) { LLP_SPMM_PEER(float); }
  if (dtype == LLP_BF16) { LLP_SPMM_PEER(__nv_bfloat16); }
#undef LLP_SPMM_PEER
  return LLP_E_BADARG;
}
