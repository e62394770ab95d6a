// Backward of the edge gather-Hadamard z[m] = h[u[m]] * h[v[m]] (autograd of the two fancy-index gathers + mul in front
// of LinkPredictor: train_teacher_gnn.py:58, main.py:186,214, models.py:140; PyTorch uses index_put_ with atomics):
//
//     gh[n,:] = sum_{m: u[m]==n} dz[m,:]*h[v[m],:]  +  sum_{m: v[m]==n} dz[m,:]*h[u[m],:]
//
// as a gather-REDUCE instead of a scatter.  `llp_edge_plan` stably sorts the 2M (node, edge) incidences of the batch by
// node with its own counting sort (keys are node ids: count per node -> exclusive scan = the row pointers -> scatter ->
// every row put back into incidence order; it depends only on u and v, so the host runs it on a side stream while the
// encoder works); the backward
// kernel then walks node rows in order, adds each row's incidences in sorted order and writes the row once in the
// activation dtype (rows without incidences are written as zeros).  No atomics, no fp32 staging buffer, no separate
// zero-fill / cast passes; the result is bit-reproducible.  Rows with more than kHubThreshold incidences (hub nodes of
// a power-law graph) are deferred to a block-per-row kernel (8 warps take every 8th incidence, fixed order combine) so
// that no single warp serialises on a long row.
#include "common.cuh"

namespace llp {
namespace eb {

constexpr int kHubThreshold = 48;
constexpr int kHubWarpsMax = 32;  // warps of a hub block (fewer when F is wide: [warps][F] floats of shared memory)
constexpr int kRowsPerWarp = 16;

// ---- incidence plan: stable counting sort of the 2M incidences (i < M: endpoint u[i] of edge i; i >= M: endpoint v[i - M])
//      by node.  Row r of the plan lists its incidences in increasing i, exactly the order of a stable sort by key.
constexpr int kScanThreads = 256, kScanItems = 8, kScanTile = kScanThreads * kScanItems;
constexpr int kSmallRow = 16;       // rows up to this length are ordered by one thread,
constexpr int kMidRow = 256;        // up to this one by a warp (anchors of a distillation batch: K + 1 incidences each), longer ones by blocks
constexpr int kBigThreads = 256, kBigTile = 4096;   // kBigTile: rows one block orders alone (16 KB of shared memory)

__device__ __forceinline__ int incidence_node(const int64_t* __restrict__ u, const int64_t* __restrict__ v, int64_t M, int64_t i) {
  return (int)(i < M ? __ldg(u + i) : __ldg(v + (i - M)));
}

// cnt[node] += 1 per incidence; the value the atomic returns is the incidence's arrival slot inside its row
// (an endpoint outside [0, N) is a caller error — the reference's fancy index would raise —; it is skipped here and in
// the scatter so that it can never write outside the plan)
__global__ void incidence_count_kernel(const int64_t* __restrict__ u, const int64_t* __restrict__ v, int64_t M, int64_t N,
                                       int32_t* __restrict__ cnt, int32_t* __restrict__ slot) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= 2 * M) return;
  const int node = incidence_node(u, v, M, i);
  slot[i] = ((unsigned)node < (unsigned)N) ? atomicAdd(cnt + node, 1) : -1;
}

// In-place exclusive scan of x[0..n) in three launches: tiles of kScanTile (thread-local prefix + block scan), the tile
// totals (one block walks them with a carry), and the add-back.
__global__ void __launch_bounds__(kScanThreads) scan_tiles_kernel(int32_t* __restrict__ x, int64_t n, int32_t* __restrict__ tile_sum) {
  __shared__ int32_t warp_sum[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int32_t v[kScanItems];
  int32_t t = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = base + k < n ? x[base + k] : 0;
    t += v[k];
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t inc = t;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) warp_sum[w] = inc;
  __syncthreads();
  int32_t before = inc - t;
  for (int q = 0; q < w; ++q) before += warp_sum[q];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) x[base + k] = before;
    before += v[k];
  }
  if (threadIdx.x == kScanThreads - 1) tile_sum[blockIdx.x] = before;
}

__global__ void __launch_bounds__(1024) scan_tile_sums_kernel(int32_t* __restrict__ tile_sum, int64_t tiles) {
  __shared__ int32_t warp_sum[32];
  __shared__ int32_t carry_s;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t t0 = 0; t0 < tiles; t0 += 1024) {
    const int64_t i = t0 + threadIdx.x;
    const int32_t t = i < tiles ? tile_sum[i] : 0;
    int32_t inc = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) warp_sum[w] = inc;
    __syncthreads();
    int32_t before = carry_s + inc - t;
    for (int q = 0; q < w; ++q) before += warp_sum[q];
    if (i < tiles) tile_sum[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = before + t;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_add_kernel(int32_t* __restrict__ x, int64_t n, const int32_t* __restrict__ tile_sum) {
  const int32_t add = tile_sum[blockIdx.x];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < n) x[base + k] += add;
}

// ids[rowptr[node] + slot] = i: every row now holds its incidences, in arrival (arbitrary) order
__global__ void incidence_scatter_kernel(const int64_t* __restrict__ u, const int64_t* __restrict__ v, int64_t M,
                                         const int32_t* __restrict__ rowptr, const int32_t* __restrict__ slot,
                                         int32_t* __restrict__ ids) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < 2 * M && slot[i] >= 0) ids[rowptr[incidence_node(u, v, M, i)] + slot[i]] = (int32_t)i;
}

// meta[p] = (edge m, the OTHER endpoint of that edge) of incidence e, so that the backward kernel has no index chains
__device__ __forceinline__ int2 incidence_meta(const int64_t* __restrict__ u, const int64_t* __restrict__ v, int64_t M, int e) {
  const int64_t m = e < M ? e : e - M;
  return make_int2((int)m, (int)(e < M ? __ldg(v + m) : __ldg(u + m)));
}

// Put every row back into increasing incidence order (ids are distinct: the rank of an id is the number of smaller ids
// in its row) and write the plan.  One thread per short row; long rows are only listed for the block kernel below.
__global__ void incidence_order_rows_kernel(const int64_t* __restrict__ u, const int64_t* __restrict__ v, int64_t M, int64_t N,
                                            const int32_t* __restrict__ rowptr, const int32_t* __restrict__ ids,
                                            int2* __restrict__ meta, int32_t* __restrict__ big_count, int32_t* __restrict__ big_list,
                                            int32_t* __restrict__ huge_list, int32_t* __restrict__ mid_list) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= N) return;
  const int s = rowptr[r], n = rowptr[r + 1] - s;
  if (n == 0) return;
  if (n > kSmallRow) {
    if (n <= kMidRow) mid_list[atomicAdd(big_count + 2, 1)] = (int32_t)r;       // slot order does not matter
    else if (n <= kBigTile) big_list[atomicAdd(big_count, 1)] = (int32_t)r;
    else huge_list[atomicAdd(big_count + 1, 1)] = (int32_t)r;
    return;
  }
  int e[kSmallRow];
#pragma unroll
  for (int a = 0; a < kSmallRow; ++a) e[a] = a < n ? ids[s + a] : INT32_MAX;
#pragma unroll
  for (int a = 0; a < kSmallRow; ++a) {
    if (a < n) {
      int rank = 0;
#pragma unroll
      for (int b = 0; b < kSmallRow; ++b) rank += e[b] < e[a] ? 1 : 0;
      meta[s + rank] = incidence_meta(u, v, M, e[a]);
    }
  }
}

// Long rows.  Rows of up to kMidRow incidences take one warp each; a row of up to kBigTile incidences is ordered by ONE block (the row sits in shared memory, every thread ranks
// its ids against it); the few rows beyond that (at most 2M / kBigTile of them: a node that owns a large share of the
// whole batch) are ranked slice by slice by all blocks together, the row streamed through shared memory in tiles.
__global__ void __launch_bounds__(kBigThreads) incidence_order_big_kernel(const int64_t* __restrict__ u, const int64_t* __restrict__ v,
                                                                          int64_t M, const int32_t* __restrict__ rowptr,
                                                                          const int32_t* __restrict__ ids, int2* __restrict__ meta,
                                                                          const int32_t* __restrict__ big_count,
                                                                          const int32_t* __restrict__ big_list,
                                                                          const int32_t* __restrict__ huge_list,
                                                                          const int32_t* __restrict__ mid_list) {
  __shared__ int32_t tile[kBigTile];
  const int n_big = big_count[0], n_huge = big_count[1], n_mid = big_count[2];
  {  // rows of 17 .. kMidRow incidences: one WARP per row (a block per row spends its time on the row's dependent index
     // loads: 5 us per row for the thousands of 21-incidence anchor rows of a student batch), 8 rows in flight per block
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int32_t* mine = tile + warp * kMidRow;
    for (int b = blockIdx.x * (kBigThreads / 32) + warp; b < n_mid; b += gridDim.x * (kBigThreads / 32)) {
      const int64_t r = mid_list[b];
      const int s = rowptr[r], n = rowptr[r + 1] - s;
      __syncwarp();
      for (int q = lane; q < n; q += 32) mine[q] = ids[s + q];
      __syncwarp();
      for (int a = lane; a < n; a += 32) {
        const int ea = mine[a];
        int rank = 0;
        for (int q = 0; q < n; ++q) rank += mine[q] < ea ? 1 : 0;
        meta[s + rank] = incidence_meta(u, v, M, ea);
      }
    }
  }
  __syncthreads();
  for (int b = blockIdx.x; b < n_big; b += gridDim.x) {
    const int64_t r = big_list[b];
    const int s = rowptr[r], n = rowptr[r + 1] - s;
    __syncthreads();
    for (int q = threadIdx.x; q < n; q += kBigThreads) tile[q] = ids[s + q];
    __syncthreads();
    for (int a = threadIdx.x; a < n; a += kBigThreads) {
      const int ea = tile[a];
      int rank = 0;
      for (int q = 0; q < n; ++q) rank += tile[q] < ea ? 1 : 0;
      meta[s + rank] = incidence_meta(u, v, M, ea);
    }
  }
  for (int b = 0; b < n_huge; ++b) {
    const int64_t r = huge_list[b];
    const int s = rowptr[r], n = rowptr[r + 1] - s;
    const int slices = (n + kBigThreads - 1) / kBigThreads;
    for (int sl = blockIdx.x; sl < slices; sl += gridDim.x) {
      const int a = sl * kBigThreads + threadIdx.x;
      const int ea = a < n ? ids[s + a] : INT32_MAX;
      int rank = 0;
      for (int t0 = 0; t0 < n; t0 += kBigTile) {
        __syncthreads();
        for (int q = threadIdx.x; q < kBigTile; q += kBigThreads) tile[q] = t0 + q < n ? ids[s + t0 + q] : INT32_MAX;
        __syncthreads();
        const int lim = n - t0 < kBigTile ? n - t0 : kBigTile;
        for (int q = 0; q < lim; ++q) rank += tile[q] < ea ? 1 : 0;
      }
      if (a < n) meta[s + rank] = incidence_meta(u, v, M, ea);
    }
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

struct Workspace {
  int32_t *slot, *ids, *tile_sum, *big_count, *big_list, *huge_list, *mid_list;
  int64_t tiles;
  size_t total;
};

static Workspace carve(char* base, int64_t M, int64_t N) {
  const int64_t E = 2 * M;
  Workspace w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align256(bytes); return p; };
  w.tiles = ceil_div(N + 1, (int64_t)kScanTile);
  w.slot = (int32_t*)take((size_t)E * 4 + 4);
  w.ids = (int32_t*)take((size_t)E * 4 + 4);
  w.tile_sum = (int32_t*)take((size_t)w.tiles * 4 + 4);
  w.big_count = (int32_t*)take(16);                                       // {rows one block orders, rows all blocks order, rows one warp orders}
  w.mid_list = (int32_t*)take((size_t)(E / (kSmallRow + 1) + 1) * 4);
  w.big_list = (int32_t*)take((size_t)(E / (kMidRow + 1) + 1) * 4);
  w.huge_list = (int32_t*)take((size_t)(E / (kBigTile + 1) + 1) * 4);
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

namespace llp {
// In-place exclusive scan of n int32 on the device (three small launches); tile_sum: scan_i32_tiles(n) int32 of scratch.
int64_t scan_i32_tiles(int64_t n) { return ceil_div(n, (int64_t)eb::kScanTile); }
int exclusive_scan_i32(int32_t* x, int64_t n, int32_t* tile_sum, cudaStream_t stream) {
  if (n <= 0) return 0;
  const int64_t tiles = scan_i32_tiles(n);
  eb::scan_tiles_kernel<<<(unsigned)tiles, eb::kScanThreads, 0, stream>>>(x, n, tile_sum);
  LLP_LAUNCH_OK();
  if (tiles > 1) {
    eb::scan_tile_sums_kernel<<<1, 1024, 0, stream>>>(tile_sum, tiles);
    LLP_LAUNCH_OK();
    eb::scan_add_kernel<<<(unsigned)tiles, eb::kScanThreads, 0, stream>>>(x, n, tile_sum);
    LLP_LAUNCH_OK();
  }
  return 0;
}
}  // namespace llp

using namespace llp;

extern "C" size_t llp_edge_plan_workspace_bytes(int64_t M, int64_t N) {
  return (M < 0 || N < 0) ? 256 : eb::carve(nullptr, M, N).total;
}

extern "C" int llp_edge_plan(const int64_t* u, const int64_t* v, int64_t M, int64_t N, int32_t* rowptr, int32_t* meta,
                             void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(M >= 0 && N >= 0 && rowptr && workspace && (M == 0 || (u && v && meta)));
  LLP_CHECK_ARG(2 * M < (int64_t)INT32_MAX && N < (int64_t)INT32_MAX);
  if (int rc = check_device()) return rc;
  if (workspace_bytes < llp_edge_plan_workspace_bytes(M, N)) return LLP_E_WORKSPACE;
  const int64_t E = 2 * M;
  eb::Workspace w = eb::carve(reinterpret_cast<char*>(workspace), M, N);
  int2* meta2 = reinterpret_cast<int2*>(meta);
  LLP_CUDA(cudaMemsetAsync(rowptr, 0, (size_t)(N + 1) * sizeof(int32_t), stream));
  if (E == 0) return 0;
  LLP_CUDA(cudaMemsetAsync(w.big_count, 0, 4 * sizeof(int32_t), stream));
  const unsigned e_blocks = (unsigned)ceil_div(E, 256);
  eb::incidence_count_kernel<<<e_blocks, 256, 0, stream>>>(u, v, M, N, rowptr, w.slot);
  LLP_LAUNCH_OK();
  if (int rc = exclusive_scan_i32(rowptr, N + 1, w.tile_sum, stream)) return rc;
  eb::incidence_scatter_kernel<<<e_blocks, 256, 0, stream>>>(u, v, M, rowptr, w.slot, w.ids);
  LLP_LAUNCH_OK();
  eb::incidence_order_rows_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, stream>>>(u, v, M, N, rowptr, w.ids, meta2, w.big_count,
                                                                                 w.big_list, w.huge_list, w.mid_list);
  LLP_LAUNCH_OK();
  eb::incidence_order_big_kernel<<<kNumSMs * 2, eb::kBigThreads, 0, stream>>>(u, v, M, rowptr, w.ids, meta2, w.big_count, w.big_list,
                                                                              w.huge_list, w.mid_list);
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
