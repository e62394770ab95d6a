// Peer memory over NVLink / NVSwitch for the node-partitioned encoder: one process per GPU
// exports a device buffer through CUDA IPC, every other rank maps it, and kernels of this library then LOAD remote rows
// themselves (spmm.cu: spmm_stream_kernel<kPeer = true>) instead of waiting for a staged all-gather.
//
// The only synchronisation the data path needs is "every rank's block is in place" / "every rank has finished reading":
// llp_peer_barrier is a one-block kernel over W flag words per rank.  Rank r stores its epoch into slot r of EVERY rank's
// flag array (st.release.sys over NVLink) and then waits until all W slots of its OWN array have reached the epoch
// (ld.acquire.sys).  The epoch lives in device memory and is advanced by the kernel, so a captured CUDA graph replays the
// barrier correctly.  The wait is bounded (~2 s of clock64): a rank that never arrives sets the sticky error word of the flag
// array (slot W + 1, read by the caller) instead of hanging the GPU.
#include <string.h>

#include "common.cuh"

namespace llp {

constexpr int kPeerMaxRanks = 32;

struct PeerFlagPtrs { unsigned long long* p[kPeerMaxRanks]; };

// flags layout per rank (uint64): [0, W) arrival epochs written by the peers, [W] this rank's epoch counter, [W+1] error word
__global__ void peer_barrier_kernel(PeerFlagPtrs peers, int rank, int world, long long timeout_clocks) {
  unsigned long long* mine = peers.p[rank];
  __shared__ unsigned long long epoch_s;
  if (threadIdx.x == 0) {
    epoch_s = mine[world] + 1;
    mine[world] = epoch_s;
  }
  __syncthreads();
  const unsigned long long epoch = epoch_s;
  const int r = threadIdx.x;
  if (r < world) {
    __threadfence_system();   // everything this rank wrote before the barrier (earlier kernels included) is visible first
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peers.p[r] + rank), "l"(epoch) : "memory");
    const long long t0 = clock64();
    unsigned long long seen = 0;
    while (true) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine + r) : "memory");
      if (seen >= epoch) break;
      if (clock64() - t0 > timeout_clocks) {
        atomicExch(mine + world + 1, 1ull);
        break;
      }
      __nanosleep(100);
    }
  }
  __syncthreads();
}

// Pull `n_rows` rows out of the peers' blocks into a local staging matrix: row i comes from rank (src[i] >> shift), row
// (src[i] & mask) of that rank's block.  This is the NVLink half of a node-partitioned aggregation: each rank fetches
// every remote row its local messages reference exactly ONCE (the list is built with the graph), then aggregates
// locally.  A row is covered by row_bytes / 16 lanes (16-byte vectors); every lane keeps kPeerRowsInFlight rows in flight.
constexpr int kPeerRowsInFlight = 8;
__global__ void __launch_bounds__(256)
peer_gather_rows_kernel(const unsigned long long* __restrict__ table, const int32_t* __restrict__ src,
                        const int32_t* __restrict__ dst_rows /* or null: row i goes to dst row i */, int shift, int64_t n_rows,
                        int row_vecs /* 16-byte vectors per row */, uint4* __restrict__ dst) {
  const int rows_per_pass = 256 / row_vecs;                  // rows one block covers with one load per thread
  const int sub = threadIdx.x / row_vecs, v = threadIdx.x % row_vecs;
  if (sub >= rows_per_pass) return;
  const int64_t stride = (int64_t)gridDim.x * rows_per_pass;
  for (int64_t i0 = (int64_t)blockIdx.x * rows_per_pass + sub; i0 < n_rows; i0 += stride * kPeerRowsInFlight) {
    uint4 t[kPeerRowsInFlight];
#pragma unroll
    for (int k = 0; k < kPeerRowsInFlight; ++k) {
      const int64_t i = i0 + k * stride;
      if (i < n_rows) {
        const int s = __ldg(src + i);
        const uint4* row = reinterpret_cast<const uint4*>(__ldg(table + (s >> shift))) + (int64_t)(s & ((1 << shift) - 1)) * row_vecs;
        asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t[k].x), "=r"(t[k].y), "=r"(t[k].z), "=r"(t[k].w) : "l"(row + v));
      }
    }
#pragma unroll
    for (int k = 0; k < kPeerRowsInFlight; ++k) {
      const int64_t i = i0 + k * stride;
      if (i < n_rows) dst[(dst_rows != nullptr ? (int64_t)__ldg(dst_rows + i) : i) * row_vecs + v] = t[k];
    }
  }
}

// ---- sparse return of the embedding gradient (the transpose of the pull by node id) -------------------------------------
// Every rank q has published (a) the node ids its edge shard scored, as {count, ids...} in an exported int32 buffer, and
// (b) its gradient rows for exactly those nodes, in their places of an exported [N_padded, F] matrix.  The OWNER of a
// node block then (1) marks, per source rank, which of its rows that rank touched and (2) pulls and adds the marked rows
// in rank order 0..W-1 (fp32 accumulation, fixed order: deterministic) — 2B/W rows per rank over NVLink instead of a dense
// reduce-scatter of all N rows.
__global__ void peer_mark_rows_kernel(const unsigned long long* __restrict__ ids_table, int64_t lo, int64_t n_loc, int64_t max_ids,
                                      uint8_t* __restrict__ mark) {
  const int q = blockIdx.y;
  const int32_t* ids = reinterpret_cast<const int32_t*>(__ldg(ids_table + q));
  const int64_t count = min((int64_t)ids[0], max_ids);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t id = ids[1 + i];
    if (id >= lo && id < lo + n_loc) mark[(int64_t)q * n_loc + (id - lo)] = 1;
  }
}

// one warp per owned row; lane l covers the 16-byte vectors l, l + 32, ... of the row
template <typename T>
__global__ void __launch_bounds__(256)
peer_reduce_rows_kernel(const unsigned long long* __restrict__ g_table, int world, const uint8_t* __restrict__ mark, int64_t lo,
                        int64_t n_loc, int F, int64_t ld, T* __restrict__ out, int64_t ldo) {
  constexpr int VE = 16 / (int)sizeof(T);
  const int lane = threadIdx.x & 31;
  const int64_t n = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (n >= n_loc) return;
  for (int c0 = lane * VE; c0 < F; c0 += 32 * VE) {
    float acc[VE];
#pragma unroll
    for (int i = 0; i < VE; ++i) acc[i] = 0.0f;
    for (int q = 0; q < world; ++q) {
      if (!mark[(int64_t)q * n_loc + n]) continue;   // warp-uniform
      const T* row = reinterpret_cast<const T*>(__ldg(g_table + q)) + (lo + n) * ld + c0;
      uint4 v;
      asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(row));
      if constexpr (sizeof(T) == 4) {
        acc[0] += __uint_as_float(v.x); acc[1] += __uint_as_float(v.y); acc[2] += __uint_as_float(v.z); acc[3] += __uint_as_float(v.w);
      } else {
        acc[0] += __uint_as_float(v.x << 16); acc[1] += __uint_as_float(v.x & 0xffff0000u);
        acc[2] += __uint_as_float(v.y << 16); acc[3] += __uint_as_float(v.y & 0xffff0000u);
        acc[4] += __uint_as_float(v.z << 16); acc[5] += __uint_as_float(v.z & 0xffff0000u);
        acc[6] += __uint_as_float(v.w << 16); acc[7] += __uint_as_float(v.w & 0xffff0000u);
      }
    }
    T* dst = out + n * ldo + c0;
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<uint4*>(dst) = make_uint4(__float_as_uint(acc[0]), __float_as_uint(acc[1]), __float_as_uint(acc[2]), __float_as_uint(acc[3]));
    } else {
      *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
    }
  }
}

}  // namespace llp

using namespace llp;

typedef unsigned int (*CuMemGetAddressRangeFn)(unsigned long long*, size_t*, unsigned long long);

// Export the allocation that holds `ptr` (any device pointer inside a cudaMalloc'ed block, e.g. a torch tensor of the
// default caching allocator): 64-byte IPC handle of the block + the offset of `ptr` inside it.
extern "C" int llp_ipc_export(const void* ptr, void* handle64, int64_t* offset) {
  LLP_CHECK_ARG(ptr && handle64 && offset);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  LLP_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q));
  if (fn == nullptr) return LLP_E_DEVICE;
  unsigned long long base = 0;
  size_t size = 0;
  if (((CuMemGetAddressRangeFn)fn)(&base, &size, (unsigned long long)ptr) != 0) return LLP_E_BADARG;
  cudaIpcMemHandle_t h;
  LLP_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
  memcpy(handle64, &h, 64);
  *offset = (int64_t)((unsigned long long)ptr - base);
  return 0;
}

// Map a peer's exported allocation into this process (peer access is enabled lazily by the driver) and return the
// address that corresponds to the exporter's `ptr`.  `base_out` is what llp_ipc_close takes.
extern "C" int llp_ipc_open(const void* handle64, int64_t offset, void** ptr_out, void** base_out) {
  LLP_CHECK_ARG(handle64 && ptr_out && base_out && offset >= 0);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* base = nullptr;
  LLP_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  *base_out = base;
  *ptr_out = reinterpret_cast<char*>(base) + offset;
  return 0;
}

extern "C" int llp_ipc_close(void* base) {
  if (base == nullptr) return 0;
  LLP_CUDA(cudaIpcCloseMemHandle(base));
  return 0;
}

// flags[r]: rank r's flag array (uint64[world + 2], zero-initialised, peer-mapped) as seen from THIS process.
extern "C" int llp_peer_barrier(void* const* flags, int rank, int world, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(flags && world >= 1 && world <= kPeerMaxRanks && rank >= 0 && rank < world);
  if (int rc = check_device()) return rc;
  PeerFlagPtrs p;
  for (int r = 0; r < kPeerMaxRanks; ++r) p.p[r] = r < world ? reinterpret_cast<unsigned long long*>(flags[r]) : nullptr;
  for (int r = 0; r < world; ++r) LLP_CHECK_ARG(p.p[r] != nullptr);
  static int khz = 0;   // one driver query per process (the attribute is slow to read: ~1 ms)
  if (khz == 0) {
    int dev = 0, v = 0;
    LLP_CUDA(cudaGetDevice(&dev));
    LLP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, dev));
    khz = v > 0 ? v : 1500000;
  }
  const long long timeout = (long long)khz * 2000ll;   // ~2 s
  peer_barrier_kernel<<<1, kPeerMaxRanks, 0, stream>>>(p, rank, world, timeout);
  LLP_LAUNCH_OK();
  return 0;
}

// dst[dst_rows ? dst_rows[i] : i, :] = row (src[i] & mask) of rank (src[i] >> shift)'s block, i < n_rows; row_bytes a
// multiple of 16 up to 4096; peer_x = DEVICE table of the blocks' base pointers as mapped into this process.  With
// dst_rows several i may name the same destination row as long as they name the same source row.
extern "C" int llp_peer_gather_rows(const void* const* peer_x, const int32_t* src, const int32_t* dst_rows, int shift,
                                    int64_t n_rows, int64_t row_bytes, void* dst, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(peer_x && n_rows >= 0 && shift >= 0 && shift < 31 && row_bytes > 0 && row_bytes % 16 == 0 && row_bytes <= 4096);
  if (int rc = check_device()) return rc;
  if (n_rows == 0) return 0;
  LLP_CHECK_ARG(src && dst && aligned(dst, 16));
  const int row_vecs = (int)(row_bytes / 16);
  const int rows_per_pass = 256 / row_vecs;
  const int64_t want = ceil_div(n_rows, (int64_t)rows_per_pass * kPeerRowsInFlight);
  const unsigned blocks = (unsigned)(want < 148 * 8 ? (want > 0 ? want : 1) : 148 * 8);
  peer_gather_rows_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const unsigned long long*>(peer_x), src, dst_rows, shift, n_rows,
                                                      row_vecs, reinterpret_cast<uint4*>(dst));
  LLP_LAUNCH_OK();
  return 0;
}

// mark[q * n_loc + (id - lo)] = 1 for every id that rank q published and that lies in [lo, lo + n_loc); ids_table = DEVICE
// array of `world` pointers to the ranks' int32 id buffers {count, id_0, id_1, ...} (peer-mapped).  `mark` must be zeroed
// by the caller.
extern "C" int llp_peer_mark_rows(const void* const* ids_table, int world, int64_t lo, int64_t n_loc, int64_t max_ids,
                                  void* mark, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(ids_table && mark && world >= 1 && world <= kPeerMaxRanks && lo >= 0 && n_loc > 0 && max_ids >= 0);
  if (int rc = check_device()) return rc;
  if (max_ids == 0) return 0;
  const int64_t bx = ceil_div(max_ids, 256);
  dim3 grid((unsigned)(bx < 256 ? bx : 256), (unsigned)world);
  peer_mark_rows_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const unsigned long long*>(ids_table), lo, n_loc, max_ids,
                                                  reinterpret_cast<uint8_t*>(mark));
  LLP_LAUNCH_OK();
  return 0;
}

// out[n, :] = sum over q = 0..world-1 with mark[q * n_loc + n] of row (lo + n) of rank q's [N_padded, ld] matrix
// (g_table = DEVICE array of peer-mapped base pointers), fp32 accumulation in rank order; rows nobody marked become zero.
extern "C" int llp_peer_reduce_rows(int dtype, const void* const* g_table, int world, const void* mark, int64_t lo, int64_t n_loc,
                                    int64_t feat, int64_t ld, void* out, int64_t ldo, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(g_table && mark && out && world >= 1 && world <= kPeerMaxRanks && lo >= 0 && n_loc > 0 && feat > 0 && ld >= feat && ldo >= feat);
  if (int rc = check_device()) return rc;
  const int64_t elt = dtype == LLP_F32 ? 4 : 2;
  if ((feat * elt) % 16 != 0 || (ld * elt) % 16 != 0 || (ldo * elt) % 16 != 0 || !aligned(out, 16)) return LLP_E_ALIGN;
  const unsigned blocks = (unsigned)ceil_div(n_loc * 32, 256);
  if (dtype == LLP_F32)
    peer_reduce_rows_kernel<float><<<blocks, 256, 0, stream>>>(reinterpret_cast<const unsigned long long*>(g_table), world,
        reinterpret_cast<const uint8_t*>(mark), lo, n_loc, (int)feat, ld, reinterpret_cast<float*>(out), ldo);
  else if (dtype == LLP_BF16)
    peer_reduce_rows_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(reinterpret_cast<const unsigned long long*>(g_table), world,
        reinterpret_cast<const uint8_t*>(mark), lo, n_loc, (int)feat, ld, reinterpret_cast<__nv_bfloat16*>(out), ldo);
  else
    return LLP_E_BADARG;
  LLP_LAUNCH_OK();
  return 0;
}
