// tcgen05 GEMMs for sm_100a: TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring ->
// tcgen05.mma (bf16 x bf16 -> fp32 in TMEM, issued by one thread) -> tcgen05.ld epilogue.
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (one TMEM lane quadrant each).  Two accumulator stages in TMEM so the
// epilogue of tile i overlaps the MMAs of tile i+1.
//
//   NT  : D[M,N]   = epi(A1[M,K1]*B1[N,K1]^T + A2[M,K2]*B2[N,K2]^T)      (both operands K-major)
//         replaces F.linear / lin_l+lin_r (models.py:48,143; PyG SAGEConv; sageconv_updated.py:71,76)
//   TN  : D[N1,N2] = A[M,N1]^T * B[M,N2]  split over M                    (both operands MN-major)
//         the weight gradient of the same layers; fp32 partials + fixed-order reduce.
#include "tc_epilogue.cuh"

namespace llp {

int splitk_reduce(const float* partial, int splits, int64_t rows, int64_t cols, float* D, int64_t ldd, int accumulate,
                  cudaStream_t stream);

namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // bf16 elements = 128 bytes = one swizzle row
constexpr int kEpiWarps = 8;             // two warps per TMEM lane quadrant, each takes half of the tile's columns
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kAccStages = 2;
constexpr int kSlabBytes = BLOCK_K * 128;  // 64 rows x 128 B: one TMA box of the MN-major layout

template <int BLOCK_N>
struct Config {
  static constexpr int kABytes = BLOCK_M * 128;
  static constexpr int kBBytes = BLOCK_N * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int kTmemCols = kAccStages * BLOCK_N;  // 128 / 256 / 512: powers of two >= 32
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// ------------------------------------------------------------------------------------------------
template <int BLOCK_N, bool kTN, typename TO>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ Maps maps, const TcParams p) {
  using Cfg = Config<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                          // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;          // [kStages]
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;      // [kAccStages]
  uint64_t* tmem_empty = tmem_full + kAccStages;      // [kAccStages]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + kAccStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int64_t m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M, n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int64_t num_tiles = m_tiles * n_tiles * p.splits;
  const int kb1 = kTN ? 0 : (int)((p.K1 + BLOCK_K - 1) / BLOCK_K);
  const int kb2 = kTN ? 0 : (int)((p.K2 + BLOCK_K - 1) / BLOCK_K);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a1);
    tma_prefetch_desc(&maps.b1);
    if (kb2 > 0) { tma_prefetch_desc(&maps.a2); tma_prefetch_desc(&maps.b2); }
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int s = 0; s < kAccStages; ++s) { mbar_init(smem_u32(&tmem_full[s]), 1); mbar_init(smem_u32(&tmem_empty[s]), kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_holder), Cfg::kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the ring, one elected lane issues) =====================
    int stage = 0; uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int64_t split = tile / (m_tiles * n_tiles);
      const int64_t mn = tile % (m_tiles * n_tiles);
      const int m0 = (int)(mn / n_tiles) * BLOCK_M, n0 = (int)(mn % n_tiles) * BLOCK_N;
      int num_kb; int64_t k_begin = 0;
      if constexpr (kTN) {
        k_begin = split * p.k_per_split;
        int64_t k_end = k_begin + p.k_per_split < p.K1 ? k_begin + p.k_per_split : p.K1;
        num_kb = (int)((k_end - k_begin + BLOCK_K - 1) / BLOCK_K);
      } else {
        num_kb = kb1 + kb2;
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        if (elect_one_sync()) {
        const uint32_t bar = smem_u32(&full_bar[stage]);
        mbar_expect_tx(bar, Cfg::kStageBytes);
        const uint32_t sa = smem_u32(smem_a + stage * Cfg::kABytes), sb = smem_u32(smem_b + stage * Cfg::kBBytes);
        if constexpr (kTN) {
          // MN-major: boxes of {64 columns, 64 reduction rows}; one box per 64-column slab
          const int k0 = (int)(k_begin + (int64_t)kb * BLOCK_K);
#pragma unroll
          for (int j = 0; j < BLOCK_M / 64; ++j) tma_load_2d(sa + j * kSlabBytes, &maps.a1, m0 + j * 64, k0, bar);
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j) tma_load_2d(sb + j * kSlabBytes, &maps.b1, n0 + j * 64, k0, bar);
        } else {
          const bool second = kb >= kb1;
          const int k0 = (second ? kb - kb1 : kb) * BLOCK_K;
          tma_load_2d(sa, second ? &maps.a2 : &maps.a1, k0, m0, bar);
          tma_load_2d(sb, second ? &maps.b2 : &maps.b1, k0, n0, bar);
        }
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp walks the pipeline, one elected lane issues) =====================
    constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, kTN);
    constexpr uint32_t kHi = desc_hi_sw128(1024);
    // K-major SW128: LBO 16 B, one UMMA_K step = 32 B inside the swizzle row.  MN-major SW128: LBO = distance between
    // 64-element MN slabs, one UMMA_K step = 16 reduction rows = 2048 B.  SBO = 8 rows (1024 B) in both.
    constexpr uint32_t kStep = (kTN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
    const uint32_t a_lo0 = desc_lo(smem_u32(smem_a), kTN ? kSlabBytes : 16);
    const uint32_t b_lo0 = desc_lo(smem_u32(smem_b), kTN ? kSlabBytes : 16);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    long long t_begin = 0, w_full = 0, w_acc = 0, t0 = 0;
    auto now = []() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    if (p.dbg) t_begin = now();
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int num_kb;
      if constexpr (kTN) {
        const int64_t split = tile / (m_tiles * n_tiles);
        int64_t k_begin = split * p.k_per_split;
        int64_t k_end = k_begin + p.k_per_split < p.K1 ? k_begin + p.k_per_split : p.K1;
        num_kb = (int)((k_end - k_begin + BLOCK_K - 1) / BLOCK_K);
      } else {
        num_kb = kb1 + kb2;
      }
      if (p.dbg) t0 = now();
      mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
      if (p.dbg) w_acc += now() - t0;
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int kb = 0; kb < num_kb; ++kb) {
        if (p.dbg) t0 = now();
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        if (p.dbg) w_full += now() - t0;
        tcgen05_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * (Cfg::kABytes >> 4), b_lo = b_lo0 + (uint32_t)stage * (Cfg::kBBytes >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16(tmem_d, desc_from(a_lo + k * kStep, kHi), desc_from(b_lo + k * kStep, kHi), idesc, (uint32_t)((kb | k) != 0));
          umma_commit(smem_u32(&empty_bar[stage]));  // frees the smem slot once these MMAs retire
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (elect_one_sync()) umma_commit(smem_u32(&tmem_full[acc]));      // accumulator ready for the epilogue
      __syncwarp();
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
    if (p.dbg && lane == 0) {
      p.dbg[blockIdx.x * 4 + 0] = now() - t_begin;
      p.dbg[blockIdx.x * 4 + 1] = w_full;
      p.dbg[blockIdx.x * 4 + 2] = w_acc;
    }
  } else if (warp >= 2) {
    // ===================== epilogue: TMEM -> registers -> global =====================
    TcParams pe = p;                      // thread-local copy: the dropout stream is resolved once (device rng_state)
    if (pe.ep.dropout_p > 0.0f) resolve_rng(pe.ep);
    const int quad = warp & 3;            // TMEM lanes [32*quad, 32*quad+32) are the only ones this warp may read
    const int half = (warp - 2) >> 2;     // which half of the tile's columns
    constexpr int kColsPerWarp = BLOCK_N / 2;
    int acc = 0; uint32_t acc_phase = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int64_t split = tile / (m_tiles * n_tiles);
      const int64_t mn = tile % (m_tiles * n_tiles);
      const int64_t m0 = (mn / n_tiles) * BLOCK_M, n0 = (mn % n_tiles) * BLOCK_N;
      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
      tcgen05_fence_after();
      const int64_t m = m0 + quad * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
      uint4 rnd128 = make_uint4(0, 0, 0, 0);
      uint32_t rnd_group = 0xffffffffu;  // which 128-column Philox block rnd128 holds (p == 0.5 fast path)
#pragma unroll 1
      for (int c0 = half * kColsPerWarp; c0 < (half + 1) * kColsPerWarp; c0 += 32) {
        if (n0 + c0 >= p.N) break;  // warp-uniform
        if constexpr (!kTN) {
          if (pe.ep.dropout_p == 0.5f && (uint32_t)((n0 + c0) >> 7) != rnd_group) {
            rnd_group = (uint32_t)((n0 + c0) >> 7);
            rnd128 = philox4x32_10(pe.ep.seed, (uint64_t)m, pe.ep.offset + (uint64_t)rnd_group);
          }
        }
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        if (m < p.M) {
          if constexpr (kTN) {
            float* dst = p.partial + ((int64_t)split * p.M + m) * p.N + n0 + c0;
            const bool vec = (p.N % 4 == 0) && (n0 + c0 + 32 <= p.N);
            if (vec) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(dst + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
            } else {
              for (int j = 0; j < 32; ++j)
                if (n0 + c0 + j < p.N) dst[j] = __uint_as_float(r[j]);
            }
          } else {
            epilogue_chunk<TO>(r, m, n0 + c0, pe, rnd128, pe.ep.seed, pe.ep.offset);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// NT with the weight operand RESIDENT in shared memory.  The plain kernel above re-reads the whole [N, K] weight
// matrix from L2 for every 128-row tile (256 KB of B per 128 KB of A at K = 512), which makes tall-skinny layer GEMMs
// L2->SM-traffic bound.  Here a persistent CTA owns ONE column tile of the output: its B operand (all K blocks, <= 128 KB)
// is loaded once, the ring only streams A tiles, and the CTA walks the row tiles m = first, first + stride, ...
// (N = 256, K = 512: two column tiles of 128, so A is read twice -- the second read hits L2 because the two CTAs of a
// row tile run side by side -- instead of B 1843 times.)
template <int BLOCK_N, typename TO>
__global__ void __launch_bounds__(kThreads, 1)
gemm_nt_resb_kernel(const __grid_constant__ Maps maps, const TcParams p, const int stages, long long* dbg) {
  constexpr int kABytes = BLOCK_M * 128, kBBytes = BLOCK_N * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int kb1 = (int)((p.K1 + BLOCK_K - 1) / BLOCK_K), kb2 = (int)((p.K2 + BLOCK_K - 1) / BLOCK_K);
  const int num_kb = kb1 + kb2;
  uint8_t* smem_b = smem;                                   // [num_kb][BLOCK_N rows x 128 B], resident
  uint8_t* smem_a = smem + (size_t)num_kb * kBBytes;        // [stages][128 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + (size_t)stages * kABytes);
  uint64_t* full_bar = bars;                     // [stages]
  uint64_t* empty_bar = bars + stages;           // [stages]
  uint64_t* tmem_full = bars + 2 * stages;       // [kAccStages]
  uint64_t* tmem_empty = tmem_full + kAccStages; // [kAccStages]
  uint64_t* b_full = tmem_empty + kAccStages;    // [1]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(b_full + 1);
  constexpr int kTmemCols = kAccStages * BLOCK_N;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (int)((p.N + BLOCK_N - 1) / BLOCK_N);
  const int n0 = (int)(blockIdx.x % n_tiles) * BLOCK_N;
  const int64_t m_first = blockIdx.x / n_tiles, m_stride = gridDim.x / n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a1);
    tma_prefetch_desc(&maps.b1);
    if (kb2 > 0) { tma_prefetch_desc(&maps.a2); tma_prefetch_desc(&maps.b2); }
    for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int s = 0; s < kAccStages; ++s) { mbar_init(smem_u32(&tmem_full[s]), 1); mbar_init(smem_u32(&tmem_empty[s]), kEpiWarps); }
    mbar_init(smem_u32(b_full), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_holder), kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer: B once, then the A ring =====================
    if (m_first < m_tiles && elect_one_sync()) {
      mbar_expect_tx(smem_u32(b_full), (uint32_t)(num_kb * kBBytes));
      for (int kb = 0; kb < num_kb; ++kb) {
        const bool second = kb >= kb1;
        tma_load_2d(smem_u32(smem_b + (size_t)kb * kBBytes), second ? &maps.b2 : &maps.b1, (second ? kb - kb1 : kb) * BLOCK_K,
                    n0, smem_u32(b_full));
      }
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    for (int64_t mt = m_first; mt < m_tiles; mt += m_stride) {
      const int m0 = (int)mt * BLOCK_M;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        if (elect_one_sync()) {
          const uint32_t bar = smem_u32(&full_bar[stage]);
          mbar_expect_tx(bar, kABytes);
          const bool second = kb >= kb1;
          tma_load_2d(smem_u32(smem_a + (size_t)stage * kABytes), second ? &maps.a2 : &maps.a1,
                      (second ? kb - kb1 : kb) * BLOCK_K, m0, bar);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, false);
    constexpr uint32_t kHi = desc_hi_sw128(1024);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem_a), 16), b_lo0 = desc_lo(smem_u32(smem_b), 16);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    long long t_begin = 0, w_full = 0, w_acc = 0, t0 = 0, t1 = 0;
    auto now = []() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    if (dbg) t_begin = now();
    if (m_first < m_tiles) {
      mbar_wait(smem_u32(b_full), 0);
      tcgen05_fence_after();
    }
    for (int64_t mt = m_first; mt < m_tiles; mt += m_stride) {
      if (dbg) t0 = now();
      mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
      if (dbg) w_acc += now() - t0;
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int kb = 0; kb < num_kb; ++kb) {
        if (dbg) t1 = now();
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        if (dbg) w_full += now() - t1;
        tcgen05_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * (kABytes >> 4), b_lo = b_lo0 + (uint32_t)kb * (kBBytes >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16(tmem_d, desc_from(a_lo + k * 2, kHi), desc_from(b_lo + k * 2, kHi), idesc, (uint32_t)((kb | k) != 0));
          umma_commit(smem_u32(&empty_bar[stage]));
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one_sync()) umma_commit(smem_u32(&tmem_full[acc]));
      __syncwarp();
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
    if (dbg && lane == 0) {
      dbg[blockIdx.x * 4 + 0] = now() - t_begin;   // issue-loop time of this CTA
      dbg[blockIdx.x * 4 + 1] = w_full;            // waiting for operands (load-bound)
      dbg[blockIdx.x * 4 + 2] = w_acc;             // waiting for a free accumulator (epilogue-bound)
    }
  } else if (warp >= 2) {
    // ===================== epilogue =====================
    TcParams pe = p;
    if (pe.ep.dropout_p > 0.0f) resolve_rng(pe.ep);
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int kColsPerWarp = BLOCK_N / 2;
    int acc = 0; uint32_t acc_phase = 0;
    for (int64_t mt = m_first; mt < m_tiles; mt += m_stride) {
      const int64_t m0 = mt * BLOCK_M;
      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
      tcgen05_fence_after();
      const int64_t m = m0 + quad * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
      uint4 rnd128 = make_uint4(0, 0, 0, 0);
      uint32_t rnd_group = 0xffffffffu;
#pragma unroll 1
      for (int c0 = half * kColsPerWarp; c0 < (half + 1) * kColsPerWarp; c0 += 32) {
        if (n0 + c0 >= p.N) break;
        if (pe.ep.dropout_p == 0.5f && (uint32_t)((n0 + c0) >> 7) != rnd_group) {
          rnd_group = (uint32_t)((n0 + c0) >> 7);
          rnd128 = philox4x32_10(pe.ep.seed, (uint64_t)m, pe.ep.offset + (uint64_t)rnd_group);
        }
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        if (m < p.M) epilogue_chunk<TO>(r, m, n0 + c0, pe, rnd128, pe.ep.seed, pe.ep.offset);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

constexpr int kResBBudget = 128 * 1024;   // bytes of shared memory the resident B operand may take
constexpr int kResBSmemMax = 226 * 1024;

template <int BLOCK_N, typename TO>
static int launch_resb(const Maps& maps, const TcParams& p, int num_kb, cudaStream_t stream) {
  auto kern = gemm_nt_resb_kernel<BLOCK_N, TO>;
  static PerDeviceOnce configured;  // cudaFuncSetAttribute is per device
  if (configured.need()) {
    LLP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kResBSmemMax));
    configured.done();
  }
  const int b_bytes = num_kb * BLOCK_N * 128;
  int stages = (kResBSmemMax - 1024 - 512 - b_bytes) / (BLOCK_M * 128);
  if (stages > 8) stages = 8;
  if (g_tuning[17] > 0 && g_tuning[17] < stages) stages = g_tuning[17];
  const int smem = b_bytes + stages * BLOCK_M * 128 + 1024 + 512;
  const int n_tiles = (int)ceil_div(p.N, BLOCK_N);
  const unsigned grid = (unsigned)(kNumSMs / n_tiles * n_tiles);
  kern<<<grid, kThreads, smem, stream>>>(maps, p, stages, g_tuning[15] ? debug_buffer() : nullptr);
  LLP_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// NT on a CTA PAIR (cta_group::2): one 256 x 256 output tile per pair of SMs, the weight operand resident ACROSS the
// pair.  tcgen05.mma.cta_group::2 reads the A operand (128 rows) from each CTA's own shared memory and the B operand
// split along N: CTA r holds the rows [128 r, 128 r + 128) of W for every K block (K <= 512: 128 KB), loaded once.
// So the ring only streams A tiles and nothing is read twice: L2 -> SM traffic = the activations (the streaming
// kernel above re-reads the 256 KB weight matrix for every 128-row tile: 2/3 of its traffic at K = 512).
// Roles per CTA: warp 0 TMA producer (its own A rows / W half; completions are signalled on the LEADER's barriers),
// warp 1 TMEM allocator (both CTAs) and, in the leader CTA only, the MMA issuer (tcgen05.commit multicasts the stage
// release and the accumulator-ready arrival to both CTAs), warps 2..9 epilogue of the CTA's own 128 rows (they release
// the accumulator on the leader's barrier).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_remote(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
// TMA load whose completion is signalled on a barrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t cluster_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives (once the issued MMAs have retired) on the barrier at this offset in every CTA of the mask
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}

constexpr int kPairN = 256;                 // output columns of the pair tile
constexpr int kPairHalfBytes = 128 * 128;   // one K block of a CTA's W half / one A stage: 128 rows x 128 B

template <typename TO>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_nt_pair_kernel(const __grid_constant__ Maps maps, const TcParams p, const int stages) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int kb1 = (int)((p.K1 + BLOCK_K - 1) / BLOCK_K), kb2 = (int)((p.K2 + BLOCK_K - 1) / BLOCK_K);
  const int num_kb = kb1 + kb2;
  uint8_t* smem_b = smem;                                        // [num_kb][128 rows of W x 128 B], resident
  uint8_t* smem_a = smem + (size_t)num_kb * kPairHalfBytes;      // [stages][128 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + (size_t)stages * kPairHalfBytes);
  uint64_t* full_bar = bars;                     // [stages]   (the leader's are used)
  uint64_t* empty_bar = bars + stages;           // [stages]
  uint64_t* tmem_full = bars + 2 * stages;       // [kAccStages]
  uint64_t* tmem_empty = tmem_full + kAccStages; // [kAccStages] (the leader's are used)
  uint64_t* b_full = tmem_empty + kAccStages;    // [1]   my W half has landed
  uint64_t* peer_full = b_full + 1;              // [stages] leader only: the peer's stage has landed (forwarded arrival)
  uint64_t* b_peer = peer_full + stages;         // [1]   leader only: the peer's W half has landed
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(b_peer + 1);
  constexpr int kTmemCols = kAccStages * kPairN;  // 512

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int64_t pair_tiles = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int64_t pt_first = blockIdx.x >> 1, pt_stride = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a1);
    tma_prefetch_desc(&maps.b1);
    if (kb2 > 0) { tma_prefetch_desc(&maps.a2); tma_prefetch_desc(&maps.b2); }
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); mbar_init(smem_u32(&peer_full[s]), 1);
    }
    for (int s = 0; s < kAccStages; ++s) { mbar_init(smem_u32(&tmem_full[s]), 1); mbar_init(smem_u32(&tmem_empty[s]), 2 * kEpiWarps); }
    mbar_init(smem_u32(b_full), 1);
    mbar_init(smem_u32(b_peer), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(smem_u32(tmem_holder), kTmemCols);
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer: my W half once, then my A rows =====================
    if (pt_first < pair_tiles && elect_one_sync()) {
      const uint32_t bar = smem_u32(b_full);
      mbar_expect_tx(bar, (uint32_t)(num_kb * kPairHalfBytes));
      for (int kb = 0; kb < num_kb; ++kb) {
        const bool second = kb >= kb1;
        tma_load_2d(smem_u32(smem_b + (size_t)kb * kPairHalfBytes), second ? &maps.b2 : &maps.b1,
                    (second ? kb - kb1 : kb) * BLOCK_K, (int)rank * 128, bar);
      }
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    for (int64_t pt = pt_first; pt < pair_tiles; pt += pt_stride) {
      const int m0 = (int)(pt * 2 * BLOCK_M) + (int)rank * BLOCK_M;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        if (elect_one_sync()) {
          const uint32_t bar = smem_u32(&full_bar[stage]);
          mbar_expect_tx(bar, kPairHalfBytes);
          const bool second = kb >= kb1;
          tma_load_2d(smem_u32(smem_a + (size_t)stage * kPairHalfBytes), second ? &maps.a2 : &maps.a1,
                      (second ? kb - kb1 : kb) * BLOCK_K, m0, bar);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (leader CTA) =====================
    constexpr uint32_t idesc = make_idesc(2 * BLOCK_M, kPairN, false);
    constexpr uint32_t kHi = desc_hi_sw128(1024);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem_a), 16), b_lo0 = desc_lo(smem_u32(smem_b), 16);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    long long t_begin = 0, w_full = 0, w_acc = 0, t0 = 0;
    auto now = []() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    if (p.dbg) t_begin = now();
    if (pt_first < pair_tiles) {
      mbar_wait(smem_u32(b_full), 0);
      mbar_wait(smem_u32(b_peer), 0);
      tcgen05_fence_after();
    }
    for (int64_t pt = pt_first; pt < pair_tiles; pt += pt_stride) {
      if (p.dbg) t0 = now();
      mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
      if (p.dbg) w_acc += now() - t0;
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kPairN);
      for (int kb = 0; kb < num_kb; ++kb) {
        if (p.dbg) t0 = now();
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        mbar_wait(smem_u32(&peer_full[stage]), phase);
        if (p.dbg) w_full += now() - t0;
        tcgen05_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * (kPairHalfBytes >> 4), b_lo = b_lo0 + (uint32_t)kb * (kPairHalfBytes >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            if (p.splits == -1) break;   // experiment (llp_set_tuning(11, 1)): no MMAs, pure load pipeline
            umma_bf16_pair(tmem_d, desc_from(a_lo + k * 2, kHi), desc_from(b_lo + k * 2, kHi), idesc, (uint32_t)((kb | k) != 0));
          }
          umma_commit_pair(smem_u32(&empty_bar[stage]), 3);   // both CTAs' stage is free once these MMAs retire
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one_sync()) umma_commit_pair(smem_u32(&tmem_full[acc]), 3);
      __syncwarp();
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
    if (p.dbg && lane == 0) {
      p.dbg[blockIdx.x * 4 + 0] = now() - t_begin;
      p.dbg[blockIdx.x * 4 + 1] = w_full;
      p.dbg[blockIdx.x * 4 + 2] = w_acc;
    }
  } else if (warp == 1) {
    // ===================== peer CTA: forward "my stage has landed" to the leader =====================
    // (TMA completions stay on local barriers; one remote arrive per stage crosses the pair)
    if (pt_first < pair_tiles) {
      mbar_wait(smem_u32(b_full), 0);
      if (lane == 0) mbar_arrive_remote(mapa_shared(smem_u32(b_peer), 0));
      __syncwarp();
    }
    int stage = 0; uint32_t phase = 0;
    for (int64_t pt = pt_first; pt < pair_tiles; pt += pt_stride) {
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        if (lane == 0) mbar_arrive_remote(mapa_shared(smem_u32(&peer_full[stage]), 0));
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 2) {
    // ===================== epilogue of this CTA's 128 rows =====================
    TcParams pe = p;
    if (pe.ep.dropout_p > 0.0f) resolve_rng(pe.ep);
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int kColsPerWarp = kPairN / 2;
    int acc = 0; uint32_t acc_phase = 0;
    for (int64_t pt = pt_first; pt < pair_tiles; pt += pt_stride) {
      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
      tcgen05_fence_after();
      const int64_t m = pt * 2 * BLOCK_M + (int64_t)rank * BLOCK_M + quad * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kPairN);
      uint4 rnd128 = make_uint4(0, 0, 0, 0);
      uint32_t rnd_group = 0xffffffffu;
#pragma unroll 1
      for (int c0 = half * kColsPerWarp; c0 < (half + 1) * kColsPerWarp; c0 += 32) {
        if (c0 >= p.N) break;
        if (pe.ep.dropout_p == 0.5f && (uint32_t)(c0 >> 7) != rnd_group) {
          rnd_group = (uint32_t)(c0 >> 7);
          rnd128 = philox4x32_10(pe.ep.seed, (uint64_t)m, pe.ep.offset + (uint64_t)rnd_group);
        }
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        if (m < p.M) epilogue_chunk<TO>(r, m, c0, pe, rnd128, pe.ep.seed, pe.ep.offset);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(smem_u32(&tmem_empty[acc]));
        else mbar_arrive_remote(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer may still signal its barriers / read its operands
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

template <typename TO>
static int launch_pair(const Maps& maps, const TcParams& p, int num_kb, cudaStream_t stream) {
  auto kern = gemm_nt_pair_kernel<TO>;
  static PerDeviceOnce configured;  // cudaFuncSetAttribute is per device
  if (configured.need()) {
    LLP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kResBSmemMax));
    configured.done();
  }
  const int b_bytes = num_kb * kPairHalfBytes;
  int stages = (kResBSmemMax - 1024 - 512 - b_bytes) / kPairHalfBytes;
  if (stages > 8) stages = 8;
  const int smem = b_bytes + stages * kPairHalfBytes + 1024 + 512;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(kNumSMs / 2 * 2));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = stream;
  LLP_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, p, stages));   // cluster dims come from __cluster_dims__
  LLP_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// NT on TALL tiles: 256 output rows per weight pass.  The streaming kernel above re-reads the whole [N, K] weight matrix
// from L2 for every 128-row tile — at K = 512 that is 256 KB of W per 128 KB of activations, and the kernels' L2 -> SM
// traffic (not HBM) is what bounds them (~9.4 TB/s measured).  Here one weight k-block in shared memory feeds TWO 128-row
// MMAs (rows [m0, m0+128) -> TMEM columns [0, 256), rows [m0+128, m0+256) -> columns [256, 512)): a third less L2 -> SM
// traffic, no second CTA involved.  The price: the two accumulators take all 512 TMEM columns, so the epilogue of a tile
// does not overlap the MMAs of the next one (the TMA ring keeps filling meanwhile) — and that price is too high on B200:
// 100 us against the streaming kernel's 76 us at M = 235,868, K = 512 (tools/kbench.py tall).  Opt-in only.  Every CTA owns a contiguous range of
// 128-row units (12 or 13 of the 1,843 at the collab size: max/mean = 1.04 instead of 7 / 6.23 whole tall tiles); an odd
// unit at the end of the range is a half tile.  Same MMA order per output element as the streaming kernel: bit-identical.
constexpr int kTallStages = 3;
constexpr int kTallStageBytes = 2 * BLOCK_M * 128 + 256 * 128;   // A rows 0..127 | A rows 128..255 | W k-block
constexpr int kTallSmemBytes = kTallStages * kTallStageBytes + 1024 + 256;

template <typename TO>
__global__ void __launch_bounds__(kThreads, 1)
gemm_nt_tall_kernel(const __grid_constant__ Maps maps, const TcParams p) {
  constexpr int BLOCK_N = 256;
  constexpr int kABytes = BLOCK_M * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTallStages * kTallStageBytes);
  uint64_t* full_bar = bars;                       // [kTallStages]
  uint64_t* empty_bar = bars + kTallStages;        // [kTallStages]
  uint64_t* tmem_full = bars + 2 * kTallStages;    // [1]
  uint64_t* tmem_empty = tmem_full + 1;            // [1]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb1 = (int)((p.K1 + BLOCK_K - 1) / BLOCK_K), kb2 = (int)((p.K2 + BLOCK_K - 1) / BLOCK_K);
  const int num_kb = kb1 + kb2;
  // my contiguous range of 128-row units
  const int64_t units = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int64_t base = units / gridDim.x, rem = units % gridDim.x;
  const int64_t u_begin = blockIdx.x * base + ((int64_t)blockIdx.x < rem ? blockIdx.x : rem);
  const int u_count = (int)(base + ((int64_t)blockIdx.x < rem ? 1 : 0));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a1);
    tma_prefetch_desc(&maps.b1);
    if (kb2 > 0) { tma_prefetch_desc(&maps.a2); tma_prefetch_desc(&maps.b2); }
    for (int s = 0; s < kTallStages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(tmem_full), 1);
    mbar_init(smem_u32(tmem_empty), kEpiWarps);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_holder), 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0; uint32_t phase = 0;
    for (int t = 0; t < u_count; t += 2) {
      const bool tall = t + 1 < u_count;
      const int m0 = (int)((u_begin + t) * BLOCK_M);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        if (elect_one_sync()) {
          const uint32_t bar = smem_u32(&full_bar[stage]);
          mbar_expect_tx(bar, (uint32_t)((tall ? 2 : 1) * kABytes + BLOCK_N * 128));
          const uint32_t sa = smem_u32(smem + stage * kTallStageBytes);
          const bool second = kb >= kb1;
          const int k0 = (second ? kb - kb1 : kb) * BLOCK_K;
          const CUtensorMap* ma = second ? &maps.a2 : &maps.a1;
          tma_load_2d(sa, ma, k0, m0, bar);
          if (tall) tma_load_2d(sa + kABytes, ma, k0, m0 + BLOCK_M, bar);
          tma_load_2d(sa + 2 * kABytes, second ? &maps.b2 : &maps.b1, k0, 0, bar);
        }
        __syncwarp();
        if (++stage == kTallStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, false);
    constexpr uint32_t kHi = desc_hi_sw128(1024);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem), 16);
    int stage = 0; uint32_t phase = 0, acc_phase = 0;
    for (int t = 0; t < u_count; t += 2) {
      const bool tall = t + 1 < u_count;
      mbar_wait(smem_u32(tmem_empty), acc_phase ^ 1);
      tcgen05_fence_after();
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tcgen05_fence_after();
        if (elect_one_sync()) {
          const uint32_t a0 = a_lo0 + (uint32_t)stage * (kTallStageBytes >> 4);
          const uint32_t a1 = a0 + (kABytes >> 4), b = a0 + (2 * kABytes >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16(tmem_base, desc_from(a0 + k * 2, kHi), desc_from(b + k * 2, kHi), idesc, (uint32_t)((kb | k) != 0));
          if (tall) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_bf16(tmem_base + BLOCK_N, desc_from(a1 + k * 2, kHi), desc_from(b + k * 2, kHi), idesc, (uint32_t)((kb | k) != 0));
          }
          umma_commit(smem_u32(&empty_bar[stage]));
        }
        __syncwarp();
        if (++stage == kTallStages) { stage = 0; phase ^= 1; }
      }
      if (elect_one_sync()) umma_commit(smem_u32(tmem_full));
      __syncwarp();
      acc_phase ^= 1;
    }
  } else {
    // ===================== epilogue =====================
    TcParams pe = p;
    if (pe.ep.dropout_p > 0.0f) resolve_rng(pe.ep);
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int kColsPerWarp = BLOCK_N / 2;
    uint32_t acc_phase = 0;
    for (int t = 0; t < u_count; t += 2) {
      const int halves = t + 1 < u_count ? 2 : 1;
      mbar_wait(smem_u32(tmem_full), acc_phase);
      tcgen05_fence_after();
      for (int rh = 0; rh < halves; ++rh) {
        const int64_t m = (u_begin + t + rh) * BLOCK_M + quad * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(rh * BLOCK_N);
        uint4 rnd128 = make_uint4(0, 0, 0, 0);
        uint32_t rnd_group = 0xffffffffu;
#pragma unroll 1
        for (int c0 = half * kColsPerWarp; c0 < (half + 1) * kColsPerWarp; c0 += 32) {
          if (c0 >= p.N) break;
          if (pe.ep.dropout_p == 0.5f && (uint32_t)(c0 >> 7) != rnd_group) {
            rnd_group = (uint32_t)(c0 >> 7);
            rnd128 = philox4x32_10(pe.ep.seed, (uint64_t)m, pe.ep.offset + (uint64_t)rnd_group);
          }
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          if (m < p.M) epilogue_chunk<TO>(r, m, c0, pe, rnd128, pe.ep.seed, pe.ep.offset);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(tmem_empty));
      acc_phase ^= 1;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <typename TO>
static int launch_tall(const Maps& maps, const TcParams& p, cudaStream_t stream) {
  auto kern = gemm_nt_tall_kernel<TO>;
  static PerDeviceOnce configured;  // cudaFuncSetAttribute is per device
  if (configured.need()) {
    LLP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTallSmemBytes));
    configured.done();
  }
  const int64_t units = ceil_div(p.M, BLOCK_M);
  const unsigned grid = (unsigned)(units < kNumSMs ? units : kNumSMs);
  kern<<<grid, kThreads, kTallSmemBytes, stream>>>(maps, p);
  LLP_LAUNCH_OK();
  return 0;
}

template <int BLOCK_N, bool kTN, typename TO>
static int launch(const Maps& maps, const TcParams& p, cudaStream_t stream) {
  using Cfg = Config<BLOCK_N>;
  auto kern = gemm_tcgen05_kernel<BLOCK_N, kTN, TO>;
  static PerDeviceOnce configured;  // cudaFuncSetAttribute is per device
  if (configured.need()) {
    LLP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured.done();
  }
  int64_t m_tiles = ceil_div(p.M, BLOCK_M), n_tiles = ceil_div(p.N, BLOCK_N);
  int64_t tiles = m_tiles * n_tiles * p.splits;
  unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
  kern<<<grid, kThreads, Cfg::kSmemBytes, stream>>>(maps, p);
  LLP_LAUNCH_OK();
  return 0;
}

static int pick_block_n(int64_t N) { return N > 128 ? 256 : (N > 64 ? 128 : 64); }

}  // namespace tc

int gemm_nt_tcgen05(const llp_gemm_nt_args& a, cudaStream_t stream) {
  using namespace tc;
  if (a.dtype != LLP_BF16) return LLP_E_SHAPE;
  const bool dual = a.A2 != nullptr && a.K2 > 0;
  int bn = pick_block_n(a.N);
  // resident-B variant: tall problems whose weight column tile (all K blocks) fits in 128 KB of shared memory
  const int num_kb = (int)(ceil_div(a.K1, BLOCK_K) + (dual ? ceil_div(a.K2, BLOCK_K) : 0));
  int res_bn = 0;
  if (g_tuning[16] == 0 && a.M >= (int64_t)BLOCK_M * kNumSMs * 2 && a.N >= 128) {
    if (a.N > 128 && num_kb * 256 * 128 <= kResBBudget) res_bn = 256;
    else if (g_tuning[18] && num_kb * 128 * 128 <= kResBBudget) res_bn = 128;  // half-rate MMAs (N = 128): experiment only
  }
  if (res_bn) bn = res_bn;
  // CTA-pair kernel (opt-in, llp_set_tuning(20, 2)): N in (128, 256], weights of all K blocks fit in 2 x 128 KB, tall M.
  // Bit-identical to the streaming kernel and it removes all weight re-reads, but on B200 it is SLOWER for the layer
  // shapes of this workload (112 us vs 73 us at M = 235,868, K = 512): the tensor pipe is fine (tools/mmabench.cu: 128
  // clk per 256 x 256 x 16 MMA, multicast commits included) and so is the remote signalling scheme (two were tried); what
  // is left is a ~5 us stage round trip across the pair with only 96 KB of A stages per CTA beside the resident weights.
  const bool pair = g_tuning[20] == 2 && !res_bn && a.M >= (int64_t)BLOCK_M * kNumSMs * 2 && a.N > 128 && a.N <= 256 &&
                    num_kb * kPairHalfBytes <= kResBBudget;
  // tall tiles (two 128-row MMAs per weight k-block; opt-in, llp_set_tuning(20, 3)): bit-identical and a third less
  // L2 -> SM traffic, but SLOWER on B200 (100 us vs 76 us at M = 235,868, K = 512): both accumulators fill the TMEM, so the
  // 3.5 us epilogue of every 128 x 256 half tile is no longer hidden behind the next tile's MMAs.  Measured negative result.
  const bool tall = !pair && !res_bn && g_tuning[20] == 3 && a.M >= (int64_t)BLOCK_M * kNumSMs * 2 && a.N > 128 && a.N <= 256;
  if (tall) bn = 256;
  Maps maps;
  memset(&maps, 0, sizeof(maps));
  if (int rc = make_map(&maps.a1, a.A1, a.M, a.K1, a.lda1, BLOCK_K, BLOCK_M)) return rc;
  if (int rc = make_map(&maps.b1, a.B1, a.N, a.K1, a.ldb1, BLOCK_K, pair ? 128 : bn)) return rc;
  if (dual) {
    if (int rc = make_map(&maps.a2, a.A2, a.M, a.K2, a.lda2, BLOCK_K, BLOCK_M)) return rc;
    if (int rc = make_map(&maps.b2, a.B2, a.N, a.K2, a.ldb2, BLOCK_K, pair ? 128 : bn)) return rc;
  }
  TcParams p{};
  p.M = a.M; p.N = a.N; p.K1 = a.K1; p.K2 = dual ? a.K2 : 0; p.splits = 1; p.k_per_split = 0;
  p.ep = EpilogueParams{a.bias, a.addend, a.ldadd, a.gate, a.ldgate, a.gate_scale, a.relu, a.dropout_p, a.seed, a.offset, a.rng_state};
  p.D = a.D; p.ldd = a.ldd; p.partial = nullptr;
  p.dbg = g_tuning[15] ? debug_buffer() : nullptr;
#ifdef LLP_EXPERIMENT
  if (g_tuning[11]) p.splits = -1;  // experiment knob read by the pair kernel only (no MMAs: pure load pipeline)
#endif
  {
    const size_t so = a.out_dtype == LLP_BF16 ? 2 : 4;
    auto ok = [&](const void* ptr, int64_t ld) { return ptr != nullptr && aligned(ptr, 16) && (ld * so) % 16 == 0; };
    auto ok32 = [&](const void* ptr, int64_t ld) { return ptr != nullptr && aligned(ptr, 32) && (ld * so) % 32 == 0; };
    p.ep_flags = (a.bias != nullptr && aligned(a.bias, 16) ? kVecBias : 0) | (ok(a.addend, a.ldadd) ? kVecAddend : 0) |
                 (ok(a.gate, a.ldgate) ? kVecGate : 0) | (ok(a.D, a.ldd) ? kVecOut : 0) |
                 (ok32(a.addend, a.ldadd) ? kVec32Addend : 0) | (ok32(a.gate, a.ldgate) ? kVec32Gate : 0) |
                 (ok32(a.D, a.ldd) ? kVec32Out : 0);
    if (g_tuning[19]) p.ep_flags &= ~(kVec32Addend | kVec32Gate | kVec32Out);  // experiment: 128-bit epilogue accesses
  }
  if (tall) {
    if (a.out_dtype == LLP_BF16) return launch_tall<__nv_bfloat16>(maps, p, stream);
    if (a.out_dtype == LLP_F32) return launch_tall<float>(maps, p, stream);
    return LLP_E_BADARG;
  }
  if (pair) {
    if (a.out_dtype == LLP_BF16) return launch_pair<__nv_bfloat16>(maps, p, num_kb, stream);
    if (a.out_dtype == LLP_F32) return launch_pair<float>(maps, p, num_kb, stream);
    return LLP_E_BADARG;
  }
  if (res_bn == 256) {
    if (a.out_dtype == LLP_BF16) return launch_resb<256, __nv_bfloat16>(maps, p, num_kb, stream);
    if (a.out_dtype == LLP_F32) return launch_resb<256, float>(maps, p, num_kb, stream);
    return LLP_E_BADARG;
  }
  if (res_bn == 128) {
    if (a.out_dtype == LLP_BF16) return launch_resb<128, __nv_bfloat16>(maps, p, num_kb, stream);
    if (a.out_dtype == LLP_F32) return launch_resb<128, float>(maps, p, num_kb, stream);
    return LLP_E_BADARG;
  }
#define LLP_TC_NT(BN)                                                                                          \
  if (bn == BN) {                                                                                              \
    if (a.out_dtype == LLP_BF16) return launch<BN, false, __nv_bfloat16>(maps, p, stream);                     \
    if (a.out_dtype == LLP_F32) return launch<BN, false, float>(maps, p, stream);                              \
    return LLP_E_BADARG;                                                                                       \
  }
  LLP_TC_NT(256)
  LLP_TC_NT(128)
  LLP_TC_NT(64)
#undef LLP_TC_NT
  return LLP_E_SHAPE;
}

// Split count over the M (reduction) rows.  Cost model in units of one k-block step of a CTA: a launch takes
// waves x (k-blocks per split + fill/epilogue) and every split adds one fp32 partial of the whole [N1,N2] result to
// write and re-read.  The previous rule ("just enough splits to cover the SMs") picked 3 splits for the 66 tiles of the
// Coauthor-Physics weight gradient (256 x 8415): 198 work units = two waves, the second one a third full.
void tn_split_plan(int64_t M, int64_t N1, int64_t N2, int* splits, int64_t* k_per_split) {
  using namespace tc;
  int bn = pick_block_n(N2);
  const int64_t tiles = ceil_div(N1, BLOCK_M) * ceil_div(N2, bn);
  const int64_t kblocks = ceil_div(M, BLOCK_K);
  const double partial_cost = (double)N1 * (double)N2 * 2.7e-6;  // 8 bytes per element at ~5 TB/s over a 0.6 us k-block
  const double fixed_cost = 10.0;                                 // pipeline fill + 128 x bn fp32 tile store
  double best = 1e300;
  int64_t best_per = kblocks;
  const int64_t smax = kblocks < 64 ? kblocks : 64;
  for (int64_t s = 1; s <= smax; ++s) {
    const int64_t per = ceil_div(kblocks, s);
    const int64_t actual = ceil_div(kblocks, per);
    const int64_t waves = ceil_div(tiles * actual, (int64_t)kNumSMs);
    const double cost = (double)waves * ((double)per + fixed_cost) + (double)actual * partial_cost;
    if (cost < best) { best = cost; best_per = per; }
  }
  *k_per_split = best_per * BLOCK_K;
  *splits = (int)ceil_div(M, *k_per_split);
}

int gemm_tn_tcgen05(int64_t M, int64_t N1, int64_t N2, const void* A, int64_t lda, const void* B, int64_t ldb, float* D,
                    int64_t ldd, int accumulate, float* ws, cudaStream_t stream) {
  using namespace tc;
  const int bn = pick_block_n(N2);
  Maps maps;
  memset(&maps, 0, sizeof(maps));
  if (int rc = make_map(&maps.a1, A, M, N1, lda, 64, BLOCK_K)) return rc;
  if (int rc = make_map(&maps.b1, B, M, N2, ldb, 64, BLOCK_K)) return rc;
  TcParams p{};
  p.M = N1; p.N = N2; p.K1 = M; p.K2 = 0;
  tn_split_plan(M, N1, N2, &p.splits, &p.k_per_split);
  p.ep = EpilogueParams{};
  p.D = nullptr; p.ldd = 0; p.partial = ws;
  int rc;
  if (bn == 256) rc = launch<256, true, float>(maps, p, stream);
  else if (bn == 128) rc = launch<128, true, float>(maps, p, stream);
  else rc = launch<64, true, float>(maps, p, stream);
  if (rc) return rc;
  return splitk_reduce(ws, p.splits, N1, N2, D, ldd, accumulate, stream);
}

}  // namespace llp
