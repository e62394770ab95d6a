// fp32-parity GEMMs on the tensor cores: 3xTF32 split accumulation with tcgen05.mma.kind::tf32 (sm_100a).
//
// The reference's dense layers are true fp32 (cuBLAS SGEMM with allow_tf32 = False: src/models.py:48,143,146,
// src/sageconv_updated.py:71,76).  A single TF32 product keeps 11 significand bits (~1e-3 relative):
// not enough for the 1e-5 parity bar.  Here every fp32 operand x is split as
//     x_hi = rna_tf32(x)          (round to nearest, 11 bits)
//     x_lo = rna_tf32(x - x_hi)   (the subtraction is exact; |x_lo| <= 2^-11 |x|)
// and  A B  is accumulated as  A_lo B_hi + A_hi B_lo + A_hi B_hi : the dropped A_lo B_lo term and the rounding of the low
// parts are ~2^-22 relative per product.
//
// Accumulator precision.  The tensor core adds every MMA's products into the fp32 TMEM accumulator with TRUNCATION
// (measured on B200, tests/test_gpu_kernels.py::test_tf32x3_is_fp32_grade: with one accumulator chain per tile the error
// grows linearly in K — 4e-6 / 1.3e-5 / 6e-5 of the result scale at K = 512 / 1433 / 8415 against 1e-6 / 1.5e-6 / 4e-6 for
// round-to-nearest FFMA).  So the chain is cut: the MMA warp accumulates ONE 32-wide k-block (8 cross-term MMAs, then 4
// hi x hi MMAs) from zero, alternating between the two TMEM accumulator stages, and the epilogue warps add each finished
// chunk into fp32 REGISTER accumulators with round-to-nearest adds on the CUDA cores (the "promotion" DeepGEMM uses for
// FP8 on Hopper); the TMEM drain of chunk i overlaps the MMAs of chunk i+1.  Measured on one C4 training step against an
// fp64 oracle (tools/fp32_accuracy.py, profiles/r02_fp32_accuracy.txt): embeddings 2.7e-7 (CUDA-core fp32 kernels
// 3.9e-7), every parameter gradient at or below the CUDA-core kernels' error; chains of 2 / 4 k-blocks are 7 % / 10 %
// faster and 2-5x / 5-50x less accurate (llp_set_tuning(23, n)).
//
// Pipeline per CTA (persistent, warp-specialised; 512 threads = 4 warpgroups, registers re-balanced with setmaxnreg):
//   warp 0        TMA producer: raw fp32 tiles of both operands -> shared memory ring (cp.async.bulk.tensor, SW128)
//   warps 12..15  splitters: rewrite the landed tile in place as x_hi and write x_lo to the twin buffer at the same
//                 (swizzled) offsets — the split is elementwise, so it is layout-agnostic — then fence.proxy.async
//   warp 1        TMEM allocator + MMA issuer: three tcgen05.mma.kind::tf32 per 8-wide K step
//   warps 4..11   chunk promotion (TMEM -> 128 register accumulators per thread) + the fused epilogue shared with the
//                 bf16 kernels (tc_epilogue.cuh)
//   NT : D[M,N]   = epi(A1[M,K1] B1[N,K1]^T + A2[M,K2] B2[N,K2]^T)   both operands K-major   (F.linear, lin_l + lin_r)
//   TN : D[N1,N2] = A[M,N1]^T B[M,N2], split over M                    both operands MN-major  (weight gradients)
#include "tc_epilogue.cuh"

namespace llp {

int splitk_reduce(const float* partial, int splits, int64_t rows, int64_t cols, float* D, int64_t ldd, int accumulate,
                  cudaStream_t stream);

namespace tf {
using namespace tc;

constexpr int BLOCK_M = 128;
constexpr int UMMA_K_TF32 = 8;     // 32 bytes of K per tcgen05.mma.kind::tf32
// k-block length kBK (template parameter): 32 fp32 = one 128-byte swizzle row (K-major SWIZZLE_128B), or 16 fp32 = 64-byte
// rows (SWIZZLE_64B): half the bytes per stage, so twice the stages fit (the hi AND lo copy of both operands live in
// shared memory: 96 KB per 32-wide stage at N = 256 leaves room for two stages only, which leaves the TMA round trip and
// the split pass exposed; 16-wide stages are 48 KB: four of them).  MN-major (TN) tiles keep 128-byte rows either way.
constexpr int kEpiWarps = 8;       // two warps per TMEM lane quadrant (warps 4..11)
constexpr int kConvWarps = 6;      // operand splitters (warps 12..15 and the two spare warps 2, 3 of warpgroup 0)
constexpr int kThreads = 512;      // warpgroup 0: TMA, MMA, 2 splitters; 1-2: epilogue; 3: 4 splitters
constexpr int kAccStages = 2;
constexpr int kChainK = 32;        // reduction elements per accumulator chain by default (llp_set_tuning(23, k-blocks) for A/B runs)

template <int BLOCK_N, int kBK>
struct Config {
  static constexpr int kABytes = BLOCK_M * kBK * 4;
  static constexpr int kBBytes = BLOCK_N * kBK * 4;
  static constexpr int kHiBytes = kABytes + kBBytes;      // what TMA lands per stage (and the x_hi operands after the split)
  static constexpr int kStageBytes = 2 * kHiBytes;        // + the x_lo twins
  static constexpr int kStages = (192 * 1024) / kStageBytes > 8 ? 8 : (192 * 1024) / kStageBytes;
  static constexpr int kSlabBytes = kBK * 128;            // MN-major: one TMA box of 32 columns x kBK reduction rows
  static constexpr int kTmemCols = kAccStages * BLOCK_N;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 512 /*barriers*/;
};

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// c = f32 (1 << 4), a = b = tf32 (2 << 7, 2 << 10), a/b major bits 15/16, N >> 3 at 17, M >> 4 at 24
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n, bool mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((mn_major ? 1u : 0u) << 15) | ((mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// Round to the nearest TF32 value (ties away from zero, like cvt.rna.tf32.f32) with two integer ALU operations: add half
// a TF32 ulp to the bit pattern and clear the 13 low bits.  The conversion instruction itself runs on the 16-lane
// conversion pipe; at 24,576 conversions per 32-wide k-block it would take as long as the k-block's MMAs.
__device__ __forceinline__ uint32_t rna_tf32_bits(uint32_t b) { return (b + 0x1000u) & 0xffffe000u; }
// kTruncHi (experiment, llp_set_tuning(25, 1)): leave the landed tile as the high operand — the tensor core then uses its
// upper 19 bits, i.e. x_hi = trunc_tf32(x) — and only write x_lo = x - trunc_tf32(x): 48 KB less shared-memory traffic per
// 32-wide k-block at N = 256 (the kernel is shared-memory-bandwidth bound), |x_lo| up to 2^-10 |x| instead of 2^-11 |x|.
__device__ __forceinline__ uint32_t lo_of_trunc(float x) {
  const float l = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  return (l == l) ? __float_as_uint(l) : 0u;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = rna_tf32_bits(__float_as_uint(x));
  // x - hi is exact (at most 13 significant bits) and is stored as it is: the tensor core reads the upper 19 bits of an
  // operand word, i.e. truncates the low part to TF32 — an error of 2^-22 of x, the size of the dropped lo x lo term.
  // A non-finite x keeps its value in the high part only (inf - inf would poison the product with NaN).
  const float l = x - __uint_as_float(hi);
  lo = (l == l) ? __float_as_uint(l) : 0u;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <uint32_t kRegs>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <uint32_t kRegs>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }
// MN-major fp32 operands: the only shared-memory layout tcgen05.mma.kind::tf32 accepts is "128B swizzle with a 32-byte
// atom" (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B on the TMA side, layout type 1 in the descriptor; CUTLASS
// sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only available smem layout"): 128-byte rows, the
// swizzle pattern repeats every 4 rows, so SBO = 512 B.
__host__ __device__ constexpr uint32_t desc_hi_sw128_base32(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14) | (1u << 29); }
// K-major operands with 64-byte rows: SWIZZLE_64B (layout type 4), 8 rows = 512 B per group
__host__ __device__ constexpr uint32_t desc_hi_sw64(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14) | (4u << 29); }

// ------------------------------------------------------------------------------------------------
template <int BLOCK_N, bool kTN, int kBK>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32x3_kernel(const __grid_constant__ Maps maps, const TcParams p) {
  using Cfg = Config<BLOCK_N, kBK>;
  constexpr int BLOCK_K = kBK;
  constexpr int kSlabBytes = Cfg::kSlabBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // stage s: [A_hi | B_hi | A_lo | B_lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                          // [kStages] raw tile landed            (TMA -> splitters)
  uint64_t* conv_bar = bars + Cfg::kStages;           // [kStages] hi / lo operands ready     (splitters -> MMA)
  uint64_t* empty_bar = bars + 2 * Cfg::kStages;      // [kStages] MMAs retired, slot free    (MMA -> TMA)
  uint64_t* tmem_full = bars + 3 * Cfg::kStages;      // [kAccStages]
  uint64_t* tmem_empty = tmem_full + kAccStages;      // [kAccStages]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + kAccStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int64_t m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M, n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int64_t num_tiles = m_tiles * n_tiles * p.splits;
  const int kb1 = kTN ? 0 : (int)((p.K1 + BLOCK_K - 1) / BLOCK_K);
  const int kb2 = kTN ? 0 : (int)((p.K2 + BLOCK_K - 1) / BLOCK_K);
  const int kChunkKB = p.chunk_kb;
  auto tile_kblocks = [&](int64_t tile) -> int {   // k-blocks of one output tile (TN: of its split of the M rows)
    if constexpr (kTN) {
      const int64_t split = tile / (m_tiles * n_tiles);
      const int64_t k_begin = split * p.k_per_split;
      const int64_t k_end = k_begin + p.k_per_split < p.K1 ? k_begin + p.k_per_split : p.K1;
      return (int)((k_end - k_begin + BLOCK_K - 1) / BLOCK_K);
    } else {
      return kb1 + kb2;
    }
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a1);
    tma_prefetch_desc(&maps.b1);
    if (kb2 > 0) { tma_prefetch_desc(&maps.a2); tma_prefetch_desc(&maps.b2); }
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&conv_bar[s]), kConvWarps);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < kAccStages; ++s) { mbar_init(smem_u32(&tmem_full[s]), 1); mbar_init(smem_u32(&tmem_empty[s]), kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_holder), Cfg::kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  // Operand split, shared by the six splitter warps (ct = 0 .. 191): the measured limiter of the first version was this
  // pass on four warps (~2,500 cycles per 32-wide k-block against 1,536 cycles of tensor time; ncu: tensor pipe 42 %).
  auto split_loop = [&](int ct) {
    constexpr int kVecs = Cfg::kHiBytes / 16;
    int stage = 0; uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int num_kb = tile_kblocks(tile);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        uint4* hi = reinterpret_cast<uint4*>(smem + stage * Cfg::kStageBytes);
        uint4* lo = reinterpret_cast<uint4*>(smem + stage * Cfg::kStageBytes + Cfg::kHiBytes);
        if (p.trunc_hi) {
#pragma unroll 4
          for (int i = ct; i < kVecs; i += 32 * kConvWarps) {
            const uint4 v = hi[i];
            lo[i] = make_uint4(lo_of_trunc(__uint_as_float(v.x)), lo_of_trunc(__uint_as_float(v.y)),
                               lo_of_trunc(__uint_as_float(v.z)), lo_of_trunc(__uint_as_float(v.w)));
          }
        } else {
#pragma unroll 4
          for (int i = ct; i < kVecs; i += 32 * kConvWarps) {
            const uint4 v = hi[i];
            uint4 h, l;
            split_tf32(__uint_as_float(v.x), h.x, l.x);
            split_tf32(__uint_as_float(v.y), h.y, l.y);
            split_tf32(__uint_as_float(v.z), h.z, l.z);
            split_tf32(__uint_as_float(v.w), h.w, l.w);
            hi[i] = h;
            lo[i] = l;
          }
        }
        fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&conv_bar[stage]));
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  };

  const int wg = warp >> 2;
  if (wg == 0) {
    setmaxnreg_dec<56>();
    if (warp >= 2) {
      split_loop(128 + threadIdx.x - 64);   // warps 2, 3: splitter threads 128 .. 191
    } else if (warp == 0) {
      // ===================== TMA producer =====================
      int stage = 0; uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t split = tile / (m_tiles * n_tiles);
        const int64_t mn = tile % (m_tiles * n_tiles);
        const int m0 = (int)(mn / n_tiles) * BLOCK_M, n0 = (int)(mn % n_tiles) * BLOCK_N;
        const int num_kb = tile_kblocks(tile);
        const int64_t k_begin = kTN ? split * p.k_per_split : 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          if (elect_one_sync()) {
            const uint32_t bar = smem_u32(&full_bar[stage]);
            mbar_expect_tx(bar, Cfg::kHiBytes);
            const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes), sb = sa + Cfg::kABytes;
            if constexpr (kTN) {
              const int k0 = (int)(k_begin + (int64_t)kb * BLOCK_K);
#pragma unroll
              for (int j = 0; j < BLOCK_M / 32; ++j) tma_load_2d(sa + j * kSlabBytes, &maps.a1, m0 + j * 32, k0, bar);
#pragma unroll
              for (int j = 0; j < BLOCK_N / 32; ++j) tma_load_2d(sb + j * kSlabBytes, &maps.b1, n0 + j * 32, k0, bar);
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
      constexpr uint32_t idesc = make_idesc_tf32(BLOCK_M, BLOCK_N, kTN);
      // K-major SW128: LBO 16 B, SBO = 8 rows (1024 B), one UMMA_K step = 32 B inside the swizzle row.
      // MN-major SW128 with 32-byte atoms: LBO = distance between the 32-column slabs, SBO = 4 reduction rows (512 B),
      // one UMMA_K step = 8 reduction rows = 1024 B.
      constexpr uint32_t kHi = kTN ? desc_hi_sw128_base32(512) : (kBK == 32 ? desc_hi_sw128(1024) : desc_hi_sw64(512));
      constexpr uint32_t kStep = (kTN ? UMMA_K_TF32 * 128 : UMMA_K_TF32 * 4) >> 4;
      constexpr uint32_t kLoOff = (uint32_t)Cfg::kHiBytes >> 4;   // x_lo twin of an operand
      const uint32_t a_lo0 = desc_lo(smem_u32(smem), kTN ? kSlabBytes : 16);
      const uint32_t b_lo0 = desc_lo(smem_u32(smem) + Cfg::kABytes, kTN ? kSlabBytes : 16);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int num_kb = tile_kblocks(tile);
        for (int kb0 = 0; kb0 < num_kb; kb0 += kChunkKB) {   // one accumulator chain per chunk
          const int kb_end = kb0 + kChunkKB < num_kb ? kb0 + kChunkKB : num_kb;
          mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
          for (int kb = kb0; kb < kb_end; ++kb) {
            mbar_wait(smem_u32(&conv_bar[stage]), phase);
            tcgen05_fence_after();
            if (elect_one_sync()) {
              const uint32_t a_hi = a_lo0 + (uint32_t)stage * ((uint32_t)Cfg::kStageBytes >> 4);
              const uint32_t b_hi = b_lo0 + (uint32_t)stage * ((uint32_t)Cfg::kStageBytes >> 4);
              // The eight cross-term MMAs of the k-block first, then its four hi x hi MMAs: every MMA's sum is TRUNCATED into
              // the accumulator, and the error of a truncation scales with the accumulator's magnitude — cross terms are
              // 2^-11 of the main terms, so while only they have been added (the first k-block of a chain) truncation
              // costs nothing, and the main terms hit the accumulator in as few steps as possible.
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K_TF32; ++k) {
                const uint64_t ah = desc_from(a_hi + k * kStep, kHi), bh = desc_from(b_hi + k * kStep, kHi);
                const uint64_t al = desc_from(a_hi + kLoOff + k * kStep, kHi), bl = desc_from(b_hi + kLoOff + k * kStep, kHi);
                umma_tf32(tmem_d, al, bh, idesc, (uint32_t)((kb > kb0) | (k != 0)));
                umma_tf32(tmem_d, ah, bl, idesc, 1u);
              }
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K_TF32; ++k)
                umma_tf32(tmem_d, desc_from(a_hi + k * kStep, kHi), desc_from(b_hi + k * kStep, kHi), idesc, 1u);
              umma_commit(smem_u32(&empty_bar[stage]));  // frees the smem slot once these MMAs retire
            }
            __syncwarp();
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
          }
          if (elect_one_sync()) umma_commit(smem_u32(&tmem_full[acc]));      // chunk ready for promotion
          __syncwarp();
          if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (wg == 3) {
    // ===================== splitters: x -> (x_hi in place, x_lo twin) =====================
    setmaxnreg_dec<56>();
    split_loop(threadIdx.x - 32 * 12);
  } else {
    // ===================== chunk promotion + epilogue: TMEM -> register accumulators -> global =====================
    setmaxnreg_inc<200>();
    uint64_t seed = p.ep.seed, offset = p.ep.offset;   // the dropout stream is resolved once (device rng_state); the rest
    if (p.ep.dropout_p > 0.0f) {                       // of the parameter block stays in constant memory: the register
      EpilogueParams e = p.ep;                         // budget belongs to the 128 accumulators
      resolve_rng(e);
      seed = e.seed; offset = e.offset;
    }
    const int quad = warp & 3;            // TMEM lanes [32*quad, 32*quad+32) are the only ones this warp may read
    const int half = (warp - 4) >> 2;     // which half of the tile's columns
    constexpr int kColsPerWarp = BLOCK_N / 2;
    constexpr int kSlices = kColsPerWarp / 32;
    int acc = 0; uint32_t acc_phase = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int64_t split = tile / (m_tiles * n_tiles);
      const int64_t mn = tile % (m_tiles * n_tiles);
      const int64_t m0 = (mn / n_tiles) * BLOCK_M, n0 = (mn % n_tiles) * BLOCK_N;
      const int64_t m = m0 + quad * 32 + lane;
      const int num_kb = tile_kblocks(tile);
      float racc[kSlices][32];
      for (int kb0 = 0; kb0 < num_kb; kb0 += kChunkKB) {
        mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N) + (uint32_t)(half * kColsPerWarp);
#pragma unroll
        for (int sl = 0; sl < kSlices; ++sl) {
          if (n0 + half * kColsPerWarp + sl * 32 < p.N) {   // warp-uniform
            uint32_t r[32];
            tmem_ld32(taddr + sl * 32, r);
            if (kb0 == 0) {
#pragma unroll
              for (int j = 0; j < 32; ++j) racc[sl][j] = __uint_as_float(r[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) racc[sl][j] += __uint_as_float(r[j]);   // round-to-nearest promotion
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));   // the stage is free before the global epilogue starts
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
      uint4 rnd128 = make_uint4(0, 0, 0, 0);
      uint32_t rnd_group = 0xffffffffu;
#pragma unroll
      for (int sl = 0; sl < kSlices; ++sl) {
        const int c0 = half * kColsPerWarp + sl * 32;
        if (n0 + c0 < p.N && m < p.M) {
          uint32_t r[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(racc[sl][j]);
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
            if (p.ep.dropout_p == 0.5f && (uint32_t)((n0 + c0) >> 7) != rnd_group) {
              rnd_group = (uint32_t)((n0 + c0) >> 7);
              rnd128 = philox4x32_10(seed, (uint64_t)m, offset + (uint64_t)rnd_group);
            }
            epilogue_chunk<float>(r, m, n0 + c0, p, rnd128, seed, offset);
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// 2-D fp32 row-major [rows, cols] with leading dimension ld; box = {box_cols (inner, 32 floats = 128 B), box_rows}
static int make_map_f32(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
                        CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return LLP_E_DEVICE;
  if (!aligned(base, 16) || (ld * 4) % 16 != 0) return LLP_E_ALIGN;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : LLP_E_BADARG;
}

template <int BLOCK_N, bool kTN, int kBK>
static int launch(const Maps& maps, const TcParams& p, cudaStream_t stream) {
  using Cfg = Config<BLOCK_N, kBK>;
  auto kern = gemm_tf32x3_kernel<BLOCK_N, kTN, kBK>;
  static PerDeviceOnce configured;  // cudaFuncSetAttribute is per device
  if (configured.need()) {
    LLP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured.done();
  }
  const int64_t tiles = ceil_div(p.M, BLOCK_M) * ceil_div(p.N, BLOCK_N) * p.splits;
  const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
  kern<<<grid, kThreads, Cfg::kSmemBytes, stream>>>(maps, p);
  LLP_LAUNCH_OK();
  return 0;
}

static int pick_block_n(int64_t N) { return N > 128 ? 256 : (N > 64 ? 128 : 64); }
// k-block length: 32 (two 96 KB stages at N = 256; SWIZZLE_128B) unless llp_set_tuning(24, 16) asks for the 16-wide ring
// (four 48 KB stages; SWIZZLE_64B).  Measured equal (436 vs 440 us at M = 235,868, K = 512): with the hi AND lo copies of
// both operands in shared memory the kernel is bound by shared-memory bandwidth — per 32-wide k-block the 12 MMAs read
// 144 KB of operands, the split pass moves another 144 KB and TMA writes 48 KB, i.e. ~2,600 cycles at 128 B/clk against
// 1,536 cycles of tensor time — not by the depth of the ring.
static int pick_block_k() { return g_tuning[24] == 16 ? 16 : 32; }
static int chain_kblocks(int bk) { return g_tuning[23] > 0 ? g_tuning[23] : kChainK / bk; }

template <bool kTN>
static int dispatch(int bn, int bk, const Maps& maps, const TcParams& p, cudaStream_t stream) {
  if (bk == 32) {
    if (bn == 256) return launch<256, kTN, 32>(maps, p, stream);
    if (bn == 128) return launch<128, kTN, 32>(maps, p, stream);
    return launch<64, kTN, 32>(maps, p, stream);
  }
  if (bn == 256) return launch<256, kTN, 16>(maps, p, stream);
  if (bn == 128) return launch<128, kTN, 16>(maps, p, stream);
  return launch<64, kTN, 16>(maps, p, stream);
}

// Split count over the reduction rows of the TN product (same cost model as the bf16 kernel's tn_split_plan: waves x
// (k-blocks per split + fill) + one fp32 partial of the result per split), with this kernel's 32-row k-blocks.
static void tn_plan(int64_t M, int64_t N1, int64_t N2, int* splits, int64_t* k_per_split) {
  const int bn = pick_block_n(N2);
  const int64_t tiles = ceil_div(N1, BLOCK_M) * ceil_div(N2, bn);
  constexpr int BLOCK_K = 32;   // planning granularity (a multiple of both k-block lengths)
  const int64_t kblocks = ceil_div(M, BLOCK_K);
  const double partial_cost = (double)N1 * (double)N2 * 2.0e-6;
  const double fixed_cost = 8.0;
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

}  // namespace tf

bool tf32_operand_ok(const void* p, int64_t ld) { return p != nullptr && aligned(p, 16) && (ld * 4) % 16 == 0; }

int gemm_tn_tf32_splits(int64_t M, int64_t N1, int64_t N2) {
  int s = 1;
  int64_t per = 0;
  tf::tn_plan(M, N1, N2, &s, &per);
  return s;
}

int gemm_nt_tf32(const llp_gemm_nt_args& a, cudaStream_t stream) {
  using namespace tf;
  if (a.dtype != LLP_F32 || a.out_dtype != LLP_F32) return LLP_E_SHAPE;
  const bool dual = a.A2 != nullptr && a.K2 > 0;
  const int bn = pick_block_n(a.N), bk = pick_block_k();
  const CUtensorMapSwizzle sw = bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  Maps maps;
  memset(&maps, 0, sizeof(maps));
  if (int rc = make_map_f32(&maps.a1, a.A1, a.M, a.K1, a.lda1, bk, BLOCK_M, sw)) return rc;
  if (int rc = make_map_f32(&maps.b1, a.B1, a.N, a.K1, a.ldb1, bk, bn, sw)) return rc;
  if (dual) {
    if (int rc = make_map_f32(&maps.a2, a.A2, a.M, a.K2, a.lda2, bk, BLOCK_M, sw)) return rc;
    if (int rc = make_map_f32(&maps.b2, a.B2, a.N, a.K2, a.ldb2, bk, bn, sw)) return rc;
  }
  TcParams p{};
  p.M = a.M; p.N = a.N; p.K1 = a.K1; p.K2 = dual ? a.K2 : 0; p.splits = 1; p.k_per_split = 0;
  p.ep = EpilogueParams{a.bias, a.addend, a.ldadd, a.gate, a.ldgate, a.gate_scale, a.relu, a.dropout_p, a.seed, a.offset, a.rng_state};
  p.D = a.D; p.ldd = a.ldd; p.partial = nullptr; p.dbg = nullptr;
  p.chunk_kb = chain_kblocks(bk);
  p.trunc_hi = g_tuning[25];
  auto ok = [&](const void* ptr, int64_t ld) { return ptr != nullptr && aligned(ptr, 16) && (ld * 4) % 16 == 0; };
  auto ok32 = [&](const void* ptr, int64_t ld) { return ptr != nullptr && aligned(ptr, 32) && (ld * 4) % 32 == 0; };
  p.ep_flags = (a.bias != nullptr && aligned(a.bias, 16) ? kVecBias : 0) | (ok(a.addend, a.ldadd) ? kVecAddend : 0) |
               (ok(a.gate, a.ldgate) ? kVecGate : 0) | (ok(a.D, a.ldd) ? kVecOut : 0) |
               (ok32(a.addend, a.ldadd) ? kVec32Addend : 0) | (ok32(a.gate, a.ldgate) ? kVec32Gate : 0) |
               (ok32(a.D, a.ldd) ? kVec32Out : 0);
  return dispatch<false>(bn, bk, maps, p, stream);
}

int gemm_tn_tf32(int64_t M, int64_t N1, int64_t N2, const void* A, int64_t lda, const void* B, int64_t ldb, float* D,
                 int64_t ldd, int accumulate, float* ws, cudaStream_t stream) {
  using namespace tf;
  const int bn = pick_block_n(N2), bk = pick_block_k();
  Maps maps;
  memset(&maps, 0, sizeof(maps));
  if (int rc = make_map_f32(&maps.a1, A, M, N1, lda, 32, bk, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
  if (int rc = make_map_f32(&maps.b1, B, M, N2, ldb, 32, bk, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
  TcParams p{};
  p.M = N1; p.N = N2; p.K1 = M; p.K2 = 0;
  tn_plan(M, N1, N2, &p.splits, &p.k_per_split);
  p.ep = EpilogueParams{};
  p.D = nullptr; p.ldd = 0; p.partial = ws; p.dbg = nullptr;
  p.chunk_kb = chain_kblocks(bk);
  p.trunc_hi = g_tuning[25];
  if (int rc = dispatch<true>(bn, bk, maps, p, stream)) return rc;
  return splitk_reduce(ws, p.splits, N1, N2, D, ldd, accumulate, stream);
}

}  // namespace llp
