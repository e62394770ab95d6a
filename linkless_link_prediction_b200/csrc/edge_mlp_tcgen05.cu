// Fused edge scorer (sm_100a, tcgen05): for a batch of edges (u[m], v[m])
//
//     z[m,:]  = h[u[m],:] * h[v[m],:]                       (the two fancy-index gathers + mul, models.py:140)
//     y[m,:]  = dropout(relu(z[m,:] W1^T + b1))             (first predictor layer, models.py:143-145)
//     prob[m] = sigmoid(y[m,:] . w2 + b2)                   (1-output last layer + sigmoid, models.py:146,150)
//
// in ONE kernel: the gather-Hadamard is the A-operand PRODUCER of the first predictor GEMM.  Eight
// producer warps gather the two embedding rows of every edge of a 128-edge tile with 128-bit loads (whole rows per
// load instruction), multiply them and write the bf16 products straight into the 128B-swizzled K-major shared-memory
// stages the tensor core reads (and, for training, also to `z`, which the weight gradient needs); W1 (<= 128 KB) is loaded into shared memory once per CTA by
// TMA; one elected lane issues tcgen05.mma into a double-buffered TMEM accumulator; eight epilogue warps apply
// bias + relu + Philox dropout, store `y` (training) and reduce the row against w2 for the sigmoid score.  With z and y
// switched off (evaluation) a scored edge costs two row gathers and four bytes of output.
#include "tcgen05.cuh"

namespace llp {
namespace em {

using namespace tc;

constexpr int kTileM = 128;
constexpr int kBK = 64;                 // bf16 elements per K block = one 128-byte swizzle row
constexpr int kGatherWarps = 8;         // 16 rows of every tile each
constexpr int kEpiWarps = 8;
constexpr int kCtrlWarps = 4;            // warpgroup 0: warp 0 = W1 loader, TMEM allocator and MMA issuer; warps 1..3 only give
                                         // their registers away (setmaxnreg works on whole warpgroups)
constexpr int kThreads = 32 * (kCtrlWarps + kGatherWarps + kEpiWarps);
constexpr int kAccStages = 2;
constexpr int kABytes = kTileM * 128;   // one A stage: 128 rows x 128 B
constexpr int kSmemMax = 226 * 1024;

struct Params {
  const __nv_bfloat16* h; int64_t ldh;
  const int64_t* u; const int64_t* v;
  int64_t M; int K, N;
  const float* bias1;
  int relu; float dropout_p; uint64_t seed, offset; const uint64_t* rng_state;
  __nv_bfloat16* z; int64_t ldz;
  __nv_bfloat16* y; int64_t ldy;
  const float* w2; const float* b2; float* prob;
  int stages;
  long long* dbg;   // instrumentation (llp_set_tuning(15, 1)): per CTA {issue loop ns, operand wait ns, accumulator wait ns}
};

__device__ __forceinline__ uint32_t mul_bf16x2(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1) edge_mlp_kernel(const __grid_constant__ CUtensorMap map_w, const Params p) {
  constexpr int kBBytes = BLOCK_N * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int num_kb = p.K / kBK;
  uint8_t* smem_b = smem;                                   // [num_kb][BLOCK_N rows x 128 B], resident
  uint8_t* smem_a = smem + (size_t)num_kb * kBBytes;        // [stages][128 rows x 128 B]
  float* head_part = reinterpret_cast<float*>(smem_a + (size_t)p.stages * kABytes);  // [kAccStages][2][kTileM]
  float* s_bias = head_part + kAccStages * 2 * kTileM;     // [256] bias1 (zeros when absent)
  float* s_w2 = s_bias + 256;                              // [256] w2 (zeros when absent)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_w2 + 256);
  uint64_t* full_bar = bars;                       // [stages], one arrival per gather warp
  uint64_t* empty_bar = bars + p.stages;           // [stages]
  uint64_t* tmem_full = bars + 2 * p.stages;       // [kAccStages]
  uint64_t* tmem_empty = tmem_full + kAccStages;   // [kAccStages]
  uint64_t* w_full = tmem_empty + kAccStages;      // [1]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(w_full + 1);
  constexpr int kTmemCols = kAccStages * BLOCK_N;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m_tiles = (p.M + kTileM - 1) / kTileM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), kGatherWarps); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int s = 0; s < kAccStages; ++s) { mbar_init(smem_u32(&tmem_full[s]), 1); mbar_init(smem_u32(&tmem_empty[s]), kEpiWarps); }
    mbar_init(smem_u32(w_full), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_holder), kTmemCols);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {   // epilogue vectors: shared-memory broadcasts, not global loads
    s_bias[i] = (p.bias1 != nullptr && i < p.N) ? __ldg(p.bias1 + i) : 0.0f;
    s_w2[i] = (p.w2 != nullptr && i < p.N) ? __ldg(p.w2 + i) : 0.0f;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  // Register budget (640 threads x 96): warpgroup 0 keeps 24 per thread, the two producer warpgroups take 128 (two load
  // groups of 16 rows x 2 operands in flight per lane = 64 registers of 128-bit loads), the epilogue warpgroups stay at 96.
  if (warp < kCtrlWarps) asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
  else if (warp < kCtrlWarps + kGatherWarps) asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
  if (warp == 0) {
    // ===================== W1: all K blocks, once; then the MMA issue loop =====================
    if ((int64_t)blockIdx.x < m_tiles && elect_one_sync()) {
      mbar_expect_tx(smem_u32(w_full), (uint32_t)(num_kb * kBBytes));
      for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(smem_u32(smem_b + (size_t)kb * kBBytes), &map_w, kb * kBK, 0, smem_u32(w_full));
    }
    __syncwarp();
    constexpr uint32_t idesc = make_idesc(kTileM, BLOCK_N, false);
    constexpr uint32_t kHi = desc_hi_sw128(1024);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem_a), 16), b_lo0 = desc_lo(smem_u32(smem_b), 16);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    long long t_begin = 0, w_full_ns = 0, w_acc_ns = 0, t0 = 0;
    auto now = []() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    if (p.dbg) t_begin = now();
    if ((int64_t)blockIdx.x < m_tiles) {
      mbar_wait(smem_u32(w_full), 0);
      tcgen05_fence_after();
    }
    for (int64_t mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
      if (p.dbg) t0 = now();
      mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
      if (p.dbg) w_acc_ns += now() - t0;
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int kb = 0; kb < num_kb; ++kb) {
        if (p.dbg) t0 = now();
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        if (p.dbg) w_full_ns += now() - t0;
        tcgen05_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * (kABytes >> 4), b_lo = b_lo0 + (uint32_t)kb * (kBBytes >> 4);
#pragma unroll
          for (int k = 0; k < kBK / UMMA_K; ++k)
            umma_bf16(tmem_d, desc_from(a_lo + k * 2, kHi), desc_from(b_lo + k * 2, kHi), idesc, (uint32_t)((kb | k) != 0));
          umma_commit(smem_u32(&empty_bar[stage]));
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one_sync()) umma_commit(smem_u32(&tmem_full[acc]));
      __syncwarp();
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
    if (p.dbg && lane == 0) {
      p.dbg[blockIdx.x * 4 + 0] = now() - t_begin;
      p.dbg[blockIdx.x * 4 + 1] = w_full_ns;
      p.dbg[blockIdx.x * 4 + 2] = w_acc_ns;
    }
  } else if (warp < kCtrlWarps) {
    // idle warps of warpgroup 0
  } else if (warp < kCtrlWarps + kGatherWarps) {
    // ===================== gather producers: z = h[u] * h[v] straight into the swizzled A stages =====================
    // The 8 warps are split over the K blocks of the tile (8 / num_kb warps per K block, each owning a slice of the
    // 128 rows), so every pipeline stage has its own producers: the stage of K block k of the NEXT tile is refilled as
    // soon as the tensor core has consumed it, while later K blocks of the current tile are still being multiplied.
    // 8 lanes cover one 128-byte K slice of a row (a full L2 line), 4 rows per instruction, 16 row loads in flight
    // per lane batch (8 x (h[u], h[v])).
    // Every producer warp owns 16 rows of each tile and fetches WHOLE embedding rows: K / 8 lanes cover one row (K = 256:
    // one 512-byte row per load instruction, like the SpMM's gathers; the first version gave each warp a 128-byte
    // K slice of 64 rows, i.e. four separate requests per row at four different times), so a lane's 8 columns belong to
    // K block (lane % (K/8)) / 8 and its product goes to that K block's stage.  All 8 warps arrive on every stage.
    // The producer is latency-bound (ncu: its warps wait on the row loads and on the endpoint indices in front of them):
    // a warp keeps TWO load groups (4 instructions x 2 operands each) in flight and fetches the endpoints of its rows of
    // the NEXT tile while it works on the current one.
    const int gw = warp - kCtrlWarps;
    const int lpr = num_kb * 8;                          // lanes per row
    const int rpi = 32 / lpr;                            // rows per load instruction: 1, 2 or 4 (num_kb in {4, 2, 1})
    const int sub = lane / lpr, jj = lane % lpr;
    const int kb_l = jj >> 3, j = jj & 7;                // K block / 16-byte chunk inside its 128-byte stage row
    constexpr int kRowsW = kTileM / kGatherWarps;        // 16 rows per warp and tile
    constexpr int kIt = 4;                               // load instructions per operand and group
    constexpr int kMaxGroups = 4;
    const int n_groups = num_kb;                         // kRowsW = n_groups * kIt * rpi
    const __nv_bfloat16* hk = p.h + jj * 8;
    // endpoints of my 16 rows: lanes 0..15 hold u of row (lane), lanes 16..31 hold v of row (lane - 16)
    auto load_idx = [&](int64_t mt) -> int {
      const int64_t m_l = mt * kTileM + gw * kRowsW + (lane & 15);
      return m_l < p.M ? (int)__ldg((lane < 16 ? p.u : p.v) + m_l) : 0;
    };
    int ci = (int64_t)blockIdx.x < m_tiles ? load_idx(blockIdx.x) : 0, ni = 0;
    int64_t slot0 = 0;                                   // ring position of K block 0 of the current tile
    for (int64_t mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, slot0 += num_kb) {
      const int64_t my_slot = slot0 + kb_l;
      uint8_t* sa = smem_a + (size_t)(my_slot % p.stages) * kABytes;
      uint4 a[2][kIt], b[2][kIt];
      auto issue = [&](int g, uint4 (&A)[kIt], uint4 (&B)[kIt]) {
#pragma unroll
        for (int it = 0; it < kIt; ++it) {
          const int rl = (g * kIt + it) * rpi + sub;     // row inside my 16
          const int ur = __shfl_sync(0xffffffffu, ci, rl), vr = __shfl_sync(0xffffffffu, ci, 16 + rl);
          A[it] = ldg_v4(hk + (int64_t)ur * p.ldh);
          B[it] = ldg_v4(hk + (int64_t)vr * p.ldh);
        }
      };
      auto consume = [&](int g, const uint4 (&A)[kIt], const uint4 (&B)[kIt]) {
#pragma unroll
        for (int it = 0; it < kIt; ++it) {
          const int row = gw * kRowsW + (g * kIt + it) * rpi + sub;   // row inside the tile
          const int64_t m = mt * kTileM + row;
          uint4 z;
          z.x = mul_bf16x2(A[it].x, B[it].x); z.y = mul_bf16x2(A[it].y, B[it].y);
          z.z = mul_bf16x2(A[it].z, B[it].z); z.w = mul_bf16x2(A[it].w, B[it].w);
          if (m >= p.M) z = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(sa + row * 128 + ((j ^ (row & 7)) << 4)) = z;   // SWIZZLE_128B, K-major
          if (p.z != nullptr && m < p.M) stg_v4(p.z + m * p.ldz + jj * 8, z);
        }
      };
      issue(0, a[0], b[0]);
      if (mt + gridDim.x < m_tiles) ni = load_idx(mt + gridDim.x);
#pragma unroll
      for (int g = 0; g < kMaxGroups; ++g) {
        if (g + 1 < kMaxGroups && g + 1 < n_groups) issue(g + 1, a[(g + 1) & 1], b[(g + 1) & 1]);
        if (g == 0) {   // first store below: every stage of this tile must have been consumed
          for (int kb = 0; kb < num_kb; ++kb) {
            const int64_t sl = slot0 + kb;
            mbar_wait(smem_u32(&empty_bar[sl % p.stages]), (uint32_t)(((sl / p.stages) & 1) ^ 1));
          }
        }
        if (g < n_groups) consume(g, a[g & 1], b[g & 1]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane < num_kb) mbar_arrive(smem_u32(&full_bar[(slot0 + lane) % p.stages]));
      ci = ni;
    }
  } else {
    // ===================== epilogue: bias + relu + dropout -> y, row . w2 -> sigmoid =====================
    const int ew = warp - (kCtrlWarps + kGatherWarps);
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
    const int half = ew >> 2;                  // which half of the columns
    constexpr int kColsPerWarp = BLOCK_N / 2;
    EpilogueParams ep{};
    ep.dropout_p = p.dropout_p; ep.seed = p.seed; ep.offset = p.offset; ep.rng_state = p.rng_state;
    if (ep.dropout_p > 0.0f) resolve_rng(ep);
    const bool y32 = p.y != nullptr && (reinterpret_cast<uintptr_t>(p.y) % 32) == 0 && (p.ldy * 2) % 32 == 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int64_t mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
      tcgen05_fence_after();
      const int row = quad * 32 + lane;
      const int64_t m = mt * kTileM + row;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
      uint4 rnd128 = make_uint4(0, 0, 0, 0);
      uint32_t rnd_group = 0xffffffffu;
      float dot = 0.0f;
#pragma unroll 1
      for (int c0 = half * kColsPerWarp; c0 < (half + 1) * kColsPerWarp; c0 += 32) {
        if (c0 >= p.N) break;
        if (ep.dropout_p == 0.5f && (uint32_t)(c0 >> 7) != rnd_group) {
          rnd_group = (uint32_t)(c0 >> 7);
          rnd128 = philox4x32_10(ep.seed, (uint64_t)m, ep.offset + (uint64_t)rnd_group);
        }
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        const int valid = p.N - c0 < 32 ? p.N - c0 : 32;
        const uint32_t bits = ep.dropout_p == 0.5f ? dropout_word(rnd128, (c0 >> 5) & 3) : 0u;
        const uint32_t thr = dropout_thr16(ep.dropout_p);
        const float scale = 1.0f / (1.0f - ep.dropout_p);
        uint32_t packed[16];
        // 8 columns at a time: bias (smem broadcast), relu, dropout, round to bf16, head dot on the rounded values
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c0 + g * 8), b1 = *reinterpret_cast<const float4*>(s_bias + c0 + g * 8 + 4);
          float f[8] = {__uint_as_float(r[g * 8 + 0]) + b0.x, __uint_as_float(r[g * 8 + 1]) + b0.y, __uint_as_float(r[g * 8 + 2]) + b0.z,
                        __uint_as_float(r[g * 8 + 3]) + b0.w, __uint_as_float(r[g * 8 + 4]) + b1.x, __uint_as_float(r[g * 8 + 5]) + b1.y,
                        __uint_as_float(r[g * 8 + 6]) + b1.z, __uint_as_float(r[g * 8 + 7]) + b1.w};
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.0f);
          }
          if (ep.dropout_p == 0.5f) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = ((bits >> (g * 8 + i)) & 1u) ? f[i] * 2.0f : 0.0f;
          } else if (ep.dropout_p > 0.0f) {
            const uint4 rnd = philox4x32_10(ep.seed, (uint64_t)m, ep.offset + (uint64_t)((c0 >> 3) + g));
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = dropout_u16(rnd, i) >= thr ? f[i] * scale : 0.0f;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) packed[g * 4 + i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
          if (p.prob != nullptr) {   // columns >= N carry zero weights (and zero accumulators: W1 rows are zero-filled)
            const float4 w0 = *reinterpret_cast<const float4*>(s_w2 + c0 + g * 8), w1 = *reinterpret_cast<const float4*>(s_w2 + c0 + g * 8 + 4);
            dot = fmaf(__uint_as_float(packed[g * 4 + 0] << 16), w0.x, dot);
            dot = fmaf(__uint_as_float(packed[g * 4 + 0] & 0xffff0000u), w0.y, dot);
            dot = fmaf(__uint_as_float(packed[g * 4 + 1] << 16), w0.z, dot);
            dot = fmaf(__uint_as_float(packed[g * 4 + 1] & 0xffff0000u), w0.w, dot);
            dot = fmaf(__uint_as_float(packed[g * 4 + 2] << 16), w1.x, dot);
            dot = fmaf(__uint_as_float(packed[g * 4 + 2] & 0xffff0000u), w1.y, dot);
            dot = fmaf(__uint_as_float(packed[g * 4 + 3] << 16), w1.z, dot);
            dot = fmaf(__uint_as_float(packed[g * 4 + 3] & 0xffff0000u), w1.w, dot);
          }
        }
        if (p.y != nullptr && m < p.M) {
          __nv_bfloat16* dst = p.y + m * p.ldy + c0;
          if (y32 && valid == 32) {
            U32x8 lo, hi;
#pragma unroll
            for (int q = 0; q < 8; ++q) { lo.v[q] = packed[q]; hi.v[q] = packed[8 + q]; }
            stg_v8(dst, lo);
            stg_v8(dst + 16, hi);
          } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              if (2 * q < valid) {   // N is a multiple of 8: pairs are whole
                __nv_bfloat162 v2 = *reinterpret_cast<__nv_bfloat162*>(&packed[q]);
                *reinterpret_cast<__nv_bfloat162*>(dst + 2 * q) = v2;
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));   // accumulator drained: the MMA warp may reuse it
      if (p.prob != nullptr) {
        float* part = head_part + acc * 2 * kTileM;
        part[half * kTileM + row] = dot;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");   // the 8 epilogue warps only
        if (half == 0 && m < p.M) {
          const float logit = part[row] + part[kTileM + row] + (p.b2 != nullptr ? __ldg(p.b2) : 0.0f);
          p.prob[m] = 1.0f / (1.0f + expf(-logit));
        }
        // the next tile uses the other head_part slot; the slot after that is separated from this read by a bar.sync
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int BLOCK_N>
static int launch(const CUtensorMap& map_w, Params& p, cudaStream_t stream) {
  auto kern = edge_mlp_kernel<BLOCK_N>;
  static PerDeviceOnce configured;  // cudaFuncSetAttribute is per device
  if (configured.need()) {
    LLP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    configured.done();
  }
  const int num_kb = p.K / kBK;
  const int b_bytes = num_kb * BLOCK_N * 128;
  const int fixed = 1024 + kAccStages * 2 * kTileM * 4 + 2 * 256 * 4 + 512;
  int stages = (kSmemMax - fixed - b_bytes) / kABytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return LLP_E_SHAPE;
  p.stages = stages;
  const int smem = b_bytes + stages * kABytes + fixed;
  const int64_t m_tiles = ceil_div(p.M, kTileM);
  const unsigned grid = (unsigned)(m_tiles < kNumSMs ? m_tiles : kNumSMs);
  kern<<<grid, kThreads, smem, stream>>>(map_w, p);
  LLP_LAUNCH_OK();
  return 0;
}

}  // namespace em
}  // namespace llp

using namespace llp;

extern "C" int llp_edge_mlp_supported(int64_t K, int64_t N) {
  return ((K == 64 || K == 128 || K == 256) && N > 0 && N % 8 == 0 && N <= 256 && K * ((N > 128) ? 256 : (N > 64 ? 128 : 64)) * 2 <= 128 * 1024)
             ? 1 : 0;
}

extern "C" int llp_edge_mlp_fused(const llp_edge_mlp_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(a != nullptr && a->M >= 0);
  if (int rc = check_device()) return rc;
  if (a->M == 0) return 0;
  LLP_CHECK_ARG(a->h && a->u && a->v && a->W1 && a->ldh >= a->K && a->ldw1 >= a->K);
  LLP_CHECK_ARG(a->z == nullptr || a->ldz >= a->K);
  LLP_CHECK_ARG(a->y == nullptr || a->ldy >= a->N);
  LLP_CHECK_ARG(a->prob == nullptr || a->w2 != nullptr);
  LLP_CHECK_ARG(a->y != nullptr || a->prob != nullptr);
  LLP_CHECK_ARG(a->dropout_p >= 0.0f && a->dropout_p < 1.0f);
  if (!llp_edge_mlp_supported(a->K, a->N)) return LLP_E_SHAPE;
  if (!aligned(a->h, 16) || (a->ldh * 2) % 16 != 0 || (a->z && (!aligned(a->z, 16) || (a->ldz * 2) % 16 != 0)) ||
      (a->bias1 && !aligned(a->bias1, 16)) || (a->w2 && !aligned(a->w2, 8)))
    return LLP_E_ALIGN;
  const int bn = a->N > 128 ? 256 : (a->N > 64 ? 128 : 64);
  CUtensorMap map_w;
  memset(&map_w, 0, sizeof(map_w));
  if (int rc = tc::make_map(&map_w, a->W1, a->N, a->K, a->ldw1, em::kBK, bn)) return rc;
  em::Params p{};
  p.h = reinterpret_cast<const __nv_bfloat16*>(a->h); p.ldh = a->ldh;
  p.u = a->u; p.v = a->v; p.M = a->M; p.K = (int)a->K; p.N = (int)a->N;
  p.bias1 = a->bias1; p.relu = a->relu; p.dropout_p = a->dropout_p; p.seed = a->seed; p.offset = a->offset;
  p.rng_state = a->rng_state;
  p.z = reinterpret_cast<__nv_bfloat16*>(a->z); p.ldz = a->ldz;
  p.y = reinterpret_cast<__nv_bfloat16*>(a->y); p.ldy = a->ldy;
  p.w2 = a->w2; p.b2 = a->b2; p.prob = a->prob;
  p.dbg = g_tuning[15] ? debug_buffer() : nullptr;
  if (bn == 256) return em::launch<256>(map_w, p, stream);
  if (bn == 128) return em::launch<128>(map_w, p, stream);
  return em::launch<64>(map_w, p, stream);
}
