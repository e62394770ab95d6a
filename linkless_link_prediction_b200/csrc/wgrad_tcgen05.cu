// Fused weight-gradient kernel of one layer (sm_100a, tcgen05):
//
//     dWa[N1,N2a] (+)= G[M,N1]^T * A[M,N2a]      dWb[N1,N2b] (+)= G^T * B[M,N2b]      dbias[N1] (+)= colsum(G)
//
// i.e. everything the backward of `lin_l(agg) + lin_r(x)` (PyG SAGEConv; sageconv_updated.py:71-76) or of one
// nn.Linear (models.py:48,143) needs from the output gradient G, in ONE pass that reads G, A and B once.  Replaces two
// split-K TN GEMMs + a column-sum kernel (autograd of F.linear: two `mm`s and a `sum(0)` per layer).
//
// One CTA owns a [128 x (N2a+N2b)] fp32 accumulator in TMEM (up to 512 columns = all of it) for one slice of the
// reduction dimension M (split-K over CTAs, grid = N1/128 tiles x splits <= 148).  All operands are read MN-major
// straight from the row-major activations: TMA boxes of {64 columns, 32 rows} land as 128B-swizzled slabs, a stage
// = 2 slabs of G + the slabs of A and B (40 KB for 256+256 columns, 5 stages).  warp 0 = TMA producer, warp 1 = MMA
// issuer, warp 2 = bias-gradient warp (column sums of the G slabs straight from shared memory while the tensor core
// works on the same stage), then all 8 epilogue warps drain TMEM into fp32 split partials; `wgrad_reduce_kernel` adds
// the partials in split order (deterministic) into the gradient buffers.
#include "tcgen05.cuh"

namespace llp {
namespace wg {

using namespace tc;

constexpr int kTileN1 = 128;        // rows of the output tile = UMMA M
constexpr int kEpiWarps = 8;
constexpr int kThreads = 96 + 32 * kEpiWarps;  // producer, MMA, colsum, 8 epilogue warps
constexpr int kSmemBudget = 220 * 1024;

struct Params {
  int64_t M, N1;
  int debug_skip_mma, debug_wide_box;
  long long* debug_times;        // nullptr, or [grid][4] globaltimer stamps (start, accumulators ready, epilogue done)
  int strided;                   // k-blocks of a split: 0 = one contiguous slice of M, 1 = every splits-th block
  int slabs_a, slabs_b;          // 64-column slabs of A and B (B may have 0)
  int64_t n2a, n2b;              // logical widths
  int splits;
  int64_t rows_per_split;        // multiple of kBK
  int stages;
  int tmem_cols;
  float* partial;                // [splits][N1][ncols]  (ncols = 64 * (slabs_a + slabs_b))
  float* partial_bias;           // [splits][N1] or nullptr
};

struct Maps {
  CUtensorMap g, a, b;
};

template <int kBK>
__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const __grid_constant__ Maps maps, const Params p) {
  constexpr int kSlab = kBK * 128;  // one {64 col, kBK row} box, 128B-swizzled
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int slabs = 2 + p.slabs_a + p.slabs_b;
  const int stage_bytes = slabs * kSlab;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full_bar = bars;                 // [stages]
  uint64_t* empty_bar = bars + p.stages;     // [stages]
  uint64_t* tmem_full = bars + 2 * p.stages; // [1]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n1_tiles = (int)((p.N1 + kTileN1 - 1) / kTileN1);
  const int tile = blockIdx.x % n1_tiles, split = blockIdx.x / n1_tiles;
  const int n1_0 = tile * kTileN1;
  int64_t m_begin, kb_stride;
  int num_kb;
  if (p.strided) {  // block kb of this CTA = global k-block (split + kb * splits): all CTAs sweep M together
    const int64_t total_kb = (p.M + kBK - 1) / kBK;
    m_begin = (int64_t)split * kBK;
    kb_stride = (int64_t)p.splits * kBK;
    num_kb = total_kb > split ? (int)((total_kb - split + p.splits - 1) / p.splits) : 0;
  } else {
    m_begin = (int64_t)split * p.rows_per_split;
    const int64_t m_end = m_begin + p.rows_per_split < p.M ? m_begin + p.rows_per_split : p.M;
    kb_stride = kBK;
    num_kb = m_end > m_begin ? (int)((m_end - m_begin + kBK - 1) / kBK) : 0;
  }
  const bool want_bias = p.partial_bias != nullptr;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.g);
    tma_prefetch_desc(&maps.a);
    if (p.slabs_b > 0) tma_prefetch_desc(&maps.b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), want_bias ? 2 : 1);
    }
    mbar_init(smem_u32(tmem_full), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_holder), (uint32_t)p.tmem_cols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the ring, one elected lane issues) =====================
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
      if (elect_one_sync()) {
      const uint32_t bar = smem_u32(&full_bar[stage]);
      mbar_expect_tx(bar, (uint32_t)stage_bytes);
      const uint32_t s0 = smem_u32(smem + (size_t)stage * stage_bytes);
      const int k0 = (int)(m_begin + (int64_t)kb * kb_stride);
      if (p.debug_wide_box) {  // load-rate experiment only: un-swizzled full-row boxes (layout unusable by the MMA)
        tma_load_2d(s0, &maps.g, n1_0, k0, bar);
        tma_load_2d(s0 + 2 * kSlab, &maps.a, 0, k0, bar);
        if (p.slabs_b > 0) tma_load_2d(s0 + (2 + p.slabs_a) * kSlab, &maps.b, 0, k0, bar);
      } else {
      tma_load_2d(s0, &maps.g, n1_0, k0, bar);
      tma_load_2d(s0 + kSlab, &maps.g, n1_0 + 64, k0, bar);
      for (int j = 0; j < p.slabs_a; ++j) tma_load_2d(s0 + (2 + j) * kSlab, &maps.a, j * 64, k0, bar);
      for (int j = 0; j < p.slabs_b; ++j) tma_load_2d(s0 + (2 + p.slabs_a + j) * kSlab, &maps.b, j * 64, k0, bar);
      }
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp walks the ring, one elected lane issues) =====================
    // MN-major SW128 operands: LBO = slab stride, SBO = 8 reduction rows (1024 B), one UMMA_K step = 16 rows = 2048 B
    const uint32_t idesc_a = make_idesc(kTileN1, p.slabs_a * 64, true);
    const uint32_t idesc_b = make_idesc(kTileN1, p.slabs_b * 64, true);
    constexpr uint32_t kHi = desc_hi_sw128(1024);
    constexpr uint32_t kStep = (UMMA_K * 128) >> 4;
    const uint32_t g_lo0 = desc_lo(smem_u32(smem), kSlab);
    const uint32_t a_off = (uint32_t)(2 * kSlab) >> 4, b_off = (uint32_t)((2 + p.slabs_a) * kSlab) >> 4;
    const uint32_t tmem_b = tmem_base + (uint32_t)(p.slabs_a * 64);
    const bool has_b = p.slabs_b > 0;
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(smem_u32(&full_bar[stage]), phase);
      tcgen05_fence_after();
      if (elect_one_sync()) {
        const uint32_t g_lo = g_lo0 + (uint32_t)stage * ((uint32_t)stage_bytes >> 4);
#pragma unroll
        for (int k = 0; k < kBK / UMMA_K; ++k) {
          if (p.debug_skip_mma) break;
          const uint64_t gdesc = desc_from(g_lo + k * kStep, kHi);
          umma_bf16(tmem_base, gdesc, desc_from(g_lo + a_off + k * kStep, kHi), idesc_a, (uint32_t)((kb | k) != 0));
          if (has_b) umma_bf16(tmem_b, gdesc, desc_from(g_lo + b_off + k * kStep, kHi), idesc_b, (uint32_t)((kb | k) != 0));
        }
        umma_commit(smem_u32(&empty_bar[stage]));
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    if (elect_one_sync()) umma_commit(smem_u32(tmem_full));
    __syncwarp();
  } else if (warp == 2 && want_bias) {
    // ===================== bias gradient: column sums of the G slabs, from shared memory =====================
    // lane -> (row half, slab, 16-byte chunk): 8 consecutive columns of G, one half of the rows of every stage
    const int half = lane >> 4, slab = (lane >> 3) & 1, chunk = lane & 7;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(smem_u32(&full_bar[stage]), phase);
      const uint8_t* base = smem + (size_t)stage * stage_bytes + slab * kSlab;
#pragma unroll 4
      for (int r = (kBK / 2) * half; r < (kBK / 2) * half + kBK / 2; ++r) {
        const uint4 v = *reinterpret_cast<const uint4*>(base + r * 128 + ((chunk ^ (r & 7)) << 4));
        float f[8];
        unpack16(v, f, __nv_bfloat16());
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += f[i];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&empty_bar[stage]));
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
    if (half == 0) {
      const int n = n1_0 + slab * 64 + chunk * 8;
      float* dst = p.partial_bias + (int64_t)split * p.N1 + n;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (n + i < p.N1) dst[i] = acc[i];
    }
  } else if (warp >= 3) {
    // ===================== epilogue: TMEM -> fp32 split partials =====================
    const int ew = warp - 3;
    const int quad = warp & 3;   // TMEM lane quadrant this warp may read
    // the two warps that share a quadrant take alternate 32-column chunks
    int first = 0;
    {
      // warps 3..10: quadrants 3,0,1,2,3,0,1,2 -> the second visitor of a quadrant starts at chunk 1
      first = ew >= 4 ? 1 : 0;
    }
    const int ncols = 64 * (p.slabs_a + p.slabs_b);
    long long t_start = 0, t_ready = 0;
    if (p.debug_times != nullptr && warp == 3 && lane == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    mbar_wait(smem_u32(tmem_full), 0);
    tcgen05_fence_after();
    if (p.debug_times != nullptr && warp == 3 && lane == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ready));
    const int64_t row = n1_0 + quad * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    for (int c0 = first * 32; c0 < ncols; c0 += 64) {
      uint32_t r[32];
      tmem_ld32(taddr + c0, r);
      if (row < p.N1 && num_kb > 0) {
        float* dst = p.partial + ((int64_t)split * p.N1 + row) * ncols + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          U32x8 v;
#pragma unroll
          for (int q = 0; q < 8; ++q) v.v[q] = r[j + q];
          stg_v8(dst + j, v);   // the workspace is 256-byte aligned and every row is a multiple of 64 floats
        }
      } else if (row < p.N1) {
        float* dst = p.partial + ((int64_t)split * p.N1 + row) * ncols + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(dst + j) = make_uint4(0, 0, 0, 0);
      }
    }
    tcgen05_fence_before();
    if (p.debug_times != nullptr && warp == 3 && lane == 0) {
      long long t_done;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_done));
      p.debug_times[blockIdx.x * 4 + 0] = t_start;
      p.debug_times[blockIdx.x * 4 + 1] = t_ready;
      p.debug_times[blockIdx.x * 4 + 2] = t_done;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// Sum of the split partials.  A block = 64 output quads (4 consecutive columns of a row, or bias elements) x 4 split
// groups: group g adds the splits g, g+4, g+8, ... (8 loads in flight per thread), the four group sums are combined in
// group order through shared memory.  A fixed association => deterministic; four times the loads in flight of a
// one-thread-per-output reduce (the partials are L2-resident, the reduce is latency-bound).
constexpr int kRedQuads = 64, kRedGroups = 4;
__global__ void __launch_bounds__(kRedQuads * kRedGroups)
wgrad_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ partial_bias, int splits, int64_t N1,
                    int ncols, int slabs_a, int64_t n2a, int64_t n2b, float* __restrict__ dWa, int64_t ldwa,
                    float* __restrict__ dWb, int64_t ldwb, float* __restrict__ dbias, int accumulate) {
  __shared__ float4 red[kRedGroups][kRedQuads];
  const int quads = ncols / 4;
  const int tx = threadIdx.x % kRedQuads, g = threadIdx.x / kRedQuads;
  const int64_t i = blockIdx.x * (int64_t)kRedQuads + tx;
  const int64_t n_quads = N1 * quads;
  const int64_t stride = N1 * (int64_t)ncols;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  bool is_bias = false;
  int64_t row = 0, j = 0, width = 0;
  bool in_a = true, live = false;
  if (i < n_quads) {
    row = i / quads;
    const int col = (int)(i % quads) * 4;
    in_a = col < slabs_a * 64;
    j = in_a ? col : col - slabs_a * 64;
    width = in_a ? n2a : n2b;
    live = j < width;  // not the zero padding of the last slab
    if (live) {
      const float* src = partial + row * ncols + col;
      int k = g;
      for (; k + 7 * kRedGroups < splits; k += 8 * kRedGroups) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(src + (int64_t)(k + u * kRedGroups) * stride));
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
      }
      for (; k < splits; k += kRedGroups) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(src + (int64_t)k * stride));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
  } else if (dbias != nullptr && i < n_quads + N1) {
    is_bias = live = true;
    row = i - n_quads;
    for (int k = g; k < splits; k += kRedGroups) acc.x += partial_bias[(int64_t)k * N1 + row];
  }
  red[g][tx] = acc;
  __syncthreads();
  if (g != 0 || !live) return;
  float4 s = red[0][tx];
#pragma unroll
  for (int q = 1; q < kRedGroups; ++q) { s.x += red[q][tx].x; s.y += red[q][tx].y; s.z += red[q][tx].z; s.w += red[q][tx].w; }
  if (is_bias) {
    dbias[row] = accumulate ? dbias[row] + s.x : s.x;
    return;
  }
  float* dst = (in_a ? dWa + row * ldwa : dWb + row * ldwb) + j;
  const float v4[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (j + q < width) dst[q] = accumulate ? dst[q] + v4[q] : v4[q];
}

static void plan(int64_t M, int64_t N1, int64_t n2a, int64_t n2b, Params* p, int kBK) {
  const int kSlab = kBK * 128;
  p->M = M; p->N1 = N1; p->n2a = n2a; p->n2b = n2b;
  p->slabs_a = (int)ceil_div(n2a, 64);
  p->slabs_b = (int)ceil_div(n2b, 64);
  const int n1_tiles = (int)ceil_div(N1, kTileN1);
  int64_t kblocks = ceil_div(M, kBK);
  int64_t want = kNumSMs / n1_tiles;
  if (want < 1) want = 1;
  int64_t s = want < kblocks ? want : kblocks;
  if (s < 1) s = 1;
  p->rows_per_split = ceil_div(kblocks, s) * kBK;
  p->splits = (int)ceil_div(M, p->rows_per_split);
  const int stage_bytes = (2 + p->slabs_a + p->slabs_b) * kSlab;
  int stages = (kSmemBudget - 2048) / stage_bytes;
  p->stages = stages > 12 ? 12 : stages;
  if (g_tuning[10] > 0 && g_tuning[10] < p->stages) p->stages = g_tuning[10];
#ifdef LLP_EXPERIMENT   // work-skipping experiments exist only in -DLLP_EXPERIMENT builds
  p->debug_skip_mma = g_tuning[11];
  p->debug_wide_box = g_tuning[13];
#else
  p->debug_skip_mma = 0;
  p->debug_wide_box = 0;
#endif
  p->strided = g_tuning[14];
  p->debug_times = g_tuning[15] ? debug_buffer() : nullptr;
  const int cols = 64 * (p->slabs_a + p->slabs_b);
  p->tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
}

}  // namespace wg

bool wgrad_tcgen05_supported(int64_t M, int64_t N1, int64_t n2a, int64_t n2b) {
  // operand B needs an N that is a legal UMMA shape (multiple of 16, <= 256): whole 64-column slabs always are
  return M > 0 && N1 > 0 && n2a > 0 && n2a <= 256 && n2b >= 0 && n2b <= 256;
}

size_t wgrad_tcgen05_workspace_bytes(int64_t M, int64_t N1, int64_t n2a, int64_t n2b) {
  wg::Params p{};
  wg::plan(M, N1, n2a, n2b, &p, 32);
  const size_t ncols = 64 * (size_t)(p.slabs_a + p.slabs_b);
  return ((size_t)p.splits * (size_t)N1 * (ncols + 1)) * sizeof(float) + 256;
}

int wgrad_tcgen05(int64_t M, int64_t N1, const void* G, int64_t ldg, int64_t n2a, const void* A, int64_t lda, float* dWa,
                  int64_t ldwa, int64_t n2b, const void* B, int64_t ldb, float* dWb, int64_t ldwb, float* dbias,
                  int accumulate, float* ws, cudaStream_t stream) {
  using namespace wg;
  const int kBK = g_tuning[12] == 64 ? 64 : 32;
  const int kSlab = kBK * 128;
  Params p{};
  plan(M, N1, n2a, n2b, &p, kBK);
  Maps maps;
  memset(&maps, 0, sizeof(maps));
  if (p.debug_wide_box) {
    p.debug_skip_mma = 1;
    if (int rc = tc::make_map(&maps.g, G, M, N1, ldg, 128, kBK, false)) return rc;
    if (int rc = tc::make_map(&maps.a, A, M, n2a, lda, p.slabs_a * 64, kBK, false)) return rc;
    if (p.slabs_b > 0)
      if (int rc = tc::make_map(&maps.b, B, M, n2b, ldb, p.slabs_b * 64, kBK, false)) return rc;
  } else {
  if (int rc = tc::make_map(&maps.g, G, M, N1, ldg, 64, kBK)) return rc;
  if (int rc = tc::make_map(&maps.a, A, M, n2a, lda, 64, kBK)) return rc;
  if (p.slabs_b > 0)
    if (int rc = tc::make_map(&maps.b, B, M, n2b, ldb, 64, kBK)) return rc;
  }
  const int ncols = 64 * (p.slabs_a + p.slabs_b);
  p.partial = ws;
  p.partial_bias = dbias != nullptr ? ws + (size_t)p.splits * (size_t)N1 * ncols : nullptr;
  const int smem = p.stages * (2 + p.slabs_a + p.slabs_b) * kSlab + 1024 + 256;
  static PerDeviceOnce configured;  // cudaFuncSetAttribute is per device
  if (configured.need()) {
    LLP_CUDA(cudaFuncSetAttribute(wgrad_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    LLP_CUDA(cudaFuncSetAttribute(wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    configured.done();
  }
  const unsigned grid = (unsigned)(ceil_div(N1, kTileN1) * p.splits);
  if (kBK == 64) wgrad_kernel<64><<<grid, kThreads, smem, stream>>>(maps, p);
  else wgrad_kernel<32><<<grid, kThreads, smem, stream>>>(maps, p);
  LLP_LAUNCH_OK();
  const int64_t work = N1 * (ncols / 4) + (dbias != nullptr ? N1 : 0);
  wgrad_reduce_kernel<<<(unsigned)ceil_div(work, kRedQuads), kRedQuads * kRedGroups, 0, stream>>>(
      p.partial, p.partial_bias, p.splits, N1, ncols, p.slabs_a, n2a, n2b, dWa, ldwa, dWb, ldwb, dbias, accumulate);
  LLP_LAUNCH_OK();
  return 0;
}

}  // namespace llp
