// CUDA-core dense kernels: the fp32-parity GEMM backend (true fp32 FFMA, as the reference's
// cuBLAS SGEMM with allow_tf32=False — SURVEY.md K4/H2), plus the small dense helpers
// (column sums, cast/transposes, relu-dropout gate).  The tensor-core path is gemm_tcgen05.cu.
#include "common.cuh"

namespace llp {

constexpr int kTM = 64, kTN = 64, kTK = 16, kSimtThreads = 256;

struct StridedOperand {
  const void* p;
  int64_t s_outer;  // stride of the M (or N) index
  int64_t s_k;      // stride of the reduction index
  int64_t K;
};

template <typename T>
__device__ __forceinline__ void load_tile(const StridedOperand& op, int64_t outer0, int64_t outer_lim, int64_t k0,
                                          int64_t k_lim, float (*tile)[kTM + 1]) {
  const T* p = reinterpret_cast<const T*>(op.p);
  const bool k_contig = op.s_k == 1;
#pragma unroll
  for (int i = 0; i < (kTM * kTK) / kSimtThreads; ++i) {
    int idx = threadIdx.x + i * kSimtThreads;
    int k = k_contig ? (idx % kTK) : (idx / kTM);
    int o = k_contig ? (idx / kTK) : (idx % kTM);
    int64_t go = outer0 + o, gk = k0 + k;
    float v = 0.0f;
    if (go < outer_lim && gk < k_lim) v = to_f32(p[go * op.s_outer + gk * op.s_k]);
    tile[k][o] = v;
  }
}

// C[m,n] = sum over operand pairs of sum_k A(m,k) * B(n,k); blockIdx.z selects a K-split.
template <typename T, typename TO, bool kPartial>
__global__ void __launch_bounds__(kSimtThreads)
gemm_simt_kernel(int64_t M, int64_t N, StridedOperand a1, StridedOperand b1, StridedOperand a2, StridedOperand b2,
                 int64_t k_per_split, EpilogueParams ep, TO* __restrict__ D, int64_t ldd, float* __restrict__ partial) {
  __shared__ float As[kTK][kTM + 1];
  __shared__ float Bs[kTK][kTN + 1];
  if (ep.dropout_p > 0.0f) resolve_rng(ep);
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int64_t m0 = (int64_t)blockIdx.y * kTM, n0 = (int64_t)blockIdx.x * kTN;
  float acc[4][4] = {};
  for (int pair = 0; pair < 2; ++pair) {
    const StridedOperand& a = pair == 0 ? a1 : a2;
    const StridedOperand& b = pair == 0 ? b1 : b2;
    if (a.p == nullptr) continue;
    int64_t kb = kPartial ? (int64_t)blockIdx.z * k_per_split : 0;
    int64_t ke = kPartial ? min(a.K, kb + k_per_split) : a.K;
    for (int64_t k0 = kb; k0 < ke; k0 += kTK) {
      load_tile<T>(a, m0, M, k0, ke, As);
      load_tile<T>(b, n0, N, k0, ke, Bs);
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kTK; ++k) {
        float av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t n = n0 + tx * 4 + j;
      if (n >= N) continue;
      if constexpr (kPartial) {
        partial[((int64_t)blockIdx.z * M + m) * N + n] = acc[i][j];
      } else {
        D[m * ldd + n] = from_f32<TO>(epilogue_apply<TO>(acc[i][j], m, n, ep));
      }
    }
  }
}

// D[i] (+)= sum_z partial[z][i]   (fixed order => deterministic)
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, int64_t rows, int64_t cols,
                                     float* __restrict__ D, int64_t ldd, int accumulate) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  float acc = 0.0f;
  for (int z = 0; z < splits; ++z) acc += partial[(int64_t)z * rows * cols + i];
  int64_t r = i / cols, c = i % cols;
  D[r * ldd + c] = accumulate ? D[r * ldd + c] + acc : acc;
}

template <typename T, typename TO>
static int gemm_nt_simt_typed(const llp_gemm_nt_args& a, cudaStream_t stream) {
  StridedOperand a1{a.A1, a.lda1, 1, a.K1}, b1{a.B1, a.ldb1, 1, a.K1};
  StridedOperand a2{a.A2, a.lda2, 1, a.K2}, b2{a.B2, a.ldb2, 1, a.K2};
  if (a.A2 == nullptr || a.K2 == 0) a2.p = b2.p = nullptr;
  EpilogueParams ep{a.bias, a.addend, a.ldadd, a.gate, a.ldgate, a.gate_scale, a.relu, a.dropout_p, a.seed, a.offset, a.rng_state};
  dim3 grid((unsigned)ceil_div(a.N, kTN), (unsigned)ceil_div(a.M, kTM), 1);
  gemm_simt_kernel<T, TO, false><<<grid, kSimtThreads, 0, stream>>>(a.M, a.N, a1, b1, a2, b2, 0, ep,
                                                                  reinterpret_cast<TO*>(a.D), a.ldd, nullptr);
  LLP_LAUNCH_OK();
  return 0;
}

int gemm_nt_simt(const llp_gemm_nt_args& a, cudaStream_t stream) {
  if (a.dtype == LLP_F32 && a.out_dtype == LLP_F32) return gemm_nt_simt_typed<float, float>(a, stream);
  if (a.dtype == LLP_BF16 && a.out_dtype == LLP_BF16) return gemm_nt_simt_typed<__nv_bfloat16, __nv_bfloat16>(a, stream);
  if (a.dtype == LLP_BF16 && a.out_dtype == LLP_F32) return gemm_nt_simt_typed<__nv_bfloat16, float>(a, stream);
  if (a.dtype == LLP_F32 && a.out_dtype == LLP_BF16) return gemm_nt_simt_typed<float, __nv_bfloat16>(a, stream);
  return LLP_E_BADARG;
}

int tn_splits(int64_t M, int64_t N1, int64_t N2) {
  int64_t tiles = ceil_div(N1, kTM) * ceil_div(N2, kTN);
  int64_t want = ceil_div((int64_t)kNumSMs * 4, tiles);
  int64_t max_by_k = ceil_div(M, 256);
  int64_t s = want < max_by_k ? want : max_by_k;
  return (int)(s < 1 ? 1 : s);
}

template <typename T>
static int gemm_tn_simt_typed(int64_t M, int64_t N1, int64_t N2, const void* A, int64_t lda, const void* B, int64_t ldb,
                              float* D, int64_t ldd, int accumulate, float* ws, cudaStream_t stream) {
  int splits = tn_splits(M, N1, N2);
  int64_t k_per = ceil_div(ceil_div(M, splits), kTK) * kTK;
  splits = (int)ceil_div(M, k_per);
  StridedOperand a1{A, 1, lda, M}, b1{B, 1, ldb, M}, none{nullptr, 0, 0, 0};
  EpilogueParams ep{};
  dim3 grid((unsigned)ceil_div(N2, kTN), (unsigned)ceil_div(N1, kTM), (unsigned)splits);
  gemm_simt_kernel<T, float, true><<<grid, kSimtThreads, 0, stream>>>(N1, N2, a1, b1, none, none, k_per, ep, nullptr, 0, ws);
  LLP_LAUNCH_OK();
  int64_t n = N1 * N2;
  splitk_reduce_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(ws, splits, N1, N2, D, ldd, accumulate);
  LLP_LAUNCH_OK();
  return 0;
}

int gemm_tn_simt(int dtype, int64_t M, int64_t N1, int64_t N2, const void* A, int64_t lda, const void* B, int64_t ldb,
                 float* D, int64_t ldd, int accumulate, float* ws, cudaStream_t stream) {
  if (dtype == LLP_F32) return gemm_tn_simt_typed<float>(M, N1, N2, A, lda, B, ldb, D, ldd, accumulate, ws, stream);
  if (dtype == LLP_BF16) return gemm_tn_simt_typed<__nv_bfloat16>(M, N1, N2, A, lda, B, ldb, D, ldd, accumulate, ws, stream);
  return LLP_E_BADARG;
}

int splitk_reduce(const float* partial, int splits, int64_t rows, int64_t cols, float* D, int64_t ldd, int accumulate,
                  cudaStream_t stream) {
  int64_t n = rows * cols;
  splitk_reduce_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(partial, splits, rows, cols, D, ldd, accumulate);
  LLP_LAUNCH_OK();
  return 0;
}

// ---- column reductions ---------------------------------------------------------------------------
// out[n] = sum_m w[m] * A[m,n]   (w == nullptr -> plain column sums: the bias gradient).
// grid (column tiles of 32*VE, row splits); every warp reads whole 16-byte vectors of consecutive rows, so a
// [M,256] bf16 matrix is streamed with fully coalesced 512-byte rows.  Partials are combined in fixed order by
// colreduce_final_kernel => deterministic.
template <typename T, bool kVec>
__global__ void __launch_bounds__(256)
colreduce_partial_kernel(const T* __restrict__ A, int64_t lda, int64_t M, int64_t N, const float* __restrict__ w,
                         int64_t rows_per_split, float* __restrict__ partial) {
  constexpr int VE = kVec ? Vec16<T>::n : 1;
  __shared__ float red[8][32 * VE + 1];
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const int64_t c = ((int64_t)blockIdx.x * 32 + lane) * VE;
  const int64_t mb = (int64_t)blockIdx.y * rows_per_split, me = min(M, mb + rows_per_split);
  float acc[VE];
#pragma unroll
  for (int i = 0; i < VE; ++i) acc[i] = 0.0f;
  if (c < N) {
    if constexpr (kVec) {
      int64_t m = mb + wv;
      for (; m + 24 < me; m += 32) {  // 4 independent 128-bit loads in flight per lane
        uint4 v[4];
        float wm[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          v[u] = ldg_nc_v4(A + (m + 8 * u) * lda + c);
          wm[u] = w != nullptr ? __ldg(w + m + 8 * u) : 1.0f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float f[VE];
          unpack16(v[u], f, T());
#pragma unroll
          for (int i = 0; i < VE; ++i) acc[i] = fmaf(wm[u], f[i], acc[i]);
        }
      }
      for (; m < me; m += 8) {
        const float wm = w != nullptr ? __ldg(w + m) : 1.0f;
        float f[VE];
        unpack16(ldg_nc_v4(A + m * lda + c), f, T());
#pragma unroll
        for (int i = 0; i < VE; ++i) acc[i] = fmaf(wm, f[i], acc[i]);
      }
    } else {
      for (int64_t m = mb + wv; m < me; m += 8) {
        const float wm = w != nullptr ? __ldg(w + m) : 1.0f;
        acc[0] = fmaf(wm, to_f32(A[m * lda + c]), acc[0]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VE; ++i) red[wv][lane * VE + i] = acc[i];
  __syncthreads();
  for (int t = threadIdx.x; t < 32 * VE; t += 256) {
    const int64_t n = (int64_t)blockIdx.x * 32 * VE + t;
    if (n < N) {
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += red[i][t];
      partial[(int64_t)blockIdx.y * N + n] = s;
    }
  }
}

// one warp per output column: lanes stride over the splits, then a shuffle tree (fixed order => deterministic)
__global__ void colreduce_final_kernel(const float* __restrict__ partial, int splits, int64_t N, float* __restrict__ out,
                                       int accumulate) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (n >= N) return;
  double s = 0.0;   // up to 2048 signed partials per column: double keeps the bias gradient at fp32-oracle accuracy
  for (int i = lane; i < splits; i += 32) s += (double)partial[(int64_t)i * N + n];
  s = warp_sum(s);
  if (lane == 0) out[n] = accumulate ? out[n] + (float)s : (float)s;
}

constexpr int kColreduceSplits = 2048;

size_t colreduce_workspace_bytes(int64_t N) { return (size_t)kColreduceSplits * (size_t)(N > 0 ? N : 1) * sizeof(float); }

template <typename T>
static int colreduce_typed(const void* A_, int64_t lda, int64_t M, int64_t N, const float* w, float* out, int accumulate,
                           float* partial, cudaStream_t stream) {
  const T* A = reinterpret_cast<const T*>(A_);
  const bool vec = aligned(A, 16) && (lda * sizeof(T)) % 16 == 0 && N % Vec16<T>::n == 0;
  const int VE = vec ? Vec16<T>::n : 1;
  const int splits = (int)imin64(kColreduceSplits, imax64(1, ceil_div(M, 96)));
  const int64_t rows_per_split = ceil_div(imax64(M, 1), splits);
  dim3 grid((unsigned)ceil_div(N, 32 * VE), (unsigned)splits);
  if (vec) colreduce_partial_kernel<T, true><<<grid, 256, 0, stream>>>(A, lda, M, N, w, rows_per_split, partial);
  else colreduce_partial_kernel<T, false><<<grid, 256, 0, stream>>>(A, lda, M, N, w, rows_per_split, partial);
  LLP_LAUNCH_OK();
  colreduce_final_kernel<<<(unsigned)ceil_div(N * 32, 256), 256, 0, stream>>>(partial, splits, N, out, accumulate);
  LLP_LAUNCH_OK();
  return 0;
}

int colreduce(int dtype, const void* A, int64_t lda, int64_t M, int64_t N, const float* w, float* out, int accumulate,
              float* partial, cudaStream_t stream) {
  if (dtype == LLP_F32) return colreduce_typed<float>(A, lda, M, N, w, out, accumulate, partial, stream);
  if (dtype == LLP_BF16) return colreduce_typed<__nv_bfloat16>(A, lda, M, N, w, out, accumulate, partial, stream);
  return LLP_E_BADARG;
}

// ---- cast / transpose ----------------------------------------------------------------------------
template <typename TS, typename TD>
__global__ void cast2d_kernel(const TS* __restrict__ src, int64_t lds, int64_t rows, int64_t cols, TD* __restrict__ dst,
                              int64_t ldd, int transpose) {
  __shared__ float tile[32][33];
  int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? to_f32(src[r * lds + c]) : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    if (transpose) {
      int64_t c = c0 + i, r = r0 + threadIdx.x;  // dst[c, r]
      if (r < rows && c < cols) dst[c * ldd + r] = from_f32<TD>(tile[threadIdx.x][i]);
    } else {
      int64_t r = r0 + i, c = c0 + threadIdx.x;
      if (r < rows && c < cols) dst[r * ldd + c] = from_f32<TD>(tile[i][threadIdx.x]);
    }
  }
}

// ---- low-precision copies of all weight matrices in one launch ------------------------------------
// After every optimiser step the fp32 master weights are re-materialised as bf16 [rows, cols] (the B operand of the
// forward GEMMs) and bf16 [cols, rows] (the B operand of the input-gradient GEMMs) for up to kMaxPrep matrices by ONE
// grid of 32x32 tiles, instead of a cast or transpose launch per weight and use.
constexpr int kMaxPrep = 16;
struct PrepTable {
  int count;
  int tile_begin[kMaxPrep + 1];
  llp_weight_desc d[kMaxPrep];
};

__global__ void weights_prep_kernel(const PrepTable t) {
  __shared__ float tile[32][33];
  int k = 0;
  while (k + 1 < t.count && (int)blockIdx.x >= t.tile_begin[k + 1]) ++k;
  const llp_weight_desc d = t.d[k];
  const int local = blockIdx.x - t.tile_begin[k];
  const int tiles_x = (int)((d.cols + 31) / 32);
  const int64_t r0 = (int64_t)(local / tiles_x) * 32, c0 = (int64_t)(local % tiles_x) * 32;
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(d.dst);
  __nv_bfloat16* dst_t = reinterpret_cast<__nv_bfloat16*>(d.dst_t);
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    const float v = (r < d.rows && c < d.cols) ? d.src[r * d.cols + c] : 0.0f;
    tile[i][threadIdx.x] = v;
    if (dst != nullptr && r < d.rows && c < d.cols) dst[r * d.ld + c] = __float2bfloat16_rn(v);
  }
  if (dst_t == nullptr) return;
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t c = c0 + i, r = r0 + threadIdx.x;  // dst_t[c, r]
    if (r < d.rows && c < d.cols) dst_t[c * d.ld_t + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

// y = gate > 0 ? g*scale : 0, one 16-byte vector per thread (rows are 16-byte multiples) or scalar fallback
// plain (non-transposing) cast: one 16-byte output vector per thread when rows allow it
template <typename TS, typename TD>
__global__ void cast_rows_kernel(const TS* __restrict__ src, int64_t lds, int64_t rows, int64_t cols, TD* __restrict__ dst,
                                 int64_t ldd) {
  constexpr int VE = Vec16<TD>::n;  // elements per 16-byte output vector
  const int64_t vec_per_row = cols / VE;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= rows * vec_per_row) return;
  const int64_t r = i / vec_per_row, c = (i % vec_per_row) * VE;
  float f[VE];
  if constexpr (sizeof(TS) == 4) {
#pragma unroll
    for (int k = 0; k < VE; k += 4) {
      const uint4 v = ldg_nc_v4(src + r * lds + c + k);
      f[k] = __uint_as_float(v.x); f[k + 1] = __uint_as_float(v.y); f[k + 2] = __uint_as_float(v.z); f[k + 3] = __uint_as_float(v.w);
    }
  } else {
#pragma unroll
    for (int k = 0; k < VE; k += 8) {
      float t[8];
      unpack16(ldg_nc_v4(src + r * lds + c + k), t, TS());
#pragma unroll
      for (int q = 0; q < 8 && k + q < VE; ++q) f[k + q] = t[q];
    }
  }
  stg_v4(dst + r * ldd + c, pack16(f, TD()));
}

template <typename T, bool kVec>
__global__ void gate_kernel(const T* __restrict__ g, int64_t ldg, const T* __restrict__ gate, int64_t ldgate, int64_t M,
                            int64_t N, float scale, T* __restrict__ y, int64_t ldy) {
  constexpr int VE = kVec ? Vec16<T>::n : 1;
  const int64_t vec_per_row = (N + VE - 1) / VE;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M * vec_per_row) return;
  const int64_t m = i / vec_per_row, n = (i % vec_per_row) * VE;
  if constexpr (kVec) {
    float a[VE], b[VE];
    unpack16(ldg_nc_v4(g + m * ldg + n), a, T());
    unpack16(ldg_nc_v4(gate + m * ldgate + n), b, T());
#pragma unroll
    for (int k = 0; k < VE; ++k) a[k] = b[k] > 0.0f ? a[k] * scale : 0.0f;
    stg_v4(y + m * ldy + n, pack16(a, T()));
  } else {
    float v = to_f32(gate[m * ldgate + n]) > 0.0f ? to_f32(g[m * ldg + n]) * scale : 0.0f;
    y[m * ldy + n] = from_f32<T>(v);
  }
}

template <typename T>
static int gate_typed(const void* g, int64_t ldg, const void* gate, int64_t ldgate, int64_t M, int64_t N, float scale,
                      void* y, int64_t ldy, cudaStream_t stream) {
  auto ok = [&](const void* p, int64_t ld) { return aligned(p, 16) && (ld * sizeof(T)) % 16 == 0; };
  const bool vec = ok(g, ldg) && ok(gate, ldgate) && ok(y, ldy) && N % Vec16<T>::n == 0;
  const int VE = vec ? Vec16<T>::n : 1;
  unsigned blocks = (unsigned)ceil_div(M * ceil_div(N, VE), 256);
  if (vec) gate_kernel<T, true><<<blocks, 256, 0, stream>>>((const T*)g, ldg, (const T*)gate, ldgate, M, N, scale, (T*)y, ldy);
  else gate_kernel<T, false><<<blocks, 256, 0, stream>>>((const T*)g, ldg, (const T*)gate, ldgate, M, N, scale, (T*)y, ldy);
  LLP_LAUNCH_OK();
  return 0;
}

}  // namespace llp

using namespace llp;

extern "C" size_t llp_colsum_workspace_bytes(int64_t N) { return colreduce_workspace_bytes(N); }

extern "C" int llp_colsum(int dtype, const void* A, int64_t lda, int64_t M, int64_t N, float* out, int accumulate,
                          void* workspace, void* stream_) {
  LLP_CHECK_ARG(A && out && workspace && M >= 0 && N > 0 && lda >= N);
  if (int rc = check_device()) return rc;
  return colreduce(dtype, A, lda, M, N, nullptr, out, accumulate, reinterpret_cast<float*>(workspace), (cudaStream_t)stream_);
}

extern "C" int llp_cast2d(int src_dtype, int dst_dtype, const void* src, int64_t lds, int64_t rows, int64_t cols,
                          void* dst, int64_t ldd, int transpose, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(src && dst && rows >= 0 && cols >= 0);
  if (int rc = check_device()) return rc;
  if (rows == 0 || cols == 0) return 0;
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32)), block(32, 8);
  {  // fast path: fp32 -> bf16 / bf16 -> fp32 without transposition on 16-byte-aligned rows (cols % 8 == 0)
    const size_t ss = src_dtype == LLP_BF16 ? 2 : 4, ds = dst_dtype == LLP_BF16 ? 2 : 4;
    const bool ok = !transpose && cols % 8 == 0 && aligned(src, 16) && aligned(dst, 16) && (lds * ss) % 16 == 0 &&
                    (ldd * ds) % 16 == 0;
    if (ok && src_dtype == LLP_F32 && dst_dtype == LLP_BF16) {
      cast_rows_kernel<float, __nv_bfloat16><<<(unsigned)ceil_div(rows * (cols / 8), 256), 256, 0, stream>>>((const float*)src, lds, rows, cols, (__nv_bfloat16*)dst, ldd);
      LLP_LAUNCH_OK();
      return 0;
    }
    if (ok && src_dtype == LLP_BF16 && dst_dtype == LLP_F32) {
      cast_rows_kernel<__nv_bfloat16, float><<<(unsigned)ceil_div(rows * (cols / 4), 256), 256, 0, stream>>>((const __nv_bfloat16*)src, lds, rows, cols, (float*)dst, ldd);
      LLP_LAUNCH_OK();
      return 0;
    }
  }
#define LLP_CAST_CASE(SD, DD, TS, TD)                                                                             \
  if (src_dtype == SD && dst_dtype == DD) {                                                                       \
    cast2d_kernel<TS, TD><<<grid, block, 0, stream>>>((const TS*)src, lds, rows, cols, (TD*)dst, ldd, transpose); \
    LLP_LAUNCH_OK();                                                                                              \
    return 0;                                                                                                     \
  }
  LLP_CAST_CASE(LLP_F32, LLP_F32, float, float)
  LLP_CAST_CASE(LLP_F32, LLP_BF16, float, __nv_bfloat16)
  LLP_CAST_CASE(LLP_BF16, LLP_F32, __nv_bfloat16, float)
  LLP_CAST_CASE(LLP_BF16, LLP_BF16, __nv_bfloat16, __nv_bfloat16)
#undef LLP_CAST_CASE
  return LLP_E_BADARG;
}

extern "C" int llp_gate(int dtype, const void* g, int64_t ldg, const void* gate, int64_t ldgate, int64_t M, int64_t N,
                        float scale, void* y, int64_t ldy, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(g && gate && y && M >= 0 && N >= 0);
  if (int rc = check_device()) return rc;
  if (M * N == 0) return 0;
  if (dtype == LLP_F32) return gate_typed<float>(g, ldg, gate, ldgate, M, N, scale, y, ldy, stream);
  if (dtype == LLP_BF16) return gate_typed<__nv_bfloat16>(g, ldg, gate, ldgate, M, N, scale, y, ldy, stream);
  return LLP_E_BADARG;
}

// y = epi(a): the GEMM epilogue (bias, addend, relu, dropout from the same Philox stream) as a stand-alone pass, for
// layers whose pre-activation is not the direct output of a GEMM (SAGEConv_updated: mean-aggregate of lin_l + lin_r).
template <typename T>
__global__ void add_act_kernel(const T* __restrict__ a, int64_t lda, int64_t M, int64_t N, EpilogueParams ep,
                               T* __restrict__ y, int64_t ldy) {
  resolve_rng(ep);
  const int64_t total = M * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / N, n = i - m * N;
    y[m * ldy + n] = from_f32<T>(epilogue_apply<T>(to_f32(a[m * lda + n]), m, n, ep));
  }
}

// 8 bf16 columns per thread (128-bit accesses); one Philox call yields the thread's eight keep decisions
__global__ void add_act_vec8_kernel(const __nv_bfloat16* __restrict__ a, int64_t lda, int64_t M, int64_t N, EpilogueParams ep,
                                    __nv_bfloat16* __restrict__ y, int64_t ldy) {
  resolve_rng(ep);
  const int64_t groups = N >> 3, total = M * groups;
  const __nv_bfloat16* addend = reinterpret_cast<const __nv_bfloat16*>(ep.addend);
  const float keep_scale = ep.dropout_p > 0.0f ? 1.0f / (1.0f - ep.dropout_p) : 1.0f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / groups, n = (i - m * groups) << 3;
    float v[8], w[8];
    unpack16(ldg_nc_v4(a + m * lda + n), v, __nv_bfloat16());
    if (ep.bias) {  // same operation order as epilogue_apply: bias, addend, relu, dropout
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += __ldg(ep.bias + n + k);
    }
    if (addend) {
      unpack16(ldg_nc_v4(addend + m * ep.ldadd + n), w, __nv_bfloat16());
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += w[k];
    }
    if (ep.relu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.0f);
    }
    if (ep.dropout_p > 0.0f) {
      uint32_t keep = 0;
      if (ep.dropout_p == 0.5f) {
        uint4 r = philox4x32_10(ep.seed, (uint64_t)m, ep.offset + (uint64_t)(n >> 7));
        keep = (dropout_word(r, (int)((n >> 5) & 3)) >> (n & 31)) & 0xffu;
      } else {
        uint4 r = philox4x32_10(ep.seed, (uint64_t)m, ep.offset + (uint64_t)(n >> 3));
        const uint32_t thr = dropout_thr16(ep.dropout_p);
#pragma unroll
        for (int k = 0; k < 8; ++k) keep |= (dropout_u16(r, k) >= thr ? 1u : 0u) << k;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = ((keep >> k) & 1u) ? v[k] * keep_scale : 0.0f;
    }
    stg_v4(y + m * ldy + n, pack16(v, __nv_bfloat16()));
  }
}

extern "C" int llp_add_act(int dtype, const void* a, int64_t lda, const void* addend, int64_t ldadd, const float* bias,
                           int64_t M, int64_t N, int relu, float dropout_p, uint64_t seed, uint64_t offset,
                           const uint64_t* rng_state, void* y, int64_t ldy, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(M >= 0 && N >= 0 && dropout_p >= 0.0f && dropout_p < 1.0f);
  if (int rc = check_device()) return rc;
  if (M * N == 0) return 0;
  LLP_CHECK_ARG(a && y && lda >= N && ldy >= N && (addend == nullptr || ldadd >= N));
  EpilogueParams ep{bias, addend, ldadd, nullptr, 0, 1.0f, relu, dropout_p, seed, offset, rng_state};
  unsigned blocks = (unsigned)imin64(ceil_div(M * N, 256), (int64_t)kNumSMs * 16);
  if (dtype == LLP_F32)
    add_act_kernel<float><<<blocks, 256, 0, stream>>>((const float*)a, lda, M, N, ep, (float*)y, ldy);
  else if (dtype == LLP_BF16) {
    const bool vec = N % 8 == 0 && lda % 8 == 0 && ldy % 8 == 0 && aligned(a, 16) && aligned(y, 16) &&
                     (addend == nullptr || (ldadd % 8 == 0 && aligned(addend, 16)));
    if (vec) {
      blocks = (unsigned)imin64(ceil_div(M * (N / 8), 256), (int64_t)kNumSMs * 16);
      add_act_vec8_kernel<<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)a, lda, M, N, ep, (__nv_bfloat16*)y, ldy);
    } else {
      add_act_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)a, lda, M, N, ep, (__nv_bfloat16*)y, ldy);
    }
  } else {
    return LLP_E_BADARG;
  }
  LLP_LAUNCH_OK();
  return 0;
}

extern "C" int llp_weights_prep(int count, const llp_weight_desc* host_descs, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(count >= 0 && (count == 0 || host_descs != nullptr));
  if (int rc = check_device()) return rc;
  for (int base = 0; base < count; base += kMaxPrep) {
    PrepTable t;
    t.count = count - base < kMaxPrep ? count - base : kMaxPrep;
    int tiles = 0;
    for (int i = 0; i < t.count; ++i) {
      const llp_weight_desc& d = host_descs[base + i];
      LLP_CHECK_ARG(d.src && d.rows > 0 && d.cols > 0 && (d.dst == nullptr || d.ld >= d.cols) &&
                    (d.dst_t == nullptr || d.ld_t >= d.rows));
      t.d[i] = d;
      t.tile_begin[i] = tiles;
      tiles += (int)(ceil_div(d.rows, 32) * ceil_div(d.cols, 32));
    }
    t.tile_begin[t.count] = tiles;
    if (tiles == 0) continue;
    weights_prep_kernel<<<(unsigned)tiles, dim3(32, 8), 0, stream>>>(t);
    LLP_LAUNCH_OK();
  }
  return 0;
}
