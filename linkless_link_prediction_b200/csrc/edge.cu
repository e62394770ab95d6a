// Edge scoring kernels: gather-Hadamard z = h[u]*h[v] (train_teacher_gnn.py:58,97; main.py:186,214;
// models.py:140) and the 1-output predictor head + sigmoid (models.py:146,150).  The backward of the gather is in
// edge_bwd.cu.
#include "common.cuh"

namespace llp {

// one warp per edge; 16-byte vectors when rows are 16-byte aligned
template <typename T, bool kVec>
__global__ void __launch_bounds__(256)
edge_hadamard_kernel(const T* __restrict__ h, int64_t ldh, int64_t F, const int64_t* __restrict__ u,
                     const int64_t* __restrict__ v, int64_t M, T* __restrict__ z, int64_t ldz) {
  int lane = threadIdx.x & 31;
  int64_t m = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (m >= M) return;
  const T* hu = h + u[m] * ldh;
  const T* hv = h + v[m] * ldh;
  T* zr = z + m * ldz;
  if constexpr (kVec) {
    constexpr int VE = Vec16<T>::n;
    for (int64_t c = lane * VE; c < F; c += 32 * VE) {
      float a[VE], b[VE];
      unpack16(ldg_v4(hu + c), a, T());
      unpack16(ldg_v4(hv + c), b, T());
#pragma unroll
      for (int i = 0; i < VE; ++i) a[i] *= b[i];
      stg_v4(zr + c, pack16(a, T()));
    }
  } else {
    for (int64_t c = lane; c < F; c += 32) zr[c] = from_f32<T>(to_f32(hu[c]) * to_f32(hv[c]));
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
score_head_kernel(const T* __restrict__ y, int64_t ldy, int64_t M, int64_t H, const float* __restrict__ w,
                  const float* __restrict__ b, float* __restrict__ prob) {
  int lane = threadIdx.x & 31;
  int64_t m = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (m >= M) return;
  const T* yr = y + m * ldy;
  float acc = 0.0f;
  for (int64_t c = lane; c < H; c += 32) acc = fmaf(to_f32(yr[c]), __ldg(w + c), acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    float logit = acc + (b != nullptr ? b[0] : 0.0f);
    prob[m] = 1.0f / (1.0f + expf(-logit));
  }
}

// dlogit = dprob*p*(1-p); gy[m,:] = dlogit*w, optionally masked by the relu/dropout gate of the layer that produced y
// (gate_scale > 0: gy = y > 0 ? dlogit*w*gate_scale : 0)
template <typename T>
__global__ void __launch_bounds__(256)
score_head_bwd_kernel(const T* __restrict__ y, int64_t ldy, int64_t M, int64_t H, const float* __restrict__ w,
                      const float* __restrict__ prob, const float* __restrict__ dprob, float gate_scale,
                      T* __restrict__ gy, int64_t ldgy, float* __restrict__ dlogit) {
  int lane = threadIdx.x & 31;
  int64_t m = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (m >= M) return;
  float p = prob[m];
  float dl = dprob[m] * p * (1.0f - p);
  if (lane == 0) dlogit[m] = dl;
  if (gy != nullptr) {
    for (int64_t c = lane; c < H; c += 32) {
      float v = dl * __ldg(w + c);
      if (gate_scale > 0.0f) v = to_f32(y[m * ldy + c]) > 0.0f ? v * gate_scale : 0.0f;
      gy[m * ldgy + c] = from_f32<T>(v);
    }
  }
}

// ---- vector paths (rows are 16-byte multiples, H <= 4 * 32 * VE): one 128-bit load per lane and row -------------
constexpr int kHeadMaxChunks = 4;   // 16-byte chunks of a row per lane
constexpr int kHeadRows = 4;        // rows in flight per warp

template <typename T, int Q>
__global__ void __launch_bounds__(256)
score_head_vec_kernel(const T* __restrict__ y, int64_t ldy, int64_t M, int H, const float* __restrict__ w,
                      const float* __restrict__ b, float* __restrict__ prob) {
  constexpr int VE = Vec16<T>::n;
  const int lane = threadIdx.x & 31;
  const int nchunk = H / VE;
  float wr[Q][VE];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int ch = lane + 32 * q;
#pragma unroll
    for (int i = 0; i < VE; ++i) wr[q][i] = ch < nchunk ? __ldg(w + ch * VE + i) : 0.0f;
  }
  const float bias = b != nullptr ? b[0] : 0.0f;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t m0 = warp * kHeadRows; m0 < M; m0 += nwarps * kHeadRows) {
    float acc[kHeadRows];
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) acc[r] = 0.0f;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int ch = lane + 32 * q;
      if (q * 32 >= nchunk) break;
      uint4 v[kHeadRows];
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r)
        v[r] = (ch < nchunk && m0 + r < M) ? ldg_nc_v4(y + (m0 + r) * ldy + ch * VE) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r) {
        float f[VE];
        unpack16(v[r], f, T());
#pragma unroll
        for (int i = 0; i < VE; ++i) acc[r] = fmaf(f[i], wr[q][i], acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) acc[r] = warp_sum(acc[r]);
    if (lane < kHeadRows && m0 + lane < M) {
      float a = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
      prob[m0 + lane] = 1.0f / (1.0f + expf(-(a + bias)));
    }
  }
}

// One pass over y: gy = dlogit*w (gated), partial gw = sum_m dlogit*y[m,:], partial gb = sum_m dlogit.  The grid is a
// fixed function of M, rows go to warps in a fixed pattern and the 8 warps of a block are combined in warp order, so the
// partials (and the final sums over blocks) are bit-reproducible.
template <typename T, int Q>
__global__ void __launch_bounds__(256)
score_head_bwd_vec_kernel(const T* __restrict__ y, int64_t ldy, int64_t M, int H, const float* __restrict__ w,
                          const float* __restrict__ prob, const float* __restrict__ dprob, float gate_scale,
                          T* __restrict__ gy, int64_t ldgy, float* __restrict__ partial_gw, float* __restrict__ partial_gb) {
  constexpr int VE = Vec16<T>::n;
  extern __shared__ float red[];  // [8][H] + [8]
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const int nchunk = H / VE;
  float wr[Q][VE], gw[Q][VE];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int ch = lane + 32 * q;
#pragma unroll
    for (int i = 0; i < VE; ++i) {
      wr[q][i] = ch < nchunk ? __ldg(w + ch * VE + i) : 0.0f;
      gw[q][i] = 0.0f;
    }
  }
  float gb = 0.0f;
  const int64_t warp = (int64_t)blockIdx.x * 8 + wv, nwarps = (int64_t)gridDim.x * 8;
  for (int64_t m0 = warp * kHeadRows; m0 < M; m0 += nwarps * kHeadRows) {
    float dl[kHeadRows];
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) {
      const float pr = m0 + r < M ? __ldg(prob + m0 + r) : 0.0f;
      dl[r] = m0 + r < M ? __ldg(dprob + m0 + r) * pr * (1.0f - pr) : 0.0f;
      gb += dl[r];
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int ch = lane + 32 * q;
      if (q * 32 >= nchunk) break;
      uint4 v[kHeadRows];
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r)
        v[r] = (ch < nchunk && m0 + r < M) ? ldg_nc_v4(y + (m0 + r) * ldy + ch * VE) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r) {
        float f[VE], o[VE];
        unpack16(v[r], f, T());
#pragma unroll
        for (int i = 0; i < VE; ++i) {
          gw[q][i] = fmaf(dl[r], f[i], gw[q][i]);
          const float t = dl[r] * wr[q][i];
          o[i] = gate_scale > 0.0f ? (f[i] > 0.0f ? t * gate_scale : 0.0f) : t;
        }
        if (gy != nullptr && ch < nchunk && m0 + r < M) stg_v4(gy + (m0 + r) * ldgy + ch * VE, pack16(o, T()));
      }
    }
  }
  // block combine in warp order
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int ch = lane + 32 * q;
    if (ch < nchunk) {
#pragma unroll
      for (int i = 0; i < VE; ++i) red[wv * H + ch * VE + i] = gw[q][i];
    }
  }
  if (lane == 0) red[8 * H + wv] = gb;   // every lane carries the same gb
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += 256) {
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k * H + c];
    partial_gw[(int64_t)blockIdx.x * H + c] = t;
  }
  if (threadIdx.x == 0) {
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[8 * H + k];
    partial_gb[blockIdx.x] = t;
  }
}

// out[n] = sum_i partial[i][n]: one warp per column, lanes stride over the partials, shuffle tree (fixed order)
__global__ void head_final_sum_kernel(const float* __restrict__ partial, int count, int64_t N, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (n >= N) return;
  float s = 0.0f;
  for (int i = lane; i < count; i += 32) s += partial[(int64_t)i * N + n];
  s = warp_sum(s);
  if (lane == 0) out[n] = s;
}

static int head_bwd_blocks(int64_t M) { return (int)imin64(kNumSMs * 8, imax64(1, ceil_div(M, 8 * kHeadRows))); }

template <typename T>
static bool rows_vec_ok(const void* p, int64_t ld, int64_t F) {
  return aligned(p, 16) && (ld * sizeof(T)) % 16 == 0 && F % Vec16<T>::n == 0;
}

template <typename T>
static int hadamard_launch(const void* h, int64_t ldh, int64_t F, const int64_t* u, const int64_t* v, int64_t M, void* z,
                           int64_t ldz, cudaStream_t stream) {
  unsigned blocks = (unsigned)ceil_div(M * 32, 256);
  if (rows_vec_ok<T>(h, ldh, F) && rows_vec_ok<T>(z, ldz, F))
    edge_hadamard_kernel<T, true><<<blocks, 256, 0, stream>>>((const T*)h, ldh, F, u, v, M, (T*)z, ldz);
  else
    edge_hadamard_kernel<T, false><<<blocks, 256, 0, stream>>>((const T*)h, ldh, F, u, v, M, (T*)z, ldz);
  LLP_LAUNCH_OK();
  return 0;
}

int colreduce(int dtype, const void* A, int64_t lda, int64_t M, int64_t N, const float* w, float* out, int accumulate,
              float* partial, cudaStream_t stream);
size_t colreduce_workspace_bytes(int64_t N);

}  // namespace llp

using namespace llp;

extern "C" int llp_edge_hadamard(int dtype, const void* h, int64_t ldh, int64_t F, const int64_t* u, const int64_t* v,
                                 int64_t M, void* z, int64_t ldz, void* stream_) {
  LLP_CHECK_ARG(F > 0 && M >= 0);
  if (int rc = check_device()) return rc;
  if (M == 0) return 0;
  LLP_CHECK_ARG(h && u && v && z && ldh >= F && ldz >= F);
  if (dtype == LLP_F32) return hadamard_launch<float>(h, ldh, F, u, v, M, z, ldz, (cudaStream_t)stream_);
  if (dtype == LLP_BF16) return hadamard_launch<__nv_bfloat16>(h, ldh, F, u, v, M, z, ldz, (cudaStream_t)stream_);
  return LLP_E_BADARG;
}

extern "C" int llp_score_head(int dtype, const void* y, int64_t ldy, int64_t M, int64_t H, const float* w, const float* b,
                              float* prob, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(y && w && prob && M >= 0 && H > 0 && ldy >= H);
  if (int rc = check_device()) return rc;
  if (M == 0) return 0;
  if (dtype != LLP_F32 && dtype != LLP_BF16) return LLP_E_BADARG;
  const size_t es = dtype == LLP_BF16 ? 2 : 4;
  const int ve = (int)(16 / es);
  if (aligned(y, 16) && (ldy * es) % 16 == 0 && H % ve == 0 && H <= kHeadMaxChunks * 32 * ve) {
    const unsigned vblocks = (unsigned)imin64(kNumSMs * 8, imax64(1, ceil_div(M, 8 * kHeadRows)));
    const int q = H <= 32 * ve ? 1 : (H <= 64 * ve ? 2 : 4);
#define LLP_HEAD_FWD(T, QQ) score_head_vec_kernel<T, QQ><<<vblocks, 256, 0, stream>>>((const T*)y, ldy, M, (int)H, w, b, prob)
    if (dtype == LLP_F32) { if (q == 1) LLP_HEAD_FWD(float, 1); else if (q == 2) LLP_HEAD_FWD(float, 2); else LLP_HEAD_FWD(float, 4); }
    else { if (q == 1) LLP_HEAD_FWD(__nv_bfloat16, 1); else if (q == 2) LLP_HEAD_FWD(__nv_bfloat16, 2); else LLP_HEAD_FWD(__nv_bfloat16, 4); }
#undef LLP_HEAD_FWD
    LLP_LAUNCH_OK();
    return 0;
  }
  unsigned blocks = (unsigned)ceil_div(M * 32, 256);
  if (dtype == LLP_F32) score_head_kernel<float><<<blocks, 256, 0, stream>>>((const float*)y, ldy, M, H, w, b, prob);
  else score_head_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)y, ldy, M, H, w, b, prob);
  LLP_LAUNCH_OK();
  return 0;
}

extern "C" size_t llp_score_head_bwd_workspace_bytes(int64_t M, int64_t H) {
  return 8192 + (size_t)(M > 0 ? M : 1) * sizeof(float) + colreduce_workspace_bytes(H);
}

extern "C" int llp_score_head_bwd(int dtype, const void* y, int64_t ldy, int64_t M, int64_t H, const float* w,
                                  const float* prob, const float* dprob, float gate_scale, void* gy, int64_t ldgy,
                                  float* gw, float* gb, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(y && w && prob && dprob && workspace && M > 0 && H > 0 && ldy >= H);
  if (workspace_bytes < llp_score_head_bwd_workspace_bytes(M, H)) return LLP_E_WORKSPACE;
  if (int rc = check_device()) return rc;
  float* dlogit = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 8192);
  float* partial = dlogit + M;
  if (dtype != LLP_F32 && dtype != LLP_BF16) return LLP_E_BADARG;
  {
    const size_t es = dtype == LLP_BF16 ? 2 : 4;
    const int ve = (int)(16 / es);
    const bool vec = aligned(y, 16) && (ldy * es) % 16 == 0 && H % ve == 0 && H <= kHeadMaxChunks * 32 * ve &&
                     (gy == nullptr || (aligned(gy, 16) && (ldgy * es) % 16 == 0)) && (size_t)(8 * H + 8) * 4 <= 48 * 1024;
    if (vec) {
      // partial_gb lives where the scalar path keeps dlogit (blocks <= M), partial_gw in the colreduce scratch
      const int nb = head_bwd_blocks(M);
      const size_t smem = (size_t)(8 * H + 8) * sizeof(float);
      const int q = H <= 32 * ve ? 1 : (H <= 64 * ve ? 2 : 4);
#define LLP_HEAD_BWD(T, QQ)                                                                                              \
  score_head_bwd_vec_kernel<T, QQ><<<nb, 256, smem, stream>>>((const T*)y, ldy, M, (int)H, w, prob, dprob, gate_scale, (T*)gy, \
                                                              ldgy, partial, dlogit)
      if (dtype == LLP_F32) { if (q == 1) LLP_HEAD_BWD(float, 1); else if (q == 2) LLP_HEAD_BWD(float, 2); else LLP_HEAD_BWD(float, 4); }
      else { if (q == 1) LLP_HEAD_BWD(__nv_bfloat16, 1); else if (q == 2) LLP_HEAD_BWD(__nv_bfloat16, 2); else LLP_HEAD_BWD(__nv_bfloat16, 4); }
#undef LLP_HEAD_BWD
      LLP_LAUNCH_OK();
      if (gw) {
        head_final_sum_kernel<<<(unsigned)ceil_div(H * 32, 256), 256, 0, stream>>>(partial, nb, H, gw);
        LLP_LAUNCH_OK();
      }
      if (gb) {
        head_final_sum_kernel<<<1, 32, 0, stream>>>(dlogit, nb, 1, gb);
        LLP_LAUNCH_OK();
      }
      return 0;
    }
  }
  unsigned blocks = (unsigned)ceil_div(M * 32, 256);
  if (dtype == LLP_F32)
    score_head_bwd_kernel<float><<<blocks, 256, 0, stream>>>((const float*)y, ldy, M, H, w, prob, dprob, gate_scale, (float*)gy, ldgy, dlogit);
  else if (dtype == LLP_BF16)
    score_head_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)y, ldy, M, H, w, prob, dprob, gate_scale, (__nv_bfloat16*)gy, ldgy, dlogit);
  else
    return LLP_E_BADARG;
  LLP_LAUNCH_OK();
  if (gw)
    if (int rc = colreduce(dtype, y, ldy, M, H, dlogit, gw, 0, partial, stream)) return rc;  // gw = sum_m dlogit[m]*y[m,:]
  if (gb) return sum_f32(dlogit, M, 1.0f, gb, workspace, stream);
  return 0;
}
