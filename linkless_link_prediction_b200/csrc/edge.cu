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
  unsigned blocks = (unsigned)ceil_div(M * 32, 256);
  if (dtype == LLP_F32) score_head_kernel<float><<<blocks, 256, 0, stream>>>((const float*)y, ldy, M, H, w, b, prob);
  else if (dtype == LLP_BF16) score_head_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)y, ldy, M, H, w, b, prob);
  else return LLP_E_BADARG;
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
