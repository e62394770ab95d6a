// GEMM entry points: backend selection between the tcgen05 tensor-core kernels (bf16 operands: kind::f16; fp32 operands:
// 3xTF32 split accumulation, kind::tf32) and the CUDA-core fp32 kernels (operands TMA cannot address).  There is no
// library (cuBLAS) call and no CPU path.
#include "common.cuh"

namespace llp {
int gemm_nt_simt(const llp_gemm_nt_args& a, cudaStream_t stream);
int gemm_nt_tcgen05(const llp_gemm_nt_args& a, cudaStream_t stream);
int gemm_tn_simt(int dtype, int64_t M, int64_t N1, int64_t N2, const void* A, int64_t lda, const void* B, int64_t ldb,
                 float* D, int64_t ldd, int accumulate, float* ws, cudaStream_t stream);
int gemm_tn_tcgen05(int64_t M, int64_t N1, int64_t N2, const void* A, int64_t lda, const void* B, int64_t ldb, float* D,
                    int64_t ldd, int accumulate, float* ws, cudaStream_t stream);
int gemm_nt_tf32(const llp_gemm_nt_args& a, cudaStream_t stream);
int gemm_tn_tf32(int64_t M, int64_t N1, int64_t N2, const void* A, int64_t lda, const void* B, int64_t ldb, float* D,
                 int64_t ldd, int accumulate, float* ws, cudaStream_t stream);
int gemm_tn_tf32_splits(int64_t M, int64_t N1, int64_t N2);
bool tf32_operand_ok(const void* p, int64_t ld);
int tn_splits(int64_t M, int64_t N1, int64_t N2);
void tn_split_plan(int64_t M, int64_t N1, int64_t N2, int* splits, int64_t* k_per_split);

bool wgrad_tcgen05_supported(int64_t M, int64_t N1, int64_t n2a, int64_t n2b);
size_t wgrad_tcgen05_workspace_bytes(int64_t M, int64_t N1, int64_t n2a, int64_t n2b);
int wgrad_tcgen05(int64_t M, int64_t N1, const void* G, int64_t ldg, int64_t n2a, const void* A, int64_t lda, float* dWa,
                  int64_t ldwa, int64_t n2b, const void* B, int64_t ldb, float* dWb, int64_t ldwb, float* dbias,
                  int accumulate, float* ws, cudaStream_t stream);
size_t colreduce_workspace_bytes(int64_t N);
int colreduce(int dtype, const void* A, int64_t lda, int64_t M, int64_t N, const float* w, float* out, int accumulate,
              float* partial, cudaStream_t stream);

static bool tma_ok(const void* p, int64_t ld) { return p != nullptr && aligned(p, 16) && (ld * 2) % 16 == 0; }
}  // namespace llp

using namespace llp;

extern "C" int llp_gemm_nt(const llp_gemm_nt_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(a != nullptr);
  LLP_CHECK_ARG(a->M >= 0 && a->N > 0 && a->K1 > 0 && a->K2 >= 0);
  LLP_CHECK_ARG(a->A1 && a->B1 && a->D && a->lda1 >= a->K1 && a->ldb1 >= a->K1 && a->ldd >= a->N);
  LLP_CHECK_ARG(a->K2 == 0 || a->A2 == nullptr || (a->B2 && a->lda2 >= a->K2 && a->ldb2 >= a->K2));
  LLP_CHECK_ARG(a->dropout_p >= 0.0f && a->dropout_p < 1.0f);
  if (int rc = check_device()) return rc;
  if (a->M == 0) return 0;
  int backend = a->backend;
  const bool dual = a->A2 != nullptr && a->K2 > 0;
  if (backend == LLP_GEMM_AUTO) {
    bool ok = a->dtype == LLP_BF16 && tma_ok(a->A1, a->lda1) && tma_ok(a->B1, a->ldb1) &&
              (!dual || (tma_ok(a->A2, a->lda2) && tma_ok(a->B2, a->ldb2)));
    bool ok32 = a->dtype == LLP_F32 && a->out_dtype == LLP_F32 && tf32_operand_ok(a->A1, a->lda1) &&
                tf32_operand_ok(a->B1, a->ldb1) && (!dual || (tf32_operand_ok(a->A2, a->lda2) && tf32_operand_ok(a->B2, a->ldb2)));
    if (g_tuning[21]) ok32 = false;   // A/B knob: fp32 operands on the CUDA cores (same result up to fp32 round-off)
    backend = ok ? LLP_GEMM_TCGEN05 : (ok32 ? LLP_GEMM_TF32X3 : LLP_GEMM_SIMT);
  }
  if (backend == LLP_GEMM_TCGEN05) return gemm_nt_tcgen05(*a, stream);
  if (backend == LLP_GEMM_TF32X3) return gemm_nt_tf32(*a, stream);
  if (backend == LLP_GEMM_SIMT) return gemm_nt_simt(*a, stream);
  return LLP_E_BADARG;
}

extern "C" size_t llp_gemm_tn_workspace_bytes(int64_t M, int64_t N1, int64_t N2) {
  if (M <= 0 || N1 <= 0 || N2 <= 0) return 256;
  int s_tc = 1;
  int64_t per = 0;
  tn_split_plan(M, N1, N2, &s_tc, &per);
  int s_simt = tn_splits(M, N1, N2);
  int s = s_tc > s_simt ? s_tc : s_simt;
  const int s_tf = gemm_tn_tf32_splits(M, N1, N2);
  if (s_tf > s) s = s_tf;
  return (size_t)(s + 1) * (size_t)N1 * (size_t)N2 * sizeof(float);
}

extern "C" int llp_gemm_tn(int dtype, int backend, int64_t M, int64_t N1, int64_t N2, const void* A, int64_t lda,
                           const void* B, int64_t ldb, float* D, int64_t ldd, int accumulate, void* workspace,
                           size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(A && B && D && workspace && M > 0 && N1 > 0 && N2 > 0 && lda >= N1 && ldb >= N2 && ldd >= N2);
  if (workspace_bytes < llp_gemm_tn_workspace_bytes(M, N1, N2)) return LLP_E_WORKSPACE;
  if (int rc = check_device()) return rc;
  if (backend == LLP_GEMM_AUTO) {
    backend = (dtype == LLP_BF16 && tma_ok(A, lda) && tma_ok(B, ldb)) ? LLP_GEMM_TCGEN05 : LLP_GEMM_SIMT;
    if (dtype == LLP_F32 && tf32_operand_ok(A, lda) && tf32_operand_ok(B, ldb) && !g_tuning[21]) backend = LLP_GEMM_TF32X3;
  }
  if (backend == LLP_GEMM_TCGEN05) {
    if (dtype != LLP_BF16) return LLP_E_SHAPE;
    return gemm_tn_tcgen05(M, N1, N2, A, lda, B, ldb, D, ldd, accumulate, reinterpret_cast<float*>(workspace), stream);
  }
  if (backend == LLP_GEMM_TF32X3) {
    if (dtype != LLP_F32) return LLP_E_SHAPE;
    return gemm_tn_tf32(M, N1, N2, A, lda, B, ldb, D, ldd, accumulate, reinterpret_cast<float*>(workspace), stream);
  }
  if (backend == LLP_GEMM_SIMT)
    return gemm_tn_simt(dtype, M, N1, N2, A, lda, B, ldb, D, ldd, accumulate, reinterpret_cast<float*>(workspace), stream);
  return LLP_E_BADARG;
}

// ---- fused weight gradient of one layer ------------------------------------------------------------
extern "C" size_t llp_wgrad_workspace_bytes(int64_t M, int64_t N1, int64_t N2a, int64_t N2b) {
  if (M <= 0 || N1 <= 0 || N2a <= 0 || N2b < 0) return 256;
  size_t fused = wgrad_tcgen05_supported(M, N1, N2a, N2b) ? wgrad_tcgen05_workspace_bytes(M, N1, N2a, N2b) : 0;
  size_t sep = llp_gemm_tn_workspace_bytes(M, N1, N2a);
  if (N2b > 0) {
    size_t t = llp_gemm_tn_workspace_bytes(M, N1, N2b);
    sep = t > sep ? t : sep;
  }
  size_t cs = colreduce_workspace_bytes(N1);
  sep = cs > sep ? cs : sep;
  return fused > sep ? fused : sep;
}

extern "C" int llp_wgrad(int dtype, int backend, int64_t M, int64_t N1, const void* G, int64_t ldg, int64_t N2a,
                         const void* A, int64_t lda, float* dWa, int64_t ldwa, int64_t N2b, const void* B, int64_t ldb,
                         float* dWb, int64_t ldwb, float* dbias, int accumulate, void* workspace, size_t workspace_bytes,
                         void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(G && A && dWa && workspace && M > 0 && N1 > 0 && N2a > 0 && N2b >= 0);
  LLP_CHECK_ARG(ldg >= N1 && lda >= N2a && ldwa >= N2a);
  LLP_CHECK_ARG(N2b == 0 || (B && dWb && ldb >= N2b && ldwb >= N2b));
  if (workspace_bytes < llp_wgrad_workspace_bytes(M, N1, N2a, N2b)) return LLP_E_WORKSPACE;
  if (int rc = check_device()) return rc;
  float* ws = reinterpret_cast<float*>(workspace);
  const bool auto_backend = backend == LLP_GEMM_AUTO;
  if (backend == LLP_GEMM_AUTO) {
    const bool ok = dtype == LLP_BF16 && wgrad_tcgen05_supported(M, N1, N2a, N2b) && tma_ok(G, ldg) && tma_ok(A, lda) &&
                    (N2b == 0 || tma_ok(B, ldb));
    backend = ok ? LLP_GEMM_TCGEN05 : LLP_GEMM_SIMT;
  }
  if (backend == LLP_GEMM_TCGEN05) {
    if (dtype != LLP_BF16 || !wgrad_tcgen05_supported(M, N1, N2a, N2b)) return LLP_E_SHAPE;
    return wgrad_tcgen05(M, N1, G, ldg, N2a, A, lda, dWa, ldwa, N2b, B, ldb, dWb, ldwb, dbias, accumulate, ws, stream);
  }
  if (backend != LLP_GEMM_SIMT && backend != LLP_GEMM_TF32X3) return LLP_E_BADARG;
  // separate launches: fp32-parity mode (3xTF32 tensor-core GEMMs; CUDA cores only for operands TMA cannot address or when
  // LLP_GEMM_SIMT was asked for explicitly) and shapes the fused bf16 kernel does not take (N2 > 256)
  const bool tc = dtype == LLP_BF16 && tma_ok(G, ldg);
  const bool tf = dtype == LLP_F32 && tf32_operand_ok(G, ldg) && !(backend == LLP_GEMM_SIMT && !auto_backend) &&
                  !(auto_backend && g_tuning[21]);
  auto one = [&](const void* X, int64_t ldx, int64_t n2, float* dW, int64_t ldw) -> int {
    if (tc && tma_ok(X, ldx)) return gemm_tn_tcgen05(M, N1, n2, G, ldg, X, ldx, dW, ldw, accumulate, ws, stream);
    if (tf && tf32_operand_ok(X, ldx)) return gemm_tn_tf32(M, N1, n2, G, ldg, X, ldx, dW, ldw, accumulate, ws, stream);
    return gemm_tn_simt(dtype, M, N1, n2, G, ldg, X, ldx, dW, ldw, accumulate, ws, stream);
  };
  if (int rc = one(A, lda, N2a, dWa, ldwa)) return rc;
  if (N2b > 0)
    if (int rc = one(B, ldb, N2b, dWb, ldwb)) return rc;
  if (dbias != nullptr) return colreduce(dtype, G, ldg, M, N1, nullptr, dbias, accumulate, ws, stream);
  return 0;
}
