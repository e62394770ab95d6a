// GEMM entry points: backend selection between the tcgen05 tensor-core kernels (bf16 operands) and the
// CUDA-core fp32 kernels (fp32-parity mode).  There is no library (cuBLAS) call and no CPU path.
#include "common.cuh"

namespace llp {
int gemm_nt_simt(const llp_gemm_nt_args& a, cudaStream_t stream);
int gemm_nt_tcgen05(const llp_gemm_nt_args& a, cudaStream_t stream);
int gemm_tn_simt(int dtype, int64_t M, int64_t N1, int64_t N2, const void* A, int64_t lda, const void* B, int64_t ldb,
                 float* D, int64_t ldd, int accumulate, float* ws, cudaStream_t stream);
int gemm_tn_tcgen05(int64_t M, int64_t N1, int64_t N2, const void* A, int64_t lda, const void* B, int64_t ldb, float* D,
                    int64_t ldd, int accumulate, float* ws, cudaStream_t stream);
int tn_splits(int64_t M, int64_t N1, int64_t N2);
void tn_split_plan(int64_t M, int64_t N1, int64_t N2, int* splits, int64_t* k_per_split);

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
    backend = ok ? LLP_GEMM_TCGEN05 : LLP_GEMM_SIMT;
  }
  if (backend == LLP_GEMM_TCGEN05) return gemm_nt_tcgen05(*a, stream);
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
  return (size_t)(s + 1) * (size_t)N1 * (size_t)N2 * sizeof(float);
}

extern "C" int llp_gemm_tn(int dtype, int backend, int64_t M, int64_t N1, int64_t N2, const void* A, int64_t lda,
                           const void* B, int64_t ldb, float* D, int64_t ldd, int accumulate, void* workspace,
                           size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(A && B && D && workspace && M > 0 && N1 > 0 && N2 > 0 && lda >= N1 && ldb >= N2 && ldd >= N2);
  if (workspace_bytes < llp_gemm_tn_workspace_bytes(M, N1, N2)) return LLP_E_WORKSPACE;
  if (int rc = check_device()) return rc;
  if (backend == LLP_GEMM_AUTO)
    backend = (dtype == LLP_BF16 && tma_ok(A, lda) && tma_ok(B, ldb)) ? LLP_GEMM_TCGEN05 : LLP_GEMM_SIMT;
  if (backend == LLP_GEMM_TCGEN05) {
    if (dtype != LLP_BF16) return LLP_E_SHAPE;
    return gemm_tn_tcgen05(M, N1, N2, A, lda, B, ldb, D, ldd, accumulate, reinterpret_cast<float*>(workspace), stream);
  }
  if (backend == LLP_GEMM_SIMT)
    return gemm_tn_simt(dtype, M, N1, N2, A, lda, B, ldb, D, ldd, accumulate, reinterpret_cast<float*>(workspace), stream);
  return LLP_E_BADARG;
}
