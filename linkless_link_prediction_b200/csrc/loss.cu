// Fused loss kernels: BCE (train_teacher_gnn.py:33,59), LLP_D = softmax/KL over anchor-context
// score rows (main.py:27-31,188), LLP_R = pairwise margin-rank loss (main.py:190-203).
// Every kernel emits the loss terms AND the gradient w.r.t. the student input in one pass; the scalar
// loss is reduced deterministically (fixed-order tree, double accumulation) — no float atomics.
#include "common.cuh"

namespace llp {

constexpr int kSumBlocks = 256, kSumThreads = 256;

__global__ void sum_stage1_kernel(const float* __restrict__ in, int64_t n, double* __restrict__ partial) {
  __shared__ double red[kSumThreads];
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kSumThreads + threadIdx.x; i < n; i += (int64_t)kSumBlocks * kSumThreads)
    acc += (double)in[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = kSumThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void sum_stage2_kernel(const double* __restrict__ partial, float scale, float* __restrict__ out) {
  __shared__ double red[kSumBlocks];
  red[threadIdx.x] = partial[threadIdx.x];
  __syncthreads();
  for (int s = kSumBlocks / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(red[0] * (double)scale);
}

int sum_f32(const float* in, int64_t n, float scale, float* out, void* workspace, cudaStream_t stream) {
  double* partial = reinterpret_cast<double*>(workspace);
  sum_stage1_kernel<<<kSumBlocks, kSumThreads, 0, stream>>>(in, n, partial);
  LLP_LAUNCH_OK();
  sum_stage2_kernel<<<1, kSumBlocks, 0, stream>>>(partial, scale, out);
  LLP_LAUNCH_OK();
  return 0;
}

constexpr size_t kSumWsBytes = 8192;

__global__ void bce_kernel(const float* __restrict__ p, int64_t n, int64_t n_pos, float inv_n, float* __restrict__ terms,
                           float* __restrict__ dprob) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float pi = p[i];
  float y = i < n_pos ? 1.0f : 0.0f;
  // ATen binary_cross_entropy: (y-1)*max(log1p(-p),-100) - y*max(log(p),-100)
  float lp = fmaxf(logf(pi), -100.0f), l1p = fmaxf(log1pf(-pi), -100.0f);
  terms[i] = (y - 1.0f) * l1p - y * lp;
  if (dprob != nullptr) dprob[i] = (pi - y) / fmaxf((1.0f - pi) * pi, 1e-12f) * inv_n;
}

// One warp per row.  loss_row = sum_j y_t[j] * (log y_t[j] - log_softmax(s/T)[j]);  ds = (softmax(s/T) - y_t) * T / rows
__global__ void kd_d_kernel(const float* __restrict__ s, const float* __restrict__ t, int64_t rows, int64_t K, float invT,
                            float ds_scale, float* __restrict__ row_loss, float* __restrict__ ds) {
  int lane = threadIdx.x & 31;
  int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const float* sr = s + r * K;
  const float* tr = t + r * K;
  float ms = -INFINITY, mt = -INFINITY;
  for (int64_t j = lane; j < K; j += 32) {
    ms = fmaxf(ms, sr[j] * invT);
    mt = fmaxf(mt, tr[j] * invT);
  }
  ms = warp_max(ms);
  mt = warp_max(mt);
  float zs = 0.0f, zt = 0.0f;
  for (int64_t j = lane; j < K; j += 32) {
    zs += expf(sr[j] * invT - ms);
    zt += expf(tr[j] * invT - mt);
  }
  zs = warp_sum(zs);
  zt = warp_sum(zt);
  float lzs = logf(zs), lzt = logf(zt);
  float acc = 0.0f;
  for (int64_t j = lane; j < K; j += 32) {
    float ls = sr[j] * invT - ms - lzs;  // log_softmax(s/T)
    float lt = tr[j] * invT - mt - lzt;
    float yt = expf(lt);
    acc += yt > 0.0f ? yt * (lt - ls) : 0.0f;  // xlogy convention of F.kl_div
    if (ds != nullptr) ds[r * K + j] = (expf(ls) - yt) * ds_scale;
  }
  acc = warp_sum(acc);
  if (lane == 0) row_loss[r] = acc;
}

// One warp per row; lane i owns score i, i+32, ... and visits every partner j (K^2 work, K <= 1024, no atomics).
constexpr int kMaxRankK = 1024;
__global__ void kd_r_kernel(const float* __restrict__ s, const float* __restrict__ t, int64_t rows, int K, float margin,
                            float ds_scale, float* __restrict__ row_loss, float* __restrict__ ds) {
  extern __shared__ float sm[];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  float* ss = sm + (size_t)w * 2 * K;
  float* ts = ss + K;
  if (r < rows) {
    for (int j = lane; j < K; j += 32) {
      ss[j] = s[r * K + j];
      ts[j] = t[r * K + j];
    }
  }
  __syncwarp();
  if (r >= rows) return;
  float loss = 0.0f;
  for (int i = lane; i < K; i += 32) {
    float si = ss[i], ti = ts[i], g = 0.0f;
    for (int j = 0; j < K; ++j) {
      if (j == i) continue;
      float sj = ss[j], tj = ts[j];
      // the pair is (a,b) = (min, max); y from the teacher, hinge on the student
      float sa = i < j ? si : sj, sb = i < j ? sj : si;
      float ta = i < j ? ti : tj, tb = i < j ? tj : ti;
      float y = ta > tb + margin ? 1.0f : (ta < tb - margin ? -1.0f : 0.0f);
      float term = -y * (sa - sb) + margin;
      if (term >= 0.0f) {
        if (i < j) { loss += term; g -= y; } else { g += y; }
      }
    }
    if (ds != nullptr) ds[r * K + i] = g * ds_scale;
  }
  loss = warp_sum(loss);
  if (lane == 0) row_loss[r] = loss;
}

// LLP_D and LLP_R of the same score rows in ONE pass (main.py:188 and :190-203 read the same s_r / t_r): the row is
// staged in shared memory once, one warp computes the softmax/KL term, the C(K,2) margin-rank term and the gradient of
// w_d * LLP_D + w_r * LLP_R.  Per-term arithmetic is the same as kd_d_kernel / kd_r_kernel (same operation order).
__global__ void kd_fused_kernel(const float* __restrict__ s, const float* __restrict__ t, int64_t rows, int K, float invT,
                                float margin, float dsd_scale, float dsr_scale, float* __restrict__ row_loss_d,
                                float* __restrict__ row_loss_r, float* __restrict__ ds) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  float* ss = sm + (size_t)w * 2 * K;
  float* ts = ss + K;
  if (r < rows) {
    for (int j = lane; j < K; j += 32) {
      ss[j] = s[r * K + j];
      ts[j] = t[r * K + j];
    }
  }
  __syncwarp();
  if (r >= rows) return;
  // ---- LLP_D: KL(softmax(t/T) || softmax(s/T)) ----
  float ms = -INFINITY, mt = -INFINITY;
  for (int j = lane; j < K; j += 32) {
    ms = fmaxf(ms, ss[j] * invT);
    mt = fmaxf(mt, ts[j] * invT);
  }
  ms = warp_max(ms);
  mt = warp_max(mt);
  float zs = 0.0f, zt = 0.0f;
  for (int j = lane; j < K; j += 32) {
    zs += expf(ss[j] * invT - ms);
    zt += expf(ts[j] * invT - mt);
  }
  zs = warp_sum(zs);
  zt = warp_sum(zt);
  const float lzs = logf(zs), lzt = logf(zt);
  float acc_d = 0.0f, acc_r = 0.0f;
  for (int i = lane; i < K; i += 32) {
    const float si = ss[i], ti = ts[i];
    const float ls = si * invT - ms - lzs;
    const float lt = ti * invT - mt - lzt;
    const float yt = expf(lt);
    acc_d += yt > 0.0f ? yt * (lt - ls) : 0.0f;
    const float gd = (expf(ls) - yt) * dsd_scale;
    // ---- LLP_R: every partner j of score i ----
    float g = 0.0f;
    for (int j = 0; j < K; ++j) {
      if (j == i) continue;
      const float sj = ss[j], tj = ts[j];
      const float sa = i < j ? si : sj, sb = i < j ? sj : si;
      const float ta = i < j ? ti : tj, tb = i < j ? tj : ti;
      const float y = ta > tb + margin ? 1.0f : (ta < tb - margin ? -1.0f : 0.0f);
      const float term = -y * (sa - sb) + margin;
      if (term >= 0.0f) {
        if (i < j) { acc_r += term; g -= y; } else { g += y; }
      }
    }
    if (ds != nullptr) ds[r * K + i] = gd + g * dsr_scale;
  }
  acc_d = warp_sum(acc_d);
  acc_r = warp_sum(acc_r);
  if (lane == 0) { row_loss_d[r] = acc_d; row_loss_r[r] = acc_r; }
}

// out[0] = scale_a * sum(a), out[1] = scale_b * sum(b), out[2] = w_a * out[0] + w_b * out[1]: two deterministic sums in
// the launches of one (same fixed-order tree as sum_f32)
__global__ void sum2_stage1_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, double* __restrict__ partial) {
  __shared__ double red[kSumThreads];
  const float* in = blockIdx.y == 0 ? a : b;
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kSumThreads + threadIdx.x; i < n; i += (int64_t)kSumBlocks * kSumThreads)
    acc += (double)in[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int st = kSumThreads / 2; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.y * kSumBlocks + blockIdx.x] = red[0];
}
__global__ void sum2_stage2_kernel(const double* __restrict__ partial, float scale_a, float scale_b, float w_a, float w_b,
                                   float* __restrict__ out) {
  __shared__ double red[2][kSumBlocks];
  red[0][threadIdx.x] = partial[threadIdx.x];
  red[1][threadIdx.x] = partial[kSumBlocks + threadIdx.x];
  __syncthreads();
  for (int st = kSumBlocks / 2; st > 0; st >>= 1) {
    if (threadIdx.x < st) { red[0][threadIdx.x] += red[0][threadIdx.x + st]; red[1][threadIdx.x] += red[1][threadIdx.x + st]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float la = (float)(red[0][0] * (double)scale_a), lb = (float)(red[1][0] * (double)scale_b);
    out[0] = la;
    out[1] = lb;
    out[2] = w_a * la + w_b * lb;
  }
}

}  // namespace llp

using namespace llp;

extern "C" size_t llp_loss_workspace_bytes(int64_t rows) { return kSumWsBytes + (size_t)(rows > 0 ? rows : 1) * sizeof(float); }

extern "C" int llp_sum(const float* in, int64_t n, float scale, float* out, void* workspace, void* stream_) {
  LLP_CHECK_ARG(in && out && workspace && n >= 0);
  if (int rc = check_device()) return rc;
  return sum_f32(in, n, scale, out, workspace, (cudaStream_t)stream_);
}

extern "C" int llp_bce(const float* prob, int64_t n, int64_t n_pos, float* loss, float* dprob, void* workspace,
                       void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(prob && loss && workspace && n > 0 && n_pos >= 0 && n_pos <= n);
  if (int rc = check_device()) return rc;
  float* terms = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kSumWsBytes);
  bce_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(prob, n, n_pos, 1.0f / (float)n, terms, dprob);
  LLP_LAUNCH_OK();
  return sum_f32(terms, n, 1.0f / (float)n, loss, workspace, stream);
}

extern "C" int llp_kd_d(const float* s, const float* t, int64_t rows, int64_t K, float T, float* loss, float* ds,
                        void* workspace, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(s && t && loss && workspace && rows > 0 && K > 0 && T > 0.0f);
  if (int rc = check_device()) return rc;
  float* row_loss = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kSumWsBytes);
  kd_d_kernel<<<(unsigned)ceil_div(rows * 32, 256), 256, 0, stream>>>(s, t, rows, K, 1.0f / T, T / (float)rows, row_loss, ds);
  LLP_LAUNCH_OK();
  return sum_f32(row_loss, rows, T * T / (float)rows, loss, workspace, stream);
}

extern "C" int llp_kd_r(const float* s, const float* t, int64_t rows, int64_t K, float margin, float* loss, float* ds,
                        void* workspace, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(s && t && loss && workspace && rows > 0 && K > 1);
  if (K > kMaxRankK) return LLP_E_SHAPE;
  if (int rc = check_device()) return rc;
  float* row_loss = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kSumWsBytes);
  double pairs = (double)rows * (double)K * (double)(K - 1) / 2.0;
  float scale = (float)(1.0 / pairs);
  int warps = 4;
  size_t smem = (size_t)warps * 2 * K * sizeof(float);
  kd_r_kernel<<<(unsigned)ceil_div(rows, warps), warps * 32, smem, stream>>>(s, t, rows, (int)K, margin, scale, row_loss, ds);
  LLP_LAUNCH_OK();
  return sum_f32(row_loss, rows, scale, loss, workspace, stream);
}

extern "C" size_t llp_kd_fused_workspace_bytes(int64_t rows) { return kSumWsBytes + 2 * (size_t)(rows > 0 ? rows : 1) * sizeof(float); }

extern "C" int llp_kd_fused(const float* s, const float* t, int64_t rows, int64_t K, float T, float margin, float w_d,
                            float w_r, float* losses, float* ds, void* workspace, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(s && t && losses && workspace && rows > 0 && K > 1 && T > 0.0f);
  if (K > kMaxRankK) return LLP_E_SHAPE;
  if (int rc = check_device()) return rc;
  float* row_d = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kSumWsBytes);
  float* row_r = row_d + rows;
  const double pairs = (double)rows * (double)K * (double)(K - 1) / 2.0;
  const float r_scale = (float)(1.0 / pairs);
  const int warps = 4;
  const size_t smem = (size_t)warps * 2 * K * sizeof(float);
  kd_fused_kernel<<<(unsigned)ceil_div(rows, warps), warps * 32, smem, stream>>>(
      s, t, rows, (int)K, 1.0f / T, margin, w_d * T / (float)rows, w_r * r_scale, row_d, row_r, ds);
  LLP_LAUNCH_OK();
  double* partial = reinterpret_cast<double*>(workspace);   // 2 x 256 doubles = 4 KiB of the 8 KiB reduction scratch
  sum2_stage1_kernel<<<dim3(kSumBlocks, 2), kSumThreads, 0, stream>>>(row_d, row_r, rows, partial);
  LLP_LAUNCH_OK();
  sum2_stage2_kernel<<<1, kSumBlocks, 0, stream>>>(partial, T * T / (float)rows, r_scale, w_d, w_r, losses);
  LLP_LAUNCH_OK();
  return 0;
}
