// Optimiser tail on flat fp32 buffers: per-group gradient-norm clipping fused with the Adam update.
// Replaces clip_grad_norm_(model) + clip_grad_norm_(predictor) + torch.optim.Adam.step foreach kernels
// (reference: src/train_teacher_gnn.py:63-67, :426-428; src/main.py:226-230).
// `grad_scale` folds the 1/world_size of the data-parallel gradient average into the same pass.
#include "common.cuh"

namespace llp {

constexpr int kMaxGroups = 8;
constexpr int kNormBlocks = 128;

struct Groups {
  int64_t begin[kMaxGroups + 1];
  int n;
};

// partial[g][b] = sum of squares of the slice of group g owned by block b (double accumulation, fixed order)
__global__ void sumsq_kernel(const float* __restrict__ grad, Groups groups, double* __restrict__ partial,
                             int64_t* __restrict__ device_step) {
  __shared__ double red[256];
  if (device_step != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *device_step += 1;
  int g = blockIdx.y;
  int64_t b = groups.begin[g], e = groups.begin[g + 1];
  double acc = 0.0;
  for (int64_t i = b + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x) {
    double v = (double)grad[i];
    acc += v * v;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[g * kNormBlocks + blockIdx.x] = red[0];
}

__global__ void clip_adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                                 float* __restrict__ v, int64_t n, Groups groups, const double* __restrict__ partial,
                                 float max_norm, float grad_scale, float step_size, float beta1, float beta2,
                                 float inv_sqrt_bc2, float eps, __nv_bfloat16* __restrict__ bf16_copy,
                                 float* __restrict__ group_norms, const int64_t* __restrict__ device_step, float lr) {
  __shared__ float coef[kMaxGroups];
  __shared__ float s_step_size, s_inv_sqrt_bc2;
  if (threadIdx.x == 0) {
    if (device_step != nullptr) {  // bias corrections from the device-side step counter
      const double t = (double)*device_step;
      s_step_size = (float)((double)lr / (1.0 - pow((double)beta1, t)));
      s_inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)beta2, t)));
    } else {
      s_step_size = step_size;
      s_inv_sqrt_bc2 = inv_sqrt_bc2;
    }
  }
  if (threadIdx.x < groups.n) {
    double s = 0.0;
    for (int b = 0; b < kNormBlocks; ++b) s += partial[threadIdx.x * kNormBlocks + b];
    float norm = (float)sqrt(s) * grad_scale;
    // torch.nn.utils.clip_grad_norm_: coef = clamp(max_norm / (total_norm + 1e-6), max=1)
    coef[threadIdx.x] = max_norm > 0.0f ? fminf(max_norm / (norm + 1e-6f), 1.0f) : 1.0f;
    if (blockIdx.x == 0 && group_norms != nullptr) group_norms[threadIdx.x] = norm;
  }
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int g = 0;
    while (g + 1 < groups.n && i >= groups.begin[g + 1]) ++g;
    float gr = grad[i] * grad_scale * coef[g];
    float mi = m[i] * beta1 + gr * (1.0f - beta1);
    float vi = v[i] * beta2 + gr * gr * (1.0f - beta2);
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) * s_inv_sqrt_bc2 + eps;
    float p = param[i] - s_step_size * (mi / denom);
    param[i] = p;
    if (bf16_copy != nullptr) bf16_copy[i] = __float2bfloat16_rn(p);
  }
}

}  // namespace llp

using namespace llp;

namespace llp {
__global__ void rng_advance_kernel(uint64_t* state) { state[1] += 1; }
}  // namespace llp

extern "C" int llp_rng_advance(uint64_t* state, void* stream_) {
  LLP_CHECK_ARG(state != nullptr);
  if (int rc = check_device()) return rc;
  llp::rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream_>>>(state);
  LLP_LAUNCH_OK();
  return 0;
}

extern "C" size_t llp_clip_adam_workspace_bytes(int num_groups) {
  return (size_t)(num_groups > 0 ? num_groups : 1) * kNormBlocks * sizeof(double);
}

extern "C" int llp_clip_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                             const int64_t* host_group_begin, int num_groups, float max_norm, float grad_scale, float lr,
                             float beta1, float beta2, float eps, int64_t step, int64_t* device_step, void* bf16_copy,
                             float* group_norms,
                             void* workspace, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LLP_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && host_group_begin && workspace && n > 0);
  LLP_CHECK_ARG(num_groups >= 1 && num_groups <= kMaxGroups && (step >= 1 || device_step != nullptr));
  if (int rc = check_device()) return rc;
  Groups groups;
  groups.n = num_groups;
  for (int g = 0; g <= num_groups; ++g) groups.begin[g] = host_group_begin[g];
  LLP_CHECK_ARG(groups.begin[0] == 0 && groups.begin[num_groups] == n);
  double* partial = reinterpret_cast<double*>(workspace);
  dim3 grid(kNormBlocks, num_groups);
  sumsq_kernel<<<grid, 256, 0, stream>>>(grad, groups, partial, device_step);
  LLP_LAUNCH_OK();
  const double tstep = (double)(step >= 1 ? step : 1);
  double bc1 = 1.0 - pow((double)beta1, tstep);
  double bc2 = 1.0 - pow((double)beta2, tstep);
  float step_size = (float)((double)lr / bc1);
  float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  unsigned blocks = (unsigned)imin64(ceil_div(n, 256), (int64_t)kNumSMs * 8);
  clip_adam_kernel<<<blocks, 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, n, groups, partial, max_norm, grad_scale,
                                               step_size, beta1, beta2, inv_sqrt_bc2, eps,
                                               reinterpret_cast<__nv_bfloat16*>(bf16_copy), group_norms, device_step, lr);
  LLP_LAUNCH_OK();
  return 0;
}
