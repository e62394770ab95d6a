// Shared helpers for the LLP B200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/llp_b200.h"

namespace llp {

extern std::atomic<int64_t> g_launch_count;

inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

#define LLP_CHECK_ARG(cond) \
  do {                      \
    if (!(cond)) return LLP_E_BADARG; \
  } while (0)

#define LLP_CUDA(expr)                         \
  do {                                         \
    cudaError_t _e = (expr);                   \
    if (_e != cudaSuccess) return (int)_e;     \
  } while (0)

// After a kernel launch: surface launch-configuration errors without synchronising.
#define LLP_LAUNCH_OK()                        \
  do {                                         \
    cudaError_t _e = cudaGetLastError();       \
    if (_e != cudaSuccess) return (int)_e;     \
    ::llp::count_launch();                     \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// Function attributes (opt-in shared memory) are per DEVICE: a process that drives several GPUs must set them on each.
struct PerDeviceOnce {
  std::atomic<uint64_t> mask{0};
  uint64_t bit = 0;
  bool need() {
    int d = 0;
    cudaGetDevice(&d);
    bit = 1ull << (d & 63);
    return (mask.load(std::memory_order_acquire) & bit) == 0;
  }
  void done() { mask.fetch_or(bit, std::memory_order_release); }
};

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
static inline int64_t imax64(int64_t a, int64_t b) { return a > b ? a : b; }

static inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

template <typename T>
struct DType;
template <>
struct DType<float> {
  static constexpr int id = LLP_F32;
};
template <>
struct DType<__nv_bfloat16> {
  static constexpr int id = LLP_BF16;
};

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- 16-byte vectors of T <-> fp32 lanes -----------------------------------------------------
template <typename T>
struct Vec16;  // elements per 16 bytes
template <>
struct Vec16<float> {
  static constexpr int n = 4;
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int n = 8;
};

__device__ __forceinline__ void unpack16(const uint4& r, float (&f)[4], float) {
  f[0] = __uint_as_float(r.x); f[1] = __uint_as_float(r.y); f[2] = __uint_as_float(r.z); f[3] = __uint_as_float(r.w);
}
__device__ __forceinline__ void unpack16(const uint4& r, float (&f)[8], __nv_bfloat16) {
  // bf16 -> fp32 is a 16-bit shift
  f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
  f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
  f[4] = __uint_as_float(r.z << 16); f[5] = __uint_as_float(r.z & 0xffff0000u);
  f[6] = __uint_as_float(r.w << 16); f[7] = __uint_as_float(r.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack16(const float (&f)[4], float) {
  return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
}
__device__ __forceinline__ uint4 pack16(const float (&f)[8], __nv_bfloat16) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {   // streaming data: read once, keep it out of L1
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// gathered embedding rows: allowed to allocate in L1 — the hub rows of a power-law graph repeat within a CTA's lifetime
// (measured on the C4 SpMM: 224 us vs 235 us forward, 249 us vs 261 us transpose)
__device__ __forceinline__ uint4 ldg_gather_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_v4(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg_v4(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

// 256-bit global accesses (sm_100: LDG.256 / STG.256): one full 32-byte sector per lane and instruction
struct alignas(32) U32x8 { uint32_t v[8]; };
__device__ __forceinline__ U32x8 ldg_v8(const void* p) {
  U32x8 r;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_v8(void* p, const U32x8& r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]),
               "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- Philox4x32-10 (dropout masks) ---------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// Dropout mask from Philox4x32-10, a pure function of (seed, offset, row, col) shared by every kernel that draws it.
//   p == 0.5 (the reference default, train_teacher_gnn.py:277): ONE random bit per element — a Philox call covers 128
//             consecutive columns of a row (call index col/128, word (col/32)%4, bit col%32), keep iff the bit is set;
//   other p : 16 random bits per element — a call covers 8 columns, keep iff u16 >= p * 65536.
__host__ __device__ __forceinline__ uint32_t dropout_thr16(float p) {
  float t = p * 65536.0f;
  return t >= 65535.0f ? 65535u : (uint32_t)t;
}
__device__ __forceinline__ uint32_t dropout_u16(const uint4& r, int idx /*0..7*/) {
  uint32_t w = (idx >> 1) == 0 ? r.x : (idx >> 1) == 1 ? r.y : (idx >> 1) == 2 ? r.z : r.w;
  return (idx & 1) ? (w >> 16) : (w & 0xffffu);
}
__device__ __forceinline__ uint32_t dropout_word(const uint4& r, int idx /*0..3*/) {
  return idx == 0 ? r.x : idx == 1 ? r.y : idx == 2 ? r.z : r.w;
}
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t offset, int64_t row, int64_t col, float p) {
  if (p == 0.5f) {
    uint4 r = philox4x32_10(seed, (uint64_t)row, offset + (uint64_t)(col >> 7));
    return (dropout_word(r, (int)((col >> 5) & 3)) >> (col & 31)) & 1u;
  }
  uint4 r = philox4x32_10(seed, (uint64_t)row, offset + (uint64_t)(col >> 3));
  return dropout_u16(r, (int)(col & 7)) >= dropout_thr16(p);
}

// ---- epilogue shared by the SIMT and tcgen05 GEMMs ---------------------------------------------
struct EpilogueParams {
  const float* bias;
  const void* addend; int64_t ldadd;
  const void* gate; int64_t ldgate;
  float gate_scale;
  int relu;
  float dropout_p;
  uint64_t seed, offset;
  const uint64_t* rng_state;  // device {seed, step}: when set, the mask stream also depends on it (CUDA-graph safe)
};

// Effective Philox key / counter base of a dropout site.  By value: key = seed, base = offset << 20.  With a device
// rng_state {s, step}: key = seed ^ s, base = (step << 32) + (offset << 20); `offset` is then the site id (< 4096), the
// low 20 bits are left for the column-group index.
__device__ __forceinline__ void resolve_rng(EpilogueParams& ep) {
  uint64_t key = ep.seed, base = ep.offset << 20;
  if (ep.rng_state != nullptr) {
    key ^= ep.rng_state[0];
    base += ep.rng_state[1] << 32;
  }
  ep.seed = key;
  ep.offset = base;
}

template <typename TO>
__device__ __forceinline__ float epilogue_apply(float acc, int64_t m, int64_t n, const EpilogueParams& ep) {
  if (ep.bias) acc += __ldg(ep.bias + n);
  if (ep.addend) acc += to_f32(reinterpret_cast<const TO*>(ep.addend)[m * ep.ldadd + n]);
  if (ep.relu) acc = fmaxf(acc, 0.0f);
  if (ep.dropout_p > 0.0f) acc = dropout_keep(ep.seed, ep.offset, m, n, ep.dropout_p) ? acc * (1.0f / (1.0f - ep.dropout_p)) : 0.0f;
  if (ep.gate) acc = to_f32(reinterpret_cast<const TO*>(ep.gate)[m * ep.ldgate + n]) > 0.0f ? acc * ep.gate_scale : 0.0f;
  return acc;
}

extern int g_tuning[32];
long long* debug_buffer();   // lazily allocated device scratch of 4096 int64 for instrumented runs (llp_debug_read)  // experiment knobs (llp_set_tuning), 0 = default

int check_device();  // 0 if the current device is sm_100, LLP_E_DEVICE otherwise (cached)

// In-place exclusive scan of n int32 (edge_bwd.cu); tile_sum: scan_i32_tiles(n) int32 of scratch
int64_t scan_i32_tiles(int64_t n);
int exclusive_scan_i32(int32_t* x, int64_t n, int32_t* tile_sum, cudaStream_t stream);

// Deterministic reduction helpers implemented in loss.cu
int sum_f32(const float* in, int64_t n, float scale, float* out, void* workspace, cudaStream_t stream);

}  // namespace llp
