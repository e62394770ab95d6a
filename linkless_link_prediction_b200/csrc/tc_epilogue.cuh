// Epilogue shared by the tcgen05 GEMM kernels (bf16 kind::f16 and fp32 3xTF32): bias + addend + relu + Philox dropout +
// relu/dropout-backward gate on one row x 32 accumulator columns read from TMEM, 128/256-bit global accesses.
#pragma once
#include "tcgen05.cuh"

namespace llp {
namespace tc {

struct Maps {
  CUtensorMap a1, b1, a2, b2;
};

struct TcParams {
  int64_t M, N;          // output tile space: rows (M) x cols (N)
  int64_t K1, K2;        // reduction lengths of the two operand pairs (K2 = 0 when unused)
  int splits;            // TN only: number of K-splits
  int64_t k_per_split;   // TN only: reduction rows per split (multiple of BLOCK_K)
  EpilogueParams ep;
  int ep_flags;          // kVec*: which epilogue operands may be accessed with 128-bit vectors
  void* D; int64_t ldd;
  float* partial;        // TN: [splits][M][N] fp32
  int chunk_kb;          // 3xTF32 kernels: k-blocks per accumulator chain (chunked promotion)
  int trunc_hi;          // 3xTF32 kernels: experiment, see gemm_tf32.cu
  long long* dbg;        // instrumentation (llp_set_tuning(15, 1)): per CTA {issue loop ns, operand wait ns, accumulator wait ns}
};

// ---- fused epilogue for one row x 32 columns ------------------------------------------------------------
// Every option is a warp-uniform branch around fully unrolled register code; bias / addend / gate are read with
// 128-bit loads when the host verified the alignment (ep_flags), the dropout mask costs one Philox call per 8 columns.
constexpr int kVecBias = 1, kVecAddend = 2, kVecGate = 4, kVecOut = 8;
constexpr int kVec32Addend = 16, kVec32Gate = 32, kVec32Out = 64;  // 32-byte alignment: 256-bit accesses

// 32 consecutive elements of a row -> fp32.  wide: 256-bit loads (32-byte aligned rows), vec: 128-bit loads.
template <typename TO>
__device__ __forceinline__ void load32(const TO* __restrict__ src, bool wide, bool vec, int valid, float (&v)[32]) {
  if (wide && valid == 32) {
    constexpr int E = 32 / sizeof(TO);  // elements per 256-bit load: 8 fp32 / 16 bf16
#pragma unroll
    for (int j = 0; j < 32; j += E) {
      const U32x8 r = ldg_v8(src + j);
      if constexpr (sizeof(TO) == 4) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[j + i] = __uint_as_float(r.v[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[j + 2 * i] = __uint_as_float(r.v[i] << 16);
          v[j + 2 * i + 1] = __uint_as_float(r.v[i] & 0xffff0000u);
        }
      }
    }
  } else if (vec && valid == 32) {
    constexpr int VE = Vec16<TO>::n;
#pragma unroll
    for (int j = 0; j < 32; j += VE) {
      float t[VE];
      unpack16(ldg_v4(src + j), t, TO());
#pragma unroll
      for (int i = 0; i < VE; ++i) v[j + i] = t[i];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = j < valid ? to_f32(src[j]) : 0.0f;
  }
}

// `seed` / `offset`: the RESOLVED Philox key / counter base of the dropout site (resolve_rng), passed separately so that
// callers can leave the parameter block in constant memory instead of holding a modified copy in registers.
template <typename TO>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&r)[32], int64_t m, int64_t n_base, const TcParams& p,
                                               const uint4& rnd128, uint64_t seed, uint64_t offset) {
  const EpilogueParams& ep = p.ep;
  const int valid = (int)(p.N - n_base < 32 ? p.N - n_base : 32);
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(r[j]);
  if (ep.bias != nullptr) {
    float b[32];
    load32<float>(ep.bias + n_base, false, (p.ep_flags & kVecBias) != 0, valid, b);
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] += b[j];
  }
  if (ep.addend != nullptr) {
    float a[32];
    load32<TO>(reinterpret_cast<const TO*>(ep.addend) + m * ep.ldadd + n_base, (p.ep_flags & kVec32Addend) != 0,
               (p.ep_flags & kVecAddend) != 0, valid, a);
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] += a[j];
  }
  if (ep.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
  }
  if (ep.dropout_p == 0.5f) {  // one random bit per element: this chunk is one 32-bit word of the row's Philox block
    const uint32_t bits = dropout_word(rnd128, (int)((n_base >> 5) & 3));
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = ((bits >> j) & 1u) ? f[j] * 2.0f : 0.0f;
  } else if (ep.dropout_p > 0.0f) {
    const uint32_t thr = dropout_thr16(ep.dropout_p);
    const float scale = 1.0f / (1.0f - ep.dropout_p);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 rnd = philox4x32_10(seed, (uint64_t)m, offset + (uint64_t)((n_base >> 3) + q));
#pragma unroll
      for (int i = 0; i < 8; ++i) f[q * 8 + i] = dropout_u16(rnd, i) >= thr ? f[q * 8 + i] * scale : 0.0f;
    }
  }
  if (ep.gate != nullptr) {
    float g[32];
    load32<TO>(reinterpret_cast<const TO*>(ep.gate) + m * ep.ldgate + n_base, (p.ep_flags & kVec32Gate) != 0,
               (p.ep_flags & kVecGate) != 0, valid, g);
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = g[j] > 0.0f ? f[j] * ep.gate_scale : 0.0f;
  }
  TO* dst = reinterpret_cast<TO*>(p.D) + m * p.ldd + n_base;
  if ((p.ep_flags & kVec32Out) && valid == 32) {  // full 32-byte sectors per lane: half the store instructions
    constexpr int E = 32 / sizeof(TO);
#pragma unroll
    for (int j = 0; j < 32; j += E) {
      U32x8 r;
      if constexpr (sizeof(TO) == 4) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = __float_as_uint(f[j + i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = pack_bf16x2(f[j + 2 * i], f[j + 2 * i + 1]);
      }
      stg_v8(dst + j, r);
    }
  } else if ((p.ep_flags & kVecOut) && valid == 32) {
    constexpr int VE = Vec16<TO>::n;
#pragma unroll
    for (int j = 0; j < 32; j += VE) {
      float g[VE];
#pragma unroll
      for (int i = 0; i < VE; ++i) g[i] = f[j + i];
      stg_v4(dst + j, pack16(g, TO()));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < valid) dst[j] = from_f32<TO>(f[j]);
  }
}

}  // namespace tc
}  // namespace llp
