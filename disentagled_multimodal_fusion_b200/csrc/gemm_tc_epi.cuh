// Epilogue shared by the tcgen05 GEMM kernels (gemm_tc.cu single-CTA, gemm_tc2.cu persistent CTA pair):
// one thread owns one accumulator row and 32 consecutive columns just read from TMEM.
//   bias / ReLU / ReLU-mask (dgrad) -> fp32 store (plain or red.add for split-K), bf16 row-major store,
//   bf16 TRANSPOSED store (out_t[n][m]: the copy the next wgrad consumes as a K-major operand, so no
//   separate transpose pass ever runs over the activations).
#pragma once
#include "common.cuh"

namespace dmf {

struct TcEpi {
  float* out_f32; long long ldo_f32;
  uint16_t* out_bf16; long long ldo_bf16;
  uint16_t* out_t; long long ldo_t;          // [N, ldo_t] bf16, element (n, m)
  const float* bias;
  const uint16_t* mask; long long ldmask;
  int M, N;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// kAtomic: fp32 output is accumulated with red.global.add (split-K partial tiles); `add_bias` tells whether
// this K-split owns the bias (only the first one does).
template <int EPI, bool kAtomic>
__device__ __forceinline__ void tc_epilogue_chunk(const TcEpi& g, const uint32_t (&r)[32], int row, int nbase,
                                                  bool add_bias) {
  if (row >= g.M || nbase >= g.N) return;
  const int nvalid = min(32, g.N - nbase);
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float x = __uint_as_float(r[j]);
    if (EPI == DMF_EPI_BIAS || EPI == DMF_EPI_BIAS_RELU) {
      if (g.bias && add_bias && j < nvalid) x += __ldg(g.bias + nbase + j);
    }
    if (EPI == DMF_EPI_BIAS_RELU) x = fmaxf(x, 0.f);
    v[j] = x;
  }
  if (EPI == DMF_EPI_RELU_MASK) {
    const uint16_t* mk = g.mask + (long long)row * g.ldmask + nbase;
    if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(mk) & 15) == 0)) {
#pragma unroll
      for (int j8 = 0; j8 < 4; ++j8) {
        const uint4 m4 = __ldg(reinterpret_cast<const uint4*>(mk) + j8);
        const uint32_t w[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          // activation is post-ReLU bf16: positive <=> nonzero magnitude with sign bit clear
          if (!(bf2f((uint16_t)(w[h] & 0xFFFFu)) > 0.f)) v[j8 * 8 + h * 2] = 0.f;
          if (!(bf2f((uint16_t)(w[h] >> 16)) > 0.f)) v[j8 * 8 + h * 2 + 1] = 0.f;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid && !(bf2f(mk[j]) > 0.f)) v[j] = 0.f;
    }
  }
  if (g.out_f32) {
    float* dst = g.out_f32 + (long long)row * g.ldo_f32 + nbase;
    const bool vec = nvalid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    if (kAtomic) {
      if (vec) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) red_add_v4(dst + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) atomicAdd(dst + j, v[j]);
      }
    } else {
      if (vec) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) dst[j] = v[j];
      }
    }
  }
  if (g.out_bf16) {
    uint16_t* dst = g.out_bf16 + (long long)row * g.ldo_bf16 + nbase;
    if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 pk;
        pk.x = pack_bf16x2(v[j], v[j + 1]);
        pk.y = pack_bf16x2(v[j + 2], v[j + 3]);
        pk.z = pack_bf16x2(v[j + 4], v[j + 5]);
        pk.w = pack_bf16x2(v[j + 6], v[j + 7]);
        *reinterpret_cast<uint4*>(dst + j) = pk;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) dst[j] = f2bf(v[j]);
    }
  }
  if (g.out_t) {
    // lanes of the warp hold consecutive rows: each store instruction writes 32 consecutive bf16 (64 B)
    uint16_t* dst = g.out_t + (long long)nbase * g.ldo_t + row;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < nvalid) dst[(long long)j * g.ldo_t] = f2bf(v[j]);
  }
}

}  // namespace dmf
