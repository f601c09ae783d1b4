// Epilogue shared by the tcgen05 GEMM kernels (gemm_tc.cu single-CTA, gemm_tc2.cu persistent CTA pair).
//
// A warp owns a 32-row x 32-column chunk of the accumulator (lane = row, as tcgen05.ld 32x32b delivers it).
// Storing from that layout would touch 32 different cache lines per instruction, so the chunk is staged
// through a per-warp shared-memory tile (pitch 36 words: conflict-free 128-bit stores by row and 128-bit
// loads by quarter-warp) and written back COALESCED: lane -> (row = lane/8 + 4*it, 4 consecutive columns),
// i.e. four full 128-byte row segments per fp32 store instruction.  Bias / ReLU / ReLU-mask (dgrad) are
// applied in the coalesced phase (bias and mask are then read coalesced too).  Outputs:
//   fp32 (plain, or red.global.add.v4.f32 for split-K), bf16 row-major, and a bf16 TRANSPOSED copy
//   out_t[n][m] -- the K-major operand of the next wgrad -- so no transpose pass ever runs over activations.
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
  int fast;        // host-checked: pointers / pitches allow the vector fast path (see tc_epi_fast_ok)
};

// The lean path of a FULL 32 x 32 chunk needs: no transposed copy, 16-byte aligned fp32 rows, 8-byte aligned bf16 / mask
// rows (pitches multiples of 4 elements).  Checked once per group on the host.
inline int tc_epi_fast_ok(const TcEpi& g) {
  if (g.out_t) return 0;
  if (g.out_f32 && ((reinterpret_cast<uintptr_t>(g.out_f32) & 15) != 0 || (g.ldo_f32 & 3) != 0)) return 0;
  if (g.out_bf16 && ((reinterpret_cast<uintptr_t>(g.out_bf16) & 7) != 0 || (g.ldo_bf16 & 3) != 0)) return 0;
  if (g.mask && ((reinterpret_cast<uintptr_t>(g.mask) & 7) != 0 || (g.ldmask & 3) != 0)) return 0;
  return 1;
}

constexpr int kEpiPitch = 36;                       // words per staged row
constexpr int kEpiStageFloats = 32 * kEpiPitch;     // per-warp staging tile

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Lean path of a full 32 x 32 chunk (no ragged edges, no transposed copy; alignment checked on the host):
// the same staging tile and lane mapping as the general path below, with every store a full-width vector store and the
// destination pointers advanced by a constant -- ~125 instructions per chunk against ~540 for the general path, which
// (ncu, K = 1024: 17 k warp instructions per 256 x 256 tile at 29 % issue utilisation) made every layer with K <= 1024
// EPILOGUE-bound: ~14 k cycles per tile next to a 4-8 k cycle main loop.
template <int EPI, bool kAtomic>
__device__ __forceinline__ void tc_epilogue_chunk_fast(const TcEpi& g, const uint32_t (&r)[32], int row0, int lane, int nbase,
                                                       bool add_bias, float* stage) {
  const int c = (lane & 7) * 4;
  const int rq = lane >> 3;
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if ((EPI == DMF_EPI_BIAS || EPI == DMF_EPI_BIAS_RELU) && g.bias && add_bias) {
    const float* bp = g.bias + nbase + c;
    b4 = make_float4(__ldg(bp), __ldg(bp + 1), __ldg(bp + 2), __ldg(bp + 3));
  }
  float4* srow = reinterpret_cast<float4*>(stage + lane * kEpiPitch);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    srow[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                          __uint_as_float(r[4 * j + 3]));
  // the ReLU-mask words of the chunk are fetched once the accumulator registers are dead (8 independent loads in flight)
  uint2 mk8[8];
  if (EPI == DMF_EPI_RELU_MASK) {
    const uint16_t* mk0 = g.mask + (long long)(row0 + rq) * g.ldmask + nbase + c;
#pragma unroll
    for (int it = 0; it < 8; ++it) mk8[it] = __ldg(reinterpret_cast<const uint2*>(mk0 + (long long)(4 * it) * g.ldmask));
  }
  __syncwarp();
  float* d32 = g.out_f32 ? g.out_f32 + (long long)(row0 + rq) * g.ldo_f32 + nbase + c : nullptr;
  uint16_t* d16 = g.out_bf16 ? g.out_bf16 + (long long)(row0 + rq) * g.ldo_bf16 + nbase + c : nullptr;
  const long long s32 = 4 * g.ldo_f32, s16 = 4 * g.ldo_bf16;
  const float* sp = stage + rq * kEpiPitch + c;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    float4 v = *reinterpret_cast<const float4*>(sp + it * 4 * kEpiPitch);
    if (EPI == DMF_EPI_BIAS || EPI == DMF_EPI_BIAS_RELU) { v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w; }
    if (EPI == DMF_EPI_BIAS_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (EPI == DMF_EPI_RELU_MASK) {
      // the mask is a post-ReLU bf16 activation (>= 0, possibly -0): strictly positive <=> its low 15 bits are not all zero
      const uint2 m2 = mk8[it];
      if ((m2.x & 0x00007FFFu) == 0u || (m2.x & 0x00008000u)) v.x = 0.f;
      if ((m2.x & 0x7FFF0000u) == 0u || (m2.x & 0x80000000u)) v.y = 0.f;
      if ((m2.y & 0x00007FFFu) == 0u || (m2.y & 0x00008000u)) v.z = 0.f;
      if ((m2.y & 0x7FFF0000u) == 0u || (m2.y & 0x80000000u)) v.w = 0.f;
    }
    if (d32) {
      if (kAtomic) red_add_v4(d32, v.x, v.y, v.z, v.w);       // split-K partial tile / accumulate-into-grad (wgrad)
      else *reinterpret_cast<float4*>(d32) = v;
      d32 += s32;
    }
    if (d16) {
      uint2 pk;
      pk.x = pack_bf16x2(v.x, v.y);
      pk.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(d16) = pk;
      d16 += s16;
    }
  }
  __syncwarp();     // the staging tile is reused by the next chunk
}

// r[32]: this lane's row (row0 + lane) of the chunk, columns nbase .. nbase+31.  `stage` = this warp's
// private smem tile (kEpiStageFloats floats, 16-byte aligned).  All 32 lanes must call (warp-uniform flow).
// kAtomic: fp32 output accumulated with red.add (split-K partial tiles); add_bias: this K-split owns the bias.
template <int EPI, bool kAtomic>
__device__ __forceinline__ void tc_epilogue_chunk(const TcEpi& g, const uint32_t (&r)[32], int row0, int lane, int nbase,
                                                  bool add_bias, float* stage) {
  if (row0 >= g.M || nbase >= g.N) return;          // warp-uniform
  const int nvalid = min(32, g.N - nbase);
  const int mvalid = min(32, g.M - row0);
  if (g.fast && nvalid == 32 && mvalid == 32) {
    tc_epilogue_chunk_fast<EPI, kAtomic>(g, r, row0, lane, nbase, add_bias, stage);
    return;
  }
  // ReLU-mask words of the whole chunk are fetched up front (8 independent 8-byte loads in flight) so the
  // coalesced pass below pays ONE global-load latency per chunk instead of one per row group
  const int c = (lane & 7) * 4;
  const int rq = lane >> 3;
  const bool full_n = nvalid == 32;
  uint2 mk8[8];
  bool mask_vec = false;
  if (EPI == DMF_EPI_RELU_MASK) {
    const uint16_t* mk0 = g.mask + (long long)(row0 + rq) * g.ldmask + nbase + c;
    mask_vec = full_n && ((reinterpret_cast<uintptr_t>(mk0) & 7) == 0) && ((g.ldmask & 3) == 0);
    if (mask_vec) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        mk8[it] = make_uint2(0x3F803F80u, 0x3F803F80u);
        if (rq + 4 * it < mvalid) mk8[it] = __ldg(reinterpret_cast<const uint2*>(mk0 + (long long)(4 * it) * g.ldmask));
      }
    }
  }
  // ---- phase 1: raw accumulators -> smem, one 144-byte row per lane
  float4* srow = reinterpret_cast<float4*>(stage + lane * kEpiPitch);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    srow[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                          __uint_as_float(r[4 * j + 3]));
  __syncwarp();
  // ---- phase 2: coalesced pass, lane -> (row rr = lane/8 + 4*it, columns c = (lane%8)*4 .. +3)
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if ((EPI == DMF_EPI_BIAS || EPI == DMF_EPI_BIAS_RELU) && g.bias && add_bias) {
    if (c + 0 < nvalid) b4.x = __ldg(g.bias + nbase + c + 0);
    if (c + 1 < nvalid) b4.y = __ldg(g.bias + nbase + c + 1);
    if (c + 2 < nvalid) b4.z = __ldg(g.bias + nbase + c + 2);
    if (c + 3 < nvalid) b4.w = __ldg(g.bias + nbase + c + 3);
  }
  const bool need_back = g.out_t != nullptr && (EPI != DMF_EPI_NONE);   // T pass must see the post-epilogue values
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = rq + 4 * it;
    float4 v = *reinterpret_cast<const float4*>(stage + rr * kEpiPitch + c);
    if (EPI == DMF_EPI_BIAS || EPI == DMF_EPI_BIAS_RELU) { v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w; }
    if (EPI == DMF_EPI_BIAS_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    const bool rvalid = rr < mvalid;
    const long long grow = row0 + rr;
    if (EPI == DMF_EPI_RELU_MASK) {
      if (rvalid) {
        if (mask_vec) {
          const uint2 m2 = mk8[it];
          // activation is post-ReLU bf16: keep the gradient where it is strictly positive
          if (!(bf2f((uint16_t)(m2.x & 0xFFFFu)) > 0.f)) v.x = 0.f;
          if (!(bf2f((uint16_t)(m2.x >> 16)) > 0.f)) v.y = 0.f;
          if (!(bf2f((uint16_t)(m2.y & 0xFFFFu)) > 0.f)) v.z = 0.f;
          if (!(bf2f((uint16_t)(m2.y >> 16)) > 0.f)) v.w = 0.f;
        } else {
          const uint16_t* mk = g.mask + grow * g.ldmask + nbase + c;
          if (c + 0 < nvalid && !(bf2f(mk[0]) > 0.f)) v.x = 0.f;
          if (c + 1 < nvalid && !(bf2f(mk[1]) > 0.f)) v.y = 0.f;
          if (c + 2 < nvalid && !(bf2f(mk[2]) > 0.f)) v.z = 0.f;
          if (c + 3 < nvalid && !(bf2f(mk[3]) > 0.f)) v.w = 0.f;
        }
      }
    }
    if (need_back) *reinterpret_cast<float4*>(stage + rr * kEpiPitch + c) = v;
    if (rvalid && g.out_f32) {
      float* dst = g.out_f32 + grow * g.ldo_f32 + nbase + c;
      const bool vec = full_n && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
      if (kAtomic) {
        if (vec) {
          red_add_v4(dst, v.x, v.y, v.z, v.w);
        } else {
          if (c + 0 < nvalid) atomicAdd(dst + 0, v.x);
          if (c + 1 < nvalid) atomicAdd(dst + 1, v.y);
          if (c + 2 < nvalid) atomicAdd(dst + 2, v.z);
          if (c + 3 < nvalid) atomicAdd(dst + 3, v.w);
        }
      } else {
        if (vec) {
          *reinterpret_cast<float4*>(dst) = v;
        } else {
          if (c + 0 < nvalid) dst[0] = v.x;
          if (c + 1 < nvalid) dst[1] = v.y;
          if (c + 2 < nvalid) dst[2] = v.z;
          if (c + 3 < nvalid) dst[3] = v.w;
        }
      }
    }
    if (rvalid && g.out_bf16) {
      uint16_t* dst = g.out_bf16 + grow * g.ldo_bf16 + nbase + c;
      if (full_n && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0)) {
        uint2 pk;
        pk.x = pack_bf16x2(v.x, v.y);
        pk.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(dst) = pk;
      } else {
        if (c + 0 < nvalid) dst[0] = f2bf(v.x);
        if (c + 1 < nvalid) dst[1] = f2bf(v.y);
        if (c + 2 < nvalid) dst[2] = f2bf(v.z);
        if (c + 3 < nvalid) dst[3] = f2bf(v.w);
      }
    }
  }
  // ---- phase 3: transposed bf16 copy; lane = row again, each store instruction writes 32 consecutive
  // bf16 (64 B) of one output row n
  if (g.out_t) {
    if (need_back) __syncwarp();
    if (lane < mvalid) {
      uint16_t* dst = g.out_t + (long long)nbase * g.ldo_t + row0 + lane;
      if (need_back) {
        const float4* sr = reinterpret_cast<const float4*>(stage + lane * kEpiPitch);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 q4 = sr[j4];
          const int j = 4 * j4;
          if (j + 0 < nvalid) dst[(long long)(j + 0) * g.ldo_t] = f2bf(q4.x);
          if (j + 1 < nvalid) dst[(long long)(j + 1) * g.ldo_t] = f2bf(q4.y);
          if (j + 2 < nvalid) dst[(long long)(j + 2) * g.ldo_t] = f2bf(q4.z);
          if (j + 3 < nvalid) dst[(long long)(j + 3) * g.ldo_t] = f2bf(q4.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) dst[(long long)j * g.ldo_t] = f2bf(__uint_as_float(r[j]));
      }
    }
  }
  __syncwarp();     // the staging tile is reused by the next chunk
}

}  // namespace dmf
