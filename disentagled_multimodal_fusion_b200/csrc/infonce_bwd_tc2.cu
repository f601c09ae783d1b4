// K2 backward, second generation (bf16 tcgen05): W never leaves the tensor-core datapath.
//
//   dA[i, :] = coef*g * ( sum_j W_ij b_j - 2 b_pos(i) ),   W_ij = exp(s_ij - lseA_i) + exp(s_ij - lseB_j)
//
// CTA = 128 anchor rows x an OW-wide slice of D (OW = 128).  10 warps:
//   warp 0      TMA producer: resident A block (128 x D), 4-stage ring of [128 x 64] column k-blocks for
//               the S MMA, and per tile the transposed column slice BmT[d0:d0+OW, j0:j0+128]
//   warp 1      MMA issuer: S(t) = A * B_t^T into a double-buffered TMEM tile, then
//               O += W(t) * BmT_t^T with W read straight from TMEM (tcgen05.mma, A operand in TMEM)
//   warps 2-9   softmax: tcgen05.ld S -> W = E * (1 + a_i b_j) with ONE exp2 per element
//               (E = exp2(s' - lseA'_i), a_i = exp2(lseA'_i - c0), b_j = exp2(c0 - lseB'_j), b_j staged in smem
//               once per tile) -> bf16 pairs -> tcgen05.st back IN PLACE over the consumed S columns.
// Compared with the first version (W through swizzled smem, 2-stage ring, 4 softmax warps, two exps and
// one global load per element) this frees 32 KB of smem for a deeper TMA ring and removes the softmax
// warps from the critical path.
#include "common.cuh"
#include "tc_common.cuh"

namespace dmf {

constexpr int B2_THREADS = 320;
constexpr int B2_TILE = 128 * 64 * 2;   // 16 KB
constexpr int B2_STAGES = 4;
constexpr int B2_OW = 128;              // output slice width
constexpr float kLog2e2 = 1.4426950408889634f;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__global__ void __launch_bounds__(B2_THREADS, 1)
infonce_bwd_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmBT, int Ma, int Nb, int D, int num_kb, float scale,
                       const float* __restrict__ lseA, const float* __restrict__ lseB, float coef,
                       const float* __restrict__ gscale, long long diag_offset, const uint16_t* __restrict__ Bm,
                       long long ldb, float* __restrict__ dA, long long ldda, int accumulate) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smemA = smem;                                   // num_kb tiles (resident anchors)
  uint8_t* smemB = smemA + num_kb * B2_TILE;               // B2_STAGES tiles (S operand ring)
  uint8_t* smemV = smemB + B2_STAGES * B2_TILE;            // 2 tiles: BmT slice [OW d x 128 j] as two k-blocks
  float* bsm = reinterpret_cast<float*>(smemV + 2 * B2_TILE);   // [2][128] per-tile column factors b_j
  uint64_t* bars = reinterpret_cast<uint64_t*>(bsm + 256);
  uint64_t* a_full = bars;
  uint64_t* full_bar = bars + 1;
  uint64_t* empty_bar = full_bar + B2_STAGES;
  uint64_t* s_full = empty_bar + B2_STAGES;    // [2]
  uint64_t* p_full = s_full + 2;               // [2] W(t) stored in TMEM (8 warp arrivals)
  uint64_t* v_full = p_full + 2;
  uint64_t* pv_done = v_full + 1;
  uint64_t* acc_full = pv_done + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int d0 = blockIdx.y * B2_OW;
  const int ntiles = (Nb + 127) / 128;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    tc::tma_prefetch_desc(&tmBT);
    tc::mbar_init(a_full, 1);
    for (int s = 0; s < B2_STAGES; ++s) { tc::mbar_init(full_bar + s, 1); tc::mbar_init(empty_bar + s, 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(s_full + b, 1); tc::mbar_init(p_full + b, 8); }
    tc::mbar_init(v_full, 1);
    tc::mbar_init(pv_done, 1);
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 256;   // S buffers at [0,128) and [128,256)

  if (warp == 0) {
    if (lane == 0) {
      tc::mbar_expect_tx(a_full, num_kb * B2_TILE);
      for (int kb = 0; kb < num_kb; ++kb) tc::tma_load_2d(smemA + kb * B2_TILE, &tmA, kb * 64, m0, a_full);
      int stage = 0;
      uint32_t phase = 0;
      auto load_b = [&](int t) {
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(empty_bar + stage, phase ^ 1);
          tc::mbar_expect_tx(full_bar + stage, B2_TILE);
          tc::tma_load_2d(smemB + stage * B2_TILE, &tmB, kb * 64, t * 128, full_bar + stage);
          if (++stage == B2_STAGES) { stage = 0; phase ^= 1; }
        }
      };
      // consumption order of the MMA thread: S(0), S(1), PV(0), S(2), PV(1), ...
      load_b(0);
      for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) load_b(t + 1);
        tc::mbar_wait(pv_done, ((uint32_t)t & 1) ^ 1);   // smemV free: PV(t-1) retired
        tc::mbar_expect_tx(v_full, 2 * B2_TILE);
        tc::tma_load_2d(smemV, &tmBT, t * 128, d0, v_full);
        tc::tma_load_2d(smemV + B2_TILE, &tmBT, t * 128 + 64, d0, v_full);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = tc::make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = tc::make_idesc_bf16(128, B2_OW, 0, 0);
      tc::mbar_wait(a_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      auto issue_s = [&](int t) {
        // buffer (t&1) holds W(t-2), consumed by PV(t-2) which was issued earlier on this (in-order) pipe
        const uint32_t d_tmem = tmem_base + (uint32_t)((t & 1) * 128);
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(full_bar + stage, phase);
          tc::tc_fence_after_sync();
          const uint32_t a_addr = tc::smem_u32(smemA + kb * B2_TILE);
          const uint32_t b_addr = tc::smem_u32(smemB + stage * B2_TILE);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::umma_ss(d_tmem, tc::make_smem_desc(a_addr + k * 32, 16, 1024),
                        tc::make_smem_desc(b_addr + k * 32, 16, 1024), idesc_s, (kb | k) != 0 ? 1u : 0u);
          tc::umma_commit(empty_bar + stage);
          if (++stage == B2_STAGES) { stage = 0; phase ^= 1; }
        }
        tc::umma_commit(s_full + (t & 1));
      };
      issue_s(0);
      for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) issue_s(t + 1);
        tc::mbar_wait(p_full + (t & 1), ((uint32_t)t >> 1) & 1);
        tc::mbar_wait(v_full, (uint32_t)t & 1);
        tc::tc_fence_after_sync();
        const uint32_t w_tmem = tmem_base + (uint32_t)((t & 1) * 128);
        const uint32_t v_addr = tc::smem_u32(smemV);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // W columns j = 16k..16k+15: column half (k>>2) keeps its packed pairs at +64*(k>>2) + 8*(k&3)
          const uint32_t a_t = w_tmem + (uint32_t)((k >> 2) * 64 + (k & 3) * 8);
          const uint32_t off = (uint32_t)(k >> 2) * B2_TILE + (uint32_t)(k & 3) * 32;
          tc::umma_ts(tmem_O, a_t, tc::make_smem_desc(v_addr + off, 16, 1024), idesc_o, (t | k) != 0 ? 1u : 0u);
        }
        tc::umma_commit(pv_done);
      }
      tc::umma_commit(acc_full);
    }
  } else {
    const int sw = warp - 2;                 // 0..7
    const int q = warp & 3;                  // TMEM lane quarter
    const int ch = sw >> 2;                  // column half of the tile handled by this warp
    const int st = threadIdx.x - 64;         // 0..255 among softmax threads
    const int rloc = q * 32 + lane;
    const int row = m0 + rloc;
    const float sl2 = scale * kLog2e2;
    const float c0 = __ldg(lseB) * kLog2e2;                         // common exponent offset
    const float la2 = (row < Ma) ? __ldg(lseA + row) * kLog2e2 : c0;
    const float ai = ex2f(la2 - c0);
    for (int t = 0; t < ntiles; ++t) {
      const int j0 = t * 128;
      float* bs = bsm + (t & 1) * 128;
      if (st < 128) {
        const int j = j0 + st;
        bs[st] = (j < Nb) ? ex2f(c0 - __ldg(lseB + j) * kLog2e2) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      tc::mbar_wait(s_full + (t & 1), ((uint32_t)t >> 1) & 1);
      tc::tc_fence_after_sync();
      const uint32_t tS = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((t & 1) * 128 + ch * 64);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tS + (uint32_t)(c * 32), r);
        tc::tmem_ld_wait();
        const float4* b4 = reinterpret_cast<const float4*>(bs + ch * 64 + c * 32);
        uint32_t pk[16];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 bb = b4[j4];
          const float e0 = ex2f(fmaf(__uint_as_float(r[j4 * 4 + 0]), sl2, -la2));
          const float e1 = ex2f(fmaf(__uint_as_float(r[j4 * 4 + 1]), sl2, -la2));
          const float e2 = ex2f(fmaf(__uint_as_float(r[j4 * 4 + 2]), sl2, -la2));
          const float e3 = ex2f(fmaf(__uint_as_float(r[j4 * 4 + 3]), sl2, -la2));
          const float w0 = fmaf(e0 * ai, bb.x, e0);
          const float w1 = fmaf(e1 * ai, bb.y, e1);
          const float w2 = fmaf(e2 * ai, bb.z, e2);
          const float w3 = fmaf(e3 * ai, bb.w, e3);
          pk[j4 * 2 + 0] = pack_bf16x2(w0, w1);
          pk[j4 * 2 + 1] = pack_bf16x2(w2, w3);
        }
        tc::tmem_st_32x16(tS + (uint32_t)(c * 16), pk);   // in place over the columns this thread has consumed
      }
      tc::tmem_st_wait();
      tc::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(p_full + (t & 1));
    }
    // epilogue: this warp stores its lane quarter x column half of the dA slice
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after_sync();
    const float cg = coef * (gscale ? __ldg(gscale) : 1.0f);
    const long long pj = diag_offset >= 0 ? diag_offset + row : -1;
#pragma unroll 1
    for (int c = 0; c < B2_OW / 64; ++c) {
      uint32_t r[32];
      const int cc = ch * (B2_OW / 64) + c;
      tc::tmem_ld_32x32(tmem_O + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 32), r);
      tc::tmem_ld_wait();
      const int dbase = d0 + cc * 32;
      if (row < Ma && dbase < D) {
        const int nvalid = min(32, D - dbase);
        float* dst = dA + (long long)row * ldda + dbase;
        const bool has_pos = pj >= 0 && pj < Nb;
        const uint16_t* bp = has_pos ? Bm + pj * ldb + dbase : nullptr;
        if (nvalid == 32 && !accumulate && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o;
            o.x = (__uint_as_float(r[j + 0]) - (has_pos ? 2.0f * bf2f(bp[j + 0]) : 0.f)) * cg;
            o.y = (__uint_as_float(r[j + 1]) - (has_pos ? 2.0f * bf2f(bp[j + 1]) : 0.f)) * cg;
            o.z = (__uint_as_float(r[j + 2]) - (has_pos ? 2.0f * bf2f(bp[j + 2]) : 0.f)) * cg;
            o.w = (__uint_as_float(r[j + 3]) - (has_pos ? 2.0f * bf2f(bp[j + 3]) : 0.f)) * cg;
            *reinterpret_cast<float4*>(dst + j) = o;
          }
        } else {
          for (int j = 0; j < nvalid; ++j) {
            float v = __uint_as_float(r[j]);
            if (has_pos) v -= 2.0f * bf2f(bp[j]);
            v *= cg;
            dst[j] = accumulate ? dst[j] + v : v;
          }
        }
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<512>(tmem_base);
}

}  // namespace dmf

using namespace dmf;

int dmf_infonce_bwd_bf16_tc2(const void* A, long long lda, int Ma, const float* lseA, const void* Bm, long long ldb,
                             const void* BmT, long long ldbt, int Nb, const float* lseB, int D, float scale, float coef,
                             const float* gscale, long long diag_offset, float* dA, long long ldda, int accumulate,
                             cudaStream_t s) {
  const int num_kb = D / 64;
  CUtensorMap tmA, tmB, tmBT;
  int rc = make_tmap_bf16_2d(&tmA, A, Ma, D, lda, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, Bm, Nb, D, ldb, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmBT, BmT, D, Nb, ldbt, B2_OW);
  if (rc) return rc;
  const size_t smem = 1024 + (size_t)(num_kb + B2_STAGES + 2) * B2_TILE + 1024 + 256;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(infonce_bwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(1024 + (size_t)(8 + B2_STAGES + 2) * B2_TILE + 1024 + 256));
    if (e != cudaSuccess) return fail((int)e, "dmf_infonce_bwd(bf16 v2): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  dim3 grid((Ma + 127) / 128, (D + B2_OW - 1) / B2_OW);
  infonce_bwd_tc2_kernel<<<grid, B2_THREADS, smem, s>>>(tmA, tmB, tmBT, Ma, Nb, D, num_kb, scale, lseA, lseB, coef, gscale,
                                                        diag_offset, (const uint16_t*)Bm, ldb, dA, ldda, accumulate);
  return launched("dmf_infonce_bwd(bf16 v2)");
}
