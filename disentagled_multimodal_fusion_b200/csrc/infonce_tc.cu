// K2 (bf16 tensor-core path): fused InfoNCE forward / backward tiles on tcgen05 + TMEM + TMA.
//
// Replaces the [2B,2B] matmul + ~25 elementwise/reduction kernels of SupConLoss.forward
// (models/losses.py:64-99) and the autograd graph behind it.  Logits are never written to memory:
// a 128x128 fp32 similarity tile lives in TMEM, softmax warps read it with tcgen05.ld and keep an
// online (max, sum-exp) per anchor row (forward) or turn it into W = exp(s-lseA)+exp(s-lseB)
// (backward) which feeds a second tcgen05.mma that accumulates dA in TMEM.
//
// Forward  rowlse_tc_kernel    CTA = 128 anchor rows x a column split.  A block (128 x D bf16) is
//   TMA-loaded once and stays resident in smem; column tiles of Bm stream through a 5-stage ring of
//   [128 x 64] k-blocks; S tiles are 4-way buffered in TMEM (4 x 128 columns).
//   warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = softmax (one row per thread).
// Backward infonce_bwd_tc_kernel  CTA = 128 anchor rows x a 128-wide slice of D.  S tile recomputed
//   as in forward; softmax warps write W (bf16) into a 128B-swizzled smem tile; second MMA
//   dA[128 x 128] += W[128 x 128] * BmT_slice^T with the transposed column block (K-major) from TMA.
// All operands are K-major / 128B swizzle, the same descriptor family as gemm_tc.cu.
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace dmf {

constexpr int NT_THREADS = 192;
constexpr int NT_TILE = 128 * 64 * 2;  // 16 KB: [128 rows x 64 bf16]
constexpr int NT_MAX_KB = 8;           // D <= 512
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------------------------------------ forward
constexpr int FW_STAGES = 5, FW_BUFS = 4, FW_THREADS = 320;

__global__ void __launch_bounds__(FW_THREADS, 1)
rowlse_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int Ma, int Nb,
                 int num_kb, float scale, int tiles_per_split, float* __restrict__ part_max,
                 float* __restrict__ part_sum, long long diag_offset, float* __restrict__ diag_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smemA = smem;                                  // num_kb tiles
  uint8_t* smemB = smem + num_kb * NT_TILE;               // FW_STAGES tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemB + FW_STAGES * NT_TILE);
  uint64_t* a_full = bars;
  uint64_t* full_bar = bars + 1;
  uint64_t* empty_bar = full_bar + FW_STAGES;
  uint64_t* s_full = empty_bar + FW_STAGES;
  uint64_t* s_empty = s_full + FW_BUFS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + FW_BUFS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int total_tiles = (Nb + 127) / 128;
  const int jt0 = blockIdx.y * tiles_per_split;
  const int ntiles = max(0, min(total_tiles, jt0 + tiles_per_split) - jt0);

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    tc::mbar_init(a_full, 1);
    for (int s = 0; s < FW_STAGES; ++s) { tc::mbar_init(full_bar + s, 1); tc::mbar_init(empty_bar + s, 1); }
    for (int b = 0; b < FW_BUFS; ++b) { tc::mbar_init(s_full + b, 1); tc::mbar_init(s_empty + b, 8); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && ntiles > 0) {
      tc::mbar_expect_tx(a_full, num_kb * NT_TILE);
      for (int kb = 0; kb < num_kb; ++kb) tc::tma_load_2d(smemA + kb * NT_TILE, &tmA, kb * 64, m0, a_full);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int j0 = (jt0 + t) * 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(empty_bar + stage, phase ^ 1);
          tc::mbar_expect_tx(full_bar + stage, NT_TILE);
          tc::tma_load_2d(smemB + stage * NT_TILE, &tmB, kb * 64, j0, full_bar + stage);
          if (++stage == FW_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && ntiles > 0) {
      constexpr uint32_t idesc = tc::make_idesc_bf16(128, 128, 0, 0);
      tc::mbar_wait(a_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int buf = t % FW_BUFS;
        tc::mbar_wait(s_empty + buf, (((uint32_t)(t / FW_BUFS)) & 1) ^ 1);
        tc::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 128);
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(full_bar + stage, phase);
          tc::tc_fence_after_sync();
          const uint32_t a_addr = tc::smem_u32(smemA + kb * NT_TILE);
          const uint32_t b_addr = tc::smem_u32(smemB + stage * NT_TILE);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::umma_ss(d_tmem, tc::make_smem_desc(a_addr + k * 32, 16, 1024),
                        tc::make_smem_desc(b_addr + k * 32, 16, 1024), idesc, (kb | k) != 0 ? 1u : 0u);
          tc::umma_commit(empty_bar + stage);
          if (++stage == FW_STAGES) { stage = 0; phase ^= 1; }
        }
        tc::umma_commit(s_full + buf);
      }
    }
  } else {
    // 8 softmax warps: lane quarter q = warp % 4 (TMEM access rule), column half ch of every tile
    const int sw = warp - 2;
    const int q = warp & 3;
    const int ch = sw >> 2;
    const int rloc = q * 32 + lane;
    const int row = m0 + rloc;
    const float sl2 = scale * kLog2e;
    float m = -INFINITY, l = 0.f, diag = 0.f;
    bool has_diag = false;
    const long long dj = (diag_offset >= 0 && row < Ma) ? diag_offset + row : -1;
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t % FW_BUFS;
      tc::mbar_wait(s_full + buf, ((uint32_t)(t / FW_BUFS)) & 1);
      tc::tc_fence_after_sync();
      const int j0 = (jt0 + t) * 128 + ch * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 128 + ch * 64 + c * 32), r);
        tc::tmem_ld_wait();
        const int nbase = j0 + c * 32;
        const int nvalid = Nb - nbase;   // columns >= Nb are TMA zero fill: excluded
        if (nvalid <= 0) continue;
        if (dj >= nbase && dj < nbase + 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nbase + j == dj) { diag = __uint_as_float(r[j]) * scale; has_diag = true; }
        }
        float cmax = -INFINITY;
        if (nvalid >= 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) cmax = fmaxf(cmax, __uint_as_float(r[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) cmax = fmaxf(cmax, __uint_as_float(r[j]));
        }
        const float mn = fmaxf(m, cmax * sl2);     // sl2 > 0: max commutes with the scaling
        float ps0 = 0.f, ps1 = 0.f;
        if (nvalid >= 32) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            ps0 += fast_exp2(fmaf(__uint_as_float(r[j]), sl2, -mn));
            ps1 += fast_exp2(fmaf(__uint_as_float(r[j + 1]), sl2, -mn));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) ps0 += fast_exp2(fmaf(__uint_as_float(r[j]), sl2, -mn));
        }
        l = l * fast_exp2(m - mn) + (ps0 + ps1);
        m = mn;
      }
      tc::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(s_empty + buf);
    }
    // merge the two column halves of every row (warps w and w+4 share a lane quarter)
    float* mrg = reinterpret_cast<float*>(tmem_slot + 4);     // [4][128]: m, l, diag, has_diag of half 1
    if (ch == 1) {
      mrg[rloc] = m;
      mrg[128 + rloc] = l;
      mrg[256 + rloc] = diag;
      mrg[384 + rloc] = has_diag ? 1.f : 0.f;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (ch == 0 && row < Ma) {
      const float m1 = mrg[rloc], l1 = mrg[128 + rloc];
      const float M = fmaxf(m, m1);
      float Lc = 0.f;
      if (m > -INFINITY) Lc += l * fast_exp2(m - M);
      if (m1 > -INFINITY) Lc += l1 * fast_exp2(m1 - M);
      part_max[(long long)blockIdx.y * Ma + row] = M;   // log2 domain
      part_sum[(long long)blockIdx.y * Ma + row] = Lc;
      if (diag_out) {
        if (has_diag) diag_out[row] = diag;
        else if (mrg[384 + rloc] != 0.f) diag_out[row] = mrg[256 + rloc];
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<512>(tmem_base);
}

// merge column-split partials; convert the log2-domain max back to natural units
__global__ void rowlse_combine_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum, int Ma,
                                      int nsplit, float* __restrict__ row_max, float* __restrict__ row_sum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Ma) return;
  float M = -INFINITY;
  for (int s = 0; s < nsplit; ++s) M = fmaxf(M, part_max[(long long)s * Ma + i]);
  float L = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float ms = part_max[(long long)s * Ma + i];
    if (ms > -INFINITY) L += part_sum[(long long)s * Ma + i] * exp2f(ms - M);
  }
  row_max[i] = M * kLn2;
  row_sum[i] = L;
}

// ------------------------------------------------------------------------------------------ backward
constexpr int BW_STAGES = 2;

// Swizzled (128B) K-major smem address of element (row, col) of a [128 x 64] bf16 tile, in bytes.
__device__ __forceinline__ uint32_t sw128_offset(int row, int col_bf16) {
  const uint32_t chunk = (uint32_t)(col_bf16 >> 3);             // 16-byte chunk within the 128 B row
  return (uint32_t)row * 128u + ((chunk ^ ((uint32_t)row & 7u)) << 4) + (uint32_t)(col_bf16 & 7) * 2u;
}

__global__ void __launch_bounds__(NT_THREADS, 1)
infonce_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmBT, int Ma, int Nb, int D, int num_kb, float scale,
                      const float* __restrict__ lseA, const float* __restrict__ lseB, float coef,
                      const float* __restrict__ gscale, long long diag_offset, const uint16_t* __restrict__ Bm,
                      long long ldb, float* __restrict__ dA, long long ldda, int accumulate) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smemA = smem;                                   // num_kb tiles (resident anchors)
  uint8_t* smemB = smemA + num_kb * NT_TILE;               // BW_STAGES tiles (S operand ring)
  uint8_t* smemP = smemB + BW_STAGES * NT_TILE;            // 2 tiles: W [128 i x 128 j] as two 64-wide k-blocks
  uint8_t* smemV = smemP + 2 * NT_TILE;                    // 2 tiles: BmT slice [128 d x 128 j]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemV + 2 * NT_TILE);
  uint64_t* a_full = bars;
  uint64_t* full_bar = bars + 1;
  uint64_t* empty_bar = full_bar + BW_STAGES;
  uint64_t* s_full = empty_bar + BW_STAGES;    // [2] S tile ready in TMEM buffer b
  uint64_t* p_full = s_full + 2;               // W tile written to smem (128 arrivals)
  uint64_t* v_full = p_full + 1;               // BmT slice landed
  uint64_t* pv_done = v_full + 1;              // second MMA retired: smemP / smemV reusable
  uint64_t* acc_full = pv_done + 1;            // all tiles accumulated
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int d0 = blockIdx.y * 128;
  const int ntiles = (Nb + 127) / 128;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    tc::tma_prefetch_desc(&tmBT);
    tc::mbar_init(a_full, 1);
    for (int s = 0; s < BW_STAGES; ++s) { tc::mbar_init(full_bar + s, 1); tc::mbar_init(empty_bar + s, 1); }
    tc::mbar_init(s_full + 0, 1);
    tc::mbar_init(s_full + 1, 1);
    tc::mbar_init(p_full, 128);
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
  const uint32_t tmem_O = tmem_base + 256;    // columns [256,384): dA slice accumulator; S buffers at [0,128),[128,256)

  if (warp == 0) {
    if (lane == 0) {
      tc::mbar_expect_tx(a_full, num_kb * NT_TILE);
      for (int kb = 0; kb < num_kb; ++kb) tc::tma_load_2d(smemA + kb * NT_TILE, &tmA, kb * 64, m0, a_full);
      int stage = 0;
      uint32_t phase = 0;
      auto load_b = [&](int t) {
        const int j0 = t * 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(empty_bar + stage, phase ^ 1);
          tc::mbar_expect_tx(full_bar + stage, NT_TILE);
          tc::tma_load_2d(smemB + stage * NT_TILE, &tmB, kb * 64, j0, full_bar + stage);
          if (++stage == BW_STAGES) { stage = 0; phase ^= 1; }
        }
      };
      // consumption order of the MMA thread: S(0), S(1), PV(0), S(2), PV(1), ...
      load_b(0);
      for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) load_b(t + 1);
        tc::mbar_wait(pv_done, ((uint32_t)t & 1) ^ 1);   // smemV free: PV(t-1) retired
        tc::mbar_expect_tx(v_full, 2 * NT_TILE);
        tc::tma_load_2d(smemV, &tmBT, t * 128, d0, v_full);
        tc::tma_load_2d(smemV + NT_TILE, &tmBT, t * 128 + 64, d0, v_full);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_bf16(128, 128, 0, 0);
      tc::mbar_wait(a_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      auto issue_s = [&](int t) {
        // buffer (t&1) was last read by the softmax warps for tile t-2; they arrived on p_full(t-2),
        // which this thread waited on before issuing PV(t-2).
        const uint32_t d_tmem = tmem_base + (uint32_t)((t & 1) * 128);
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(full_bar + stage, phase);
          tc::tc_fence_after_sync();
          const uint32_t a_addr = tc::smem_u32(smemA + kb * NT_TILE);
          const uint32_t b_addr = tc::smem_u32(smemB + stage * NT_TILE);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::umma_ss(d_tmem, tc::make_smem_desc(a_addr + k * 32, 16, 1024),
                        tc::make_smem_desc(b_addr + k * 32, 16, 1024), idesc, (kb | k) != 0 ? 1u : 0u);
          tc::umma_commit(empty_bar + stage);
          if (++stage == BW_STAGES) { stage = 0; phase ^= 1; }
        }
        tc::umma_commit(s_full + (t & 1));
      };
      issue_s(0);
      for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) issue_s(t + 1);
        // second MMA: dA_slice += W * BmT_slice^T   (K = the 128 columns j of tile t)
        tc::mbar_wait(p_full, (uint32_t)t & 1);
        tc::mbar_wait(v_full, (uint32_t)t & 1);
        tc::tc_fence_after_sync();
        const uint32_t p_addr = tc::smem_u32(smemP);
        const uint32_t v_addr = tc::smem_u32(smemV);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t off = (uint32_t)(k >> 2) * NT_TILE + (uint32_t)(k & 3) * 32;
          tc::umma_ss(tmem_O, tc::make_smem_desc(p_addr + off, 16, 1024), tc::make_smem_desc(v_addr + off, 16, 1024),
                      idesc, (t | k) != 0 ? 1u : 0u);
        }
        tc::umma_commit(pv_done);
      }
      tc::umma_commit(acc_full);
    }
  } else {
    const int q = warp & 3;
    const int rloc = q * 32 + lane;
    const int row = m0 + rloc;
    const float sl2 = scale * kLog2e;
    const float la2 = (row < Ma) ? lseA[row] * kLog2e : 0.f;
    for (int t = 0; t < ntiles; ++t) {
      const int j0 = t * 128;
      tc::mbar_wait(s_full + (t & 1), ((uint32_t)t >> 1) & 1);
      tc::mbar_wait(pv_done, ((uint32_t)t & 1) ^ 1);   // smemP free: PV(t-1) retired
      tc::tc_fence_after_sync();
      const uint32_t tmem_S = tmem_base + (uint32_t)((t & 1) * 128);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem_S + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
        tc::tmem_ld_wait();
        const int nbase = j0 + c * 32;
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float w0 = 0.f, w1 = 0.f;
          if (row < Ma) {
            if (nbase + j < Nb) {
              const float tv = __uint_as_float(r[j]) * sl2;
              w0 = fast_exp2(tv - la2) + fast_exp2(tv - __ldg(lseB + nbase + j) * kLog2e);
            }
            if (nbase + j + 1 < Nb) {
              const float tv = __uint_as_float(r[j + 1]) * sl2;
              w1 = fast_exp2(tv - la2) + fast_exp2(tv - __ldg(lseB + nbase + j + 1) * kLog2e);
            }
          }
          pk[j >> 1] = (uint32_t)f2bf(w0) | ((uint32_t)f2bf(w1) << 16);
        }
        // columns c*32 .. c*32+31 of W -> k-block (c>>1), local columns (c&1)*32 ..
        uint8_t* tile = smemP + (c >> 1) * NT_TILE;
        const int cb = (c & 1) * 32;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const uint32_t off = sw128_offset(rloc, cb + v * 8);
          *reinterpret_cast<uint4*>(tile + off) = make_uint4(pk[v * 4], pk[v * 4 + 1], pk[v * 4 + 2], pk[v * 4 + 3]);
        }
      }
      tc::fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the MMA (async proxy)
      tc::tc_fence_before_sync();
      tc::mbar_arrive(p_full);
    }
    // epilogue: dA slice
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after_sync();
    const float cg = coef * (gscale ? __ldg(gscale) : 1.0f);
    const long long pj = diag_offset >= 0 ? diag_offset + row : -1;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tc::tmem_ld_32x32(tmem_O + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
      tc::tmem_ld_wait();
      const int dbase = d0 + c * 32;
      if (row < Ma && dbase < D) {
        const int nvalid = min(32, D - dbase);
        float* dst = dA + (long long)row * ldda + dbase;
        for (int j = 0; j < nvalid; ++j) {
          float v = __uint_as_float(r[j]);
          if (pj >= 0 && pj < Nb) v -= 2.0f * bf2f(Bm[pj * ldb + dbase + j]);
          v *= cg;
          dst[j] = accumulate ? dst[j] + v : v;
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
                             cudaStream_t s);
int dmf_infonce_bwd_bf16_tc3(const void* A, long long lda, int Ma, const float* lseA, const void* Bm, long long ldb,
                             const void* BmT, long long ldbt, int Nb, const float* lseB, int D, float scale, float coef,
                             const float* gscale, long long diag_offset, float* dA, long long ldda, int accumulate,
                             cudaStream_t s);

int dmf_infonce_bwd_bf16_tc5(const void* A, long long lda, int Ma, const float* lseA, const void* Bm, long long ldb,
                             const void* BmT, long long ldbt, int Nb, const float* lseB, int D, float scale, float coef,
                             const float* gscale, long long diag_offset, float* dA, long long ldda, cudaStream_t s);

int dmf_infonce_bwd_bf16_tc6(const void* A, long long lda, int Ma, const float* lseA, const void* Bm, long long ldb,
                             const void* BmT, long long ldbt, int Nb, const float* lseB, int D, float scale, float coef,
                             const float* gscale, long long diag_offset, float* dA, long long ldda, int accumulate,
                             cudaStream_t s);

static int pick_nsplit(int row_blocks, int col_tiles) {
  int best = 1;
  double best_eff = 0.0;
  for (int ns = 1; ns <= 8; ++ns) {
    if (ns > 1 && col_tiles / ns < 8) break;
    const double waves = (double)row_blocks * ns / kNumSMs;
    const double eff = waves / (double)((long long)(waves + 0.999999));
    if (eff > best_eff + 0.02) { best_eff = eff; best = ns; }
  }
  return best;
}

// 1 when dmf_infonce_bwd (bf16) needs the transposed column block BmT for width D (the M = 128 pair kernel, D = 256 or
// 512 with a 16-byte aligned dA whose row pitch is a multiple of 4, reads Bm alone), 0 otherwise.
extern "C" int dmf_infonce_bwd_needs_transposed(int D) {
  if (getenv("DMF_BWD_V1") || getenv("DMF_BWD_V2") || getenv("DMF_BWD_V3")) return 1;
  return (D == 256 || D == 512) ? 0 : 1;
}

extern "C" size_t dmf_rowlse_workspace_bytes(int Ma, int Nb) {
  (void)Nb;
  return (size_t)8 * 2 * sizeof(float) * (size_t)(Ma > 0 ? Ma : 0);
}

int dmf_rowlse_bf16_tc2(const void* A, long long lda, int Ma, const void* Bm, long long ldb, int Nb, int D, float scale,
                        float* pm, float* ps, int nsplit, int tiles_per_split, long long diag_offset, float* diag_out,
                        cudaStream_t s);

int dmf_rowlse_bf16_tc(const void* A, long long lda, int Ma, const void* Bm, long long ldb, int Nb, int D, float scale,
                       float* row_max, float* row_sum, long long diag_offset, float* diag_out, void* workspace,
                       size_t workspace_bytes, cudaStream_t s) {
  DMF_REQUIRE(D % 64 == 0 && D >= 64 && D <= 64 * NT_MAX_KB, "dmf_rowlse(bf16): D=%d must be a multiple of 64 in [64,%d]", D,
              64 * NT_MAX_KB);
  static int use_v1 = -1;
  if (use_v1 < 0) use_v1 = getenv("DMF_FWD_V1") ? 1 : 0;
  if (!use_v1) {
    // CTA-pair kernel: 256-row anchor blocks x 256-column tiles
    const int pairs = (Ma + 255) / 256, col_tiles2 = (Nb + 255) / 256;
    int ns = pick_nsplit(2 * pairs, col_tiles2);
    const size_t need2 = (size_t)ns * 2 * sizeof(float) * (size_t)Ma;
    if (ns > 1 && (!workspace || workspace_bytes < need2)) ns = 1;
    const int tps = (col_tiles2 + ns - 1) / ns;
    float* pm2 = ns > 1 ? (float*)workspace : row_max;
    float* ps2 = ns > 1 ? (float*)workspace + (size_t)ns * Ma : row_sum;
    int rc2 = dmf_rowlse_bf16_tc2(A, lda, Ma, Bm, ldb, Nb, D, scale, pm2, ps2, ns, tps, diag_offset, diag_out, s);
    if (rc2) return rc2;
    rowlse_combine_kernel<<<(Ma + 255) / 256, 256, 0, s>>>(pm2, ps2, Ma, ns, row_max, row_sum);
    return launched("dmf_rowlse(bf16) combine");
  }
  const int num_kb = D / 64;
  const int row_blocks = (Ma + 127) / 128, col_tiles = (Nb + 127) / 128;
  int nsplit = pick_nsplit(row_blocks, col_tiles);
  const size_t need = (size_t)nsplit * 2 * sizeof(float) * (size_t)Ma;
  if (nsplit > 1 && (!workspace || workspace_bytes < need)) nsplit = 1;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16_2d(&tmA, A, Ma, D, lda, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, Bm, Nb, D, ldb, 128);
  if (rc) return rc;
  const size_t smem = 1024 + (size_t)(num_kb + FW_STAGES) * NT_TILE + 256 + 2048;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(rowlse_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(1024 + (size_t)(NT_MAX_KB + FW_STAGES) * NT_TILE + 256 + 2048));
    if (e != cudaSuccess) return fail((int)e, "dmf_rowlse(bf16): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  const int tiles_per_split = (col_tiles + nsplit - 1) / nsplit;
  float* pm = nsplit > 1 ? (float*)workspace : row_max;
  float* ps = nsplit > 1 ? (float*)workspace + (size_t)nsplit * Ma : row_sum;
  dim3 grid(row_blocks, nsplit);
  rowlse_tc_kernel<<<grid, FW_THREADS, smem, s>>>(tmA, tmB, Ma, Nb, num_kb, scale, tiles_per_split, pm, ps, diag_offset,
                                                  diag_out);
  rc = launched("dmf_rowlse(bf16)");
  if (rc) return rc;
  // log2-domain partials -> natural-unit (max, sum); also the nsplit == 1 case (in place)
  rowlse_combine_kernel<<<(Ma + 255) / 256, 256, 0, s>>>(pm, ps, Ma, nsplit, row_max, row_sum);
  return launched("dmf_rowlse(bf16) combine");
}

int dmf_infonce_bwd_bf16_tc(const void* A, long long lda, int Ma, const float* lseA, const void* Bm, long long ldb,
                            const void* BmT, long long ldbt, int Nb, const float* lseB, int D, float scale, float coef,
                            const float* gscale, long long diag_offset, float* dA, long long ldda, int accumulate,
                            cudaStream_t s) {
  DMF_REQUIRE(D % 64 == 0 && D >= 64 && D <= 64 * NT_MAX_KB, "dmf_infonce_bwd(bf16): D=%d must be a multiple of 64 in [64,%d]",
              D, 64 * NT_MAX_KB);
  {
    static int use_v1 = -1, use_v2 = -1, use_v3 = -1;
    if (use_v1 < 0) { use_v1 = getenv("DMF_BWD_V1") ? 1 : 0; use_v2 = getenv("DMF_BWD_V2") ? 1 : 0; use_v3 = getenv("DMF_BWD_V3") ? 1 : 0; }
    if (!use_v1 && !use_v2 && !use_v3) {
      // M = 128 pair kernel: S computed once per column tile, full output row resident in TMEM (executed = 2x algorithmic)
      const int rc6 = dmf_infonce_bwd_bf16_tc6(A, lda, Ma, lseA, Bm, ldb, BmT, ldbt, Nb, lseB, D, scale, coef, gscale,
                                               diag_offset, dA, ldda, accumulate, s);
      if (rc6 != -100) return rc6;
    }
    DMF_REQUIRE(BmT, "dmf_infonce_bwd(bf16): this shape (D not in {256, 512} or unaligned dA) needs the transposed column "
                     "block BmT [D, Nb]; see dmf_infonce_bwd_needs_transposed");
    if (!use_v1 && !use_v2 && !accumulate && D == 512) {
      // 4-CTA clusters: S recomputed once per row block, W exchanged between the two slice pairs over DSMEM
      const int rc5 = dmf_infonce_bwd_bf16_tc5(A, lda, Ma, lseA, Bm, ldb, BmT, ldbt, Nb, lseB, D, scale, coef, gscale,
                                               diag_offset, dA, ldda, s);
      if (rc5 != -100) return rc5;
    }
    if (!use_v1 && !use_v2)
      return dmf_infonce_bwd_bf16_tc3(A, lda, Ma, lseA, Bm, ldb, BmT, ldbt, Nb, lseB, D, scale, coef, gscale, diag_offset,
                                      dA, ldda, accumulate, s);
    if (!use_v1)
      return dmf_infonce_bwd_bf16_tc2(A, lda, Ma, lseA, Bm, ldb, BmT, ldbt, Nb, lseB, D, scale, coef, gscale, diag_offset,
                                      dA, ldda, accumulate, s);
  }
  const int num_kb = D / 64;
  CUtensorMap tmA, tmB, tmBT;
  int rc = make_tmap_bf16_2d(&tmA, A, Ma, D, lda, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, Bm, Nb, D, ldb, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmBT, BmT, D, Nb, ldbt, 128);
  if (rc) return rc;
  const size_t smem = 1024 + (size_t)(num_kb + BW_STAGES + 4) * NT_TILE + 256;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(infonce_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(1024 + (size_t)(NT_MAX_KB + BW_STAGES + 4) * NT_TILE + 256));
    if (e != cudaSuccess) return fail((int)e, "dmf_infonce_bwd(bf16): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  dim3 grid((Ma + 127) / 128, (D + 127) / 128);
  infonce_bwd_tc_kernel<<<grid, NT_THREADS, smem, s>>>(tmA, tmB, tmBT, Ma, Nb, D, num_kb, scale, lseA, lseB, coef, gscale,
                                                       diag_offset, (const uint16_t*)Bm, ldb, dA, ldda, accumulate);
  return launched("dmf_infonce_bwd(bf16)");
}
