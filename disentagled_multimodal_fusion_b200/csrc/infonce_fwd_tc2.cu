// K2 forward, CTA-pair generation: rowlse with tcgen05.mma.cta_group::2 (M = 256 across two SMs).
//
// ncu on the single-CTA kernel (profiles/r01_ncu_rowlse_v1.txt) showed the tensor pipe "active" ~79% of
// the time while delivering ~37% of peak: a 128x128x16 SS-mode MMA pulls 8 KB of operands from shared
// memory per 64-cycle instruction = 128 B/clk, the whole smem port.  Pairing two CTAs fixes the ratio:
// each CTA keeps ITS 128 anchor rows resident (A half of the M=256 tile) and loads only HALF of every
// 256-row column tile; the pair's tensor cores read the other half across the TPC.  Per CTA and per
// 128-cycle 256x256x16 instruction that is 4 KB (A) + 4 KB (B half) = 64 B/clk of smem, and half the L2
// traffic per FLOP.
//
// Cluster = 2 CTAs (rank 0 = leader).  Per CTA: warp 0 TMA producer (own A block, own half of each B
// tile; completes the LEADER's full barriers), warp 1 TMEM alloc (+ MMA issue in the leader),
// warps 2-9 softmax on the CTA's own 128 TMEM lanes (S tiles double-buffered: 2 x 256 columns).
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_pair.cuh"

namespace dmf {

constexpr int F2_THREADS = 320;
constexpr int F2_TILE = 128 * 64 * 2;   // 16 KB: [128 rows x 64 bf16]
constexpr int F2_STAGES = 5;
constexpr int F2_BN = 256;              // column tile of the pair
constexpr float kLog2eF = 1.4426950408889634f;



__device__ __forceinline__ float f2_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(F2_THREADS, 1)
rowlse_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int Ma, int Nb,
                  int num_kb, float scale, int tiles_per_split, float* __restrict__ part_max,
                  float* __restrict__ part_sum, long long diag_offset, float* __restrict__ diag_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smemA = smem;                                  // num_kb tiles: this CTA's 128 anchor rows
  uint8_t* smemB = smem + num_kb * F2_TILE;               // F2_STAGES tiles: this CTA's half of the column tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemB + F2_STAGES * F2_TILE);
  uint64_t* a_full = bars;                    // leader: both A blocks landed
  uint64_t* full_bar = bars + 1;              // leader: both halves of a stage landed
  uint64_t* empty_bar = full_bar + F2_STAGES; // per CTA: stage consumed (multicast commit)
  uint64_t* s_full = empty_bar + F2_STAGES;   // per CTA [2]: S tile ready (multicast commit)
  uint64_t* s_empty = s_full + 2;             // leader [2]: 16 softmax warps of the pair drained the buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 2);
  float* mrg = reinterpret_cast<float*>(tmem_slot + 4);   // [4][128] merge scratch

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc2::cluster_ctarank();
  const bool leader = rank == 0;
  const int m0 = blockIdx.x * 128;
  const int total_tiles = (Nb + F2_BN - 1) / F2_BN;
  const int jt0 = blockIdx.y * tiles_per_split;
  const int ntiles = max(0, min(total_tiles, jt0 + tiles_per_split) - jt0);

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    tc::mbar_init(a_full, 1);
    for (int s = 0; s < F2_STAGES; ++s) { tc::mbar_init(full_bar + s, 1); tc::mbar_init(empty_bar + s, 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(s_full + b, 1); tc::mbar_init(s_empty + b, 16); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc2::tmem_alloc2<512>(tmem_slot);
  tc::tc_fence_before_sync();
  tc2::cluster_sync_all();
  tc::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (ntiles > 0) {
      // whole warp, uniform control flow; one elected lane issues the TMA instructions
      if (tc::elect_one()) {
        if (leader) tc::mbar_expect_tx(a_full, 2 * num_kb * F2_TILE);        // bytes of BOTH CTAs
        for (int kb = 0; kb < num_kb; ++kb) tc2::tma_load_2d_pair(smemA + kb * F2_TILE, &tmA, kb * 64, m0, a_full);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int j0 = (jt0 + t) * F2_BN + (int)rank * 128;               // this CTA's half of the tile
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(empty_bar + stage, phase ^ 1);
          if (tc::elect_one()) {
            if (leader) tc::mbar_expect_tx(full_bar + stage, 2 * F2_TILE);
            tc2::tma_load_2d_pair(smemB + stage * F2_TILE, &tmB, kb * 64, j0, full_bar + stage);
          }
          __syncwarp();
          if (++stage == F2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && ntiles > 0) {
      // whole warp, uniform control flow; one elected lane issues the tcgen05 instructions
      constexpr uint32_t idesc = tc::make_idesc_bf16(256, F2_BN, 0, 0);
      const uint64_t adesc0 = tc::make_smem_desc(tc::smem_u32(smemA), 16, 1024);
      const uint64_t bdesc0 = tc::make_smem_desc(tc::smem_u32(smemB), 16, 1024);
      tc::mbar_wait(a_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        tc::mbar_wait(s_empty + buf, (((uint32_t)t >> 1) & 1) ^ 1);
        tc::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * F2_BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(full_bar + stage, phase);
          tc::tc_fence_after_sync();
          const uint64_t ad = adesc0 + (uint64_t)((kb * F2_TILE) >> 4);
          const uint64_t bd = bdesc0 + (uint64_t)((stage * F2_TILE) >> 4);
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) tc2::umma_ss2(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            tc2::umma_commit2(empty_bar + stage);
          }
          __syncwarp();
          if (++stage == F2_STAGES) { stage = 0; phase ^= 1; }
        }
        if (tc::elect_one()) tc2::umma_commit2(s_full + buf);
        __syncwarp();
      }
    }
  } else {
    const int sw = warp - 2;
    const int q = warp & 3;
    const int ch = sw >> 2;                      // column half (128 columns) of every 256-wide tile
    const int rloc = q * 32 + lane;
    const int row = m0 + rloc;
    const float sl2 = scale * kLog2eF;
    float m = -INFINITY, l = 0.f, diag = 0.f;
    bool has_diag = false;
    const long long dj = (diag_offset >= 0 && row < Ma) ? diag_offset + row : -1;
    const uint32_t s_empty_leader0 = tc2::mapa(tc::smem_u32(s_empty), 0);
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      tc::mbar_wait(s_full + buf, ((uint32_t)t >> 1) & 1);
      tc::tc_fence_after_sync();
      const int j0 = (jt0 + t) * F2_BN + ch * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * F2_BN + ch * 128 + c * 32), r);
        tc::tmem_ld_wait();
        const int nbase = j0 + c * 32;
        const int nvalid = Nb - nbase;
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
        const float mn = fmaxf(m, cmax * sl2);
        float ps0 = 0.f, ps1 = 0.f;
        if (nvalid >= 32) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            ps0 += f2_exp2(fmaf(__uint_as_float(r[j]), sl2, -mn));
            ps1 += f2_exp2(fmaf(__uint_as_float(r[j + 1]), sl2, -mn));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) ps0 += f2_exp2(fmaf(__uint_as_float(r[j]), sl2, -mn));
        }
        l = l * f2_exp2(m - mn) + (ps0 + ps1);
        m = mn;
      }
      tc::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) tc2::mbar_arrive_cluster(s_empty_leader0 + (uint32_t)(buf * 8));
    }
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
      if (m > -INFINITY) Lc += l * f2_exp2(m - M);
      if (m1 > -INFINITY) Lc += l1 * f2_exp2(m1 - M);
      part_max[(long long)blockIdx.y * Ma + row] = M;   // log2 domain
      part_sum[(long long)blockIdx.y * Ma + row] = Lc;
      if (diag_out) {
        if (has_diag) diag_out[row] = diag;
        else if (mrg[384 + rloc] != 0.f) diag_out[row] = mrg[256 + rloc];
      }
    }
  }
  __syncwarp();
  tc::tc_fence_before_sync();
  tc2::cluster_sync_all();          // the peer may still target this CTA's barriers / TMEM until here
  if (warp == 1) tc2::tmem_dealloc2<512>(tmem_base);
}

}  // namespace dmf

using namespace dmf;

int dmf_rowlse_bf16_tc2(const void* A, long long lda, int Ma, const void* Bm, long long ldb, int Nb, int D, float scale,
                        float* pm, float* ps, int nsplit, int tiles_per_split, long long diag_offset, float* diag_out,
                        cudaStream_t s) {
  const int num_kb = D / 64;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16_2d(&tmA, A, Ma, D, lda, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, Bm, Nb, D, ldb, 128);
  if (rc) return rc;
  const size_t smem = 1024 + (size_t)(num_kb + F2_STAGES) * F2_TILE + 256 + 2048;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(rowlse_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(1024 + (size_t)(8 + F2_STAGES) * F2_TILE + 256 + 2048));
    if (e != cudaSuccess) return fail((int)e, "dmf_rowlse(bf16 pair): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  const int pairs = (Ma + 255) / 256;
  dim3 grid(2 * pairs, nsplit);
  rowlse_tc2_kernel<<<grid, F2_THREADS, smem, s>>>(tmA, tmB, Ma, Nb, num_kb, scale, tiles_per_split, pm, ps, diag_offset,
                                                   diag_out);
  return launched("dmf_rowlse(bf16 pair)");
}
