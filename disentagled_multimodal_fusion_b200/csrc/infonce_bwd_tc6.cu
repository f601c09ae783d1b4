// K2 backward, "S once" generation: tcgen05.mma.cta_group::2 with M = 128 (64 anchor rows per CTA).
//
//   dA[i, :] = coef*g * ( sum_j W_ij b_j - 2 b_pos(i) ),   W_ij = exp(s_ij - lseA_i) + exp(s_ij - lseB_j)
//
// Why M = 128: tensor memory (512 columns x 128 lanes per SM) is what forced the pair kernel (infonce_bwd_tc3.cu,
// 128 anchor rows per CTA) to produce the D = 512 output in two 256-wide slices and to recompute S for each slice
// (executed = 3x algorithmic FLOPs).  With 64 rows per CTA the 2-SM accumulator layout puts row m of the CTA on
// lanes m (columns [0, N/2)) AND on lanes 64 + m (columns [N/2, N)), so a 64 x 512 fp32 output occupies 256
// columns and leaves 256 columns for a double-buffered 64 x 256 S tile: S is computed ONCE per column tile and
// the whole output row is accumulated in one pass (executed = 2x algorithmic: S + W*B).  The price is operand
// traffic: 128 anchor rows share every streamed column tile instead of 256.
//
// Per 256-column tile t and CTA rank r (j0 = first column, all MMAs M = 128 / N = 256 / K = 16):
//   S(t)   = anchors[64 x D] (smem, resident)  x  Bm[j0 + 128 r .. +128, D]^T       8 ring stages [128 j x 64 d]
//   W(t)   = softmax warps: TMEM -> exp2 -> bf16 -> smem (K-major, 128B swizzle), the A operand of
//   O     += W(t)[64 x 256]  x  Bm[j0 .. j0 + 256, d]      two N = 256 halves h    8 ring stages [128 j x 64 d]
// The second product reads Bm through the SAME tensor map as the first: its B operand is taken MN-major (the 64
// contiguous d of a tile row are the N dimension, the 128 tile rows j the K dimension; two adjacent stages = the
// CTA's 128 d of one output half), so no transposed copy of the column block is ever needed.
// One unified 8-stage ring of 16 KB operand tiles, consumed in the order S(0) S(1) PV(0) S(2) PV(1) ...
// W goes through shared memory (not TMEM): the TMEM A operand of a 2-SM M = 128 MMA must be duplicated on both
// lane halves, i.e. every softmax warp would have to hand its half tile to the warp of the opposite lane half.
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_pair.cuh"

namespace dmf {

constexpr int B6_THREADS = 320;
constexpr int B6_STAGE = 128 * 64 * 2;    // 16 KB: [128 x 64] bf16 operand tile
constexpr int B6_KB = 64 * 64 * 2;        // 8 KB: [64 rows x 64] bf16 (anchor / W k-block of this CTA)
constexpr int B6_STAGES = 8;
constexpr int B6_NT = 256;                // columns per tile
constexpr int B6_MAX_KB = 8;              // D <= 512
constexpr int B6_WBUF = 4 * B6_KB;        // one W tile: [64 x 256] bf16
constexpr int B6_SMEM_USED = B6_MAX_KB * B6_KB + B6_WBUF + B6_STAGES * B6_STAGE + 2 * B6_NT * 4 + 256;
constexpr int B6_SMEM = 232448;           // the whole opt-in window; the kernel checks that its carve-up fits
constexpr float kLog2e6 = 1.4426950408889634f;

__device__ __forceinline__ float ex2f6(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2_6(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// DBG (timing experiments, tools builds only): 1 = no softmax math / W stores, 2 = no PV MMAs, 4 = no S MMAs
// CL = CTAs per cluster: 2 = one pair; 4 = TWO pairs (256 anchor rows) that consume the SAME column-tile stream: every ring
// stage is fetched from L2 once and multicast into both pairs (the pairs take turns issuing), which halves the L2 -> SM
// operand traffic -- the pair kernel needs 64 B/clk/SM of it at full MMA rate, more than the L2 delivers chip-wide.
template <int DBG, int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(B6_THREADS, 1)
infonce_bwd_tc6_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int Ma, int Nb, int D, int num_kb, float scale,
                       const float* __restrict__ lseA, const float* __restrict__ lseB, float coef,
                       const float* __restrict__ gscale, long long diag_offset, const uint16_t* __restrict__ Bm,
                       long long ldb, float* __restrict__ dA, long long ldda, int accumulate) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  if (threadIdx.x == 0 && (smem - smem_raw) + B6_SMEM_USED > B6_SMEM) {
    printf("dmf: infonce_bwd_tc6: shared-memory carve-up does not fit (base misaligned by %d)\n", (int)(smem - smem_raw));
    __trap();
  }
  uint8_t* smemA = smem;                                     // B6_MAX_KB k-blocks [64 x 64] (resident anchors)
  uint8_t* smemW = smemA + B6_MAX_KB * B6_KB;                // ONE W tile: 4 k-blocks [64 rows x 64 j]
  uint8_t* ring = smemW + B6_WBUF;                       // B6_STAGES operand tiles [128 x 64]
  float* bsm = reinterpret_cast<float*>(ring + B6_STAGES * B6_STAGE);   // [2][256] column factors of the tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(bsm + 2 * B6_NT);
  uint64_t* a_full = bars;
  uint64_t* full_bar = bars + 1;
  uint64_t* empty_bar = full_bar + B6_STAGES;
  uint64_t* s_full = empty_bar + B6_STAGES;    // [2] S(t) complete in TMEM
  uint64_t* w_full = s_full + 2;               // [2] W(t) in shared memory, S buffer drained (16 warp arrivals)
  uint64_t* pv_done = w_full + 2;              // PV(t) retired: the W tile may be overwritten
  uint64_t* acc_full = pv_done + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = tc2::cluster_ctarank();
  const uint32_t rank = crank & 1;                           // rank inside the CTA pair (cta_group::2)
  const uint32_t pairi = crank >> 1;                         // pair inside the cluster (0 when CL == 2)
  const bool leader = rank == 0;
  const uint16_t pair_mask = (uint16_t)(3u << (2 * pairi));  // commit targets: both CTAs of this pair
  constexpr uint16_t all_mask = (uint16_t)((1u << CL) - 1);  // ring-stage release: every CTA of the cluster
  const uint16_t mc_mask = (uint16_t)((1u << rank) | (1u << (rank + 2)));   // CL == 4: same-rank CTA of both pairs
  const int nh = D >> 8;                                     // N = 256 halves of the output row (D = 256 or 512)
  const int m0 = ((int)blockIdx.x / CL) * (64 * CL) + (int)crank * 64;   // first anchor row of THIS CTA
  // column split (gridDim.z): this cluster covers tiles [tz0, tz0 + ntiles) and red.adds its partial rows
  const int total_tiles = (Nb + B6_NT - 1) / B6_NT;
  const int tiles_per_split = (total_tiles + (int)gridDim.z - 1) / (int)gridDim.z;
  const int tz0 = blockIdx.z * tiles_per_split;
  const int ntiles = max(0, min(total_tiles, tz0 + tiles_per_split) - tz0);
  const bool split = gridDim.z > 1 || accumulate != 0;   // red.add into dA (zeroed by the host unless accumulating)

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    tc::mbar_init(a_full, 1);
    for (int s = 0; s < B6_STAGES; ++s) { tc::mbar_init(full_bar + s, 1); tc::mbar_init(empty_bar + s, CL / 2); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(s_full + b, 1); tc::mbar_init(w_full + b, 16); }
    tc::mbar_init(pv_done, 1);
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc2::tmem_alloc2<512>(tmem_slot);
  tc::tc_fence_before_sync();
  tc2::cluster_sync_all();
  tc::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 256;     // S buffers at [0,128) and [128,256); O halves at [256,384), [384,512)

  if (warp == 0) {
    // TMA producer: whole warp, uniform control flow; one elected lane issues
    if (tc::elect_one()) {
      if (leader) tc::mbar_expect_tx(a_full, 2 * num_kb * B6_KB);          // bytes of BOTH CTAs
      for (int kb = 0; kb < num_kb; ++kb) tc2::tma_load_2d_pair(smemA + kb * B6_KB, &tmA, kb * 64, m0, a_full);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    auto load_s = [&](int t) {
      const int jrow = (tz0 + t) * B6_NT + (int)rank * 128;
      for (int kb = 0; kb < num_kb; ++kb) {
        tc::mbar_wait(empty_bar + stage, phase ^ 1);
        if (tc::elect_one()) {
          if (leader) tc::mbar_expect_tx(full_bar + stage, 2 * B6_STAGE);
          if (CL == 2) tc2::tma_load_2d_pair(ring + stage * B6_STAGE, &tmB, kb * 64, jrow, full_bar + stage);
          else if ((uint32_t)(stage & 1) == pairi)
            tc2::tma_load_2d_pair_mc(ring + stage * B6_STAGE, &tmB, kb * 64, jrow, full_bar + stage, mc_mask);
        }
        __syncwarp();
        if (++stage == B6_STAGES) { stage = 0; phase ^= 1; }
      }
    };
    auto load_v = [&](int t) {
      // PV operand: for each half jh of the tile rows and each output half h, the CTA's 128 d-columns as two
      // [128 j x 64 d] tiles in ADJACENT stages (stage index even: every phase uses an even number of stages)
      const int j0 = (tz0 + t) * B6_NT;
      for (int jh = 0; jh < 2; ++jh)
        for (int h = 0; h < nh; ++h)
          for (int c = 0; c < 2; ++c) {
            tc::mbar_wait(empty_bar + stage, phase ^ 1);
            if (tc::elect_one()) {
              // DBG & 8 (timing only): skip the block this CTA already fetched as an S operand (j half == rank, own d set)
              if (leader) tc::mbar_expect_tx(full_bar + stage, (DBG & 8) ? B6_STAGE : 2 * B6_STAGE);
              if ((DBG & 8) && jh == (int)rank) {
              } else if (CL == 2)
                tc2::tma_load_2d_pair(ring + stage * B6_STAGE, &tmB, h * 256 + (int)rank * 128 + c * 64, j0 + jh * 128,
                                      full_bar + stage);
              else if ((uint32_t)(stage & 1) == pairi)
                tc2::tma_load_2d_pair_mc(ring + stage * B6_STAGE, &tmB, h * 256 + (int)rank * 128 + c * 64, j0 + jh * 128,
                                         full_bar + stage, mc_mask);
            }
            __syncwarp();
            if (++stage == B6_STAGES) { stage = 0; phase ^= 1; }
          }
    };
    // consumption order of the MMA thread: S(0), S(1), PV(0), S(2), PV(1), ...
    if (ntiles > 0) load_s(0);
    for (int t = 0; t < ntiles; ++t) {
      if (t + 1 < ntiles) load_s(t + 1);
      load_v(t);
    }
  } else if (warp == 1) {
    if (leader) {
      // MMA issuer: whole warp, uniform control flow; one elected lane issues the tcgen05 instructions
      constexpr uint32_t idesc = tc::make_idesc_bf16(128, 256, 0, 0);
      constexpr uint32_t idesc_pv = tc::make_idesc_bf16(128, 256, 0, 1);        // B operand MN-major
      const uint64_t adesc0 = tc::make_smem_desc(tc::smem_u32(smemA), 16, 1024);
      const uint64_t wdesc0 = tc::make_smem_desc(tc::smem_u32(smemW), 16, 1024);
      const uint64_t rdesc0 = tc::make_smem_desc(tc::smem_u32(ring), 16, 1024);
      const uint64_t vdesc0 = tc::make_smem_desc(tc::smem_u32(ring), B6_STAGE, 1024);   // MN-major: LBO = next 64-d chunk
      tc::mbar_wait(a_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      auto issue_s = [&](int t) {
        const uint32_t d_tmem = tmem_base + (uint32_t)((t & 1) * 128);
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(full_bar + stage, phase);
          tc::tc_fence_after_sync();
          const uint64_t ad = adesc0 + (uint64_t)((kb * B6_KB) >> 4);
          const uint64_t bd = rdesc0 + (uint64_t)((stage * B6_STAGE) >> 4);
          if (tc::elect_one()) {
            if (!(DBG & 4)) {
#pragma unroll
              for (int k = 0; k < 4; ++k) tc2::umma_ss2(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            tc2::umma_commit2(empty_bar + stage, all_mask);
          }
          __syncwarp();
          if (++stage == B6_STAGES) { stage = 0; phase ^= 1; }
        }
        if (tc::elect_one()) tc2::umma_commit2(s_full + (t & 1), pair_mask);
        __syncwarp();
      };
      if (ntiles > 0) issue_s(0);
      for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) issue_s(t + 1);
        tc::mbar_wait(w_full + (t & 1), ((uint32_t)t >> 1) & 1);
        tc::tc_fence_after_sync();
        for (int jh = 0; jh < 2; ++jh)
          for (int h = 0; h < nh; ++h) {
            // two adjacent stages = [128 j] x [2 x 64 d] of this output half: MN-major B operand, LBO = stage pitch
            tc::mbar_wait(full_bar + stage, phase);
            tc::mbar_wait(full_bar + stage + 1, phase);
            tc::tc_fence_after_sync();
            const uint64_t vd = vdesc0 + (uint64_t)((stage * B6_STAGE) >> 4);
            if (tc::elect_one()) {
              if (!(DBG & 2)) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  // W columns j = 128 jh + 16 k .. +15: k-block 2 jh + (k >> 2), 32-byte step (k & 3) inside it;
                  // B rows j advance by 16 x 128 B = 2048 B per step
                  const uint64_t wd = wdesc0 + (uint64_t)((((2 * jh + (k >> 2)) * B6_KB) + (k & 3) * 32) >> 4);
                  tc2::umma_ss2(tmem_O + (uint32_t)(h * 128), wd, vd + (uint64_t)((k * 2048) >> 4), idesc_pv,
                                (t | jh | k) != 0 ? 1u : 0u);
                }
              }
              tc2::umma_commit2(empty_bar + stage, all_mask);
              tc2::umma_commit2(empty_bar + stage + 1, all_mask);
            }
            __syncwarp();
            stage += 2;
            if (stage == B6_STAGES) { stage = 0; phase ^= 1; }
          }
        if (tc::elect_one()) tc2::umma_commit2(pv_done, pair_mask);
        __syncwarp();
      }
      if (tc::elect_one()) tc2::umma_commit2(acc_full, pair_mask);
      __syncwarp();
    }
  } else {
    const int sw = warp - 2;                 // 0..7
    const int q = warp & 3;                  // TMEM lane quarter: q & 1 = row half, q >> 1 = column half of the tile
    const int ch = sw >> 2;                  // 64-column sub-half handled by this warp
    const int st = threadIdx.x - 64;         // 0..255 among softmax threads
    const int rloc = (q & 1) * 32 + lane;    // row within this CTA's 64
    const int row = m0 + rloc;
    const int cb = (q >> 1) * 128 + ch * 64; // first tile column of this warp
    const float sl2 = scale * kLog2e6;
    const float c0 = __ldg(lseB) * kLog2e6;                         // common exponent offset
    const float la2 = (row < Ma) ? __ldg(lseA + row) * kLog2e6 : c0;
    const float ai = ex2f6(la2 - c0);
    const uint32_t w_full_leader = tc2::mapa(tc::smem_u32(w_full), crank & ~1u);
    // this thread's row inside a W k-block (K-major, 128-byte rows, 16-byte chunks XOR-swizzled by row & 7)
    uint8_t* wrow = smemW + ((cb >> 6) * B6_KB) + rloc * 128;
    const int sx = rloc & 7;
    for (int t = 0; t < ntiles; ++t) {
      const int j0 = (tz0 + t) * B6_NT;
      float* bs = bsm + (t & 1) * B6_NT;
      {
        const int j = j0 + st;
        bs[st] = (j < Nb) ? ex2f6(c0 - __ldg(lseB + j) * kLog2e6) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      tc::mbar_wait(s_full + (t & 1), ((uint32_t)t >> 1) & 1);
      tc::tc_fence_after_sync();
      const uint32_t tS = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((t & 1) * 128 + ch * 64);
      // W(t) is formed in registers while PV(t-1) still reads the (single) W tile; it is stored once PV(t-1) retired
      uint32_t pk[32];
      if (!(DBG & 1)) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tc::tmem_ld_32x32(tS + (uint32_t)(c * 32), r);
          tc::tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(bs + cb + c * 32);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 bb = b4[g];
            const int i = g * 4;
            const float e0 = ex2f6(fmaf(__uint_as_float(r[i + 0]), sl2, -la2));
            const float e1 = ex2f6(fmaf(__uint_as_float(r[i + 1]), sl2, -la2));
            const float e2 = ex2f6(fmaf(__uint_as_float(r[i + 2]), sl2, -la2));
            const float e3 = ex2f6(fmaf(__uint_as_float(r[i + 3]), sl2, -la2));
            pk[c * 16 + g * 2 + 0] = pack_bf16x2_6(fmaf(e0 * ai, bb.x, e0), fmaf(e1 * ai, bb.y, e1));
            pk[c * 16 + g * 2 + 1] = pack_bf16x2_6(fmaf(e2 * ai, bb.z, e2), fmaf(e3 * ai, bb.w, e3));
          }
        }
      }
      tc::tc_fence_before_sync();
      if (t >= 1) tc::mbar_wait(pv_done, ((uint32_t)t - 1) & 1);
      if (!(DBG & 1)) {
#pragma unroll
        for (int cg4 = 0; cg4 < 8; ++cg4)
          *reinterpret_cast<uint4*>(wrow + ((cg4 ^ sx) << 4)) =
              make_uint4(pk[cg4 * 4 + 0], pk[cg4 * 4 + 1], pk[cg4 * 4 + 2], pk[cg4 * 4 + 3]);
      }
      tc::fence_proxy_async_smem();          // generic-proxy W stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) tc2::mbar_arrive_cluster(w_full_leader + (uint32_t)((t & 1) * 8));
    }
    // epilogue: lanes m / 64 + m hold row m; this warp stores 64 columns of each output half
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after_sync();
    const float cg = coef * (gscale ? __ldg(gscale) : 1.0f);
    const long long pj = diag_offset >= 0 ? diag_offset + row : -1;
    const bool has_pos = pj >= 0 && pj < Nb && blockIdx.z == 0;   // the positive term is added by split 0 only
#pragma unroll 1
    for (int hc = 0; hc < nh * 2; ++hc) {
      const int h = hc >> 1, c = hc & 1;
      uint32_t r[32];
      tc::tmem_ld_32x32(tmem_O + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 128 + ch * 64 + c * 32), r);
      tc::tmem_ld_wait();
      const int dbase = h * 256 + cb + c * 32;
      if (row < Ma && ntiles > 0) {
        float* dst = dA + (long long)row * ldda + dbase;
        const uint16_t* bp = has_pos ? Bm + pj * ldb + dbase : nullptr;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o;
          o.x = (__uint_as_float(r[j + 0]) - (has_pos ? 2.0f * bf2f(bp[j + 0]) : 0.f)) * cg;
          o.y = (__uint_as_float(r[j + 1]) - (has_pos ? 2.0f * bf2f(bp[j + 1]) : 0.f)) * cg;
          o.z = (__uint_as_float(r[j + 2]) - (has_pos ? 2.0f * bf2f(bp[j + 2]) : 0.f)) * cg;
          o.w = (__uint_as_float(r[j + 3]) - (has_pos ? 2.0f * bf2f(bp[j + 3]) : 0.f)) * cg;
          if (split)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
          else
            *reinterpret_cast<float4*>(dst + j) = o;
        }
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

// Returns -100 when the shape is not eligible (caller falls back to the pair kernel).
int dmf_infonce_bwd_bf16_tc6(const void* A, long long lda, int Ma, const float* lseA, const void* Bm, long long ldb,
                             const void* BmT, long long ldbt, int Nb, const float* lseB, int D, float scale, float coef,
                             const float* gscale, long long diag_offset, float* dA, long long ldda, int accumulate,
                             cudaStream_t s) {
  if (D != 256 && D != 512) return -100;
  if ((reinterpret_cast<uintptr_t>(dA) & 15) != 0 || (ldda & 3) != 0) return -100;
  const int num_kb = D / 64;
  (void)BmT; (void)ldbt;                                       // not needed: the second product reads Bm MN-major
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16_2d(&tmA, A, Ma, D, lda, 64);        // each CTA keeps 64 anchor rows
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, Bm, Nb, D, ldb, 128);          // [128 rows x 64 cols] boxes for both products
  if (rc) return rc;
  // cluster size: two pairs sharing one multicast column-tile stream unless the row set is a single pair's worth
  static int cl_env = -1;
  if (cl_env < 0) { const char* e = getenv("DMF_TC6_CL"); cl_env = e ? atoi(e) : 0; }
  // Measured (B = 65536, D = 512): CL = 4 is SLOWER, 6.43 ms against 5.75 ms, and so is its TMA-only variant (5.29 against
  // 4.17 ms; 33 resident 4-CTA clusters = 132 of 148 SMs).  The stream is bound by what one SM can TAKE IN (~58 B/clk/SM
  // with all MMAs compiled out), not by the L2 slices, so halving the L2 reads buys nothing.  CL = 4 stays opt-in.
  const int CL = (cl_env == 4 && Ma > 128) ? 4 : 2;
  int dbg = 0;
#ifdef DMF_TC6_DBG
  { const char* e = getenv("DMF_TC6_DBG"); dbg = e ? atoi(e) : 0; }   // tools builds only (wrong results)
#endif
  using kern_t = decltype(&infonce_bwd_tc6_kernel<0, 2>);
  kern_t kern = CL == 4 ? infonce_bwd_tc6_kernel<0, 4> : infonce_bwd_tc6_kernel<0, 2>;
#ifdef DMF_TC6_DBG
  switch (dbg) {
    case 1: kern = CL == 4 ? infonce_bwd_tc6_kernel<1, 4> : infonce_bwd_tc6_kernel<1, 2>; break;
    case 2: kern = CL == 4 ? infonce_bwd_tc6_kernel<2, 4> : infonce_bwd_tc6_kernel<2, 2>; break;
    case 4: kern = CL == 4 ? infonce_bwd_tc6_kernel<4, 4> : infonce_bwd_tc6_kernel<4, 2>; break;
    case 6: kern = CL == 4 ? infonce_bwd_tc6_kernel<6, 4> : infonce_bwd_tc6_kernel<6, 2>; break;
    case 7: kern = CL == 4 ? infonce_bwd_tc6_kernel<7, 4> : infonce_bwd_tc6_kernel<7, 2>; break;
    case 8: kern = infonce_bwd_tc6_kernel<8, 2>; break;      // CL = 2 only: duplicate PV operand blocks not fetched
    case 15: kern = infonce_bwd_tc6_kernel<15, 2>; break;
    default: break;
  }
#endif
  // resident clusters (set the attribute, then ask the occupancy API once per variant)
  static int slots_tab[2][16] = {{0}};
  int& slots = slots_tab[CL == 4][dbg & 15];
  if (!slots) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, B6_SMEM);
    if (e != cudaSuccess) return fail((int)e, "dmf_infonce_bwd(bf16 m128): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL * 64, 1, 1);
    cfg.blockDim = dim3(B6_THREADS, 1, 1);
    cfg.dynamicSmemBytes = B6_SMEM;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = 0;
    e = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
    if (e != cudaSuccess || nc <= 0) { cudaGetLastError(); nc = CL == 4 ? 32 : 74; }
    slots = nc;
    if (getenv("DMF_VERBOSE")) fprintf(stderr, "dmf: infonce_bwd m128 CL=%d: %d resident clusters\n", CL, nc);
  }
  const int blocks = (Ma + 64 * CL - 1) / (64 * CL);
  // Column split: when the row blocks do not fill whole waves of the resident clusters (small local batches of a
  // data-parallel run), split the column set over gridDim.z; partial rows are accumulated with red.add.
  int nsplit = 1;
  const int total_tiles = (Nb + B6_NT - 1) / B6_NT;
  {
    double best = (double)blocks / (double)(((blocks + slots - 1) / slots) * slots);
    for (int ns = 2; ns <= 16 && best < 0.97; ++ns) {
      if (total_tiles / ns < 8) break;
      const int items = blocks * ns;
      const double eff = (double)items / (double)(((items + slots - 1) / slots) * slots);
      if (eff > best + 0.03) { best = eff; nsplit = ns; }
    }
    const int tps = (total_tiles + nsplit - 1) / nsplit;
    nsplit = (total_tiles + tps - 1) / tps;
  }
  if (nsplit > 1 && !accumulate) {
    cudaError_t e = cudaMemset2DAsync(dA, (size_t)ldda * sizeof(float), 0, (size_t)D * sizeof(float), (size_t)Ma, s);
    if (e != cudaSuccess) return fail((int)e, "dmf_infonce_bwd(bf16 m128): memset: %s", cudaGetErrorString(e));
  }
  dim3 grid(CL * blocks, 1, nsplit);
  kern<<<grid, B6_THREADS, B6_SMEM, s>>>(tmA, tmB, Ma, Nb, D, num_kb, scale, lseA, lseB, coef, gscale, diag_offset,
                                         (const uint16_t*)Bm, ldb, dA, ldda, accumulate);
  return launched("dmf_infonce_bwd(bf16 m128)");
}
