// K2 backward, CTA-pair generation (tcgen05.mma.cta_group::2, M = 256 across two SMs).
//
//   dA[i, :] = coef*g * ( sum_j W_ij b_j - 2 b_pos(i) ),   W_ij = exp(s_ij - lseA_i) + exp(s_ij - lseB_j)
//
// Same dataflow as infonce_bwd_tc2.cu (W stays in TMEM as the A operand of the second MMA, one exp2 per
// element), but a pair of CTAs owns 256 anchor rows: each CTA keeps ITS 128 rows resident, loads half of
// every 128-row column tile (8 KB k-blocks, 8-stage ring) and half (128 of 256 d-rows) of the transposed
// slice, so the output slice can be 256 wide (S is recomputed 2x instead of 4x for D = 512) while smem
// operand traffic per tensor-core cycle and L2 traffic per FLOP both drop.  MMA issue: leader CTA only;
// full barriers live in the leader (TMA of both CTAs completes them), consumer-release barriers are
// multicast-committed to both CTAs, softmax warps of both CTAs arrive on the leader's p_full.
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_pair.cuh"

namespace dmf {

constexpr int B3_THREADS = 320;
constexpr int B3_TILE = 128 * 64 * 2;   // 16 KB: [128 x 64] bf16
constexpr int B3_HALF = 64 * 64 * 2;    // 8 KB: this CTA's half of a column k-block
constexpr int B3_STAGES = 8;
constexpr int B3_OW = 256;              // output slice width of the pair
constexpr float kLog2e3 = 1.4426950408889634f;

__device__ __forceinline__ float ex2f3(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2_3(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(B3_THREADS, 1)
infonce_bwd_tc3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmBT, int Ma, int Nb, int D, int num_kb, float scale,
                       const float* __restrict__ lseA, const float* __restrict__ lseB, float coef,
                       const float* __restrict__ gscale, long long diag_offset, const uint16_t* __restrict__ Bm,
                       long long ldb, float* __restrict__ dA, long long ldda, int accumulate, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smemA = smem;                                   // num_kb tiles (resident anchors)
  uint8_t* smemB = smemA + num_kb * B3_TILE;               // B3_STAGES half k-blocks (S operand ring)
  uint8_t* smemV = smemB + B3_STAGES * B3_HALF;            // 2 tiles: this CTA's 128 d-rows of the BmT slice x 128 j
  float* bsm = reinterpret_cast<float*>(smemV + 2 * B3_TILE);   // [2][128] per-tile column factors b_j
  uint64_t* bars = reinterpret_cast<uint64_t*>(bsm + 256);
  uint64_t* a_full = bars;
  uint64_t* full_bar = bars + 1;
  uint64_t* empty_bar = full_bar + B3_STAGES;
  uint64_t* s_full = empty_bar + B3_STAGES;    // [2]
  uint64_t* p_full = s_full + 2;               // [2] W(t) stored in TMEM (8 warp arrivals)
  uint64_t* v_full = p_full + 2;
  uint64_t* pv_done = v_full + 1;
  uint64_t* acc_full = pv_done + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc2::cluster_ctarank();
  const bool leader = rank == 0;
  const int m0 = blockIdx.x * 128;
  const int d0 = blockIdx.y * B3_OW;
  // column split (gridDim.z): this cluster covers tiles [tz0, tz0 + ntiles) of the column set and, when the
  // set is split, accumulates its partial dA slice with red.add (the caller zeroes dA)
  const int total_tiles = (Nb + 127) / 128;
  const int tiles_per_split = (total_tiles + (int)gridDim.z - 1) / (int)gridDim.z;
  const int tz0 = blockIdx.z * tiles_per_split;
  const int ntiles = max(0, min(total_tiles, tz0 + tiles_per_split) - tz0);
  const bool split = gridDim.z > 1;
  // All clusters walk the column tiles in the SAME order: the 74 resident clusters then stay roughly in lockstep
  // and re-use each other's tiles out of L2 (Bm + BmT = 128 MB at B = 65 536 does not fit the 126 MB L2; a
  // per-cluster rotation of the start tile measured 6 % slower: 13.07 vs 12.29 ms).  DMF_BWD_DBG=8 rotates.
  const int trot = (!(dbg & 8) || ntiles == 0) ? 0 : (int)(((long long)(blockIdx.x >> 1) * 37 + blockIdx.y * 17) % ntiles);
  auto tile_of = [&](int t) { int x = t + trot; return tz0 + (x >= ntiles ? x - ntiles : x); };

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    tc::tma_prefetch_desc(&tmBT);
    tc::mbar_init(a_full, 1);
    for (int s = 0; s < B3_STAGES; ++s) { tc::mbar_init(full_bar + s, 1); tc::mbar_init(empty_bar + s, 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(s_full + b, 1); tc::mbar_init(p_full + b, 16); }
    tc::mbar_init(v_full, 1);
    tc::mbar_init(pv_done, 1);
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc2::tmem_alloc2<512>(tmem_slot);
  tc::tc_fence_before_sync();
  tc2::cluster_sync_all();
  tc::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 256;   // S buffers at [0,128) and [128,256)

  if (warp == 0) {
    {
      // whole warp, uniform control flow; one elected lane issues the TMA instructions
      if (tc::elect_one()) {
        if (leader) tc::mbar_expect_tx(a_full, 2 * num_kb * B3_TILE);      // bytes of BOTH CTAs
        for (int kb = 0; kb < num_kb; ++kb) tc2::tma_load_2d_pair(smemA + kb * B3_TILE, &tmA, kb * 64, m0, a_full);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      auto load_b = [&](int t) {
        const int jrow = tile_of(t) * 128 + (int)rank * 64;
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(empty_bar + stage, phase ^ 1);
          if (tc::elect_one()) {
            if (leader) tc::mbar_expect_tx(full_bar + stage, 2 * B3_HALF);
            tc2::tma_load_2d_pair(smemB + stage * B3_HALF, &tmB, kb * 64, jrow, full_bar + stage);
          }
          __syncwarp();
          if (++stage == B3_STAGES) { stage = 0; phase ^= 1; }
        }
      };
      // consumption order of the MMA thread: S(0), S(1), PV(0), S(2), PV(1), ...
      if (ntiles > 0) load_b(0);
      for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) load_b(t + 1);
        tc::mbar_wait(pv_done, ((uint32_t)t & 1) ^ 1);   // smemV free: PV(t-1) retired
        if (tc::elect_one()) {
          if (leader) tc::mbar_expect_tx(v_full, 4 * B3_TILE);
          tc2::tma_load_2d_pair(smemV, &tmBT, tile_of(t) * 128, d0 + (int)rank * 128, v_full);
          tc2::tma_load_2d_pair(smemV + B3_TILE, &tmBT, tile_of(t) * 128 + 64, d0 + (int)rank * 128, v_full);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // whole warp, uniform control flow; one elected lane issues the tcgen05 instructions
      constexpr uint32_t idesc_s = tc::make_idesc_bf16(256, 128, 0, 0);
      constexpr uint32_t idesc_o = tc::make_idesc_bf16(256, B3_OW, 0, 0);
      const uint64_t adesc0 = tc::make_smem_desc(tc::smem_u32(smemA), 16, 1024);
      const uint64_t bdesc0 = tc::make_smem_desc(tc::smem_u32(smemB), 16, 1024);
      const uint64_t vdesc0 = tc::make_smem_desc(tc::smem_u32(smemV), 16, 1024);
      tc::mbar_wait(a_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      auto issue_s = [&](int t) {
        // buffer (t&1) holds W(t-2), consumed by PV(t-2) which was issued earlier on this (in-order) pipe
        const uint32_t d_tmem = tmem_base + (uint32_t)((t & 1) * 128);
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(full_bar + stage, phase);
          tc::tc_fence_after_sync();
          // descriptor start-address field is (byte address >> 4): advance by whole tiles / 32-byte K steps
          const uint64_t ad = adesc0 + (uint64_t)((kb * B3_TILE) >> 4);
          const uint64_t bd = bdesc0 + (uint64_t)((stage * B3_HALF) >> 4);
          if (tc::elect_one()) {
            if (!(dbg & 4)) {
#pragma unroll
              for (int k = 0; k < 4; ++k) tc2::umma_ss2(d_tmem, ad + 2 * k, bd + 2 * k, idesc_s, (kb | k) != 0 ? 1u : 0u);
            }
            tc2::umma_commit2(empty_bar + stage);
          }
          __syncwarp();
          if (++stage == B3_STAGES) { stage = 0; phase ^= 1; }
        }
        if (tc::elect_one()) tc2::umma_commit2(s_full + (t & 1));
        __syncwarp();
      };
      issue_s(0);
      for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) issue_s(t + 1);
        tc::mbar_wait(p_full + (t & 1), ((uint32_t)t >> 1) & 1);
        tc::mbar_wait(v_full, (uint32_t)t & 1);
        tc::tc_fence_after_sync();
        const uint32_t w_tmem = tmem_base + (uint32_t)((t & 1) * 128);
        if (tc::elect_one()) {
          if (!(dbg & 2)) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              // W columns j = 16k..16k+15: column half (k>>2) keeps its packed pairs at +64*(k>>2) + 8*(k&3)
              const uint32_t a_t = w_tmem + (uint32_t)((k >> 2) * 64 + (k & 3) * 8);
              const uint64_t vd = vdesc0 + (uint64_t)((((k >> 2) * B3_TILE) + (k & 3) * 32) >> 4);
              tc2::umma_ts2(tmem_O, a_t, vd, idesc_o, (t | k) != 0 ? 1u : 0u);
            }
          }
          tc2::umma_commit2(pv_done);
        }
        __syncwarp();
      }
      if (tc::elect_one()) tc2::umma_commit2(acc_full);
      __syncwarp();
    }
  } else {
    const int sw = warp - 2;                 // 0..7
    const int q = warp & 3;                  // TMEM lane quarter
    const int ch = sw >> 2;                  // column half of the tile handled by this warp
    const int st = threadIdx.x - 64;         // 0..255 among softmax threads
    const int rloc = q * 32 + lane;
    const int row = m0 + rloc;
    const float sl2 = scale * kLog2e3;
    const float c0 = __ldg(lseB) * kLog2e3;                         // common exponent offset
    const float la2 = (row < Ma) ? __ldg(lseA + row) * kLog2e3 : c0;
    const float ai = ex2f3(la2 - c0);
    const uint32_t p_full_leader = tc2::mapa(tc::smem_u32(p_full), 0);
    for (int t = 0; t < ntiles; ++t) {
      const int j0 = tile_of(t) * 128;
      float* bs = bsm + (t & 1) * 128;
      if (st < 128) {
        const int j = j0 + st;
        bs[st] = (j < Nb) ? ex2f3(c0 - __ldg(lseB + j) * kLog2e3) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      tc::mbar_wait(s_full + (t & 1), ((uint32_t)t >> 1) & 1);
      tc::tc_fence_after_sync();
      const uint32_t tS = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((t & 1) * 128 + ch * 64);
      if (!(dbg & 1))
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
          const float e0 = ex2f3(fmaf(__uint_as_float(r[j4 * 4 + 0]), sl2, -la2));
          const float e1 = ex2f3(fmaf(__uint_as_float(r[j4 * 4 + 1]), sl2, -la2));
          const float e2 = ex2f3(fmaf(__uint_as_float(r[j4 * 4 + 2]), sl2, -la2));
          const float e3 = ex2f3(fmaf(__uint_as_float(r[j4 * 4 + 3]), sl2, -la2));
          const float w0 = fmaf(e0 * ai, bb.x, e0);
          const float w1 = fmaf(e1 * ai, bb.y, e1);
          const float w2 = fmaf(e2 * ai, bb.z, e2);
          const float w3 = fmaf(e3 * ai, bb.w, e3);
          pk[j4 * 2 + 0] = pack_bf16x2_3(w0, w1);
          pk[j4 * 2 + 1] = pack_bf16x2_3(w2, w3);
        }
        tc::tmem_st_32x16(tS + (uint32_t)(c * 16), pk);   // in place over the columns this thread has consumed
      }
      tc::tmem_st_wait();
      tc::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) tc2::mbar_arrive_cluster(p_full_leader + (uint32_t)((t & 1) * 8));
    }
    // epilogue: this warp stores its lane quarter x column half of the dA slice
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after_sync();
    const float cg = coef * (gscale ? __ldg(gscale) : 1.0f);
    const long long pj = diag_offset >= 0 ? diag_offset + row : -1;
#pragma unroll 1
    for (int c = 0; c < B3_OW / 64; ++c) {
      uint32_t r[32];
      const int cc = ch * (B3_OW / 64) + c;
      tc::tmem_ld_32x32(tmem_O + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 32), r);
      tc::tmem_ld_wait();
      const int dbase = d0 + cc * 32;
      if (row < Ma && dbase < D) {
        const int nvalid = min(32, D - dbase);
        float* dst = dA + (long long)row * ldda + dbase;
        const bool has_pos = pj >= 0 && pj < Nb && blockIdx.z == 0;   // the positive term is added by split 0 only
        const uint16_t* bp = has_pos ? Bm + pj * ldb + dbase : nullptr;
        if (nvalid == 32 && !accumulate && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
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
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < nvalid) {
              float v = __uint_as_float(r[j]);
              if (has_pos) v -= 2.0f * bf2f(bp[j]);
              v *= cg;
              if (split) atomicAdd(dst + j, v);
              else dst[j] = accumulate ? dst[j] + v : v;
            }
          }
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

int dmf_infonce_bwd_bf16_tc3(const void* A, long long lda, int Ma, const float* lseA, const void* Bm, long long ldb,
                             const void* BmT, long long ldbt, int Nb, const float* lseB, int D, float scale, float coef,
                             const float* gscale, long long diag_offset, float* dA, long long ldda, int accumulate,
                             cudaStream_t s) {
  const int num_kb = D / 64;
  CUtensorMap tmA, tmB, tmBT;
  int rc = make_tmap_bf16_2d(&tmA, A, Ma, D, lda, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, Bm, Nb, D, ldb, 64);          // each CTA loads 64 of the 128 tile rows
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmBT, BmT, D, Nb, ldbt, 128);      // each CTA loads 128 of the 256 slice rows
  if (rc) return rc;
  const size_t smem = 1024 + (size_t)num_kb * B3_TILE + (size_t)B3_STAGES * B3_HALF + 2 * B3_TILE + 1024 + 256;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(infonce_bwd_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(1024 + (size_t)8 * B3_TILE + (size_t)B3_STAGES * B3_HALF + 2 * B3_TILE + 1024 + 256));
    if (e != cudaSuccess) return fail((int)e, "dmf_infonce_bwd(bf16 pair): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  const int pairs = (Ma + 255) / 256;
  const int slices = (D + B3_OW - 1) / B3_OW;
  // Column split: when pairs * slices does not fill whole waves of the 74 resident clusters (small local batches
  // of a data-parallel run), split the column set over gridDim.z; partial slices are accumulated with red.add.
  int nsplit = 1;
  const int total_tiles = (Nb + 127) / 128;
  if (!accumulate) {
    const int base = pairs * slices;
    double best = (double)base / (double)(((base + 73) / 74) * 74);
    for (int ns = 2; ns <= 16 && best < 0.97; ++ns) {
      if (total_tiles / ns < 16) break;
      const int items = base * ns;
      const double eff = (double)items / (double)(((items + 73) / 74) * 74);
      if (eff > best + 0.03) { best = eff; nsplit = ns; }
    }
    const int tps = (total_tiles + nsplit - 1) / nsplit;
    nsplit = (total_tiles + tps - 1) / tps;
  }
  if (nsplit > 1) {
    cudaError_t e = cudaMemset2DAsync(dA, (size_t)ldda * sizeof(float), 0, (size_t)D * sizeof(float), (size_t)Ma, s);
    if (e != cudaSuccess) return fail((int)e, "dmf_infonce_bwd(bf16 pair): memset: %s", cudaGetErrorString(e));
  }
  dim3 grid(2 * pairs, slices, nsplit);
  int dbg = 0;
#ifdef DMF_BWD_DBG_BUILD   // timing experiments only (wrong results): never compiled into the shipped library
  { const char* e = getenv("DMF_BWD_DBG"); dbg = e ? atoi(e) : 0; }
#endif
  infonce_bwd_tc3_kernel<<<grid, B3_THREADS, smem, s>>>(tmA, tmB, tmBT, Ma, Nb, D, num_kb, scale, lseA, lseB, coef, gscale,
                                                        diag_offset, (const uint16_t*)Bm, ldb, dA, ldda, accumulate, dbg);
  return launched("dmf_infonce_bwd(bf16 pair)");
}
