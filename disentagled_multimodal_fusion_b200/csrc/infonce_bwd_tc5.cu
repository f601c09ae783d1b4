// K2 backward, 4-CTA cluster generation: the similarity tiles are recomputed ONCE per row block instead of once
// per 256-wide output slice.
//
//   dA[i, :] = coef*g * ( sum_j W_ij b_j - 2 b_pos(i) ),   W_ij = exp(s_ij - lseA_i) + exp(s_ij - lseB_j)
//
// TMEM cannot hold 256 anchor rows x D = 512 fp32 outputs next to the S buffers, so infonce_bwd_tc3.cu produces
// the output in two 256-wide slices and recomputes S for each (executed FLOPs = 3x algorithmic).  Here a cluster
// of FOUR CTAs = two CTA pairs owns the same 256 anchor rows; pair p accumulates output slice p
// (tcgen05.mma.cta_group::2, O = 256 TMEM columns).  The column tiles are dealt alternately: pair p recomputes S
// and the weights W only for tiles t with t % 2 == p, keeps W in TMEM as the A operand of its own second MMA
// (as before) AND stores the packed bf16 W tile into the shared memory of the other pair (st.shared::cluster,
// K-major 128B-swizzled, 32 KB per CTA and tile), which consumes it as an SS-mode A operand.  Per two column
// tiles a pair issues one S tile (2048 MMA cycles) + two PV tiles (2 x 1024) instead of 2 x 3072, and pulls
// 128 KB instead of 192 KB of operands from L2.
//
// Protocol per pair (leader = even CTA rank of the pair):
//   warp 0  TMA: resident anchor block, then ONE 8-slot ring (8 KB slots) filled in MMA issue order with the half
//           k-blocks of the OWN S tiles and the V chunks of EVERY tile (V of tile t+1 is in flight while PV(t) runs)
//   warp 1  MMA (leader): S(own k+1), then per tile PV from TMEM (own) or from the received smem tile (foreign)
//   warps 2-9 softmax on own tiles: TMEM in-place W + remote copy (st.shared::cluster); arrive p_full (own leader) and,
//           with release.cluster, wr_full (leader of the other pair); the generic->async proxy fence is executed once
//           by the consuming MMA thread after its acquire, not by the 512 writers
//   wr_empty (in the PRODUCER CTAs) is arrived by the consumer's tcgen05.commit multicast once its PV has read
//   the received tile.
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_pair.cuh"

namespace dmf {

constexpr int B5_THREADS = 320;
constexpr int B5_TILE = 128 * 64 * 2;   // 16 KB: [128 x 64] bf16
constexpr int B5_HALF = 64 * 64 * 2;    // 8 KB: this CTA's half of a column k-block
constexpr int B5_SLOTS = 8;             // unified operand ring: 8 slots of 8 KB (S half k-blocks AND V half chunks)
constexpr int B5_OW = 256;
constexpr int B5_KB = 8;                // D = 512 only
constexpr float kLog2e5 = 1.4426950408889634f;
constexpr size_t B5_SMEM = 1024 + (size_t)B5_KB * B5_TILE + (size_t)B5_SLOTS * B5_HALF /*ring*/ +
                           2 * B5_TILE /*W recv*/ + 1024 /*bs*/ + 256 /*barriers*/;

__device__ __forceinline__ float ex2f5(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2_5(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// Tried and rejected (9.9 ms vs 9.1 ms): st.async, whose every 16-byte store also completes tx bytes on an mbarrier of
// the receiving CTA -- no writer-side fences, but 2048 barrier updates per tile and CTA.
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint32_t mbar, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%2, %3, %4, %5}, [%1];" ::"r"(addr),
               "r"(mbar), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// cluster-scope acquire wait (pairs with the release.cluster arrive of the remote writers)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = tc::smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) {
      printf("dmf: cluster mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}
// arrive (once all prior MMAs of this thread retired) on the barrier at the same offset in the CTAs of `mask`
__device__ __forceinline__ void umma_commit_mask(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          tc::smem_u32(bar)),
      "h"(mask)
      : "memory");
}

#ifdef DMF_TC5_TRACE
__device__ unsigned long long dmf_tc5_trace[8][2048];
__device__ __forceinline__ void trace(int stream, int idx) {
  if (blockIdx.x < 4 && blockIdx.z == 0 && idx < 2048 && (threadIdx.x & 31) == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    dmf_tc5_trace[stream][idx] = t;
  }
}
#define TRACE(st, idx) trace(st, idx)
#else
#define TRACE(st, idx)
#endif

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(B5_THREADS, 1)
infonce_bwd_tc5_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmBT, int Ma, int Nb, float scale,
                       const float* __restrict__ lseA, const float* __restrict__ lseB, float coef,
                       const float* __restrict__ gscale, long long diag_offset, const uint16_t* __restrict__ Bm,
                       long long ldb, float* __restrict__ dA, long long ldda) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smemA = smem;                                   // 8 tiles (resident anchors)
  uint8_t* smemR = smemA + B5_KB * B5_TILE;                // operand ring, consumed in MMA issue order: a slot holds this
                                                           // CTA's half of an S k-block ([64 j x 64 k]) or half of a V
                                                           // chunk ([64 d-rows x 64 j]; a chunk = 2 consecutive slots)
  uint8_t* smemW = smemR + B5_SLOTS * B5_HALF;             // 2 tiles: W tile received from the other pair
  float* bsm = reinterpret_cast<float*>(smemW + 2 * B5_TILE);   // [2][128] per-tile column factors b_j
  uint64_t* bars = reinterpret_cast<uint64_t*>(bsm + 256);
  uint64_t* a_full = bars;
  uint64_t* full_bar = bars + 1;
  uint64_t* empty_bar = full_bar + B5_SLOTS;
  uint64_t* s_full = empty_bar + B5_SLOTS;     // [2]
  uint64_t* p_full = s_full + 2;               // [2] own W(t) stored in TMEM (16 warp arrivals of the pair)
  uint64_t* acc_full = p_full + 2;
  uint64_t* wr_full = acc_full + 1;            // per CTA: the partner CTA (rank ^ 2) stored its W tile here (32 KB of st.async tx)
  uint64_t* wr_empty = wr_full + 1;            // per CTA: the other pair's PV has consumed the tile we sent
  uint64_t* wr_peer = wr_empty + 1;            // leader: the non-leader CTA of this pair has received its tile (relay)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wr_peer + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank4 = tc2::cluster_ctarank();      // 0..3
  const uint32_t pair = rank4 >> 1, prank = rank4 & 1u;
  const bool leader = prank == 0;
  const uint32_t leader_rank = rank4 & ~1u, partner = rank4 ^ 2u;
  const uint16_t pmask = (uint16_t)(3u << (pair * 2)), qmask = (uint16_t)(3u << ((pair ^ 1u) * 2));
  const int m0 = (blockIdx.x >> 2) * 256 + (int)prank * 128;
  const int d0 = (int)pair * B5_OW;
  const int total_tiles = (Nb + 127) / 128;
  const int tiles_per_split = (total_tiles + (int)gridDim.z - 1) / (int)gridDim.z;
  const int tz0 = blockIdx.z * tiles_per_split;
  const int ntiles = max(0, min(total_tiles, tz0 + tiles_per_split) - tz0);
  const bool split = gridDim.z > 1;
  const int n_own = ntiles > (int)pair ? (ntiles - (int)pair + 1) / 2 : 0;   // tiles t = pair, pair+2, ...
  const int n_foreign = ntiles - n_own;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    tc::tma_prefetch_desc(&tmBT);
    tc::mbar_init(a_full, 1);
    for (int s = 0; s < B5_SLOTS; ++s) { tc::mbar_init(full_bar + s, 1); tc::mbar_init(empty_bar + s, 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(s_full + b, 1); tc::mbar_init(p_full + b, 16); }
    tc::mbar_init(acc_full, 1);
    tc::mbar_init(wr_full, 16);
    tc::mbar_init(wr_empty, 1);
    tc::mbar_init(wr_peer, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc2::tmem_alloc2<512>(tmem_slot);
  tc::tc_fence_before_sync();
  tc2::cluster_sync_all();
  tc::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 256;   // S buffers at [0,128) and [128,256)

  if (warp == 0) {
    // ---- TMA producer (whole warp, one elected lane issues)
    if (tc::elect_one()) {
      if (leader) tc::mbar_expect_tx(a_full, 2 * B5_KB * B5_TILE);
      for (int kb = 0; kb < B5_KB; ++kb) tc2::tma_load_2d_pair(smemA + kb * B5_TILE, &tmA, kb * 64, m0, a_full);
    }
    __syncwarp();
    int slot = 0;
    uint32_t phase = 0;
    auto load_slot = [&](const CUtensorMap* tm, int c0, int c1) {
      tc::mbar_wait(empty_bar + slot, phase ^ 1);
      if (tc::elect_one()) {
        if (leader) tc::mbar_expect_tx(full_bar + slot, 2 * B5_HALF);
        tc2::tma_load_2d_pair(smemR + slot * B5_HALF, tm, c0, c1, full_bar + slot);
      }
      __syncwarp();
      if (++slot == B5_SLOTS) { slot = 0; phase ^= 1; }
    };
    auto load_s = [&](int k) {      // S operand of own tile k (global tile t = pair + 2k): 8 slots
      const int jrow = (tz0 + (int)pair + 2 * k) * 128 + (int)prank * 64;
      for (int kb = 0; kb < B5_KB; ++kb) load_slot(&tmB, kb * 64, jrow);
    };
    auto load_v = [&](int t) {      // V operand of tile t: 2 chunks of 64 j, each 2 slots (d-rows 0-63 / 64-127 of this CTA)
      for (int c = 0; c < 2; ++c) {
        load_slot(&tmBT, (tz0 + t) * 128 + c * 64, d0 + (int)prank * 128);
        load_slot(&tmBT, (tz0 + t) * 128 + c * 64, d0 + (int)prank * 128 + 64);
      }
    };
    int k_next = 0;
    if (n_own > 0) { load_s(0); k_next = 1; }
    for (int t = 0; t < ntiles; ++t) {
      const bool own = ((uint32_t)t & 1u) == pair;
      if (own && k_next < n_own) load_s(k_next++);
      load_v(t);
    }
  } else if (warp == 1) {
    if (leader) {
      // ---- MMA issuer (whole warp, uniform control flow, one elected lane issues)
      constexpr uint32_t idesc_s = tc::make_idesc_bf16(256, 128, 0, 0);
      constexpr uint32_t idesc_o = tc::make_idesc_bf16(256, B5_OW, 0, 0);
      const uint64_t adesc0 = tc::make_smem_desc(tc::smem_u32(smemA), 16, 1024);
      const uint64_t rdesc0 = tc::make_smem_desc(tc::smem_u32(smemR), 16, 1024);
      const uint64_t wdesc0 = tc::make_smem_desc(tc::smem_u32(smemW), 16, 1024);
      tc::mbar_wait(a_full, 0);
      int slot = 0;
      uint32_t phase = 0;
      auto issue_s = [&](int k) {
        const uint32_t d_tmem = tmem_base + (uint32_t)((k & 1) * 128);
        for (int kb = 0; kb < B5_KB; ++kb) {
          tc::mbar_wait(full_bar + slot, phase);
          tc::tc_fence_after_sync();
          const uint64_t ad = adesc0 + (uint64_t)((kb * B5_TILE) >> 4);
          const uint64_t bd = rdesc0 + (uint64_t)((slot * B5_HALF) >> 4);
          if (tc::elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) tc2::umma_ss2(d_tmem, ad + 2 * kk, bd + 2 * kk, idesc_s, (kb | kk) != 0 ? 1u : 0u);
            umma_commit_mask(empty_bar + slot, pmask);
          }
          __syncwarp();
          if (++slot == B5_SLOTS) { slot = 0; phase ^= 1; }
        }
        if (tc::elect_one()) umma_commit_mask(s_full + (k & 1), pmask);
        __syncwarp();
        TRACE(pair * 4 + 0, k);
      };
      // PV of one tile: two 64-column chunks, each = two consecutive ring slots (an even slot index, so never wrapping)
      auto issue_pv = [&](int t, bool from_tmem, uint32_t w_tmem) {
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          tc::mbar_wait(full_bar + slot, phase);
          tc::mbar_wait(full_bar + slot + 1, phase);
          tc::tc_fence_after_sync();
          const uint64_t vd = rdesc0 + (uint64_t)((slot * B5_HALF) >> 4);
          if (tc::elect_one()) {
            if (from_tmem) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                tc2::umma_ts2(tmem_O, w_tmem + (uint32_t)(c * 64 + kk * 8), vd + 2 * kk, idesc_o, (t | c | kk) != 0 ? 1u : 0u);
            } else {
              const uint64_t wd = wdesc0 + (uint64_t)((c * B5_TILE) >> 4);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) tc2::umma_ss2(tmem_O, wd + 2 * kk, vd + 2 * kk, idesc_o, (t | c | kk) != 0 ? 1u : 0u);
            }
            umma_commit_mask(empty_bar + slot, pmask);
            umma_commit_mask(empty_bar + slot + 1, pmask);
          }
          __syncwarp();
          slot += 2;
          if (slot == B5_SLOTS) { slot = 0; phase ^= 1; }
        }
      };
      int k_next = 0, own_k = 0, f = 0;
      if (n_own > 0) { issue_s(0); k_next = 1; }
      for (int t = 0; t < ntiles; ++t) {
        const bool own = ((uint32_t)t & 1u) == pair;
        if (own) {
          if (k_next < n_own) issue_s(k_next++);
          tc::mbar_wait(p_full + (own_k & 1), ((uint32_t)own_k >> 1) & 1);
          TRACE(pair * 4 + 1, own_k);
          issue_pv(t, true, tmem_base + (uint32_t)((own_k & 1) * 128));
          ++own_k;
        } else {
          // received tile f: own half (st.async tx on the local barrier) and the non-leader's half (relayed)
          mbar_wait_cluster(wr_full, (uint32_t)f & 1);
          TRACE(pair * 4 + 2, f);
          tc::fence_proxy_async_smem();          // generic-proxy (st.async) writes -> tensor-core (async proxy) reads
          issue_pv(t, false, 0u);
          if (tc::elect_one()) umma_commit_mask(wr_empty, qmask);     // the producers of this tile may overwrite our smemW
          __syncwarp();
          ++f;
        }
      }
      if (tc::elect_one()) umma_commit_mask(acc_full, pmask);
      __syncwarp();
    } else {
      // ---- non-leader: relay "my half of the received W tile is complete" to the leader's MMA warp
      (void)n_foreign;
    }
  } else {
    // ---- softmax warps: own tiles only
    const int sw = warp - 2;                 // 0..7
    const int q = warp & 3;                  // TMEM lane quarter
    const int ch = sw >> 2;                  // column half of the tile handled by this warp
    const int st = threadIdx.x - 64;         // 0..255 among softmax threads
    const int rloc = q * 32 + lane;
    const int row = m0 + rloc;
    const float sl2 = scale * kLog2e5;
    const float c0 = __ldg(lseB) * kLog2e5;                         // common exponent offset
    const float la2 = (row < Ma) ? __ldg(lseA + row) * kLog2e5 : c0;
    const float ai = ex2f5(la2 - c0);
    const uint32_t p_full_leader = tc2::mapa(tc::smem_u32(p_full), leader_rank);
    const uint32_t wr_full_remote = tc2::mapa(tc::smem_u32(wr_full), leader_rank ^ 2u);
    const uint32_t w_remote = tc2::mapa(tc::smem_u32(smemW), partner) + (uint32_t)(ch * B5_TILE + rloc * 128);
    // column factors b_j = exp(c0 - lseB_j) of own tile k live in bsm[(k & 1) * 128 ...]; the global loads for tile k + 1
    // are issued at the top of tile k and consumed at its end (an L2 round trip at the head of every tile sat on the
    // softmax -> PV critical path of the first version)
    auto col_index = [&](int k) { return (tz0 + (int)pair + 2 * k) * 128 + st; };
    if (st < 128 && n_own > 0) {
      const int j = col_index(0);
      bsm[st] = (j < Nb) ? ex2f5(c0 - __ldg(lseB + j) * kLog2e5) : 0.f;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    for (int k = 0; k < n_own; ++k) {
      const float* bs = bsm + (k & 1) * 128;
      float lb_next = 0.f;
      bool next_ok = false;
      if (st < 128 && k + 1 < n_own) {
        const int j = col_index(k + 1);
        next_ok = j < Nb;
        if (next_ok) lb_next = __ldg(lseB + j);
      }
      tc::mbar_wait(s_full + (k & 1), ((uint32_t)k >> 1) & 1);
      tc::tc_fence_after_sync();
      if (warp == 2 && leader) TRACE(pair * 4 + 3, 2 * k);
      // the other pair must have consumed the tile we sent last time before its buffer is overwritten
      tc::mbar_wait(wr_empty, ((uint32_t)k & 1) ^ 1);
      const uint32_t tS = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((k & 1) * 128 + ch * 64);
      // both 32-column chunks are requested before the first wait (one TMEM read latency per tile instead of two)
      uint32_t rr[2][32];
      tc::tmem_ld_32x32(tS, rr[0]);
      tc::tmem_ld_32x32(tS + 32u, rr[1]);
      tc::tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t (&r)[32] = rr[c];
        const float4* b4 = reinterpret_cast<const float4*>(bs + ch * 64 + c * 32);
        uint32_t pk[16];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 bb = b4[j4];
          const float e0 = ex2f5(fmaf(__uint_as_float(r[j4 * 4 + 0]), sl2, -la2));
          const float e1 = ex2f5(fmaf(__uint_as_float(r[j4 * 4 + 1]), sl2, -la2));
          const float e2 = ex2f5(fmaf(__uint_as_float(r[j4 * 4 + 2]), sl2, -la2));
          const float e3 = ex2f5(fmaf(__uint_as_float(r[j4 * 4 + 3]), sl2, -la2));
          const float w0 = fmaf(e0 * ai, bb.x, e0);
          const float w1 = fmaf(e1 * ai, bb.y, e1);
          const float w2 = fmaf(e2 * ai, bb.z, e2);
          const float w3 = fmaf(e3 * ai, bb.w, e3);
          pk[j4 * 2 + 0] = pack_bf16x2_5(w0, w1);
          pk[j4 * 2 + 1] = pack_bf16x2_5(w2, w3);
        }
        tc::tmem_st_32x16(tS + (uint32_t)(c * 16), pk);   // in place: A operand of our own PV
        // remote copy: row rloc of k-block ch, 16-byte chunks (c*4 + i) ^ (rloc & 7)  (128B swizzle, K-major)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          st_cluster_v4(w_remote + (uint32_t)((((c * 4 + i) ^ (rloc & 7))) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2],
                        pk[4 * i + 3]);
      }
      tc::tmem_st_wait();
      tc::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        tc2::mbar_arrive_cluster(p_full_leader + (uint32_t)((k & 1) * 8));
        mbar_arrive_cluster_release(wr_full_remote);     // orders this warp's remote stores before the arrival
      }
      if (warp == 2 && leader) TRACE(pair * 4 + 3, 2 * k + 1);
      // publish the column factors of the next own tile (its other buffer was last read one tile ago, before the
      // barrier that ended that tile)
      if (st < 128 && k + 1 < n_own) bsm[((k + 1) & 1) * 128 + st] = next_ok ? ex2f5(c0 - lb_next * kLog2e5) : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    // epilogue: this warp stores its lane quarter x column half of the dA slice
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after_sync();
    const float cg = coef * (gscale ? __ldg(gscale) : 1.0f);
    const long long pj = diag_offset >= 0 ? diag_offset + row : -1;
#pragma unroll 1
    for (int c = 0; c < B5_OW / 64; ++c) {
      uint32_t r[32];
      const int cc = ch * (B5_OW / 64) + c;
      tc::tmem_ld_32x32(tmem_O + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 32), r);
      tc::tmem_ld_wait();
      const int dbase = d0 + cc * 32;
      if (row < Ma && ntiles > 0) {
        float* dst = dA + (long long)row * ldda + dbase;
        const bool has_pos = pj >= 0 && pj < Nb && blockIdx.z == 0;   // the positive term is added by split 0 only
        const uint16_t* bp = has_pos ? Bm + pj * ldb + dbase : nullptr;
        const bool vec = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o;
          o.x = (__uint_as_float(r[j + 0]) - (has_pos ? 2.0f * bf2f(bp[j + 0]) : 0.f)) * cg;
          o.y = (__uint_as_float(r[j + 1]) - (has_pos ? 2.0f * bf2f(bp[j + 1]) : 0.f)) * cg;
          o.z = (__uint_as_float(r[j + 2]) - (has_pos ? 2.0f * bf2f(bp[j + 2]) : 0.f)) * cg;
          o.w = (__uint_as_float(r[j + 3]) - (has_pos ? 2.0f * bf2f(bp[j + 3]) : 0.f)) * cg;
          if (split) {
            if (vec) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
            } else {
              atomicAdd(dst + j, o.x); atomicAdd(dst + j + 1, o.y); atomicAdd(dst + j + 2, o.z); atomicAdd(dst + j + 3, o.w);
            }
          } else if (vec) {
            *reinterpret_cast<float4*>(dst + j) = o;
          } else {
            dst[j] = o.x; dst[j + 1] = o.y; dst[j + 2] = o.z; dst[j + 3] = o.w;
          }
        }
      }
    }
  }
  __syncwarp();
  tc::tc_fence_before_sync();
  tc2::cluster_sync_all();          // peers may still target this CTA's barriers / smem / TMEM until here
  if (warp == 1) tc2::tmem_dealloc2<512>(tmem_base);
}

}  // namespace dmf

using namespace dmf;

// Returns -100 if the 4-CTA cluster configuration cannot be scheduled on this device (caller falls back).
int dmf_infonce_bwd_bf16_tc5(const void* A, long long lda, int Ma, const float* lseA, const void* Bm, long long ldb,
                             const void* BmT, long long ldbt, int Nb, const float* lseB, int D, float scale, float coef,
                             const float* gscale, long long diag_offset, float* dA, long long ldda, cudaStream_t s) {
  if (D != 512) return -100;
  CUtensorMap tmA, tmB, tmBT;
  int rc = make_tmap_bf16_2d(&tmA, A, Ma, D, lda, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, Bm, Nb, D, ldb, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmBT, BmT, D, Nb, ldbt, 64);
  if (rc) return rc;
  static int usable = -1;       // -1 unknown, 0 no, else max active 4-CTA clusters
  if (usable < 0) {
    cudaError_t e = cudaFuncSetAttribute(infonce_bwd_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B5_SMEM);
    if (e != cudaSuccess) { cudaGetLastError(); usable = 0; }
    else {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(4 * 64, 1, 1);
      cfg.blockDim = dim3(B5_THREADS, 1, 1);
      cfg.dynamicSmemBytes = B5_SMEM;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int ncl = 0;
      e = cudaOccupancyMaxActiveClusters(&ncl, infonce_bwd_tc5_kernel, &cfg);
      if (e != cudaSuccess) {       // the kernel carries __cluster_dims__: retry without the launch attribute
        cudaGetLastError();
        cfg.attrs = nullptr; cfg.numAttrs = 0;
        e = cudaOccupancyMaxActiveClusters(&ncl, infonce_bwd_tc5_kernel, &cfg);
      }
      if (e != cudaSuccess) { cudaGetLastError(); ncl = 0; }
      // Measured on B200 (B = 65 536, D = 512): 9.09 ms vs 8.64 ms for the pair kernel alone, 115.5 vs 115.0 ms per
      // training step (the pair kernel sits at the power cap, 1.55 GHz; this one runs at 1.96 GHz with the tensor pipe
      // only 50 % active).  A %globaltimer trace of one cluster (-DDMF_TC5_TRACE, tools/trace_tc5.py) shows a latency-
      // bound ping-pong between the two pairs: softmax + remote copy of a 128x128 W tile takes 2.6 us (TMEM read
      // 64 B/clk, DSMEM store round trip inside the release), and with smem (227 KB) and TMEM (512 columns) both full
      // there is no room for a third S buffer, a second receive buffer or a ring deeper than 8 x 8 KB to hide it.
      // st.async (tx-counting remote stores, no writer fences) was slower (9.9 ms); prefetching the column factors and
      // requesting both TMEM chunks before the first wait changed nothing (9.14 ms): the chain is bound by the DSMEM
      // hand-off and the shallow ring, not by the softmax warps' own latencies.  Handing W over per 64-column chunk (all
      // eight softmax warps on chunk 0 first, PV of chunk 0 under the softmax of chunk 1) was slower too (10.3 ms: twice
      // the release fences and TMEM store waits per tile).  Opt-in: DMF_BWD_TC5=1.
      {
        const char* ev = getenv("DMF_BWD_TC5");
        if (!ev || !atoi(ev)) ncl = 0;
      }
      if (getenv("DMF_DEBUG")) fprintf(stderr, "dmf: 4-CTA cluster backward: max active clusters = %d\n", ncl);
      usable = ncl;
    }
  }
  // a 4-CTA cluster needs four SMs of one GPC: only worth it if (almost) all SMs stay usable
  if (usable < 30) return -100;
  const int blocks256 = (Ma + 255) / 256;
  int nsplit = 1;
  const int total_tiles = (Nb + 127) / 128;
  {
    const int cap = usable;
    double best = (double)blocks256 / (double)(((blocks256 + cap - 1) / cap) * cap);
    for (int ns = 2; ns <= 16 && best < 0.97; ++ns) {
      if (total_tiles / ns < 16) break;
      const int items = blocks256 * ns;
      const double eff = (double)items / (double)(((items + cap - 1) / cap) * cap);
      if (eff > best + 0.03) { best = eff; nsplit = ns; }
    }
    int tps = (total_tiles + nsplit - 1) / nsplit;
    tps = (tps + 1) & ~1;                       // even tile counts keep both pairs equally loaded
    nsplit = (total_tiles + tps - 1) / tps;
  }
  if (nsplit > 1) {
    cudaError_t e = cudaMemset2DAsync(dA, (size_t)ldda * sizeof(float), 0, (size_t)D * sizeof(float), (size_t)Ma, s);
    if (e != cudaSuccess) return fail((int)e, "dmf_infonce_bwd(bf16 quad): memset: %s", cudaGetErrorString(e));
  }
  dim3 grid(4 * blocks256, 1, nsplit);
  infonce_bwd_tc5_kernel<<<grid, B5_THREADS, B5_SMEM, s>>>(tmA, tmB, tmBT, Ma, Nb, scale, lseA, lseB, coef, gscale, diag_offset,
                                                           (const uint16_t*)Bm, ldb, dA, ldda);
  return launched("dmf_infonce_bwd(bf16 quad)");
}

#ifdef DMF_TC5_TRACE
extern "C" int dmf_tc5_trace_read(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, dmf::dmf_tc5_trace, sizeof(unsigned long long) * 8 * 2048);
}
#endif
