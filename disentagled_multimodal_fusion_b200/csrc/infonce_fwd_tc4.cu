// K2 forward for L2-normalised embeddings: ONE pass over the similarity tiles yields both the row sums and
// the column sums of exp(s - shift), and symmetric (intra-view) blocks are only computed on a half window.
//
// SupConLoss.forward (models/losses.py:64-99) needs, for the anchors of view 0, the LSE of every row of
// S01 = z0 z1^T / T and, for the anchors of view 1, the LSE of every COLUMN of the same matrix; the no-grad
// diagnostics need the row LSEs of the symmetric blocks S00 and S11.  DisentangledSSL feeds unit vectors
// (vMF samples / F.normalize, models/disentangledssl.py:134-140), so |s| <= 1/T: a FIXED shift replaces the
// online max (no rescaling, one ex2 per element) and makes column sums plain additions:
//   row_sum[i] += sum_j e_ij      (per-thread accumulation, one atomicAdd per row and CTA)
//   col_sum[j] += sum_i e_ij      (32x32 butterfly transpose-reduce in the warp, one 128-byte red.add per chunk)
// with e_ij = 2^(s_ij * scale * log2e - shift2).  For a symmetric block only tiles J in the cyclic half
// window [I, I + T/2] of row block I are visited; an off-diagonal tile feeds row sums of block I AND (as
// column sums) the row sums of block J, the diagonal tile feeds row sums only.  Work per critic call drops
// from four B x B blocks (rows of S01, S10, S00, S11) to two.
//
// CTA pair, cta_group::2, M = 256 anchor rows resident in smem, 256-column tiles through a 5-stage TMA ring,
// S tiles double-buffered in TMEM (2 x 256 columns), 8 softmax warps per CTA -- same skeleton as
// infonce_fwd_tc2.cu (which remains the exact online-max path for arbitrary inputs).
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_pair.cuh"

namespace dmf {

#ifndef DMF_F4_SW
#define DMF_F4_SW 16
#endif
constexpr int F4_SW = DMF_F4_SW;                      // softmax warps per CTA (8 or 16): 4 lane quarters x F4_SW / 4 column groups
constexpr int F4_THREADS = 64 + 32 * F4_SW;
constexpr int F4_WCOLS = 256 / (F4_SW / 4);           // columns of every 256-wide tile handled by one softmax warp
constexpr int F4_TILE = 128 * 64 * 2;   // 16 KB: [128 rows x 64 bf16]
constexpr int F4_STAGES = 5;          // ring depth (4 when E is kept with double-buffered staging tiles)
constexpr int F4_BN = 256;              // column tile of the pair
constexpr float kLog2eF4 = 1.4426950408889634f;
// anchors (8 slots) + ring + 2 KB (barriers, TMEM slot, merge buffer) + 8 x 2 KB E staging + 1 KB alignment slack
constexpr size_t F4_SMEM_MAX = 1024 + (size_t)(8 + F4_STAGES - 1) * F4_TILE + 2048 + 16 * 2048;   // E: 4-stage ring + 16 staging tiles
static_assert(F4_SMEM_MAX <= 232448, "forward carve-up exceeds the opt-in shared-memory window");

__device__ __forceinline__ float f4_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// packed fp32x2 (FFMA2 / FADD2: one issue slot for two elements; the softmax warps are issue / latency bound)
__device__ __forceinline__ unsigned long long f4_pk2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f4_unpk2(unsigned long long a, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
}
__device__ __forceinline__ unsigned long long f4_add2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f4_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// Sum over the 32 lanes of e[j] for every j; lane L returns the total of column L.  The additions of each butterfly level
// run two columns per instruction.
__device__ __forceinline__ float warp_colsum32(float (&e)[32], int lane) {
  float t16[16];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
      const float s0 = up ? e[k] : e[k + 16], s1 = up ? e[k + 1] : e[k + 17];
      const float k0 = up ? e[k + 16] : e[k], k1 = up ? e[k + 17] : e[k + 1];
      f4_unpk2(f4_add2(f4_pk2(k0, k1), f4_pk2(__shfl_xor_sync(0xffffffffu, s0, 16), __shfl_xor_sync(0xffffffffu, s1, 16))),
               t16[k], t16[k + 1]);
    }
  }
  float t8[8];
  {
    const bool up = lane & 8;
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      const float s0 = up ? t16[k] : t16[k + 8], s1 = up ? t16[k + 1] : t16[k + 9];
      const float k0 = up ? t16[k + 8] : t16[k], k1 = up ? t16[k + 9] : t16[k + 1];
      f4_unpk2(f4_add2(f4_pk2(k0, k1), f4_pk2(__shfl_xor_sync(0xffffffffu, s0, 8), __shfl_xor_sync(0xffffffffu, s1, 8))),
               t8[k], t8[k + 1]);
    }
  }
  float t4[4];
  {
    const bool up = lane & 4;
#pragma unroll
    for (int k = 0; k < 4; k += 2) {
      const float s0 = up ? t8[k] : t8[k + 4], s1 = up ? t8[k + 1] : t8[k + 5];
      const float k0 = up ? t8[k + 4] : t8[k], k1 = up ? t8[k + 5] : t8[k + 1];
      f4_unpk2(f4_add2(f4_pk2(k0, k1), f4_pk2(__shfl_xor_sync(0xffffffffu, s0, 4), __shfl_xor_sync(0xffffffffu, s1, 4))),
               t4[k], t4[k + 1]);
    }
  }
  float t2[2];
  {
    const bool up = lane & 2;
    const float s0 = up ? t4[0] : t4[2], s1 = up ? t4[1] : t4[3];
    const float k0 = up ? t4[2] : t4[0], k1 = up ? t4[3] : t4[1];
    f4_unpk2(f4_add2(f4_pk2(k0, k1), f4_pk2(__shfl_xor_sync(0xffffffffu, s0, 2), __shfl_xor_sync(0xffffffffu, s1, 2))), t2[0], t2[1]);
  }
  const bool up = lane & 1;
  const float send = up ? t2[0] : t2[1];
  const float keep = up ? t2[1] : t2[0];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

struct F4Args {
  int Ma, Nb, num_kb;
  float scale;        // 1 / temperature
  float sl2;          // scale * log2(e)
  float shift2;       // fixed shift in the log2 domain
  int sym;            // 1: A rows are rows [row0_global, row0_global + Ma) of Bm (symmetric block, half window)
  int row0_global;    // global index of local row 0 (multiple of 256 when sym)
  int total_tiles;    // T = ceil(Nb / 256)
  int tiles_per_split;
  float* row_sum;     // [Ma]  += (caller zeroes)
  float* col_sum;     // [Nb]  += (caller zeroes); may be NULL when !sym (rows only)
  long long diag_offset;
  float* diag_out;    // [Ma] raw scaled similarity s_{i, diag_offset + i}
  int store_e;        // sym = 0 only: keep e_ij as bf16 in [128 x 64] blocks, block (ib, jb) at (ib * njb + jb) * 16 KB,
  int njb;            //   written through the tensor map tmE (the operand of the stored-probability backward, infonce_bwd_e.cu)
  int stages;         // ring depth: F4_STAGES, or F4_STAGES - 1 with two staging tiles per warp (estage_bufs = 2)
  int estage_bufs;
};

// 32 consecutive e values of one row -> bf16 -> this thread's 64-byte row of the warp's [32 x 64 B] staging tile (16-byte
// chunk index XOR ((row >> 1) & 3): the 64B TMA swizzle, conflict-free for 16-byte stores of 8 consecutive rows)
#ifndef DMF_F4_COLSUM_MMA
// 1: column sums as a ones x E product on the tensor core (ldmatrix.trans + mma.sync m16n8k16 from the staged bf16 tile:
// 4 LDSM + 8 HMMA per chunk instead of 31 SHFL + 62 FSEL + 16 FADD2).  Parity green, but SLOWER on B200: cross block
// 3.80 ms against 2.90 ms (with the E store 4.60 against 3.86 ms) -- the legacy mma.sync path shares the tensor pipe with
// the tcgen05 stream that bounds the kernel.  Kept for the record; the fp32 warp butterfly (0) is the product path.
#define DMF_F4_COLSUM_MMA 0
#endif
// 1: the two softmax warps that hold the two 32-column halves of the same [32 rows x 64 columns] piece of an E block share
// one 4 KB staging tile and ONE TMA store of full 128-byte rows (two 64-thread named barriers per chunk).  Tried because
// the per-warp variant (0) stores 64-byte row segments (134 M store requests per launch next to 131 M operand-tile
// requests); parity green, NOT faster: 4.05 ms against 3.86-3.90 ms -- halving the request count buys nothing, the pair
// barriers cost a little.  Kept for the record.
#ifndef DMF_F4_PAIR_STORE
#define DMF_F4_PAIR_STORE 0
#endif
#ifndef DMF_F4_VAR
#define DMF_F4_VAR 0      // timing variants, tools builds only: 1 = stage only (no TMA store), 2 = truncating pack, 3 = no column sums, 4 = no bulk-group wait, 5 = no proxy fence, 6 = neither (4-6: WRONG results)
#endif
__device__ __forceinline__ void f4_stage_e(uint32_t stage_row, int lane, const float (&e)[32]) {
  const uint32_t x = (uint32_t)((lane >> 1) & 3);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#if DMF_F4_VAR == 2
      w[k] = __byte_perm(__float_as_uint(e[g * 8 + 2 * k]), __float_as_uint(e[g * 8 + 2 * k + 1]), 0x7632);
#else
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[k]) : "f"(e[g * 8 + 2 * k + 1]), "f"(e[g * 8 + 2 * k]));
#endif
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_row + (((uint32_t)g ^ x) << 4)), "r"(w[0]), "r"(w[1]),
                 "r"(w[2]), "r"(w[3])
                 : "memory");
  }
}
// [32 rows x 32 columns] box of the staging tile -> E block rows (c0 = column inside the 64-wide block, c1 = block row)
__device__ __forceinline__ void f4_tma_store_2d(const CUtensorMap* d, uint32_t smem_src, int c0, int c1, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(
                   reinterpret_cast<uint64_t>(d)),
               "r"(smem_src), "r"(c0), "r"(c1), "l"(policy)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void f4_pair_bar(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
// Pair variant: warp `hf` (0 / 1) of the pair writes its 64-byte half of every row of the pair's [32 x 128 B] tile (128B TMA
// swizzle: 16-byte chunk k of row r sits at chunk k ^ (r & 7): conflict-free), warp 0 issues the store of the whole tile.
__device__ __forceinline__ void f4_store_wait_read();
__device__ __forceinline__ void f4_tma_store_2d(const CUtensorMap* d, uint32_t smem_src, int c0, int c1, uint64_t policy);
__device__ __forceinline__ void f4_store_chunk_pair(const CUtensorMap* tmE, uint32_t tile, int bar_id, int hf, int lane,
                                                    const float (&e)[32], int e_c1, uint64_t policy) {
  if (hf == 0 && tc::elect_one()) f4_store_wait_read();      // the last store has read the tile (the same lane issues)
  f4_pair_bar(bar_id);
  const uint32_t rowp = tile + (uint32_t)(lane * 128);
  const uint32_t x = (uint32_t)(lane & 7);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[k]) : "f"(e[g * 8 + 2 * k + 1]), "f"(e[g * 8 + 2 * k]));
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + ((((uint32_t)(hf * 4 + g)) ^ x) << 4)), "r"(w[0]),
                 "r"(w[1]), "r"(w[2]), "r"(w[3])
                 : "memory");
  }
  tc::fence_proxy_async_smem();            // generic-proxy stores -> visible to the TMA engine (async proxy)
  f4_pair_bar(bar_id);
  if (hf == 0 && tc::elect_one()) f4_tma_store_2d(tmE, tile, 0, e_c1, policy);
}
__device__ __forceinline__ void f4_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void f4_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void f4_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(F4_THREADS, 1)
rowcol_sum_tc4_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmE, const F4Args P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smemA = smem;                                    // num_kb tiles: this CTA's 128 anchor rows
  uint8_t* smemB = smem + P.num_kb * F4_TILE;               // F4_STAGES tiles: this CTA's half of the column tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemB + P.stages * F4_TILE);
  uint64_t* a_full = bars;
  uint64_t* full_bar = bars + 1;
  uint64_t* empty_bar = full_bar + F4_STAGES;
  uint64_t* s_full = empty_bar + F4_STAGES;   // per CTA [2]
  uint64_t* s_empty = s_full + 2;             // leader [2]: 16 softmax warps of the pair drained the buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 2);
  uint8_t* estage = smem + (8 + P.stages) * F4_TILE + 2048;    // 8 x estage_bufs x 2 KB staging tiles of the E stores

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc2::cluster_ctarank();
  const bool leader = rank == 0;
  const int blk = blockIdx.x >> 1;                          // 256-row anchor block (local)
  const int m0 = blk * 256 + (int)rank * 128;               // first local row of this CTA
  const int T = P.total_tiles;
  // tile window of this anchor block: cyclic [w0, w0 + wcnt) mod T
  int w0 = 0, wcnt = T;
  const int Ig = (P.row0_global >> 8) + blk;                // global row-block index (sym only)
  if (P.sym) {
    w0 = Ig;
    if (T & 1) wcnt = (T + 1) / 2;
    else wcnt = T / 2 + (Ig < T / 2 ? 1 : 0);
  }
  const int t_begin = blockIdx.y * P.tiles_per_split;
  const int ntiles = max(0, min(wcnt, t_begin + P.tiles_per_split) - t_begin);

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    tc::mbar_init(a_full, 1);
    for (int s = 0; s < P.stages; ++s) { tc::mbar_init(full_bar + s, 1); tc::mbar_init(empty_bar + s, 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(s_full + b, 1); tc::mbar_init(s_empty + b, 2 * F4_SW); }
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
        if (leader) tc::mbar_expect_tx(a_full, 2 * P.num_kb * F4_TILE);
        for (int kb = 0; kb < P.num_kb; ++kb) tc2::tma_load_2d_pair(smemA + kb * F4_TILE, &tmA, kb * 64, m0, a_full);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        int J = w0 + t_begin + t;
        if (J >= T) J -= T;
        const int j0 = J * F4_BN + (int)rank * 128;
        for (int kb = 0; kb < P.num_kb; ++kb) {
          tc::mbar_wait(empty_bar + stage, phase ^ 1);
          if (tc::elect_one()) {
            if (leader) tc::mbar_expect_tx(full_bar + stage, 2 * F4_TILE);
            tc2::tma_load_2d_pair(smemB + stage * F4_TILE, &tmB, kb * 64, j0, full_bar + stage);
          }
          __syncwarp();
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && ntiles > 0) {
      // whole warp, uniform control flow; one elected lane issues the tcgen05 instructions
      constexpr uint32_t idesc = tc::make_idesc_bf16(256, F4_BN, 0, 0);
      const uint64_t adesc0 = tc::make_smem_desc(tc::smem_u32(smemA), 16, 1024);
      const uint64_t bdesc0 = tc::make_smem_desc(tc::smem_u32(smemB), 16, 1024);
      tc::mbar_wait(a_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        tc::mbar_wait(s_empty + buf, (((uint32_t)t >> 1) & 1) ^ 1);
        tc::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * F4_BN);
        for (int kb = 0; kb < P.num_kb; ++kb) {
          tc::mbar_wait(full_bar + stage, phase);
          tc::tc_fence_after_sync();
          const uint64_t ad = adesc0 + (uint64_t)((kb * F4_TILE) >> 4);
          const uint64_t bd = bdesc0 + (uint64_t)((stage * F4_TILE) >> 4);
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) tc2::umma_ss2(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            tc2::umma_commit2(empty_bar + stage);
          }
          __syncwarp();
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
        if (tc::elect_one()) tc2::umma_commit2(s_full + buf);
        __syncwarp();
      }
    }
  } else {
    const int sw = warp - 2;
    const int q = warp & 3;
    const int ch = sw >> 2;                      // column group (F4_WCOLS columns) of every 256-wide tile
    const int rloc = q * 32 + lane;
    const int row = m0 + rloc;
    const bool rvalid = row < P.Ma;
    // invalid rows (TMA zero fill) must contribute nothing to the column sums: shift = +inf -> e = 0
    const float my_shift = rvalid ? P.shift2 : INFINITY;
    float l = 0.f, diag = 0.f;
    bool has_diag = false;
    const long long dj = (P.diag_offset >= 0 && rvalid) ? P.diag_offset + row : -1;
    const uint32_t s_empty_leader0 = tc2::mapa(tc::smem_u32(s_empty), 0);
    const uint32_t my_stage0 = tc::smem_u32(estage) + (uint32_t)(sw * 2048 * P.estage_bufs);
#if DMF_F4_PAIR_STORE
    // chunk c of warp (q, ch) = columns (ch / 2) * 128 + c * 64 + (ch & 1) * 32: warps ch and ch ^ 1 hold the two halves of
    // the same 64-column E block at the same time
    static_assert(F4_SW == 16, "the pair store maps 4 column groups onto 2 pairs");
    const int cofs = (ch >> 1) * 128 + (ch & 1) * 32, cstep = 64;
    const int hf = ch & 1, pair_bar = 1 + q * 2 + (ch >> 1);
    const uint32_t pair_tile = tc::smem_u32(estage) + (uint32_t)((q * 2 + (ch >> 1)) * 4096);
#else
    const int cofs = ch * F4_WCOLS, cstep = 32;
#endif
    uint32_t ebuf = 0;                                          // staging tile of the next E store
    const uint64_t e_policy = tc::l2_policy_evict_first();      // E is written once and read much later
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      int J = w0 + t_begin + t;
      if (J >= T) J -= T;
      const bool want_cols = P.col_sum != nullptr && !(P.sym && J == Ig);
      tc::mbar_wait(s_full + buf, ((uint32_t)t >> 1) & 1);
      tc::tc_fence_after_sync();
      const int j0 = J * F4_BN + cofs;
#pragma unroll 1
      for (int c = 0; c < F4_WCOLS / 32; ++c) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * F4_BN + cofs + c * cstep), r);
        tc::tmem_ld_wait();
        const int nbase = j0 + c * cstep;
        const int nvalid = P.Nb - nbase;
        // E: the warp's [32 rows x 32 columns] piece goes to block rows q*32.. of block (m0 / 128, nbase / 64), column
        // (c & 1) * 32 inside the block
        const int e_c0 = (c & 1) * 32, e_c1 = ((m0 >> 7) * P.njb + (nbase >> 6)) * 128 + q * 32;
        if (nvalid <= 0) {                         // warp-uniform
#if DMF_F4_PAIR_STORE
          if (P.store_e) {
            float z[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) z[j] = 0.f;
            f4_store_chunk_pair(&tmE, pair_tile, pair_bar, hf, lane, z, e_c1, e_policy);
          }
          continue;
#endif
          if (P.store_e) {
            float z[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) z[j] = 0.f;
            const uint32_t my_stage = my_stage0 + ebuf * 2048;
            if (tc::elect_one()) {                 // the last store from this staging tile has read it (same lane issues)
              if (P.estage_bufs == 2) f4_store_wait_read1(); else f4_store_wait_read();
            }
            __syncwarp();
            f4_stage_e(my_stage + (uint32_t)(lane * 64), lane, z);
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (tc::elect_one()) f4_tma_store_2d(&tmE, my_stage, e_c0, e_c1, e_policy);
            if (P.estage_bufs == 2) ebuf ^= 1;
          }
          continue;
        }
        // positive / self term of this row, if it lies in this chunk: compare against CONSTANT indices (a loop over
        // nbase + j == dj made ptxas carry 32 induction variables through the chunk loop: 38 VIADD per chunk, 10 % of
        // all warp instructions of the kernel in the ncu source view)
        const long long jd = dj - nbase;
        if (jd >= 0 && jd < 32) {
          const int jdi = (int)jd;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j == jdi) { diag = __uint_as_float(r[j]) * P.scale; has_diag = true; }
        }
        float e[32];
        if (nvalid >= 32) {
          const unsigned long long sl2p = f4_pk2(P.sl2, P.sl2), shp = f4_pk2(-my_shift, -my_shift);
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float a0, a1;
            f4_unpk2(f4_fma2(f4_pk2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), sl2p, shp), a0, a1);
            e[j] = f4_exp2(a0);
            e[j + 1] = f4_exp2(a1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) e[j] = j < nvalid ? f4_exp2(fmaf(__uint_as_float(r[j]), P.sl2, -my_shift)) : 0.f;
        }
        {
          unsigned long long pa = f4_pk2(e[0], e[1]), pb = f4_pk2(e[2], e[3]);
#pragma unroll
          for (int j = 4; j < 32; j += 4) {
            pa = f4_add2(pa, f4_pk2(e[j], e[j + 1]));
            pb = f4_add2(pb, f4_pk2(e[j + 2], e[j + 3]));
          }
          float q0, q1;
          f4_unpk2(f4_add2(pa, pb), q0, q1);
          l += q0 + q1;
        }
#if DMF_F4_COLSUM_MMA
        // The staged bf16 tile serves BOTH consumers: the TMA store of E (when kept) and the column sums, which are a
        // ones x E product on the tensor core (ldmatrix.trans + mma.sync m16n8k16 from the staging tile) instead of a
        // 32-lane butterfly: 4 LDSM + 8 HMMA + 1 red.v2 against 31 SHFL + 62 FSEL + 16 FADD2 per chunk.
        if (P.store_e || want_cols) {
          const uint32_t my_stage = my_stage0;
          if (P.store_e && tc::elect_one()) f4_store_wait_read();     // the last store has read the staging tile
          __syncwarp();
          f4_stage_e(my_stage + (uint32_t)(lane * 64), lane, e);
          if (P.store_e) tc::fence_proxy_async_smem();   // generic-proxy stores -> visible to the TMA engine (async proxy)
          __syncwarp();
          if (P.store_e && tc::elect_one()) f4_tma_store_2d(&tmE, my_stage, e_c0, e_c1, e_policy);
          if (want_cols) {
            float c0 = 0.f, c1 = 0.f;
            const int g = lane >> 2;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              // lane L supplies row L of the 8-column group b: matrices = rows 0-7, 8-15, 16-23, 24-31
              uint32_t m0r, m1r, m2r, m3r;
              asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                           : "=r"(m0r), "=r"(m1r), "=r"(m2r), "=r"(m3r)
                           : "r"(my_stage + (uint32_t)(lane * 64) + (((uint32_t)b ^ (uint32_t)((lane >> 1) & 3)) << 4)));
              float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
              const uint32_t ones = 0x3F803F80u;           // bf16 (1, 1)
              asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %4, %4, %4}, {%5, %6}, "
                           "{%0, %1, %2, %3};"
                           : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3)
                           : "r"(ones), "r"(m0r), "r"(m1r));
              asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %4, %4, %4}, {%5, %6}, "
                           "{%0, %1, %2, %3};"
                           : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3)
                           : "r"(ones), "r"(m2r), "r"(m3r));
              // every row of D holds the column sums of columns 8 b + 2 (lane % 4) + {0, 1}; lane group g == b keeps them
              if (g == b) { c0 = d0; c1 = d1; }
            }
            if (lane < 16) {
              const int col = 8 * g + 2 * (lane & 3);
              float* dst = P.col_sum + nbase + col;
              if (col + 1 < nvalid)
                asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(c0), "f"(c1) : "memory");
              else if (col < nvalid)
                atomicAdd(dst, c0);
            }
          }
        }
#else
#if DMF_F4_PAIR_STORE
        if (P.store_e) f4_store_chunk_pair(&tmE, pair_tile, pair_bar, hf, lane, e, e_c1, e_policy);
        if (false) {
#else
        if (P.store_e) {
#endif
          const uint32_t my_stage = my_stage0 + ebuf * 2048;
#if DMF_F4_VAR == 1
          f4_stage_e(my_stage + (uint32_t)(lane * 64), lane, e);
#else
#if DMF_F4_VAR != 4 && DMF_F4_VAR != 6
          if (tc::elect_one()) {                   // the last store from this staging tile has read it (same lane issues)
            if (P.estage_bufs == 2) f4_store_wait_read1(); else f4_store_wait_read();
          }
#endif
          __syncwarp();
          f4_stage_e(my_stage + (uint32_t)(lane * 64), lane, e);
#if DMF_F4_VAR != 5 && DMF_F4_VAR != 6
          tc::fence_proxy_async_smem();            // generic-proxy stores -> visible to the TMA engine (async proxy)
#endif
          __syncwarp();
          if (tc::elect_one()) f4_tma_store_2d(&tmE, my_stage, e_c0, e_c1, e_policy);
#endif
          if (P.estage_bufs == 2) ebuf ^= 1;
        }
        if (want_cols && DMF_F4_VAR != 3) {
          const float cs = warp_colsum32(e, lane);
          if (lane < nvalid) atomicAdd(P.col_sum + nbase + lane, cs);
        }
#endif
      }
      tc::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) tc2::mbar_arrive_cluster(s_empty_leader0 + (uint32_t)(buf * 8));
    }
    if (P.store_e && tc::elect_one()) f4_store_wait_all();   // the E stores must have left shared memory before the CTA exits
    // one atomicAdd per row and softmax warp (once per CTA: column groups and column splits share rows)
    if (rvalid && ntiles > 0) {
      atomicAdd(P.row_sum + row, l);
      if (P.diag_out && has_diag) P.diag_out[row] = diag;
    }
  }
  __syncwarp();
  tc::tc_fence_before_sync();
  tc2::cluster_sync_all();          // the peer may still target this CTA's barriers / TMEM until here
  if (warp == 1) tc2::tmem_dealloc2<512>(tmem_base);
}

}  // namespace dmf

using namespace dmf;

// Sums of exp(scale * <a_i, b_j> - shift) over the rows / columns of the [Ma x Nb] similarity block (bf16 operands).
//   sym = 0: every tile; row_sum[i] += sum_j, col_sum[j] += sum_i (col_sum may be NULL).
//   sym = 1: A must be rows [row0_global, row0_global + Ma) of Bm (a symmetric Gram block); only the cyclic half
//            window of tiles is computed and the FULL row sum of global row g is row_sum[g - row0_global] + col_sum[g]
//            (col_sum summed over all ranks).  Requires row0_global % 256 == 0.
// row_sum / col_sum are accumulated (caller zeroes).  diag_out[i] = scale * <a_i, b_{diag_offset+i}> (natural units).
extern "C" int dmf_infonce_rowcol_sums(const void* A, long long lda, int Ma, const void* Bm, long long ldb, int Nb, int D,
                                       float scale, float shift, int sym, int row0_global, float* row_sum, float* col_sum,
                                       long long diag_offset, float* diag_out, dmf_stream_t s) {
  return dmf_infonce_rowcol_sums_store(A, lda, Ma, Bm, ldb, Nb, D, scale, shift, sym, row0_global, row_sum, col_sum,
                                       diag_offset, diag_out, nullptr, s);
}

// The same pass; E != NULL (sym = 0 only) additionally keeps e_ij = exp(s_ij - shift) as bf16 in the blocked layout of
// dmf_infonce_bwd_stored (dmf_infonce_e_bytes(Ma, Nb) bytes, every byte written: zeros outside [Ma x Nb]).
extern "C" int dmf_infonce_rowcol_sums_store(const void* A, long long lda, int Ma, const void* Bm, long long ldb, int Nb,
                                             int D, float scale, float shift, int sym, int row0_global, float* row_sum,
                                             float* col_sum, long long diag_offset, float* diag_out, void* E,
                                             dmf_stream_t s) {
  DMF_REQUIRE(A && Bm && row_sum, "dmf_infonce_rowcol_sums: null argument");
  DMF_REQUIRE(!E || (!sym && (reinterpret_cast<uintptr_t>(E) & 15) == 0),
              "dmf_infonce_rowcol_sums_store: E needs sym = 0 and a 16-byte aligned base");
  DMF_REQUIRE(Ma >= 0 && Nb >= 1, "dmf_infonce_rowcol_sums: bad shape Ma=%d Nb=%d", Ma, Nb);
  DMF_REQUIRE(D % 64 == 0 && D >= 64 && D <= 512, "dmf_infonce_rowcol_sums: D=%d must be a multiple of 64 in [64,512]", D);
  DMF_REQUIRE(!sym || (col_sum && (row0_global % 256) == 0 && row0_global + Ma <= Nb),
              "dmf_infonce_rowcol_sums: symmetric mode needs col_sum, row0_global %% 256 == 0 and the rows inside Bm");
  if (Ma == 0) return 0;
  const int num_kb = D / 64;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16_2d(&tmA, A, Ma, D, lda, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, Bm, Nb, D, ldb, 128);
  if (rc) return rc;
  // carve-up: anchors, ring, barriers + merge buffer (2 KB); with E the ring sits after 8 anchor slots' worth of space
  // so that the 8 x 2 KB staging tiles have a fixed, 1024-aligned offset
  // with E: 4-stage ring + one 2 KB staging tile per softmax warp (F4_SW = 16); F4_SW = 8 keeps two tiles per warp
  const int ebufs = 16 / F4_SW;
#if DMF_F4_COLSUM_MMA
  const size_t smem = F4_SMEM_MAX;       // the staging tiles also feed the column sums
#else
  const size_t smem = E ? F4_SMEM_MAX : 1024 + (size_t)(num_kb + F4_STAGES) * F4_TILE + 256 + 2048;
#endif
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(rowcol_sum_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)F4_SMEM_MAX);
    if (e != cudaSuccess) return fail((int)e, "dmf_infonce_rowcol_sums: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  F4Args P;
  P.estage_bufs = E ? ebufs : 1;
  P.stages = (E || DMF_F4_COLSUM_MMA) ? F4_STAGES - 1 : F4_STAGES;
  P.Ma = Ma; P.Nb = Nb; P.num_kb = num_kb;
  P.scale = scale;
  P.sl2 = scale * kLog2eF4;
  P.shift2 = shift * kLog2eF4;
  P.sym = sym ? 1 : 0;
  P.row0_global = row0_global;
  P.total_tiles = (Nb + F4_BN - 1) / F4_BN;
  P.row_sum = row_sum; P.col_sum = col_sum;
  P.diag_offset = diag_offset; P.diag_out = diag_out;
  P.store_e = E ? 1 : 0;
  P.njb = 4 * P.total_tiles;
  CUtensorMap tmE;
  if (E) {
#if DMF_F4_PAIR_STORE
    rc = make_tmap_bf16_2d(&tmE, E, 2LL * ((Ma + 255) / 256) * P.njb * 128, 64, 64, 32);    // [32 rows x 64 cols] boxes, 128B swizzle
#else
    rc = make_tmap_bf16_2d_box32_sw64(&tmE, E, 2LL * ((Ma + 255) / 256) * P.njb * 128);
#endif
    if (rc) return rc;
  } else {
    tmE = tmA;      // never dereferenced
  }
  const int pairs = (Ma + 255) / 256;
  const int window = sym ? P.total_tiles / 2 + 1 : P.total_tiles;
  // split the tile window over blockIdx.y so that pairs * nsplit fills whole waves of the 74 clusters
  int nsplit = 1;
  {
    double best = 0.0;
    for (int ns = 1; ns <= 16; ++ns) {
      if (ns > 1 && window / ns < 4) break;
      const int items = pairs * ns;
      const double eff = (double)items / (double)(((items + 73) / 74) * 74);
      if (eff > best + 0.02) { best = eff; nsplit = ns; }
      if (eff >= 0.97) break;
    }
  }
  P.tiles_per_split = (window + nsplit - 1) / nsplit;
  dim3 grid(2 * pairs, nsplit);
  rowcol_sum_tc4_kernel<<<grid, F4_THREADS, smem, (cudaStream_t)s>>>(tmA, tmB, tmE, P);
  return launched("dmf_infonce_rowcol_sums");
}
