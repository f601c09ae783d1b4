// K1, CTA-pair generation: persistent grouped  C = epi(A * B^T)  with tcgen05.mma.cta_group::2.
//
// Replaces the cuBLAS addmm/mm chain of Linear.forward (models/classifiers.py:43-48) and the mm pairs that
// autograd runs for it in backward, for ALL groups (modalities x {orig, augmented} streams) of one MLP layer
// in one launch.  Both operands K-major bf16 in global memory (A [M,K], B [N,K]).
//
// One cluster = 2 CTAs = one 256 x 256 output tile at a time (M split 128/128 across the pair, each CTA
// stages HALF of the B tile: 64 B/clk of smem operand traffic per CTA instead of 128).  Clusters are
// persistent: work items (group, tile_m, tile_n, k_split) are dealt round-robin, tile_n fastest so that
// concurrently running clusters share A rows in L2.  Per CTA:
//   warp 0     TMA producer: 5-stage ring of {A 128x64, B-half 128x64} bf16 tiles (128B swizzle); all loads
//              complete on the LEADER's full barrier
//   warp 1     (leader only) MMA issuer: 4 x tcgen05.mma 256x256x16 per stage, commit -> empty barrier of
//              both CTAs; accumulators DOUBLE-BUFFERED in TMEM (2 x 256 columns) so the epilogue of item i
//              overlaps the main loop of item i+1
//   warps 2-9  epilogue (G2_EW = 8): tcgen05.ld (lane quarter x column half), bias / ReLU / ReLU-mask, fp32 / bf16 /
//              transposed-bf16 stores; split-K items accumulate with red.global.add.v4.f32
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_pair.cuh"
#include "gemm_tc_epi.cuh"

namespace dmf {

// 16 epilogue warps (tried: 96 registers with small spills, 4-stage ring): K = 512 layers 934 -> 999 TFLOP/s, but K = 1024 / 1536
// 1 223 -> 1 183 / 1 278 -> 1 259 and the MLP phases of the step unchanged within noise (6.97 vs 7.1 ms): 8 stays the default.
#ifndef DMF_G2_EW
#define DMF_G2_EW 8
#endif
constexpr int G2_EW = DMF_G2_EW;                      // epilogue warps per CTA (8 or 16): 4 lane quarters x G2_EW / 4 column groups
constexpr int G2_THREADS = 64 + 32 * G2_EW;
constexpr int G2_WCOLS = 256 / (G2_EW / 4);           // columns of the 256-wide tile per epilogue warp
constexpr int G2_TILE = 128 * 64 * 2;   // 16 KB
constexpr int G2_STAGES = G2_EW == 16 ? 4 : 5;        // 16 staging tiles (74 KB) leave room for a 4-stage ring
constexpr int G2_BN = 256;
constexpr int kMaxG2Groups = 8;
constexpr size_t G2_SMEM_BYTES = 1024 + (size_t)G2_STAGES * 2 * G2_TILE + G2_EW * kEpiStageFloats * 4 + 512;

struct G2Group {
  TcEpi epi;
  int K;
  int mn;              // both operands MN-major (A [K, M], B [K, N]): 64 x 64 boxes, two per operand and stage
  int tiles_n, splits, kb_per_split, num_kb;
  int atomic;          // fp32 output accumulated with red.add (split-K partial tiles, or accumulate-into-C mode)
  int item_start;
};
struct alignas(64) G2Params {
  CUtensorMap tmA[kMaxG2Groups];
  CUtensorMap tmB[kMaxG2Groups];
  G2Group g[kMaxG2Groups];
  int n;
  int total_items;
};

struct G2Item {
  int gi, m0, n0, kb0, kb1, ks;
};

__device__ __forceinline__ G2Item g2_decode(const G2Params& P, int item) {
  int gi = 0;
  while (gi + 1 < P.n && item >= P.g[gi + 1].item_start) ++gi;
  const G2Group& g = P.g[gi];
  int lt = item - g.item_start;
  const int per_m = g.tiles_n * g.splits;
  const int tm = lt / per_m;
  lt -= tm * per_m;
  const int ks = lt / g.tiles_n;
  const int tn = lt - ks * g.tiles_n;
  G2Item it;
  it.gi = gi;
  it.m0 = tm * 256;
  it.n0 = tn * G2_BN;
  it.ks = ks;
  it.kb0 = ks * g.kb_per_split;
  it.kb1 = min(g.num_kb, it.kb0 + g.kb_per_split);
  return it;
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm_bf16_tc2_kernel(const __grid_constant__ G2Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + G2_STAGES * G2_TILE;
  float* epi_stage = reinterpret_cast<float*>(smem + 2 * G2_STAGES * G2_TILE);      // [G2_EW warps][32 x 36]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + G2_EW * kEpiStageFloats);
  uint64_t* empty_bar = full_bar + G2_STAGES;
  uint64_t* acc_full = empty_bar + G2_STAGES;   // [2] per CTA (multicast commit)
  uint64_t* acc_empty = acc_full + 2;           // [2] leader: 16 epilogue warps of the pair drained the buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc2::cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < P.n; ++i) {
      tc::tma_prefetch_desc(&P.tmA[i]);
      tc::tma_prefetch_desc(&P.tmB[i]);
    }
    for (int s = 0; s < G2_STAGES; ++s) { tc::mbar_init(full_bar + s, 1); tc::mbar_init(empty_bar + s, 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(acc_full + b, 1); tc::mbar_init(acc_empty + b, 2 * G2_EW); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc2::tmem_alloc2<512>(tmem_slot);
  tc::tc_fence_before_sync();
  tc2::cluster_sync_all();
  tc::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    {
      // whole warp, uniform control flow; one elected lane issues the TMA instructions
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < P.total_items; item += num_clusters) {
        const G2Item it = g2_decode(P, item);
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          tc::mbar_wait(empty_bar + stage, phase ^ 1);
          if (tc::elect_one()) {
            if (leader) tc::mbar_expect_tx(full_bar + stage, 4 * G2_TILE);      // A + B-half of BOTH CTAs
            if (!P.g[it.gi].mn) {
              tc2::tma_load_2d_pair(smemA + stage * G2_TILE, &P.tmA[it.gi], kb * 64, it.m0 + (int)rank * 128, full_bar + stage);
              tc2::tma_load_2d_pair(smemB + stage * G2_TILE, &P.tmB[it.gi], kb * 64, it.n0 + (int)rank * 128, full_bar + stage);
            } else {
              // MN-major: the stage holds [64 k-rows] x [2 x 64 contiguous m (or n)] per operand, chunks 8 KB apart
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                tc2::tma_load_2d_pair(smemA + stage * G2_TILE + c * (G2_TILE / 2), &P.tmA[it.gi],
                                      it.m0 + (int)rank * 128 + c * 64, kb * 64, full_bar + stage);
                tc2::tma_load_2d_pair(smemB + stage * G2_TILE + c * (G2_TILE / 2), &P.tmB[it.gi],
                                      it.n0 + (int)rank * 128 + c * 64, kb * 64, full_bar + stage);
              }
            }
          }
          __syncwarp();
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // whole warp, uniform control flow; one elected lane issues the tcgen05 instructions
      constexpr uint32_t idesc = tc::make_idesc_bf16(256, G2_BN, 0, 0);
      constexpr uint32_t idesc_mn = tc::make_idesc_bf16(256, G2_BN, 1, 1);
      const uint64_t adesc0 = tc::make_smem_desc(tc::smem_u32(smemA), 16, 1024);
      const uint64_t bdesc0 = tc::make_smem_desc(tc::smem_u32(smemB), 16, 1024);
      // MN-major view of a stage: LBO = distance of the two 64-wide chunks, SBO = next 8 k-rows; UMMA_K = 16 rows = 2 KB
      const uint64_t adesc0_mn = tc::make_smem_desc(tc::smem_u32(smemA), G2_TILE / 2, 1024);
      const uint64_t bdesc0_mn = tc::make_smem_desc(tc::smem_u32(smemB), G2_TILE / 2, 1024);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t n = 0;
      for (int item = cluster_id; item < P.total_items; item += num_clusters, ++n) {
        const G2Item it = g2_decode(P, item);
        const uint32_t buf = n & 1;
        tc::mbar_wait(acc_empty + buf, ((n >> 1) & 1) ^ 1);
        tc::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + buf * G2_BN;
        const bool mn = P.g[it.gi].mn != 0;
        const uint32_t id = mn ? idesc_mn : idesc;
        const uint64_t kstep = mn ? (uint64_t)(2048 >> 4) : (uint64_t)2;
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          tc::mbar_wait(full_bar + stage, phase);
          tc::tc_fence_after_sync();
          const uint64_t ad = (mn ? adesc0_mn : adesc0) + (uint64_t)((stage * G2_TILE) >> 4);
          const uint64_t bd = (mn ? bdesc0_mn : bdesc0) + (uint64_t)((stage * G2_TILE) >> 4);
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc2::umma_ss2(d_tmem, ad + kstep * k, bd + kstep * k, id, (kb > it.kb0 || k > 0) ? 1u : 0u);
            tc2::umma_commit2(empty_bar + stage);
          }
          __syncwarp();
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
        if (tc::elect_one()) tc2::umma_commit2(acc_full + buf);
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;                  // TMEM lane quarter
    const int ch = (warp - 2) >> 2;          // column group (G2_WCOLS columns) of the 256-wide tile
    const uint32_t acc_empty_leader = tc2::mapa(tc::smem_u32(acc_empty), 0);
    float* my_stage = epi_stage + (warp - 2) * kEpiStageFloats;
    uint32_t n = 0;
    for (int item = cluster_id; item < P.total_items; item += num_clusters, ++n) {
      const G2Item it = g2_decode(P, item);
      const G2Group& g = P.g[it.gi];
      const uint32_t buf = n & 1;
      const int row0 = it.m0 + (int)rank * 128 + q * 32;
      tc::mbar_wait(acc_full + buf, (n >> 1) & 1);
      tc::tc_fence_after_sync();
      if (it.kb1 > it.kb0) {
        // the TMEM read of chunk c + 1 is in flight while chunk c goes through the staging tile
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + buf * G2_BN + (uint32_t)(ch * G2_WCOLS);
        const int nb0 = it.n0 + ch * G2_WCOLS;
        uint32_t ra[32], rb[32];
        auto emit = [&](const uint32_t (&r)[32], int c) {
          if (g.atomic) tc_epilogue_chunk<EPI, true>(g.epi, r, row0, lane, nb0 + c * 32, it.ks == 0, my_stage);
          else tc_epilogue_chunk<EPI, false>(g.epi, r, row0, lane, nb0 + c * 32, true, my_stage);
        };
        tc::tmem_ld_32x32(t0, ra);
#pragma unroll 1
        for (int c2 = 0; c2 < G2_WCOLS / 64; ++c2) {
          tc::tmem_ld_wait();
          tc::tmem_ld_32x32(t0 + (uint32_t)(c2 * 64 + 32), rb);
          emit(ra, 2 * c2);
          tc::tmem_ld_wait();
          if (c2 + 1 < G2_WCOLS / 64) tc::tmem_ld_32x32(t0 + (uint32_t)((c2 + 1) * 64), ra);
          emit(rb, 2 * c2 + 1);
        }
      }
      tc::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) tc2::mbar_arrive_cluster(acc_empty_leader + buf * 8);
    }
  }
  __syncwarp();
  tc::tc_fence_before_sync();
  tc2::cluster_sync_all();          // the peer may still target this CTA's barriers / TMEM until here
  if (warp == 1) tc2::tmem_dealloc2<512>(tmem_base);
}

template <int EPI>
static int launch_g2(const G2Params& P, int clusters, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tc2_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)G2_SMEM_BYTES);
    if (e != cudaSuccess) return fail((int)e, "gemm_bf16_tc2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  gemm_bf16_tc2_kernel<EPI><<<2 * clusters, G2_THREADS, G2_SMEM_BYTES, st>>>(P);
  return launched("dmf_grouped_gemm_bf16_tc(pair)");
}

// Host entry used by dmf_grouped_gemm_bf16_tc for groups large enough to feed CTA pairs.
int launch_gemm_tc2(const dmf_tc_gemm_desc* groups, int n_groups, int epilogue, cudaStream_t st) {
  constexpr int kClusters = kNumSMs / 2;
  for (int base = 0; base < n_groups; base += kMaxG2Groups) {
    G2Params P;
    P.n = 0;
    int items = 0, tiles_total = 0;
    const int cnt = n_groups - base < kMaxG2Groups ? n_groups - base : kMaxG2Groups;
    for (int i = 0; i < cnt; ++i) {
      const dmf_tc_gemm_desc& d = groups[base + i];
      if (d.M == 0 || d.N == 0) continue;
      tiles_total += ((d.M + 255) / 256) * ((d.N + G2_BN - 1) / G2_BN);
    }
    for (int i = 0; i < cnt; ++i) {
      const dmf_tc_gemm_desc& d = groups[base + i];
      if (d.M == 0 || d.N == 0) continue;
      // MN-major operands: the global tensors are [K, M] / [K, N]; boxes of 64 k-rows x 64 columns
      int rc = d.mn_major ? make_tmap_bf16_2d(&P.tmA[P.n], d.A, d.K, d.M, d.lda, 64)
                          : make_tmap_bf16_2d(&P.tmA[P.n], d.A, d.M, d.K, d.lda, 128);
      if (rc) return rc;
      rc = d.mn_major ? make_tmap_bf16_2d(&P.tmB[P.n], d.B, d.K, d.N, d.ldb, 64)
                      : make_tmap_bf16_2d(&P.tmB[P.n], d.B, d.N, d.K, d.ldb, 128);
      if (rc) return rc;
      G2Group& g = P.g[P.n];
      g.mn = d.mn_major ? 1 : 0;
      g.epi.out_f32 = d.out_f32; g.epi.ldo_f32 = d.ldo_f32;
      g.epi.out_bf16 = d.out_bf16; g.epi.ldo_bf16 = d.ldo_bf16;
      g.epi.out_t = d.out_bf16_t; g.epi.ldo_t = d.ldo_t;
      g.epi.bias = d.bias; g.epi.mask = d.mask_bf16; g.epi.ldmask = d.ldmask;
      g.epi.M = d.M; g.epi.N = d.N;
      g.epi.fast = tc_epi_fast_ok(g.epi);
      g.K = d.K;
      g.num_kb = (d.K + 63) / 64;
      g.tiles_n = (d.N + G2_BN - 1) / G2_BN;
      const int tiles = ((d.M + 255) / 256) * g.tiles_n;
      // split-K: only for plain fp32 accumulation (wgrad: tiny output, K = batch); the caller zeroes out_f32
      int splits = 1;
      const bool accum = d.split_k < 0 && epilogue == DMF_EPI_NONE && d.out_f32 && !d.out_bf16 && !d.out_bf16_t;
      if (d.split_k != 1 && epilogue == DMF_EPI_NONE && d.out_f32 && !d.out_bf16 && !d.out_bf16_t) {
        const int max_splits = g.num_kb / 8 > 0 ? (g.num_kb / 8 < 64 ? g.num_kb / 8 : 64) : 1;
        if (d.split_k > 1) {
          splits = d.split_k < max_splits ? d.split_k : max_splits;
        } else {
          // smallest split count whose item total fills whole waves of the 74 clusters to >= 95 %
          double best = 0.0;
          for (int sp = 1; sp <= max_splits; ++sp) {
            const int it = tiles_total * sp;
            const double eff = (double)it / (double)(((it + kClusters - 1) / kClusters) * kClusters);
            if (eff > best + 1e-9) { best = eff; splits = sp; }
            if (eff >= 0.95) break;
          }
        }
      }
      g.kb_per_split = (g.num_kb + splits - 1) / splits;
      g.splits = (g.num_kb + g.kb_per_split - 1) / g.kb_per_split;
      g.atomic = (g.splits > 1 || accum) ? 1 : 0;
      g.item_start = items;
      items += tiles * g.splits;
      ++P.n;
    }
    if (P.n == 0) continue;
    P.total_items = items;
    const int clusters = items < kClusters ? items : kClusters;
    int rc;
    switch (epilogue) {
      case DMF_EPI_NONE: rc = launch_g2<DMF_EPI_NONE>(P, clusters, st); break;
      case DMF_EPI_BIAS: rc = launch_g2<DMF_EPI_BIAS>(P, clusters, st); break;
      case DMF_EPI_BIAS_RELU: rc = launch_g2<DMF_EPI_BIAS_RELU>(P, clusters, st); break;
      case DMF_EPI_RELU_MASK: rc = launch_g2<DMF_EPI_RELU_MASK>(P, clusters, st); break;
      default: return fail(-1, "dmf_grouped_gemm_bf16_tc: unsupported epilogue %d", epilogue);
    }
    if (rc) return rc;
  }
  return 0;
}

}  // namespace dmf
