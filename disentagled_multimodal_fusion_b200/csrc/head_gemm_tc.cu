// K1 head variant: the LAST encoder layer with the per-row head fused into its epilogue.
//
//   X = A W^T + b                      (Linear.forward, models/classifiers.py:43-48)
//   head 0:  Y = X / max(|X|, eps)     (F.normalize, models/disentangledssl.py:139-140)
//   head 1:  z = vMF rsample given noise (Householder reflection of x = [w, sqrt(1-w^2) v] from e1 onto mu = X/|X|,
//            models/classifiers.py:433-437 with the sample of :314-431 drawn by dmf_vmf_draw)
//
// Both heads need the WHOLE output row, and a row is N = 256 or 512 wide: the pair GEMM (gemm_tc2.cu, 256-column tiles,
// double-buffered accumulators) would put the two halves of a row on different clusters.  Here one CTA pair owns 256
// full rows: tcgen05.mma.cta_group::2, M = 256, N = 256 per instruction, the two column halves of the 128 x N fp32
// accumulator of each CTA fill up to 512 TMEM columns (no double buffering: the layer is 0.14 ms of MMA at C5, its epilogue
// is what matters).  3-stage ring of {A 128 x 64, W-half 128 x 64 per column half} bf16 tiles.
// Epilogue, 8 warps (lane quarter x column half), thread = row:
//   pass 1  TMEM -> x = acc + bias -> row sums (sum x^2; vMF: sum x_noise * x, x_0); the two column halves of a row meet
//           through shared memory
//   pass 2  TMEM again (no HBM re-read of the pre-activation) -> head value -> per-warp staging tile -> COALESCED stores of
//           fp32 + bf16 (gemm_tc_epi.cuh); vMF also stores X itself (fp32 for its backward and the ortho term, bf16 into the
//           conditioning columns of the private encoders' input)
// The vMF noise rows are read coalesced (lane -> 4 rows x 8 x 4 columns) and transposed to the row-per-lane layout through
// the same staging tile.
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_pair.cuh"
#include "gemm_tc_epi.cuh"

namespace dmf {

constexpr int HG_THREADS = 320;
constexpr int HG_TILE = 128 * 64 * 2;     // 16 KB
constexpr int HG_STAGES = 3;
constexpr int HG_MAX_GROUPS = 4;
constexpr int HG_STAGE_BYTES = 3 * HG_TILE;   // A + two W halves
constexpr size_t HG_SMEM = 1024 + (size_t)HG_STAGES * HG_STAGE_BYTES + 8 * kEpiStageFloats * 4 + 128 * 3 * 4 + 256;

struct HGGroup {
  TcEpi pre;            // X outputs (fp32 / bf16), may both be NULL
  TcEpi out;            // head outputs (fp32 / bf16)
  const float* bias;
  float* inv_norm;
  const float* noise_w;
  const float* noise_v;
  float eps;
  int M, N, K;
};
struct alignas(64) HGParams {
  CUtensorMap tmA[HG_MAX_GROUPS];
  CUtensorMap tmW[HG_MAX_GROUPS];
  HGGroup g[HG_MAX_GROUPS];
};

// Coalesced read of a [32 rows x 32 cols] fp32 chunk src[(row0 + r) * ld + col0 + c] (c may start unaligned; elements with
// col0 + c < 0 or row >= M read as 0) into the row-per-lane layout: v[j] = chunk[lane][j].  `stage` = the warp's staging tile.
__device__ __forceinline__ void hg_load_rows(const float* __restrict__ src, long long ld, int row0, int M, int col0, int lane,
                                             float* stage, float (&v)[32]) {
  const int c = (lane & 7) * 4, rq = lane >> 3;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = rq + 4 * it;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + rr < M) {
      const float* p = src + (long long)(row0 + rr) * ld + col0 + c;
      if (col0 + c + 0 >= 0) t.x = __ldg(p + 0);
      if (col0 + c + 1 >= 0) t.y = __ldg(p + 1);
      if (col0 + c + 2 >= 0) t.z = __ldg(p + 2);
      if (col0 + c + 3 >= 0) t.w = __ldg(p + 3);
    }
    *reinterpret_cast<float4*>(stage + rr * kEpiPitch + c) = t;
  }
  __syncwarp();
  const float4* sr = reinterpret_cast<const float4*>(stage + lane * kEpiPitch);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = sr[j];
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
  __syncwarp();
}

template <int HEAD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(HG_THREADS, 1)
head_gemm_tc_kernel(const __grid_constant__ HGParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;
  float* epi_stage = reinterpret_cast<float*>(ring + HG_STAGES * HG_STAGE_BYTES);     // [8 warps][32 x 36]
  float* red = epi_stage + 8 * kEpiStageFloats;                                        // [3][128] partials of column half 1
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(red + 3 * 128);
  uint64_t* empty_bar = full_bar + HG_STAGES;
  uint64_t* acc_full = empty_bar + HG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int gi = blockIdx.y;
  const HGGroup& G = P.g[gi];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc2::cluster_ctarank();
  const bool leader = rank == 0;
  const int m0 = ((int)blockIdx.x >> 1) * 256 + (int)rank * 128;
  const int nh = G.N >> 8;                       // column halves of 256 (N = 256 or 512)
  const int num_kb = (G.K + 63) / 64;
  const bool active = ((int)blockIdx.x >> 1) * 256 < G.M;      // groups may have fewer row blocks than gridDim.x / 2

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&P.tmA[gi]);
    tc::tma_prefetch_desc(&P.tmW[gi]);
    for (int s = 0; s < HG_STAGES; ++s) { tc::mbar_init(full_bar + s, 1); tc::mbar_init(empty_bar + s, 1); }
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc2::tmem_alloc2<512>(tmem_slot);
  tc::tc_fence_before_sync();
  tc2::cluster_sync_all();
  tc::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (active) {
    if (warp == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        tc::mbar_wait(empty_bar + stage, phase ^ 1);
        if (tc::elect_one()) {
          uint8_t* sb = ring + stage * HG_STAGE_BYTES;
          if (leader) tc::mbar_expect_tx(full_bar + stage, 2 * (1 + nh) * HG_TILE);      // bytes of BOTH CTAs
          tc2::tma_load_2d_pair(sb, &P.tmA[gi], kb * 64, m0, full_bar + stage);
          for (int h = 0; h < nh; ++h)
            tc2::tma_load_2d_pair(sb + (1 + h) * HG_TILE, &P.tmW[gi], kb * 64, h * 256 + (int)rank * 128, full_bar + stage);
        }
        __syncwarp();
        if (++stage == HG_STAGES) { stage = 0; phase ^= 1; }
      }
    } else if (warp == 1) {
      if (leader) {
        constexpr uint32_t idesc = tc::make_idesc_bf16(256, 256, 0, 0);
        const uint64_t adesc0 = tc::make_smem_desc(tc::smem_u32(ring), 16, 1024);
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(full_bar + stage, phase);
          tc::tc_fence_after_sync();
          const uint64_t ad = adesc0 + (uint64_t)((stage * HG_STAGE_BYTES) >> 4);
          if (tc::elect_one()) {
            for (int h = 0; h < nh; ++h) {
              const uint64_t bd = ad + (uint64_t)(((1 + h) * HG_TILE) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                tc2::umma_ss2(tmem_base + (uint32_t)(h * 256), ad + 2 * k, bd + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            tc2::umma_commit2(empty_bar + stage);
          }
          __syncwarp();
          if (++stage == HG_STAGES) { stage = 0; phase ^= 1; }
        }
        if (tc::elect_one()) tc2::umma_commit2(acc_full);
        __syncwarp();
      }
    } else {
      const int sw = warp - 2;
      const int q = warp & 3, ch = sw >> 2;
      const int st = threadIdx.x - 64;
      const int rloc = q * 32 + lane;
      const int row0 = m0 + q * 32;
      const int row = row0 + lane;
      const int wcols = G.N >> 1;                  // columns of this warp: [ch * wcols, +wcols)
      float* my_stage = epi_stage + sw * kEpiStageFloats;
      (void)st;
      tc::mbar_wait(acc_full, 0);
      tc::tc_fence_after_sync();
      // vMF noise of this row: x = [w, wt * v]
      float nw = 0.f, wt = 0.f;
      if (HEAD == 1 && row < G.M) {
        nw = __ldg(G.noise_w + row);
        wt = sqrtf(fmaxf(1.0f - nw * nw, 1e-10f));
      }
      // ---- pass 1: row sums over this warp's columns
      float ss = 0.f, dot = 0.f, e0 = 0.f;
#pragma unroll 1
      for (int c = 0; c < wcols / 32; ++c) {
        const int nbase = ch * wcols + c * 32;
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)nbase, r);
        tc::tmem_ld_wait();
        float xn[32];
        if (HEAD == 1) hg_load_rows(G.noise_v, G.N - 1, row0, G.M, nbase - 1, lane, my_stage, xn);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = __uint_as_float(r[j]) + __ldg(G.bias + nbase + j);
          ss = fmaf(x, x, ss);
          if (HEAD == 1) {
            const float xj = (nbase + j == 0) ? nw : wt * xn[j];
            dot = fmaf(xj, x, dot);
            if (nbase + j == 0) e0 = x;
          }
        }
      }
      if (ch == 1) {
        red[rloc] = ss;
        red[128 + rloc] = dot;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (ch == 0) {
        ss += red[rloc];
        dot += red[128 + rloc];
        red[rloc] = ss;
        red[128 + rloc] = dot;
        red[256 + rloc] = e0;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      ss = red[rloc];
      dot = red[128 + rloc];
      e0 = red[256 + rloc];
      // per-row scalars
      float inv, up0 = 0.f, c2 = 0.f;
      if (HEAD == 0) {
        inv = 1.0f / fmaxf(sqrtf(ss), G.eps);
        if (ch == 0 && row < G.M && G.inv_norm) G.inv_norm[row] = inv;
      } else {
        inv = 1.0f / sqrtf(ss);
        up0 = 1.0f - e0 * inv;                                   // u' = e1 - mu
        const float nu2 = fmaf(up0, up0, inv * inv * fmaxf(ss - e0 * e0, 0.f));
        const float xu = nw * up0 - inv * (dot - nw * e0);       // <x, u'>
        const float inv_d = 1.0f / (sqrtf(nu2) + 1e-5f);
        c2 = 2.0f * xu * inv_d * inv_d;
      }
      // ---- pass 2: head values, coalesced stores through the staging tile
#pragma unroll 1
      for (int c = 0; c < wcols / 32; ++c) {
        const int nbase = ch * wcols + c * 32;
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)nbase, r);
        tc::tmem_ld_wait();
        float xn[32];
        if (HEAD == 1) hg_load_rows(G.noise_v, G.N - 1, row0, G.M, nbase - 1, lane, my_stage, xn);
        uint32_t o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = __uint_as_float(r[j]) + __ldg(G.bias + nbase + j);
          r[j] = __float_as_uint(x);
          if (HEAD == 0) {
            o[j] = __float_as_uint(x * inv);
          } else {
            const bool first = nbase + j == 0;
            const float xj = first ? nw : wt * xn[j];
            const float up = (first ? 1.0f : 0.0f) - x * inv;
            o[j] = __float_as_uint(xj - c2 * up);
          }
        }
        if (G.pre.out_f32 || G.pre.out_bf16) tc_epilogue_chunk<DMF_EPI_NONE, false>(G.pre, r, row0, lane, nbase, false, my_stage);
        tc_epilogue_chunk<DMF_EPI_NONE, false>(G.out, o, row0, lane, nbase, false, my_stage);
      }
    }
  }
  __syncwarp();
  tc::tc_fence_before_sync();
  tc2::cluster_sync_all();
  if (warp == 1) tc2::tmem_dealloc2<512>(tmem_base);
}

}  // namespace dmf

using namespace dmf;

extern "C" int dmf_head_gemm_bf16(const dmf_head_gemm_desc* groups, int n_groups, int head, dmf_stream_t s) {
  DMF_REQUIRE(groups && n_groups >= 1 && n_groups <= HG_MAX_GROUPS, "dmf_head_gemm_bf16: 1..%d groups", HG_MAX_GROUPS);
  DMF_REQUIRE(head == 0 || head == 1, "dmf_head_gemm_bf16: head must be 0 (row-normalise) or 1 (vMF sample)");
  HGParams P;
  int max_blocks = 0;
  for (int i = 0; i < n_groups; ++i) {
    const dmf_head_gemm_desc& d = groups[i];
    DMF_REQUIRE(d.A && d.W && d.bias && d.M >= 1 && d.K >= 1, "dmf_head_gemm_bf16: bad operands in group %d", i);
    DMF_REQUIRE(d.N == 256 || d.N == 512, "dmf_head_gemm_bf16: N=%d must be 256 or 512 (one CTA pair owns whole rows)", d.N);
    DMF_REQUIRE(d.out_f32 || d.out_bf16, "dmf_head_gemm_bf16: group %d has no head output", i);
    DMF_REQUIRE(head == 0 || (d.noise_w && d.noise_v), "dmf_head_gemm_bf16: the vMF head needs noise_w [M] and noise_v [M, N-1]");
    int rc = make_tmap_bf16_2d(&P.tmA[i], d.A, d.M, d.K, d.lda, 128);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&P.tmW[i], d.W, d.N, d.K, d.ldw, 128);
    if (rc) return rc;
    HGGroup& g = P.g[i];
    g.pre = TcEpi{d.pre_f32, d.ld_pre_f32, d.pre_bf16, d.ld_pre_bf16, nullptr, 0, nullptr, nullptr, 0, d.M, d.N, 0};
    g.pre.fast = tc_epi_fast_ok(g.pre);
    g.out = TcEpi{d.out_f32, d.ld_out_f32, d.out_bf16, d.ld_out_bf16, nullptr, 0, nullptr, nullptr, 0, d.M, d.N, 0};
    g.out.fast = tc_epi_fast_ok(g.out);
    g.bias = d.bias; g.inv_norm = d.inv_norm; g.noise_w = d.noise_w; g.noise_v = d.noise_v;
    g.eps = d.eps; g.M = d.M; g.N = d.N; g.K = d.K;
    const int blocks = (d.M + 255) / 256;
    if (blocks > max_blocks) max_blocks = blocks;
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(head_gemm_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HG_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(head_gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HG_SMEM);
    if (e != cudaSuccess) return fail((int)e, "dmf_head_gemm_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  dim3 grid(2 * max_blocks, n_groups);
  if (head == 0) head_gemm_tc_kernel<0><<<grid, HG_THREADS, HG_SMEM, (cudaStream_t)s>>>(P);
  else head_gemm_tc_kernel<1><<<grid, HG_THREADS, HG_SMEM, (cudaStream_t)s>>>(P);
  return launched("dmf_head_gemm_bf16");
}
