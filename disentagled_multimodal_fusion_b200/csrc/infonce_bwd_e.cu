// K2 backward from STORED probabilities: the forward pass (infonce_fwd_tc4.cu) keeps e_ij = exp(s_ij - shift) of the
// cross block as bf16 in HBM (B x B x 2 bytes: 8.6 GB per critic call at B = 65536 -- this is what 180 GB of HBM3e
// buys), so the backward never recomputes S = A B^T: it is ONE product per gradient,
//
//   dir 0:  dOut[i, :] = coef*g * ( sum_j e_ij (fa_i + fb_j) Z[j, :]  -  2 Z[i + diag_offset, :] )     Z = column block
//   dir 1:  dOut[j, :] = coef*g * ( sum_i e_ij (fa_i + fb_j) Z[i, :]  -  2 Z[j - diag_offset, :] )     Z = anchor block
//
// with fa_i = exp(shift - lseA_i), fb_j = exp(shift - lseB_j), i.e. e_ij (fa_i + fb_j) = exp(s_ij - lseA_i) +
// exp(s_ij - lseB_j) of SURVEY App. B (models/losses.py:64-99 under autograd).  Executed FLOPs = algorithmic 2*M*N*D
// (the recompute kernels execute 2x: infonce_bwd_tc6.cu, or 3x: infonce_bwd_tc3.cu).
//
// E layout: blocks of [128 rows i x 64 columns j] bf16 (16 KB, row-major inside), block (ib, jb) at
// ((ib * njb + jb) * 16 KB): ONE block is the K-major A operand of dir 0 (M = 128 i, K = 64 j) and, read as two
// [64 i x 64 j] halves, the MN-major A operand of dir 1 (M = 64 j per block, K = 64 i), so both directions stream
// 8-16 KB contiguous pieces and no transposed copy of E exists.
//
// CTA pair, tcgen05.mma.cta_group::2, M = 256 (128 output rows per CTA), N = 256 per instruction, the D = 512 wide
// fp32 output row block fills the 512 TMEM columns.  Per ring stage (K = 64 of the contraction index):
//   TMA:    E piece 16 KB (local mbarrier) + 64 streamed factors (256 B bulk copy) + Z piece [64 k x 128 d] per output
//           half as two MN-major [64 x 64] boxes (pair loads completing the leader's mbarrier)
//   scale:  8 warps multiply the E piece IN PLACE by (f_own + f_str) -> bf16 (conflict-free 16-byte accesses on the
//           128B-swizzled tile), fence.proxy.async, arrive on the leader's w_full
//   MMA:    4 k-steps x (D / 256) halves, commit -> stage empty in both CTAs
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_pair.cuh"

namespace dmf {

constexpr int BE_THREADS = 320;
constexpr int BE_EBYTES = 16384;       // E piece of a stage
constexpr int BE_ZHALF = 16384;        // Z piece of one output half: 2 x [64 k x 64 d]
constexpr int BE_FBYTES = 256;         // 64 streamed factors
constexpr int BE_SMEM = 232448;        // the whole opt-in window
constexpr int BE_MAX_STAGES = 8;

struct BEArgs {
  int n_own, n_str;          // output rows / contraction length
  int nh;                    // D / 256
  int njb;                   // E column blocks (64 wide) per row block
  int stages;
  int ksteps_total;          // ceil(n_str / 64)
  int ksteps_per_split;
  const float* f_own;        // padded to a multiple of 256, zeros beyond n_own
  const float* f_str;        // padded to a multiple of 64, zeros beyond n_str
  float coef;
  const float* gscale;
  long long pos_off;         // positive of output row r is streamed index r + pos_off
  const uint16_t* Z;
  long long ldz;
  float* dOut;
  long long ldo;
  int accumulate;
};

__device__ __forceinline__ uint32_t be_pack(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void be_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tc::smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}

// explicit shared-window accesses (the carve-up pointer went through integer alignment arithmetic, so plain
// dereferences compile to generic LD / ST)
__device__ __forceinline__ uint4 be_lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 be_lds128f(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float be_lds32f(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void be_sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

template <int DIR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BE_THREADS, 1)
infonce_bwd_e_kernel(const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmZ, const BEArgs P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = BE_EBYTES + P.nh * BE_ZHALF;
  uint8_t* ring = smem;
  float* fsm = reinterpret_cast<float*>(ring + P.stages * stage_bytes);     // [stages][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(fsm + P.stages * 64);
  uint64_t* z_full = bars;                          // leader: Z pieces of both CTAs landed
  uint64_t* e_full = z_full + BE_MAX_STAGES;        // local: E piece + factors landed
  uint64_t* w_full = e_full + BE_MAX_STAGES;        // leader: 16 scale warps of the pair finished the stage
  uint64_t* empty_bar = w_full + BE_MAX_STAGES;     // both: MMAs of the stage retired
  uint64_t* acc_full = empty_bar + BE_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  if (threadIdx.x == 0 &&
      (reinterpret_cast<uint8_t*>(tmem_slot + 4) - smem_raw) > BE_SMEM) {
    printf("dmf: infonce_bwd_e: shared-memory carve-up does not fit\n");
    __trap();
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc2::cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = (int)blockIdx.x >> 1;
  const int m0 = pair * 256 + (int)rank * 128;          // first output row of THIS CTA
  const int k0 = (int)blockIdx.z * P.ksteps_per_split;
  const int k1 = min(P.ksteps_total, k0 + P.ksteps_per_split);
  const int nk = max(0, k1 - k0);
  const bool split = gridDim.z > 1 || P.accumulate != 0;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmE);
    tc::tma_prefetch_desc(&tmZ);
    for (int s = 0; s < P.stages; ++s) {
      tc::mbar_init(z_full + s, 1);
      tc::mbar_init(e_full + s, 1);
      tc::mbar_init(w_full + s, 16);
      tc::mbar_init(empty_bar + s, 1);
    }
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc2::tmem_alloc2<512>(tmem_slot);
  tc::tc_fence_before_sync();
  tc2::cluster_sync_all();
  tc::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // TMA producer: whole warp, uniform control flow; one elected lane issues
    int stage = 0;
    uint32_t phase = 0;
    const uint64_t e_policy = tc::l2_policy_evict_first();      // E is read exactly once: keep the Z block in L2 instead
    for (int k = k0; k < k1; ++k) {
      tc::mbar_wait(empty_bar + stage, phase ^ 1);
      if (tc::elect_one()) {
        uint8_t* sb = ring + stage * stage_bytes;
        if (leader) tc::mbar_expect_tx(z_full + stage, 2 * P.nh * BE_ZHALF);     // bytes of BOTH CTAs
        tc::mbar_expect_tx(e_full + stage, BE_EBYTES + BE_FBYTES);
        if (DIR == 0) {
          // block (ib = m0 / 128, jb = k): rows 0-63 and 64-127 are adjacent in shared memory = one [128 x 64] K-major tile
          const int row = ((m0 >> 7) * P.njb + k) * 128;
          tc::tma_load_2d_hint(sb, &tmE, 0, row, e_full + stage, e_policy);
          tc::tma_load_2d_hint(sb + 8192, &tmE, 0, row + 64, e_full + stage, e_policy);
        } else {
          // rows i = 64 k .. +63 of the two blocks that hold this CTA's 128 columns j: two MN-major 64-j chunks
          const int ib = k >> 1, hh = k & 1, jb0 = m0 >> 6;
#pragma unroll
          for (int c = 0; c < 2; ++c)
            tc::tma_load_2d_hint(sb + c * 8192, &tmE, 0, (ib * P.njb + jb0 + c) * 128 + hh * 64, e_full + stage, e_policy);
        }
        be_bulk_load(fsm + stage * 64, P.f_str + (size_t)k * 64, BE_FBYTES, e_full + stage);
        for (int h = 0; h < P.nh; ++h)
#pragma unroll
          for (int c = 0; c < 2; ++c)
            tc2::tma_load_2d_pair(sb + BE_EBYTES + h * BE_ZHALF + c * 8192, &tmZ, h * 256 + (int)rank * 128 + c * 64, k * 64,
                                  z_full + stage);
      }
      __syncwarp();
      if (++stage == P.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    if (leader) {
      // MMA issuer: whole warp, uniform control flow; one elected lane issues the tcgen05 instructions
      constexpr uint32_t idesc = tc::make_idesc_bf16(256, 256, DIR, 1);
      // A: dir 0 K-major [128 x 64] (k-step = 32 B); dir 1 MN-major, 64-j chunks 8 KB apart (k-step = 16 rows = 2 KB)
      const uint64_t adesc0 = DIR == 0 ? tc::make_smem_desc(tc::smem_u32(ring), 16, 1024)
                                       : tc::make_smem_desc(tc::smem_u32(ring), 8192, 1024);
      const uint64_t astep = DIR == 0 ? 2 : (2048 >> 4);
      const uint64_t bdesc0 = tc::make_smem_desc(tc::smem_u32(ring + BE_EBYTES), 8192, 1024);
      int stage = 0;
      uint32_t phase = 0;
      for (int k = k0; k < k1; ++k) {
        tc::mbar_wait(z_full + stage, phase);
        tc::mbar_wait(w_full + stage, phase);
        tc::tc_fence_after_sync();
        const uint64_t ad = adesc0 + (uint64_t)((stage * stage_bytes) >> 4);
        if (tc::elect_one()) {
          for (int h = 0; h < P.nh; ++h) {
            const uint64_t bd = bdesc0 + (uint64_t)((stage * stage_bytes + h * BE_ZHALF) >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              tc2::umma_ss2(tmem_base + (uint32_t)(h * 256), ad + astep * kk, bd + (uint64_t)((2048 >> 4) * kk), idesc,
                            (k > k0 || kk > 0) ? 1u : 0u);
          }
          tc2::umma_commit2(empty_bar + stage);
        }
        __syncwarp();
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
      if (tc::elect_one()) tc2::umma_commit2(acc_full);
      __syncwarp();
    }
  } else {
    const int sw = warp - 2;                 // 0..7
    const int st = threadIdx.x - 64;         // 0..255
    // ---- scale pass: thread = (sub-block sbk of 64 tile rows, row r, half = 4 of the row's 8 16-byte chunks)
    const int sbk = st >> 7, r = (st & 127) >> 1, half = st & 1;
    const int sx = r & 7;
    const uint32_t toff = (uint32_t)(sbk * 8192 + r * 128);
    // dir 0: tile rows are OWN rows (one factor per thread), columns streamed (32 factors per stage from smem)
    // dir 1: tile rows are streamed (one factor per stage), columns OWN (32 factors per thread, constant)
    float fo1 = 0.f;
    float fo[32];
    if (DIR == 0) {
      fo1 = __ldg(P.f_own + m0 + sbk * 64 + r);
    } else {
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(P.f_own + m0 + sbk * 64 + half * 32 + e));
        fo[e] = v.x; fo[e + 1] = v.y; fo[e + 2] = v.z; fo[e + 3] = v.w;
      }
    }
    const uint32_t w_full_leader = tc2::mapa(tc::smem_u32(w_full), 0);
    const uint32_t ring_u32 = tc::smem_u32(ring), fsm_u32 = tc::smem_u32(fsm);
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int k = k0; k < k1; ++k) {
        tc::mbar_wait(e_full + stage, phase);
        const uint32_t eb = ring_u32 + (uint32_t)(stage * stage_bytes) + toff;
        const uint32_t fs = fsm_u32 + (uint32_t)(stage * 256);
        float a;
        float b[32];
        if (DIR == 0) {
          a = fo1;
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 v = be_lds128f(fs + (uint32_t)((half * 32 + e) * 4));
            b[e] = v.x; b[e + 1] = v.y; b[e + 2] = v.z; b[e + 3] = v.w;
          }
        } else {
          a = be_lds32f(fs + (uint32_t)(r * 4));
#pragma unroll
          for (int e = 0; e < 32; ++e) b[e] = fo[e];
        }
        uint4 v[4];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) v[c4] = be_lds128(eb + (uint32_t)(((half * 4 + c4) ^ sx) << 4));
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const uint32_t w[4] = {v[c4].x, v[c4].y, v[c4].z, v[c4].w};
          uint32_t o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float lo = __uint_as_float(w[q] << 16);
            const float hi = __uint_as_float(w[q] & 0xffff0000u);
            o[q] = be_pack(lo * (a + b[c4 * 8 + 2 * q]), hi * (a + b[c4 * 8 + 2 * q + 1]));
          }
          be_sts128(eb + (uint32_t)(((half * 4 + c4) ^ sx) << 4), o[0], o[1], o[2], o[3]);
        }
        tc::fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) tc2::mbar_arrive_cluster(w_full_leader + (uint32_t)(stage * 8));
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
    }
    // ---- epilogue: warp (q, ch) stores columns [ch * D/2, +D/2) of TMEM lanes [32 q, +32)
    const int q = warp & 3, ch = sw >> 2;
    const int row = m0 + q * 32 + lane;
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after_sync();
    const float cg = P.coef * (P.gscale ? __ldg(P.gscale) : 1.0f);
    const long long pj = (long long)row + P.pos_off;
    const bool has_pos = pj >= 0 && pj < P.n_str && blockIdx.z == 0;   // the positive term is added by split 0 only
    const int wcols = P.nh * 128;                                      // columns of this warp
#pragma unroll 1
    for (int c = 0; c < wcols / 32; ++c) {
      const int dbase = ch * wcols + c * 32;
      uint32_t rr[32];
      tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)dbase, rr);
      tc::tmem_ld_wait();
      if (row < P.n_own && nk > 0) {
        float* dst = P.dOut + (long long)row * P.ldo + dbase;
        const uint16_t* bp = has_pos ? P.Z + pj * P.ldz + dbase : nullptr;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o;
          o.x = (__uint_as_float(rr[j + 0]) - (has_pos ? 2.0f * bf2f(bp[j + 0]) : 0.f)) * cg;
          o.y = (__uint_as_float(rr[j + 1]) - (has_pos ? 2.0f * bf2f(bp[j + 1]) : 0.f)) * cg;
          o.z = (__uint_as_float(rr[j + 2]) - (has_pos ? 2.0f * bf2f(bp[j + 2]) : 0.f)) * cg;
          o.w = (__uint_as_float(rr[j + 3]) - (has_pos ? 2.0f * bf2f(bp[j + 3]) : 0.f)) * cg;
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

// f[i] = exp(shift - lse[i]) for i < n, 0 for the padding up to n_pad
__global__ void infonce_factors_kernel(const float* __restrict__ lse, int n, int n_pad, float shift, float* __restrict__ f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) f[i] = i < n ? expf(shift - lse[i]) : 0.f;
}

}  // namespace dmf

using namespace dmf;

static inline long long be_nib(int Ma) { return 2LL * ((Ma + 255) / 256); }
static inline long long be_njb(int Nb) { return 4LL * ((Nb + 255) / 256); }

extern "C" size_t dmf_infonce_e_bytes(int Ma, int Nb) {
  if (Ma <= 0 || Nb <= 0) return 0;
  return (size_t)(be_nib(Ma) * be_njb(Nb)) * 16384u;
}

extern "C" size_t dmf_infonce_bwd_stored_work_floats(int Ma, int Nb) {
  return (size_t)(((Ma + 255) / 256) * 256 + ((Nb + 255) / 256) * 256);
}

extern "C" int dmf_infonce_bwd_stored(const void* E, int Ma, int Nb, const float* lseA, const float* lseB, float shift,
                                      const void* Z, long long ldz, int D, int dir, float coef, const float* gscale,
                                      long long diag_offset, float* dOut, long long ldo, int accumulate, float* work,
                                      dmf_stream_t s_) {
  cudaStream_t s = (cudaStream_t)s_;
  DMF_REQUIRE(E && lseA && lseB && Z && dOut && work, "dmf_infonce_bwd_stored: null argument");
  DMF_REQUIRE(Ma >= 1 && Nb >= 1, "dmf_infonce_bwd_stored: bad shape Ma=%d Nb=%d", Ma, Nb);
  DMF_REQUIRE(D == 256 || D == 512, "dmf_infonce_bwd_stored: D=%d must be 256 or 512", D);
  DMF_REQUIRE(dir == 0 || dir == 1, "dmf_infonce_bwd_stored: dir must be 0 (rows) or 1 (columns)");
  DMF_REQUIRE((reinterpret_cast<uintptr_t>(dOut) & 15) == 0 && (ldo & 3) == 0,
              "dmf_infonce_bwd_stored: dOut needs a 16-byte aligned base and a row pitch that is a multiple of 4");
  DMF_REQUIRE((reinterpret_cast<uintptr_t>(work) & 15) == 0, "dmf_infonce_bwd_stored: work must be 16-byte aligned");
  const long long nib = be_nib(Ma), njb = be_njb(Nb);
  DMF_REQUIRE(nib * njb * 128 < 2147483647LL, "dmf_infonce_bwd_stored: E has too many block rows for one tensor map");
  const int n_own = dir == 0 ? Ma : Nb, n_str = dir == 0 ? Nb : Ma;
  // factor arrays (zero padded): fa over the anchors, fb over the columns
  const int pa = ((Ma + 255) / 256) * 256, pb = ((Nb + 255) / 256) * 256;
  float* fa = work;
  float* fb = work + pa;
  infonce_factors_kernel<<<(pa + 255) / 256, 256, 0, s>>>(lseA, Ma, pa, shift, fa);
  int rc = launched("dmf_infonce_bwd_stored factors");
  if (rc) return rc;
  infonce_factors_kernel<<<(pb + 255) / 256, 256, 0, s>>>(lseB, Nb, pb, shift, fb);
  rc = launched("dmf_infonce_bwd_stored factors");
  if (rc) return rc;

  CUtensorMap tmE, tmZ;
  rc = make_tmap_bf16_2d(&tmE, E, nib * njb * 128, 64, 64, 64);          // [64 x 64] boxes of the block rows
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmZ, Z, n_str, D, ldz, 64);                    // [64 k x 64 d] boxes, read MN-major
  if (rc) return rc;

  BEArgs P;
  P.n_own = n_own; P.n_str = n_str;
  P.nh = D / 256;
  P.njb = (int)njb;
  const int stage_bytes = BE_EBYTES + P.nh * BE_ZHALF;
  P.stages = (BE_SMEM - 1024 - 2048) / (stage_bytes + BE_FBYTES);
  if (P.stages > BE_MAX_STAGES) P.stages = BE_MAX_STAGES;
  P.ksteps_total = (n_str + 63) / 64;
  P.f_own = dir == 0 ? fa : fb;
  P.f_str = dir == 0 ? fb : fa;
  P.coef = coef; P.gscale = gscale;
  P.pos_off = dir == 0 ? diag_offset : -diag_offset;
  P.Z = (const uint16_t*)Z; P.ldz = ldz;
  P.dOut = dOut; P.ldo = ldo;
  P.accumulate = accumulate;

  static int slots = 0;
  if (!slots) {
    cudaError_t e = cudaFuncSetAttribute(infonce_bwd_e_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BE_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(infonce_bwd_e_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BE_SMEM);
    if (e != cudaSuccess) return fail((int)e, "dmf_infonce_bwd_stored: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(128, 1, 1);
    cfg.blockDim = dim3(BE_THREADS, 1, 1);
    cfg.dynamicSmemBytes = BE_SMEM;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = 0;
    e = cudaOccupancyMaxActiveClusters(&nc, infonce_bwd_e_kernel<0>, &cfg);
    if (e != cudaSuccess || nc <= 0) { cudaGetLastError(); nc = 74; }
    slots = nc;
  }
  const int pairs = (n_own + 255) / 256;
  // Contraction split (gridDim.z) so that pairs * nsplit fills whole waves of the resident clusters; partial rows
  // are accumulated with red.add into a zeroed dOut.
  int nsplit = 1;
  {
    double best = (double)pairs / (double)(((pairs + slots - 1) / slots) * slots);
    for (int ns = 2; ns <= 16 && best < 0.97; ++ns) {
      if (P.ksteps_total / ns < 32) break;
      const int items = pairs * ns;
      const double eff = (double)items / (double)(((items + slots - 1) / slots) * slots);
      if (eff > best + 0.03) { best = eff; nsplit = ns; }
    }
    // dir 1 streams E in steps of 64 rows of a 128-row block: keep every split on whole blocks
    int kps = (P.ksteps_total + nsplit - 1) / nsplit;
    kps = (kps + 1) & ~1;
    nsplit = (P.ksteps_total + kps - 1) / kps;
    P.ksteps_per_split = kps;
  }
  if (nsplit > 1 && !accumulate) {
    cudaError_t e = cudaMemset2DAsync(dOut, (size_t)ldo * sizeof(float), 0, (size_t)D * sizeof(float), (size_t)n_own, s);
    if (e != cudaSuccess) return fail((int)e, "dmf_infonce_bwd_stored: memset: %s", cudaGetErrorString(e));
  }
  dim3 grid(2 * pairs, 1, nsplit);
  if (dir == 0) infonce_bwd_e_kernel<0><<<grid, BE_THREADS, BE_SMEM, s>>>(tmE, tmZ, P);
  else infonce_bwd_e_kernel<1><<<grid, BE_THREADS, BE_SMEM, s>>>(tmE, tmZ, P);
  return launched("dmf_infonce_bwd_stored");
}
