// K1 (bf16 tensor-core path): grouped  C = epi(A * B^T)  with tcgen05.mma + TMEM + TMA.
//
// Replaces the cuBLAS addmm/mm calls behind Linear.forward (models/classifiers.py:43-48) for all
// groups (modalities x {orig, augmented} streams) of one MLP layer in ONE launch, with bias,
// ReLU, ReLU-mask (dgrad) and bf16 packing fused into the epilogue.  dgrad and wgrad reuse the same
// kernel on transposed bf16 copies (see host code in ops.py), so every contraction is "TN":
// both operands K-major in global memory.
//
// CTA = one 128x128 output tile; 6 warps:
//   warp 0      TMA producer     cp.async.bulk.tensor 2D, 128B swizzle, 64-wide K blocks, 4-stage ring
//   warp 1      MMA issuer       one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (128x128x16),
//                                accumulator = 128 TMEM columns; also owns TMEM alloc/dealloc
//   warps 2-5   epilogue         tcgen05.ld 32x32b (lane quarter = warp%4), bias/activation, stores
// Ragged M/N/K are handled by TMA zero fill + masked stores.  ~131 KB smem/CTA -> 1 CTA/SM.
#include "common.cuh"
#include "tc_common.cuh"
#include "gemm_tc_epi.cuh"

namespace dmf {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64, TC_STAGES = 4, TC_THREADS = 192;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 2;  // 16 KB per operand tile
constexpr int kMaxTcGroups = 8;
constexpr size_t TC_SMEM_BYTES = 1024 /*align slack*/ + (size_t)TC_STAGES * 2 * TC_TILE_BYTES + 4 * kEpiStageFloats * 4 + 256;

struct TcGroup {
  TcEpi epi;
  int K;
};
int launch_gemm_tc2(const dmf_tc_gemm_desc* groups, int n_groups, int epilogue, cudaStream_t st);  // gemm_tc2.cu
struct alignas(64) TcGemmParams {
  CUtensorMap tmA[kMaxTcGroups];
  CUtensorMap tmB[kMaxTcGroups];
  TcGroup g[kMaxTcGroups];
  int tile_start[kMaxTcGroups + 1];
  int n;
};

template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ TcGemmParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + TC_STAGES * TC_TILE_BYTES;
  float* epi_stage = reinterpret_cast<float*>(smem + 2 * TC_STAGES * TC_TILE_BYTES);   // [4 warps][32 x 36]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + 4 * kEpiStageFloats);
  uint64_t* empty_bar = full_bar + TC_STAGES;
  uint64_t* tmem_full = empty_bar + TC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int gi = 0;
  const int tile = blockIdx.x;
  while (gi + 1 < P.n && tile >= P.tile_start[gi + 1]) ++gi;
  const TcGroup& g = P.g[gi];
  const int lt = tile - P.tile_start[gi];
  const int tiles_n = (g.epi.N + TC_BN - 1) / TC_BN;
  const int tm = lt / tiles_n, tn = lt - tm * tiles_n;
  const int m0 = tm * TC_BM, n0 = tn * TC_BN;
  const int num_kb = (g.K + TC_BK - 1) / TC_BK;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&P.tmA[gi]);
    tc::tma_prefetch_desc(&P.tmB[gi]);
    for (int s = 0; s < TC_STAGES; ++s) {
      tc::mbar_init(full_bar + s, 1);
      tc::mbar_init(empty_bar + s, 1);
    }
    tc::mbar_init(tmem_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<TC_BN>(tmem_slot);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    {
      // whole warp, uniform control flow; one elected lane issues the TMA instructions
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        tc::mbar_wait(empty_bar + stage, phase ^ 1);
        if (tc::elect_one()) {
          tc::mbar_expect_tx(full_bar + stage, 2 * TC_TILE_BYTES);
          tc::tma_load_2d(smemA + stage * TC_TILE_BYTES, &P.tmA[gi], kb * TC_BK, m0, full_bar + stage);
          tc::tma_load_2d(smemB + stage * TC_TILE_BYTES, &P.tmB[gi], kb * TC_BK, n0, full_bar + stage);
        }
        __syncwarp();
        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    {
      // whole warp, uniform control flow; one elected lane issues the tcgen05 instructions
      constexpr uint32_t idesc = tc::make_idesc_bf16(TC_BM, TC_BN, 0, 0);
      const uint64_t adesc0 = tc::make_smem_desc(tc::smem_u32(smemA), 16, 1024);
      const uint64_t bdesc0 = tc::make_smem_desc(tc::smem_u32(smemB), 16, 1024);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        tc::mbar_wait(full_bar + stage, phase);
        tc::tc_fence_after_sync();
        const uint64_t off = (uint64_t)((stage * TC_TILE_BYTES) >> 4);
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            tc::umma_ss(tmem_base, adesc0 + off + 2 * k, bdesc0 + off + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          tc::umma_commit(empty_bar + stage);   // frees the smem stage when these MMAs retire
        }
        __syncwarp();
        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
      }
      if (tc::elect_one()) tc::umma_commit(tmem_full);             // accumulator complete
      __syncwarp();
    }
  } else {
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const int row0 = m0 + q * 32;
    float* my_stage = epi_stage + (warp - 2) * kEpiStageFloats;
    tc::mbar_wait(tmem_full, 0);
    tc::tc_fence_after_sync();
#pragma unroll 1
    for (int c = 0; c < TC_BN / 32; ++c) {
      uint32_t r[32];
      tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
      tc::tmem_ld_wait();
      tc_epilogue_chunk<EPI, false>(g.epi, r, row0, lane, n0 + c * 32, true, my_stage);
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<TC_BN>(tmem_base);
}

// ---------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return fail(-3, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld * 2) & 15) != 0)
    return fail(-1, "TMA operand needs a 16B-aligned base and a leading dimension that is a multiple of 8 bf16 (ld=%lld)", ld);
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled failed (CUresult %d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld);
  return 0;
}

// [rows, 64] bf16 block-row view of the stored InfoNCE probabilities, box = 32 rows x 32 columns (64-byte rows in
// shared memory, 64B swizzle): the store side of infonce_fwd_tc4.cu
int make_tmap_bf16_2d_box32_sw64(CUtensorMap* out, const void* base, long long rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return fail(-3, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(-1, "TMA store target needs a 16B-aligned base");
  cuuint64_t gdim[2] = {64u, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {128u};
  cuuint32_t box[2] = {32u, 32u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled (store map) failed (CUresult %d) rows=%lld", (int)r, rows);
  return 0;
}

}  // namespace dmf

using namespace dmf;

template <int EPI>
static int launch_tc(const TcGemmParams& P, int tiles, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tc_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)TC_SMEM_BYTES);
    if (e != cudaSuccess) return fail((int)e, "gemm_bf16_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  gemm_bf16_tc_kernel<EPI><<<tiles, TC_THREADS, TC_SMEM_BYTES, st>>>(P);
  return launched("dmf_grouped_gemm_bf16_tc");
}

extern "C" int dmf_grouped_gemm_bf16_tc(const dmf_tc_gemm_desc* groups, int n_groups, int epilogue, dmf_stream_t s) {
  DMF_REQUIRE(groups && n_groups >= 1, "dmf_grouped_gemm_bf16_tc: no groups");
  // validate, then route: CTA-pair persistent kernel when every group can feed 256-row x 256-col tiles
  bool pair_ok = true, any_mn = false;
  for (int i = 0; i < n_groups; ++i) {
    const dmf_tc_gemm_desc& d = groups[i];
    DMF_REQUIRE(d.M >= 0 && d.N >= 0 && d.K >= 1, "dmf_grouped_gemm_bf16_tc: bad dims in group %d", i);
    if (d.M == 0 || d.N == 0) continue;
    DMF_REQUIRE(d.A && d.B && (d.out_f32 || d.out_bf16 || d.out_bf16_t), "dmf_grouped_gemm_bf16_tc: null pointer in group %d", i);
    DMF_REQUIRE(epilogue != DMF_EPI_RELU_MASK || d.mask_bf16, "dmf_grouped_gemm_bf16_tc: RELU_MASK needs mask (group %d)", i);
    // skinny wgrad shapes (tiny output, K = batch: the probe heads' [128, 512] weight gradients at K = 65536 ran for
    // 0.3 ms on 16 single-CTA tiles) also go to the pair kernel: it splits K over the 74 clusters, and the unused
    // second 128 rows of a tile are TMA zero fill
    const bool skinny = epilogue == DMF_EPI_NONE && d.out_f32 && !d.out_bf16 && !d.out_bf16_t && d.split_k != 1 &&
                        d.K >= 4096 && d.M >= 128 && d.N >= 128;
    if ((d.M < 512 || d.N < 128) && !skinny) pair_ok = false;
    any_mn = any_mn || d.mn_major != 0;
    DMF_REQUIRE(!d.mn_major || ((d.lda & 7) == 0 && (d.ldb & 7) == 0), "dmf_grouped_gemm_bf16_tc: mn_major needs row pitches that are multiples of 8 (group %d)", i);
  }
  DMF_REQUIRE(!any_mn || pair_ok, "dmf_grouped_gemm_bf16_tc: mn_major operands need M >= 512 and N >= 128 in every group");
  if (pair_ok) return launch_gemm_tc2(groups, n_groups, epilogue, (cudaStream_t)s);
  for (int base = 0; base < n_groups; base += kMaxTcGroups) {
    TcGemmParams P;
    P.n = 0;
    int tiles = 0;
    const int cnt = n_groups - base < kMaxTcGroups ? n_groups - base : kMaxTcGroups;
    for (int i = 0; i < cnt; ++i) {
      const dmf_tc_gemm_desc& d = groups[base + i];
      DMF_REQUIRE(d.M >= 0 && d.N >= 0 && d.K >= 1, "dmf_grouped_gemm_bf16_tc: bad dims in group %d", base + i);
      if (d.M == 0 || d.N == 0) continue;
      DMF_REQUIRE(d.A && d.B && (d.out_f32 || d.out_bf16), "dmf_grouped_gemm_bf16_tc: null pointer in group %d", base + i);
      DMF_REQUIRE(epilogue != DMF_EPI_RELU_MASK || d.mask_bf16, "dmf_grouped_gemm_bf16_tc: RELU_MASK needs mask (group %d)", base + i);
      int rc = make_tmap_bf16_2d(&P.tmA[P.n], d.A, d.M, d.K, d.lda, TC_BM);
      if (rc) return rc;
      rc = make_tmap_bf16_2d(&P.tmB[P.n], d.B, d.N, d.K, d.ldb, TC_BN);
      if (rc) return rc;
      TcGroup& g = P.g[P.n];
      g.epi.out_f32 = d.out_f32; g.epi.ldo_f32 = d.ldo_f32;
      g.epi.out_bf16 = d.out_bf16; g.epi.ldo_bf16 = d.ldo_bf16;
      g.epi.out_t = d.out_bf16_t; g.epi.ldo_t = d.ldo_t;
      g.epi.bias = d.bias; g.epi.mask = d.mask_bf16; g.epi.ldmask = d.ldmask;
      g.epi.M = d.M; g.epi.N = d.N;
      g.epi.fast = tc_epi_fast_ok(g.epi); g.K = d.K;
      P.tile_start[P.n] = tiles;
      tiles += ((d.M + TC_BM - 1) / TC_BM) * ((d.N + TC_BN - 1) / TC_BN);
      ++P.n;
    }
    if (P.n == 0) continue;
    P.tile_start[P.n] = tiles;
    int rc;
    switch (epilogue) {
      case DMF_EPI_NONE: rc = launch_tc<DMF_EPI_NONE>(P, tiles, (cudaStream_t)s); break;
      case DMF_EPI_BIAS: rc = launch_tc<DMF_EPI_BIAS>(P, tiles, (cudaStream_t)s); break;
      case DMF_EPI_BIAS_RELU: rc = launch_tc<DMF_EPI_BIAS_RELU>(P, tiles, (cudaStream_t)s); break;
      case DMF_EPI_RELU_MASK: rc = launch_tc<DMF_EPI_RELU_MASK>(P, tiles, (cudaStream_t)s); break;
      default: return fail(-1, "dmf_grouped_gemm_bf16_tc: unsupported epilogue %d", epilogue);
    }
    if (rc) return rc;
  }
  return 0;
}
