// K2 (fp32 parity path): tiled InfoNCE pieces on FFMA units -- never materialises the logits.
//
// Replaces the matmul/div/max/sub/exp/sum/log chain of SupConLoss.forward
// (models/losses.py:64-99): `rowlse` streams 64x64 similarity tiles s = scale * A B^T and keeps
// an online (max, sum-exp) per anchor row; `infonce_bwd` recomputes the tiles and contracts
// W_ij = exp(s_ij - lseA_i) + exp(s_ij - lseB_j) with the column block.
#include "common.cuh"

namespace dmf {

constexpr int TI = 64, TJ = 64, TK = 16, NCE_THREADS = 256, NPAD = 4;

// 64x64 tile of A B^T accumulated over D in chunks of TK; result in acc (4x4 per thread)
__device__ __forceinline__ void s_tile_f32(const float* __restrict__ A, long long lda, int Ma, int i0,
                                           const float* __restrict__ Bm, long long ldb, int Nb, int j0, int D,
                                           float (*As)[TI + NPAD], float (*Bs)[TJ + NPAD], float acc[4][4]) {
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < D; k0 += TK) {
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int k = tid & 15, r = (tid >> 4) + p * 16;
      const int gk = k0 + k;
      As[k][r] = (i0 + r < Ma && gk < D) ? __ldg(A + (long long)(i0 + r) * lda + gk) : 0.f;
      Bs[k][r] = (j0 + r < Nb && gk < D) ? __ldg(Bm + (long long)(j0 + r) * ldb + gk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
}

__device__ __forceinline__ float half_warp_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(NCE_THREADS)
rowlse_f32_kernel(const float* __restrict__ A, long long lda, int Ma, const float* __restrict__ Bm, long long ldb,
                  int Nb, int D, float scale, float* __restrict__ row_max, float* __restrict__ row_sum,
                  long long diag_offset, float* __restrict__ diag_out) {
  __shared__ __align__(16) float As[TK][TI + NPAD];
  __shared__ __align__(16) float Bs[TK][TJ + NPAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.x * TI;
  float m[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; l[i] = 0.f; }
  float acc[4][4];
  for (int j0 = 0; j0 < Nb; j0 += TJ) {
    s_tile_f32(A, lda, Ma, i0, Bm, ldb, Nb, j0, D, As, Bs, acc);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gi = i0 + ty * 4 + i;
      float tmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gj = j0 + tx * 4 + j;
        acc[i][j] *= scale;
        if (gj < Nb) tmax = fmaxf(tmax, acc[i][j]);
        if (diag_out && diag_offset >= 0 && gi < Ma && (long long)gj == diag_offset + gi && gj < Nb) diag_out[gi] = acc[i][j];
      }
      tmax = half_warp_max(tmax);
      const float mn = fmaxf(m[i], tmax);
      float ps = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gj = j0 + tx * 4 + j;
        if (gj < Nb) ps += expf(acc[i][j] - mn);
      }
      ps = half_warp_sum(ps);
      l[i] = l[i] * expf(m[i] - mn) + ps;   // exp(-inf - finite) = 0 on the first tile
      m[i] = mn;
    }
  }
  if (tx == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gi = i0 + ty * 4 + i;
      if (gi < Ma) { row_max[gi] = m[i]; row_sum[gi] = l[i]; }
    }
  }
}

// dA[i, d0:d0+64] for a 64-row block: loops column tiles, recomputes s, contracts W with Bm.
__global__ void __launch_bounds__(NCE_THREADS)
infonce_bwd_f32_kernel(const float* __restrict__ A, long long lda, int Ma, const float* __restrict__ lseA,
                       const float* __restrict__ Bm, long long ldb, int Nb, const float* __restrict__ lseB, int D,
                       float scale, float coef, const float* __restrict__ gscale, long long diag_offset,
                       float* __restrict__ dA, long long ldda, int accumulate) {
  __shared__ __align__(16) float As[TK][TI + NPAD];
  __shared__ __align__(16) float Bs[TK][TJ + NPAD];
  __shared__ __align__(16) float Ws[TJ][TI + NPAD];   // W^T tile: [j][i]
  __shared__ __align__(16) float Vs[TJ][64 + NPAD];   // Bm[j0+j, d0+d]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.x * TI;
  const int d0 = blockIdx.y * 64;
  float la[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = i0 + ty * 4 + i;
    la[i] = gi < Ma ? lseA[gi] : 0.f;
  }
  float out[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[i][j] = 0.f;
  float acc[4][4];
  for (int j0 = 0; j0 < Nb; j0 += TJ) {
    s_tile_f32(A, lda, Ma, i0, Bm, ldb, Nb, j0, D, As, Bs, acc);
    // W tile -> smem (transposed), column block of Bm -> smem
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gj = j0 + tx * 4 + j;
      const float lb = gj < Nb ? __ldg(lseB + gj) : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gi = i0 + ty * 4 + i;
        float w = 0.f;
        if (gj < Nb && gi < Ma) {
          const float sv = acc[i][j] * scale;
          w = expf(sv - la[i]) + expf(sv - lb);
        }
        Ws[tx * 4 + j][ty * 4 + i] = w;
      }
    }
#pragma unroll
    for (int p = 0; p < 16; ++p) {
      const int d = tid & 63, j = (tid >> 6) + p * 4;
      Vs[j][d] = (j0 + j < Nb && d0 + d < D) ? __ldg(Bm + (long long)(j0 + j) * ldb + d0 + d) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < TJ; ++j) {
      const float4 w4 = *reinterpret_cast<const float4*>(&Ws[j][ty * 4]);
      const float4 v4 = *reinterpret_cast<const float4*>(&Vs[j][tx * 4]);
      const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
      const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int d = 0; d < 4; ++d) out[i][d] = fmaf(wv[i], vv[d], out[i][d]);
    }
    // next s_tile_f32 begins with __syncthreads(), protecting Ws/Vs reuse
  }
  const float cg = coef * (gscale ? __ldg(gscale) : 1.0f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = i0 + ty * 4 + i;
    if (gi >= Ma) continue;
    const long long pj = diag_offset + gi;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int gd = d0 + tx * 4 + d;
      if (gd >= D) continue;
      float v = out[i][d];
      if (diag_offset >= 0 && pj < Nb) v -= 2.0f * __ldg(Bm + pj * ldb + gd);
      v *= cg;
      float* dst = dA + (long long)gi * ldda + gd;
      *dst = accumulate ? *dst + v : v;
    }
  }
}

// per-anchor finalisation of one SupConLoss call (models/losses.py:68-99)
__global__ void infonce_finalize_kernel(const float* __restrict__ m_cross, const float* __restrict__ l_cross,
                                        const float* __restrict__ m_intra, const float* __restrict__ l_intra,
                                        const float* __restrict__ pos, const float* __restrict__ self, int n,
                                        float inv_count, float inv_count_diag, int which,
                                        float* __restrict__ lse_eff, float* __restrict__ out3) {
  __shared__ float red[32];
  float a_loss = 0.f, a_diag = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float mc = m_cross[i], mi = m_intra[i];
    const float mf = fmaxf(mc, mi);                         // full 2B-row max, :68
    const float sc = l_cross[i] * expf(mc - mf) + 1e-12f;   // exp_logits.sum + 1e-12, :79-80
    const float lg = logf(sc);
    a_loss += -(pos[i] - mf - lg);
    lse_eff[i] = mf + lg;
    // diagnostics (:89-99): same reference max, intra-view columns, no epsilon
    const float si = l_intra[i] * expf(mi - mf);
    a_diag += -((self[i] - mf) - logf(si));
  }
  const float s1 = block_sum(a_loss, red);
  const float s2 = block_sum(a_diag, red);
  if (threadIdx.x == 0) {
    atomicAdd(out3 + 0, s1 * inv_count);
    atomicAdd(out3 + 1 + which, s2 * inv_count_diag);
  }
}

}  // namespace dmf

using namespace dmf;

int dmf_rowlse_bf16_tc(const void* A, long long lda, int Ma, const void* Bm, long long ldb, int Nb, int D, float scale,
                       float* row_max, float* row_sum, long long diag_offset, float* diag_out, void* workspace,
                       size_t workspace_bytes, cudaStream_t s);
int dmf_infonce_bwd_bf16_tc(const void* A, long long lda, int Ma, const float* lseA, const void* Bm, long long ldb,
                            const void* BmT, long long ldbt, int Nb, const float* lseB, int D, float scale, float coef,
                            const float* gscale, long long diag_offset, float* dA, long long ldda, int accumulate,
                            cudaStream_t s);

extern "C" int dmf_rowlse(const void* A, long long lda, int Ma, const void* Bm, long long ldb, int Nb, int D,
                          float scale, float* row_max, float* row_sum, long long diag_offset, float* diag_out,
                          void* workspace, size_t workspace_bytes, int dtype, dmf_stream_t s) {
  DMF_REQUIRE(A && Bm && row_max && row_sum, "dmf_rowlse: null argument");
  DMF_REQUIRE(Ma >= 0 && Nb >= 1 && D >= 1, "dmf_rowlse: bad shape Ma=%d Nb=%d D=%d", Ma, Nb, D);
  if (Ma == 0) return 0;
  if (dtype == 1)
    return dmf_rowlse_bf16_tc(A, lda, Ma, Bm, ldb, Nb, D, scale, row_max, row_sum, diag_offset, diag_out, workspace,
                              workspace_bytes, (cudaStream_t)s);
  DMF_REQUIRE(dtype == 0, "dmf_rowlse: unknown dtype %d", dtype);
  rowlse_f32_kernel<<<(Ma + TI - 1) / TI, NCE_THREADS, 0, (cudaStream_t)s>>>(
      (const float*)A, lda, Ma, (const float*)Bm, ldb, Nb, D, scale, row_max, row_sum, diag_offset, diag_out);
  return launched("dmf_rowlse(f32)");
}

extern "C" int dmf_infonce_finalize(const float* m_cross, const float* l_cross, const float* m_intra,
                                    const float* l_intra, const float* pos, const float* self, int n, float inv_count,
                                    float inv_count_diag, int which, float* lse_eff, float* out3, dmf_stream_t s) {
  DMF_REQUIRE(m_cross && l_cross && m_intra && l_intra && pos && self && lse_eff && out3, "dmf_infonce_finalize: null argument");
  DMF_REQUIRE(which == 0 || which == 1, "dmf_infonce_finalize: which must be 0 or 1");
  if (n <= 0) return 0;
  const int blocks = min((n + 255) / 256, kNumSMs * 4);
  infonce_finalize_kernel<<<blocks, 256, 0, (cudaStream_t)s>>>(m_cross, l_cross, m_intra, l_intra, pos, self, n,
                                                               inv_count, inv_count_diag, which, lse_eff, out3);
  return launched("dmf_infonce_finalize");
}

extern "C" int dmf_infonce_bwd(const void* A, long long lda, int Ma, const float* lseA, const void* Bm, long long ldb,
                               const void* BmT, long long ldbt, int Nb, const float* lseB, int D, float scale, float coef,
                               const float* gscale, long long diag_offset, float* dA, long long ldda, int accumulate,
                               int dtype, dmf_stream_t s) {
  DMF_REQUIRE(A && Bm && lseA && lseB && dA, "dmf_infonce_bwd: null argument");
  DMF_REQUIRE(Ma >= 0 && Nb >= 1 && D >= 1, "dmf_infonce_bwd: bad shape");
  if (Ma == 0) return 0;
  if (dtype == 1)
    return dmf_infonce_bwd_bf16_tc(A, lda, Ma, lseA, Bm, ldb, BmT, ldbt, Nb, lseB, D, scale, coef, gscale, diag_offset, dA,
                                   ldda, accumulate, (cudaStream_t)s);
  DMF_REQUIRE(dtype == 0, "dmf_infonce_bwd: unknown dtype %d", dtype);
  dim3 grid((Ma + TI - 1) / TI, (D + 63) / 64);
  infonce_bwd_f32_kernel<<<grid, NCE_THREADS, 0, (cudaStream_t)s>>>((const float*)A, lda, Ma, lseA, (const float*)Bm, ldb,
                                                                     Nb, lseB, D, scale, coef, gscale, diag_offset, dA,
                                                                     ldda, accumulate);
  return launched("dmf_infonce_bwd(f32)");
}
