// K1 (fp32 parity path): grouped strided GEMM on FFMA units.
//
// One launch covers every group (view x stream) of one MLP layer: replaces the per-view
// addmm + relu kernels behind Linear.forward (models/classifiers.py:43-48), called N + N^2 times
// per DMVAE step (models/dmvae.py:139,154,162).  Strides are general so the same kernel is
// forward (Y = X W^T + b), dgrad (dX = dY W, ReLU mask fused) and wgrad (dW = dY^T X, bias
// gradient = row sums of the A operand, fused).  fp32 accumulate in a fixed k order => results
// are deterministic and within ~1e-6 of torch's fp32 ('highest') matmul.
//
// Tiling: 64x64 output tile per CTA, BK = 16, 256 threads, 4x4 register micro-tile; tiles of all
// groups are flattened into blockIdx.x (group found by a short prefix scan of tile counts).
#include "common.cuh"
#include "special_math.cuh"

namespace dmf {

constexpr int BM = 64, BN = 64, BK = 16, GEMM_THREADS = 256, PAD = 4;
constexpr int kMaxGroups = 64;

struct GemmGroups {
  dmf_gemm_desc g[kMaxGroups];
  int tile_start[kMaxGroups + 1];
  int n;
};

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS)
grouped_gemm_f32_kernel(const __grid_constant__ GemmGroups G) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  __shared__ float rs[BM];

  int gi = 0;
  const int tile = blockIdx.x;
  while (gi + 1 < G.n && tile >= G.tile_start[gi + 1]) ++gi;
  const dmf_gemm_desc& d = G.g[gi];
  const int lt = tile - G.tile_start[gi];
  const int tiles_n = (d.N + BN - 1) / BN;
  const int tm = lt / tiles_n, tn = lt - tm * tiles_n;
  const int m0 = tm * BM, n0 = tn * BN;
  const int M = d.M, N = d.N, K = d.K;
  const float* __restrict__ A = static_cast<const float*>(d.A);
  const float* __restrict__ Bp = static_cast<const float*>(d.B);
  const long long a_rs = d.a_rs, a_cs = d.a_cs, b_rs = d.b_rs, b_cs = d.b_cs;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // micro-tile: rows ty*4.., cols tx*4..
  // loader mappings: pick the one whose fastest thread index walks the contiguous stride
  const bool a_kfast = (a_cs == 1);
  const bool b_nfast = (b_cs == 1);
  const bool do_rowsum = (d.rowsum_a != nullptr) && (tn == 0);
  float rowsum = 0.f;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      int m, k;
      if (a_kfast) { k = tid & 15; m = (tid >> 4) + p * 16; }
      else { m = tid & 63; k = (tid >> 6) + p * 4; }
      const int gm = m0 + m, gk = k0 + k;
      ra[p] = (gm < M && gk < K) ? __ldg(A + gm * a_rs + gk * a_cs) : 0.f;
      int n, kb;
      if (b_nfast) { n = tid & 63; kb = (tid >> 6) + p * 4; }
      else { kb = tid & 15; n = (tid >> 4) + p * 16; }
      const int gn = n0 + n, gkb = k0 + kb;
      rb[p] = (gn < N && gkb < K) ? __ldg(Bp + gkb * b_rs + gn * b_cs) : 0.f;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      int m, k;
      if (a_kfast) { k = tid & 15; m = (tid >> 4) + p * 16; }
      else { m = tid & 63; k = (tid >> 6) + p * 4; }
      As[buf][k][m] = ra[p];
      int n, kb;
      if (b_nfast) { n = tid & 63; kb = (tid >> 6) + p * 4; }
      else { kb = tid & 15; n = (tid >> 4) + p * 16; }
      Bs[buf][kb][n] = rb[p];
    }
  };

  const int nk = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
    if (do_rowsum && tid < BM) {
#pragma unroll
      for (int k = 0; k < BK; ++k) rowsum += As[buf][k][tid];
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }
  (void)rs;
  if (do_rowsum && tid < BM && m0 + tid < M) d.rowsum_a[m0 + tid] = rowsum;

  // epilogue
  float* __restrict__ C = static_cast<float*>(d.C);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (EPI == DMF_EPI_BIAS || EPI == DMF_EPI_BIAS_RELU || EPI == DMF_EPI_BIAS_EVIDENCE) {
        if (d.bias) v += __ldg(d.bias + gn);
      }
      if (EPI == DMF_EPI_BIAS_RELU) v = fmaxf(v, 0.f);
      if (EPI == DMF_EPI_RELU_MASK) {
        const float a = static_cast<const float*>(d.aux)[(long long)gm * d.ldaux + gn];
        v = a > 0.f ? v : 0.f;
      }
      if (EPI == DMF_EPI_BIAS_EVIDENCE) {
        if (d.aux) static_cast<float*>(d.aux)[(long long)gm * d.ldaux + gn] = v;
        v = evidence_act(v);
      }
      float* dst = C + (long long)gm * d.ldc + gn;
      if (EPI == DMF_EPI_NONE && d.accumulate) v += *dst;
      *dst = v;
    }
  }
}

__global__ void colsum_kernel(const float* __restrict__ X, long long ld, int M, int N, float* __restrict__ out,
                              int rows_per_block) {
  // block = 32 columns x 8 row lanes; grid.y splits the rows; partial sums land with atomicAdd
  __shared__ float part[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int mbeg = blockIdx.y * rows_per_block;
  const int mend = min(M, mbeg + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int m = mbeg + threadIdx.y; m < mend; m += 8) s += X[(long long)m * ld + n];
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
    atomicAdd(out + n, t);
  }
}

}  // namespace dmf

using namespace dmf;

extern "C" int dmf_grouped_gemm_f32(const dmf_gemm_desc* groups, int n_groups, int epilogue, dmf_stream_t s) {
  DMF_REQUIRE(groups && n_groups >= 1 && n_groups <= kMaxGroups, "dmf_grouped_gemm_f32: n_groups=%d out of [1,%d]",
              n_groups, kMaxGroups);
  GemmGroups G;
  G.n = 0;
  int tiles = 0;
  for (int i = 0; i < n_groups; ++i) {
    const dmf_gemm_desc& d = groups[i];
    DMF_REQUIRE(d.M >= 0 && d.N >= 0 && d.K >= 0, "dmf_grouped_gemm_f32: negative dim in group %d", i);
    if (d.M == 0 || d.N == 0) continue;
    DMF_REQUIRE(d.A && d.B && d.C, "dmf_grouped_gemm_f32: null pointer in group %d", i);
    DMF_REQUIRE(epilogue != DMF_EPI_RELU_MASK || d.aux, "dmf_grouped_gemm_f32: RELU_MASK needs aux (group %d)", i);
    G.g[G.n] = d;
    G.tile_start[G.n] = tiles;
    tiles += ((d.M + BM - 1) / BM) * ((d.N + BN - 1) / BN);
    ++G.n;
  }
  if (G.n == 0) return 0;
  G.tile_start[G.n] = tiles;
  cudaStream_t st = (cudaStream_t)s;
  switch (epilogue) {
    case DMF_EPI_NONE: grouped_gemm_f32_kernel<DMF_EPI_NONE><<<tiles, GEMM_THREADS, 0, st>>>(G); break;
    case DMF_EPI_BIAS: grouped_gemm_f32_kernel<DMF_EPI_BIAS><<<tiles, GEMM_THREADS, 0, st>>>(G); break;
    case DMF_EPI_BIAS_RELU: grouped_gemm_f32_kernel<DMF_EPI_BIAS_RELU><<<tiles, GEMM_THREADS, 0, st>>>(G); break;
    case DMF_EPI_RELU_MASK: grouped_gemm_f32_kernel<DMF_EPI_RELU_MASK><<<tiles, GEMM_THREADS, 0, st>>>(G); break;
    case DMF_EPI_BIAS_EVIDENCE: grouped_gemm_f32_kernel<DMF_EPI_BIAS_EVIDENCE><<<tiles, GEMM_THREADS, 0, st>>>(G); break;
    default: return fail(-1, "dmf_grouped_gemm_f32: unknown epilogue %d", epilogue);
  }
  return launched("dmf_grouped_gemm_f32");
}

extern "C" int dmf_colsum_f32(const float* X, long long ld, int M, int N, float* out, int accumulate, dmf_stream_t s) {
  DMF_REQUIRE(X && out && M >= 0 && N >= 0, "dmf_colsum_f32: bad arguments");
  if (N == 0) return 0;
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N, (cudaStream_t)s);
    if (e != cudaSuccess) return fail((int)e, "dmf_colsum_f32: memset: %s", cudaGetErrorString(e));
  }
  if (M == 0) return 0;
  const int strips = (N + 31) / 32;
  int ysplit = (kNumSMs * 4 + strips - 1) / strips;
  const int max_split = (M + 63) / 64;
  if (ysplit > max_split) ysplit = max_split;
  if (ysplit < 1) ysplit = 1;
  const int rows_per_block = (M + ysplit - 1) / ysplit;
  dim3 block(32, 8), grid(strips, (M + rows_per_block - 1) / rows_per_block);
  colsum_kernel<<<grid, block, 0, (cudaStream_t)s>>>(X, ld, M, N, out, rows_per_block);
  return launched("dmf_colsum_f32");
}
