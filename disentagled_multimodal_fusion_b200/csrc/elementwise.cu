// Memory-bound row / elementwise kernels of the hot path (all fp32 math, coalesced access):
//   row normalisation (F.normalize, models/losses.py:105-106, disentangledssl.py:139-140)
//   vMF reparameterised sample given noise + its backward (models/classifiers.py:314-335,433-466)
//   device-side vMF noise draw (Wood's rejection sampler)
//   DMVAE head: chunk / reparameterise / PoE / KL and backward (models/dmvae.py:74-112,142-176)
//   DMVAE reconstruction MSE fused with its gradient (models/dmvae.py:155,164)
//   fused flat-buffer Adam/AdamW (torch.optim semantics; models/dmvae.py:204-210)
//   fp32->bf16 casts / transposes feeding the tensor-core path
#include "common.cuh"
#include <curand_kernel.h>

namespace dmf {

// ------------------------------------------------------------------------------------------
// row normalisation: one warp per row
// ------------------------------------------------------------------------------------------
__global__ void row_normalize_fwd_kernel(const float* __restrict__ X, long long ldx, int rows, int D, float eps,
                                         float* __restrict__ Y, long long ldy, uint16_t* __restrict__ Yb,
                                         long long ldyb, float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* x = X + (long long)row * ldx;
  float ss = 0.f;
  for (int k = lane; k < D; k += 32) { const float v = x[k]; ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
  if (lane == 0 && inv_norm) inv_norm[row] = inv;
  for (int k = lane; k < D; k += 32) {
    const float y = x[k] * inv;
    if (Y) Y[(long long)row * ldy + k] = y;
    if (Yb) Yb[(long long)row * ldyb + k] = f2bf(y);
  }
}

__global__ void row_normalize_bwd_kernel(const float* __restrict__ Y, long long ldy, const float* __restrict__ inv_norm,
                                         const float* __restrict__ dY, long long lddy, int rows, int D,
                                         float* __restrict__ dX, long long lddx, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* y = Y + (long long)row * ldy;
  const float* dy = dY + (long long)row * lddy;
  float dot = 0.f;
  for (int k = lane; k < D; k += 32) dot = fmaf(y[k], dy[k], dot);
  dot = warp_sum(dot);
  const float inv = inv_norm[row];
  float* dx = dX + (long long)row * lddx;
  for (int k = lane; k < D; k += 32) {
    const float v = (dy[k] - y[k] * dot) * inv;
    dx[k] = accumulate ? dx[k] + v : v;
  }
}

__global__ void sumsq_kernel(const float* __restrict__ G, long long n, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = G[i];
    s = fmaf(v, v, s);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

// ------------------------------------------------------------------------------------------
// vMF rsample given noise: one warp per row
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float vmf_x(const float* __restrict__ v, float w, float wt, int k) {
  return k == 0 ? w : wt * v[k - 1];
}

__global__ void vmf_fwd_kernel(const float* __restrict__ E, long long lde, const float* __restrict__ nw,
                               const float* __restrict__ nv, int rows, int D, float* __restrict__ Z, long long ldz,
                               uint16_t* __restrict__ Zb, long long ldzb) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* e = E + (long long)row * lde;
  const float* v = nv + (long long)row * (D - 1);
  const float w = nw[row];
  const float wt = sqrtf(fmaxf(1.0f - w * w, 1e-10f));
  float ss = 0.f;
  for (int k = lane; k < D; k += 32) ss = fmaf(e[k], e[k], ss);
  ss = warp_sum(ss);
  const float inv_n = 1.0f / sqrtf(ss);
  float nu2 = 0.f, xu = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float up = (k == 0 ? 1.0f : 0.0f) - e[k] * inv_n;
    nu2 = fmaf(up, up, nu2);
    xu = fmaf(vmf_x(v, w, wt, k), up, xu);
  }
  nu2 = warp_sum(nu2);
  xu = warp_sum(xu);
  const float inv_d = 1.0f / (sqrtf(nu2) + 1e-5f);
  const float c2 = 2.0f * xu * inv_d * inv_d;   // 2 <x,u> / (|u'|+eps)
  for (int k = lane; k < D; k += 32) {
    const float up = (k == 0 ? 1.0f : 0.0f) - e[k] * inv_n;
    const float z = vmf_x(v, w, wt, k) - c2 * up;
    if (Z) Z[(long long)row * ldz + k] = z;
    if (Zb) Zb[(long long)row * ldzb + k] = f2bf(z);
  }
}

__global__ void vmf_bwd_kernel(const float* __restrict__ E, long long lde, const float* __restrict__ nw,
                               const float* __restrict__ nv, const float* __restrict__ dZ, long long lddz, int rows,
                               int D, float* __restrict__ dE, long long ldde, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* e = E + (long long)row * lde;
  const float* v = nv + (long long)row * (D - 1);
  const float* g = dZ + (long long)row * lddz;
  const float w = nw[row];
  const float wt = sqrtf(fmaxf(1.0f - w * w, 1e-10f));
  float ss = 0.f;
  for (int k = lane; k < D; k += 32) ss = fmaf(e[k], e[k], ss);
  ss = warp_sum(ss);
  const float inv_n = 1.0f / sqrtf(ss);
  float nu2 = 0.f, xu = 0.f, gu = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float up = (k == 0 ? 1.0f : 0.0f) - e[k] * inv_n;
    nu2 = fmaf(up, up, nu2);
    xu = fmaf(vmf_x(v, w, wt, k), up, xu);
    gu = fmaf(g[k], up, gu);
  }
  nu2 = warp_sum(nu2);
  xu = warp_sum(xu);
  gu = warp_sum(gu);
  const float nu = fmaxf(sqrtf(nu2), 1e-30f);
  const float inv_d = 1.0f / (nu + 1e-5f);
  const float c = xu * inv_d;      // <x,u>
  const float gdu = gu * inv_d;    // <g,u>
  // du_k = -2 (gdu x_k + c g_k);  accumulate <du,u'>, <du,loc>, <u',loc>
  float du_up = 0.f, du_loc = 0.f, up_loc = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float loc = e[k] * inv_n;
    const float up = (k == 0 ? 1.0f : 0.0f) - loc;
    const float du = -2.0f * (gdu * vmf_x(v, w, wt, k) + c * g[k]);
    du_up = fmaf(du, up, du_up);
    du_loc = fmaf(du, loc, du_loc);
    up_loc = fmaf(up, loc, up_loc);
  }
  du_up = warp_sum(du_up);
  du_loc = warp_sum(du_loc);
  up_loc = warp_sum(up_loc);
  const float duu = du_up * inv_d;              // <du,u>
  const float r = duu / nu;                     // coefficient of u' in du'
  // du'_k = (du_k - r u'_k) * inv_d ; dloc = -du' ; <dloc,loc> = -(du_loc - r up_loc) * inv_d
  const float dloc_loc = -(du_loc - r * up_loc) * inv_d;
  float* de = dE + (long long)row * ldde;
  for (int k = lane; k < D; k += 32) {
    const float loc = e[k] * inv_n;
    const float up = (k == 0 ? 1.0f : 0.0f) - loc;
    const float du = -2.0f * (gdu * vmf_x(v, w, wt, k) + c * g[k]);
    const float dloc = -(du - r * up) * inv_d;
    const float val = (dloc - dloc_loc * loc) * inv_n;
    de[k] = accumulate ? de[k] + val : val;
  }
}

// Counter-hash RNG (splitmix64): every thread starts at a hashed position of the 2^64 cycle.  The device draws are
// distribution-equal to the reference's samplers, never stream-equal, so the generator only has to be good and cheap:
// one 64-bit mix (two 64-bit multiplies) per 32 random bits, against ten Philox rounds per four words plus the
// full-precision Box-Muller of curand_normal.
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
struct HashRng {
  unsigned long long s;
  __device__ __forceinline__ HashRng(unsigned long long seed, unsigned long long offset, unsigned long long stream) {
    s = mix64(mix64(seed + 0x9E3779B97F4A7C15ull * (offset + 1)) ^ (stream * 0xD1B54A32D192ED03ull + 0x8CB92BA72F3D8DD7ull));
  }
  __device__ __forceinline__ unsigned long long next64() {
    s += 0x9E3779B97F4A7C15ull;
    return mix64(s);
  }
  __device__ __forceinline__ float uniform() {            // (0, 1), 24 bits
    return ((float)(next64() >> 40) + 0.5f) * (1.0f / 16777216.0f);
  }
  // two independent N(0,1) from one 64-bit word (Box-Muller on fast intrinsics: abs. error ~1e-6, far below the
  // resolution any consumer of this noise has)
  __device__ __forceinline__ float2 normal2() {
    const unsigned long long h = next64();
    const float u1 = ((float)(unsigned)(h >> 40) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = (float)(unsigned)(h & 0xFFFFFFull) * (1.0f / 16777216.0f);
    const float r = sqrtf(-2.0f * __logf(u1));
    float sn, cs;
    __sincosf(6.283185307179586f * u2, &sn, &cs);
    return make_float2(r * cs, r * sn);
  }
};

// Device-side vMF noise: distribution-equal to VonMisesFisher.__sample_w_rej + the tangent normal
// draw (models/classifiers.py:349-431), not stream-equal.  One warp per row.
__device__ float gamma_mt(HashRng* st, float alpha) {
  // Marsaglia-Tsang; alpha < 1 boosted
  float boost = 1.0f;
  if (alpha < 1.0f) {
    boost = powf(st->uniform(), 1.0f / alpha);
    alpha += 1.0f;
  }
  const float d = alpha - 1.0f / 3.0f;
  const float c = rsqrtf(9.0f * d);
  for (int it = 0; it < 64; ++it) {
    const float x = st->normal2().x;
    float vv = 1.0f + c * x;
    if (vv <= 0.f) continue;
    vv = vv * vv * vv;
    const float u = st->uniform();
    if (logf(u) < 0.5f * x * x + d - d * vv + d * logf(vv)) return d * vv * boost;
  }
  return d * boost;
}

__global__ void __launch_bounds__(256) vmf_draw_kernel(float* __restrict__ nw, float* __restrict__ nv, int rows, int D, float kappa,
                                                       unsigned long long seed, unsigned long long offset,
                                                       const unsigned long long* __restrict__ ctr) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  if (ctr) seed += *ctr * 0xA0761D6478BD642Full;      // device-resident draw counter (CUDA-graph replays draw fresh noise)
  HashRng st(seed, offset, (unsigned long long)row * 32 + lane);
  float* v = nv + (long long)row * (D - 1);
  const int n = D - 1;
  if (n <= 1024) {
    // the lane's <= 32 tangent coordinates stay in registers: one pass, one coalesced write
    float z[32];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      if (32 * i < n) {                                   // warp-uniform
        const float2 g = st.normal2();
        z[i] = (lane + 32 * i < n) ? g.x : 0.f;
        z[i + 1] = (lane + 32 * (i + 1) < n) ? g.y : 0.f;
        ss = fmaf(z[i], z[i], fmaf(z[i + 1], z[i + 1], ss));
      } else {
        z[i] = z[i + 1] = 0.f;
      }
    }
    ss = warp_sum(ss);
    const float inv = rsqrtf(fmaxf(ss, 1e-30f));
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int k = lane + 32 * i;
      if (k < n) v[k] = z[i] * inv;
    }
  } else {
    float ss = 0.f;
    for (int k = lane; k < n; k += 32) {
      const float zz = st.normal2().x;
      v[k] = zz;
      ss = fmaf(zz, zz, ss);
    }
    ss = warp_sum(ss);
    const float inv = rsqrtf(fmaxf(ss, 1e-30f));
    for (int k = lane; k < n; k += 32) v[k] *= inv;
  }
  // rejection sampler for w: every lane runs an independent trial per round and the lowest accepting lane wins
  // (the first accepted draw of an iid sequence in a fixed order has exactly the target distribution)
  const float m1 = (float)(D - 1);
  const float c = sqrtf(4.0f * kappa * kappa + m1 * m1);
  const float b_true = (-2.0f * kappa + c) / m1;
  const float b_app = m1 / (4.0f * kappa);
  const float sgate = fminf(fmaxf(kappa - 10.0f, 0.f), 1.0f);
  const float b = b_app * sgate + b_true * (1.0f - sgate);
  const float a = (m1 + 2.0f * kappa + c) / 4.0f;
  const float d = (4.0f * a * b) / (1.0f + b) - m1 * logf(m1);
  float w = 0.f;
  for (int it = 0; it < 64; ++it) {
    const float g1 = gamma_mt(&st, 0.5f * m1), g2 = gamma_mt(&st, 0.5f * m1);
    const float eb = g1 / (g1 + g2);
    const float u = st.uniform();
    const float den = 1.0f - (1.0f - b) * eb;
    w = (1.0f - (1.0f + b) * eb) / den;
    const float t = (2.0f * a * b) / den;
    const bool acc = m1 * logf(t) - t + d > logf(u);
    const unsigned am = __ballot_sync(0xffffffffu, acc);
    if (am) {
      w = __shfl_sync(0xffffffffu, w, __ffs(am) - 1);
      break;
    }
  }
  if (lane == 0) nw[row] = w;
}


// ------------------------------------------------------------------------------------------
// device-side augmentation (utils.py:118-151): per row one of {add N(0, noise_scale^2) noise,
// zero floor(D / drop_scale) distinct random columns, identity}, each with probability 1/3.  One warp per row,
// Philox counter RNG.  Distribution-equal to the reference's numpy/torch host loop, not stream-equal.
// The column subset is the set of the n_drop smallest of D iid uniform keys (a uniform random subset without
// replacement); the key threshold is found by a 32-step bisection on the key value with warp-wide counts.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned aug_key(unsigned long long seed, unsigned long long row, int k) {
  // stateless 32-bit key of (seed, row, column): two rounds of a 64-bit mix (splitmix64 finaliser)
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (row * 0x100000001B3ull + (unsigned long long)k + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return (unsigned)((z ^ (z >> 31)) >> 32);
}

__global__ void __launch_bounds__(256) augment_kernel(const float* __restrict__ X, long long ldx, float* __restrict__ Y, long long ldy, int rows,
                               int D, float noise_scale, int n_drop, unsigned long long seed, unsigned long long offset,
                               int* __restrict__ choice_out, const unsigned long long* __restrict__ ctr) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  if (ctr) seed += *ctr * 0xA0761D6478BD642Full;      // device-resident draw counter (CUDA-graph replays draw fresh noise)
  HashRng st(seed, offset, (unsigned long long)row * 32 + lane);
  // per-row choice: one hashed word per row, mapped to {0,1,2} by a multiply-shift (bias 2^-32)
  const unsigned rw = (unsigned)(mix64(mix64(seed ^ 0xC2B2AE3D27D4EB4Full) + offset * 0x9E3779B97F4A7C15ull +
                                       (unsigned long long)row * 0xD1B54A32D192ED03ull) >> 32);
  const int t = (int)(((unsigned long long)rw * 3ull) >> 32);
  if (choice_out && lane == 0) choice_out[row] = t;
  const float* x = X + (long long)row * ldx;
  float* y = Y + (long long)row * ldy;
  // D <= 1024: the lane's <= 32 elements are loaded up front (32 independent loads in flight: one row per warp
  // with a load -> store dependency per iteration left the kernel latency-bound at 16 warps per SM)
  float xv[32];
  const bool cached = D <= 1024;
  if (cached) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int k = lane + 32 * i;
      xv[i] = k < D ? x[k] : 0.f;
    }
  }
  if (t == 0) {
    if (cached) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        if (32 * i < D) {                                 // warp-uniform
          const float2 g = st.normal2();
          const int k = lane + 32 * i;
          if (k < D) y[k] = fmaf(noise_scale, g.x, xv[i]);
          if (k + 32 < D) y[k + 32] = fmaf(noise_scale, g.y, xv[i + 1]);
        }
      }
    } else {
      for (int k = lane; k < D; k += 32) y[k] = fmaf(noise_scale, st.normal2().x, x[k]);
    }
  } else if (t == 1 && n_drop > 0 && D <= 1024) {
    // D <= 1024: the <= 32 keys of this lane are hashed ONCE into registers; every bisection step is then 32
    // compares instead of 32 hashes (two 64-bit multiplies each): the drop rows cost 1/20 of the generic path below
    const unsigned long long rs = seed ^ (offset * 0xD1B54A32D192ED03ull);
    unsigned key[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int k = lane + 32 * i;
      key[i] = k < D ? aug_key(rs, (unsigned long long)row, k) : 0xFFFFFFFFu;
    }
    // columns k >= D carry the maximum key; they are only counted when mid = 2^32 - 1, which the bisection never
    // needs unless n_drop > D
    // Bisection on the key value, stopped as soon as EXACTLY n_drop keys lie at or below the probe: any threshold
    // between the n_drop-th and the next key will do, and that gap is ~2^32 / D wide -- log2(D) + 2 probes on
    // average instead of 32.  Only a tie at the threshold (probability ~ D^2 / 2^32) runs the bisection to the end.
    unsigned lo = 0u, hi = 0xFFFFFFFFu;
    bool exact = false;
    while (lo < hi) {
      const unsigned mid = lo + ((hi - lo) >> 1);
      int cnt = 0;
#pragma unroll
      for (int i = 0; i < 32; ++i) cnt += key[i] <= mid ? 1 : 0;
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if (cnt == n_drop) { lo = mid; exact = true; break; }
      if (cnt > n_drop) hi = mid; else lo = mid + 1u;
    }
    if (exact) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int k = lane + 32 * i;
        if (k < D) y[k] = key[i] <= lo ? 0.f : xv[i];
      }
    } else {
      int below = 0;
#pragma unroll
      for (int i = 0; i < 32; ++i) below += (key[i] < lo && lane + 32 * i < D) ? 1 : 0;
      below = __reduce_add_sync(0xffffffffu, below);
      int ties_left = n_drop - below;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int k = lane + 32 * i;
        if (32 * i < D) {                                   // warp-uniform
          const bool tie = k < D && key[i] == lo;
          const unsigned tm = __ballot_sync(0xffffffffu, tie);
          const int rank_in = __popc(tm & ((1u << lane) - 1u));
          const bool drop = k < D && (key[i] < lo || (tie && rank_in < ties_left));
          if (k < D) y[k] = drop ? 0.f : xv[i];
          ties_left -= __popc(tm);
          if (ties_left < 0) ties_left = 0;
        }
      }
    }
  } else if (t == 1 && n_drop > 0) {
    const unsigned long long rs = seed ^ (offset * 0xD1B54A32D192ED03ull);
    // smallest threshold thr with count(key < thr) >= n_drop  (bisection on the 32-bit key value)
    unsigned lo = 0u, hi = 0xFFFFFFFFu;
    while (lo < hi) {
      const unsigned mid = lo + ((hi - lo) >> 1);
      int cnt = 0;
      for (int k = lane; k < D; k += 32) cnt += aug_key(rs, (unsigned long long)row, k) <= mid ? 1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (cnt >= n_drop) hi = mid; else lo = mid + 1u;
    }
    // keys <= lo are dropped; ties at the threshold (probability ~ D^2 / 2^32) are broken by column order
    int below = 0;
    for (int k = lane; k < D; k += 32) below += aug_key(rs, (unsigned long long)row, k) < lo ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
    int ties_left = n_drop - below;
    for (int k0 = 0; k0 < D; k0 += 32) {
      const int k = k0 + lane;
      const unsigned key = k < D ? aug_key(rs, (unsigned long long)row, k) : 0xFFFFFFFFu;
      const bool tie = k < D && key == lo;
      const unsigned tm = __ballot_sync(0xffffffffu, tie);
      const int rank_in = __popc(tm & ((1u << lane) - 1u));
      const bool drop = k < D && (key < lo || (tie && rank_in < ties_left));
      if (k < D) y[k] = drop ? 0.f : x[k];
      ties_left -= __popc(tm);
      if (ties_left < 0) ties_left = 0;
    }
  } else if (cached) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int k = lane + 32 * i;
      if (k < D) y[k] = xv[i];
    }
  } else {
    for (int k = lane; k < D; k += 32) y[k] = x[k];
  }
}

// ------------------------------------------------------------------------------------------
// DMVAE head
// ------------------------------------------------------------------------------------------
constexpr int kMaxViews = 16;
struct ViewPtrs {
  const float* in[kMaxViews];
  const float* in2[kMaxViews];
  float* out[kMaxViews];
};

__global__ void dmvae_head_fwd_kernel(ViewPtrs P, const float* __restrict__ noise, int N, int B, int e, float poeT,
                                      float* __restrict__ kl3) {
  __shared__ float red[32];
  const long long total = (long long)B * e;
  float kl_p = 0.f, kl_poe = 0.f, kl_u = 0.f;
  const float invT = 1.0f / fmaxf(poeT, 1e-8f);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / e), k = (int)(idx - (long long)b * e);
    float psum = invT;   // prior expert: exp(-0)/T
    float num = 0.f;
    float zsu[kMaxViews];
    for (int i = 0; i < N; ++i) {
      const float* st = P.in[i] + (long long)b * 4 * e;
      const float mu_s = st[k], lv_s = st[e + k], mu_p = st[2 * e + k], lv_p = st[3 * e + k];
      const float zp = mu_p + noise[((long long)i * B + b) * e + k] * expf(0.5f * lv_p);
      zsu[i] = mu_s + noise[((long long)(N + i) * B + b) * e + k] * expf(0.5f * lv_s);
      kl_p += -0.5f * (1.0f + lv_p - mu_p * mu_p - expf(lv_p));
      kl_u += -0.5f * (1.0f + lv_s - mu_s * mu_s - expf(lv_s));
      const float prec = expf(-lv_s) * invT;
      psum += prec;
      num = fmaf(prec, mu_s, num);
      float* di = P.out[i];
      for (int j = 0; j < N; ++j) di[((long long)j * B + b) * 2 * e + k] = zp;
    }
    psum += 1e-8f;
    const float var = 1.0f / psum;
    const float mu = var * num;
    const float lv = logf(var);
    kl_poe += -0.5f * (1.0f + lv - mu * mu - expf(lv));
    const float zs = mu + noise[((long long)(2 * N) * B + b) * e + k] * expf(0.5f * lv);
    for (int i = 0; i < N; ++i) {
      float* di = P.out[i];
      for (int j = 0; j < N; ++j) di[((long long)j * B + b) * 2 * e + e + k] = (j == i) ? zs : zsu[j];
    }
  }
  const float invB = 1.0f / (float)B;
  const float s0 = block_sum(kl_p, red), s1 = block_sum(kl_poe, red), s2 = block_sum(kl_u, red);
  if (threadIdx.x == 0) {
    atomicAdd(kl3 + 0, s0 * invB);
    atomicAdd(kl3 + 1, s1 * invB);
    atomicAdd(kl3 + 2, s2 * invB);
  }
}

// P.in = stats, P.in2 = d_dec_in, P.out = d_stats
__global__ void dmvae_head_bwd_kernel(ViewPtrs P, const float* __restrict__ noise, int N, int B, int e, float poeT,
                                      const float* __restrict__ kl_grad3) {
  const long long total = (long long)B * e;
  const float invT = 1.0f / fmaxf(poeT, 1e-8f);
  const float invB = 1.0f / (float)B;
  const float wkl_p = __ldg(kl_grad3 + 0) * invB, wkl_poe = __ldg(kl_grad3 + 1) * invB, wkl_u = __ldg(kl_grad3 + 2) * invB;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / e), k = (int)(idx - (long long)b * e);
    // PoE forward pieces
    float psum = invT, num = 0.f;
    for (int i = 0; i < N; ++i) {
      const float* st = P.in[i] + (long long)b * 4 * e;
      const float prec = expf(-st[e + k]) * invT;
      psum += prec;
      num = fmaf(prec, st[k], num);
    }
    psum += 1e-8f;
    const float var = 1.0f / psum, mu = var * num, lv = logf(var);
    const float eps_s = noise[((long long)(2 * N) * B + b) * e + k];
    float dzs = 0.f;
    for (int i = 0; i < N; ++i) dzs += P.in2[i][((long long)i * B + b) * 2 * e + e + k];
    const float dmu = dzs + wkl_poe * mu;
    const float dlv = dzs * eps_s * 0.5f * expf(0.5f * lv) + wkl_poe * 0.5f * (expf(lv) - 1.0f);
    const float dnum = dmu * var;
    const float dpsum = -dlv * var - dmu * mu * var;
    for (int i = 0; i < N; ++i) {
      const float* st = P.in[i] + (long long)b * 4 * e;
      const float mu_s = st[k], lv_s = st[e + k], mu_p = st[2 * e + k], lv_p = st[3 * e + k];
      float dzp = 0.f, dzu = 0.f;
      for (int j = 0; j < N; ++j) dzp += P.in2[i][((long long)j * B + b) * 2 * e + k];
      for (int j = 0; j < N; ++j)
        if (j != i) dzu += P.in2[j][((long long)i * B + b) * 2 * e + e + k];
      const float eps_p = noise[((long long)i * B + b) * e + k];
      const float eps_u = noise[((long long)(N + i) * B + b) * e + k];
      const float prec = expf(-lv_s) * invT;
      const float dprec = dpsum + dnum * mu_s;
      float* ds = P.out[i] + (long long)b * 4 * e;
      ds[k] = dzu + wkl_u * mu_s + dnum * prec;
      ds[e + k] = dzu * eps_u * 0.5f * expf(0.5f * lv_s) + wkl_u * 0.5f * (expf(lv_s) - 1.0f) - dprec * prec;
      ds[2 * e + k] = dzp + wkl_p * mu_p;
      ds[3 * e + k] = dzp * eps_p * 0.5f * expf(0.5f * lv_p) + wkl_p * 0.5f * (expf(lv_p) - 1.0f);
    }
  }
}

__global__ void dmvae_poe_mean_kernel(ViewPtrs P, int N, int B, int e, float poeT, float* __restrict__ mu_out) {
  const long long total = (long long)B * e;
  const float invT = 1.0f / fmaxf(poeT, 1e-8f);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / e), k = (int)(idx - (long long)b * e);
    float psum = invT, num = 0.f;
    for (int i = 0; i < N; ++i) {
      const float* st = P.in[i] + (long long)b * 4 * e;
      const float prec = expf(-st[e + k]) * invT;
      psum += prec;
      num = fmaf(prec, st[k], num);
    }
    psum += 1e-8f;
    mu_out[idx] = (1.0f / psum) * num;
  }
}

__global__ void dmvae_mse_kernel(const float* __restrict__ recon, long long ldr, const float* __restrict__ x,
                                 long long ldx, int N, int B, int d, int view, float w_joint, float w_cross,
                                 const float* __restrict__ gscale, float* __restrict__ out2,
                                 float* __restrict__ d_recon, long long lddr) {
  __shared__ float red[32];
  const long long total = (long long)N * B * d;
  const float g = gscale ? __ldg(gscale) : 1.0f;
  const float inv = 1.0f / ((float)B * (float)d);
  float sj = 0.f, sc = 0.f;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / d;
    const int c = (int)(idx - row * d);
    const int j = (int)(row / B), b = (int)(row - (long long)j * B);
    const float diff = recon[row * ldr + c] - x[(long long)b * ldx + c];
    const float w = (j == view) ? w_joint : w_cross;
    if (j == view) sj = fmaf(diff, diff, sj); else sc = fmaf(diff, diff, sc);
    if (d_recon) d_recon[row * lddr + c] = g * w * 2.0f * diff * inv;
  }
  const float s0 = block_sum(sj, red), s1 = block_sum(sc, red);
  if (threadIdx.x == 0) {
    atomicAdd(out2 + 0, w_joint * s0 * inv);
    atomicAdd(out2 + 1, w_cross * s1 * inv);
  }
}

// ------------------------------------------------------------------------------------------
// Adam / AdamW over a flat buffer
// ------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                            float wd, int decoupled, float step_size, float inv_bc2_sqrt, float gscale,
                            uint16_t* __restrict__ pb) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float pi = p[i];
    float gi = g[i] * gscale;
    if (wd != 0.f) {
      if (decoupled) pi *= (1.0f - lr * wd);
      else gi = fmaf(wd, pi, gi);
    }
    float mi = m[i], vi = v[i];
    mi = mi + (gi - mi) * (1.0f - beta1);
    vi = vi * beta2 + (1.0f - beta2) * gi * gi;
    const float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
    pi = pi - step_size * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (pb) pb[i] = f2bf(pi);
  }
}


// Graph-capturable variant: the 1-based step count and the learning rate live in DEVICE memory, so a captured
// CUDA graph of the whole training step can be replayed without re-baking host scalars.  state[0] = step count
// (as float, exact up to 2^24), state[1] = learning rate.  The bias corrections are formed per thread with
// exp2/log2 in double precision of the same quantities torch uses; adam_tick_kernel bumps the count afterwards.
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, long long n, const float* __restrict__ state, float beta1,
                                float beta2, float eps, float wd, int decoupled, float gscale,
                                uint16_t* __restrict__ pb) {
  const double step = (double)state[0] + 1.0;
  const float lr = state[1];
  const double bc1 = 1.0 - pow((double)beta1, step);
  const double bc2 = 1.0 - pow((double)beta2, step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float pi = p[i];
    float gi = g[i] * gscale;
    if (wd != 0.f) {
      if (decoupled) pi *= (1.0f - lr * wd);
      else gi = fmaf(wd, pi, gi);
    }
    float mi = m[i], vi = v[i];
    mi = mi + (gi - mi) * (1.0f - beta1);
    vi = vi * beta2 + (1.0f - beta2) * gi * gi;
    const float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
    pi = pi - step_size * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (pb) pb[i] = f2bf(pi);
  }
}
__global__ void adam_tick_kernel(float* state) { state[0] += 1.0f; }

__global__ void fill_kernel(float* __restrict__ p, long long n, float value) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = value;
}

// ------------------------------------------------------------------------------------------
// casts / transposes
// ------------------------------------------------------------------------------------------
__global__ void cast_bf16_kernel(const float* __restrict__ src, long long lds, uint16_t* __restrict__ dst,
                                 long long ldd, int rows, int cols) {
  const long long total = (long long)rows * cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cols;
    const int c = (int)(idx - r * cols);
    dst[r * ldd + c] = f2bf(src[r * lds + c]);
  }
}
// 4 elements per thread (128-bit load, 64-bit store); requires cols % 4 == 0, lds % 4 == 0, ldd % 4 == 0 and
// 16-byte / 8-byte aligned bases
__global__ void cast_bf16_vec4_kernel(const float* __restrict__ src, long long lds, uint16_t* __restrict__ dst,
                                      long long ldd, int rows, int cols4) {
  const long long total = (long long)rows * cols4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cols4;
    const int c = (int)(idx - r * cols4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + r * lds + c));
    uint2 pk;
    pk.x = (uint32_t)f2bf(v.x) | ((uint32_t)f2bf(v.y) << 16);
    pk.y = (uint32_t)f2bf(v.z) | ((uint32_t)f2bf(v.w) << 16);
    *reinterpret_cast<uint2*>(dst + r * ldd + c) = pk;
  }
}

template <typename TIn>
__global__ void transpose_to_bf16_kernel(const TIn* __restrict__ src, long long lds, uint16_t* __restrict__ dst,
                                         long long ldd, int rows, int cols) {
  __shared__ uint16_t tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    uint16_t v = 0;
    if (r < rows && c < cols) {
      if (sizeof(TIn) == 4) v = f2bf(reinterpret_cast<const float*>(src)[(long long)r * lds + c]);
      else v = reinterpret_cast<const uint16_t*>(src)[(long long)r * lds + c];
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[(long long)c * ldd + r] = tile[threadIdx.x][i];
  }
}


// bf16 [rows, cols] -> bf16 [cols, rows] with 128-byte accesses on both sides (64 x 64 tile, 32-bit = two bf16 per
// thread and access).  The 32 x 32 / 16-bit version above moved 64-byte segments and reached 1.3 TB/s on the
// [65536, 512] embedding matrices that every InfoNCE backward transposes (a per-rank cost that does not shrink
// with the data-parallel width).
__global__ void __launch_bounds__(256)
transpose_bf16_64_kernel(const uint16_t* __restrict__ src, long long lds, uint16_t* __restrict__ dst, long long ldd,
                         int rows, int cols) {
  __shared__ uint16_t tile[64][66];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = warp * 8 + i;
    const uint32_t v = *reinterpret_cast<const uint32_t*>(src + (long long)(r0 + r) * lds + c0 + 2 * lane);
    *reinterpret_cast<uint32_t*>(&tile[r][2 * lane]) = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = warp * 8 + i;
    const uint32_t v = (uint32_t)tile[2 * lane][c] | ((uint32_t)tile[2 * lane + 1][c] << 16);
    *reinterpret_cast<uint32_t*>(dst + (long long)(c0 + c) * ldd + r0 + 2 * lane) = v;
  }
}

// fp32 [rows, cols] -> bf16 row-major copy AND bf16 transposed copy in ONE pass over the source, plus
// optional column sums (bias gradient).  The transposed copy is what the wgrad GEMM consumes as a K-major
// operand (K = batch).  64 x 64 tile per CTA; fp32 tile in smem (65-word pitch).
__global__ void __launch_bounds__(256)
cast_dual_kernel(const float* __restrict__ src, long long lds, uint16_t* __restrict__ dst, long long ldd,
                 uint16_t* __restrict__ dstT, long long ldt, float* __restrict__ colsum, int rows, int cols) {
  __shared__ float tile[64][65];
  __shared__ float cacc[64];
  const int t = threadIdx.x;
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int cg = (t & 15) * 4, rl = t >> 4;              // 16 column groups of 4, 16 row lanes
  if (colsum && t < 64) cacc[t] = 0.f;
  float cs0 = 0.f, cs1 = 0.f, cs2 = 0.f, cs3 = 0.f;
  const bool vec = ((lds & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (c0 + 64 <= cols);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int r = rl + p * 16;
    const int gr = r0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr < rows) {
      const float* sp = src + (long long)gr * lds + c0 + cg;
      if (vec) {
        v = __ldg(reinterpret_cast<const float4*>(sp));
      } else {
        if (c0 + cg + 0 < cols) v.x = sp[0];
        if (c0 + cg + 1 < cols) v.y = sp[1];
        if (c0 + cg + 2 < cols) v.z = sp[2];
        if (c0 + cg + 3 < cols) v.w = sp[3];
      }
      if (dst) {
        uint16_t* dp = dst + (long long)gr * ldd + c0 + cg;
        if (vec && ((ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0)) {
          uint2 pk;
          pk.x = (uint32_t)f2bf(v.x) | ((uint32_t)f2bf(v.y) << 16);
          pk.y = (uint32_t)f2bf(v.z) | ((uint32_t)f2bf(v.w) << 16);
          *reinterpret_cast<uint2*>(dp) = pk;
        } else {
          if (c0 + cg + 0 < cols) dp[0] = f2bf(v.x);
          if (c0 + cg + 1 < cols) dp[1] = f2bf(v.y);
          if (c0 + cg + 2 < cols) dp[2] = f2bf(v.z);
          if (c0 + cg + 3 < cols) dp[3] = f2bf(v.w);
        }
      }
    }
    cs0 += v.x; cs1 += v.y; cs2 += v.z; cs3 += v.w;
    tile[r][cg + 0] = v.x; tile[r][cg + 1] = v.y; tile[r][cg + 2] = v.z; tile[r][cg + 3] = v.w;
  }
  __syncthreads();
  if (colsum) {
    atomicAdd(&cacc[cg + 0], cs0); atomicAdd(&cacc[cg + 1], cs1);
    atomicAdd(&cacc[cg + 2], cs2); atomicAdd(&cacc[cg + 3], cs3);
  }
  if (dstT) {
    const int rg = (t & 15) * 4, cl = t >> 4;            // 16 row groups of 4, 16 column lanes
    const bool vt = ((ldt & 3) == 0) && ((reinterpret_cast<uintptr_t>(dstT) & 7) == 0) && (r0 + 64 <= rows);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int c = cl + p * 16;
      if (c0 + c >= cols) continue;
      uint16_t* dp = dstT + (long long)(c0 + c) * ldt + r0 + rg;
      const float a = tile[rg + 0][c], b = tile[rg + 1][c], cc = tile[rg + 2][c], d = tile[rg + 3][c];
      if (vt) {
        uint2 pk;
        pk.x = (uint32_t)f2bf(a) | ((uint32_t)f2bf(b) << 16);
        pk.y = (uint32_t)f2bf(cc) | ((uint32_t)f2bf(d) << 16);
        *reinterpret_cast<uint2*>(dp) = pk;
      } else {
        if (r0 + rg + 0 < rows) dp[0] = f2bf(a);
        if (r0 + rg + 1 < rows) dp[1] = f2bf(b);
        if (r0 + rg + 2 < rows) dp[2] = f2bf(cc);
        if (r0 + rg + 3 < rows) dp[3] = f2bf(d);
      }
    }
  }
  if (colsum) {
    __syncthreads();
    if (t < 64 && c0 + t < cols) atomicAdd(colsum + c0 + t, cacc[t]);
  }
}

// column sums of a bf16 matrix (bias gradient of a hidden layer from the bf16 dgrad output):
// out[n] += sum_m X[m, n]; CTA = 64 columns x 256 rows, 4-byte loads (two columns per thread): ragged / unaligned shapes
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const uint16_t* __restrict__ X, long long ld, int M, int N, float* __restrict__ out) {
  __shared__ float red[8][64];
  const int cp = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + cp * 2;
  const int r0 = blockIdx.y * 256;
  float s0 = 0.f, s1 = 0.f;
  const bool pair_ok = ((ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(X) & 3) == 0) && (c + 1 < N);
  for (int r = r0 + rl; r < min(M, r0 + 256); r += 8) {
    const uint16_t* p = X + (long long)r * ld + c;
    if (pair_ok) {
      const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p));
      s0 += bf2f((uint16_t)(w & 0xFFFFu));
      s1 += bf2f((uint16_t)(w >> 16));
    } else {
      if (c < N) s0 += bf2f(p[0]);
      if (c + 1 < N) s1 += bf2f(p[1]);
    }
  }
  red[rl][cp * 2] = s0;
  red[rl][cp * 2 + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    const int cc = blockIdx.x * 64 + threadIdx.x;
    if (cc < N) atomicAdd(out + cc, s);
  }
}

// Aligned shapes (N % 256 == 0, 16-byte aligned rows): CTA = 256 columns x `rows_per_cta` rows; a warp reads one 512-byte
// row segment per instruction (16 bytes = 8 columns per lane), four rows in flight per lane.
__global__ void __launch_bounds__(256)
colsum_bf16_wide_kernel(const uint16_t* __restrict__ X, long long ld, int M, int rows_per_cta, float* __restrict__ out) {
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(M, r0 + rows_per_cta);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  auto add8 = [&](const uint4 w) {
    const uint32_t u[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[2 * j] += __uint_as_float(u[j] << 16);
      acc[2 * j + 1] += __uint_as_float(u[j] & 0xFFFF0000u);
    }
  };
  const uint16_t* p = X + (long long)(r0 + rl) * ld + c;
  const long long step = 8 * ld;
  int r = r0 + rl;
  for (; r + 24 < r1; r += 32, p += 4 * step) {
    const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(p));
    const uint4 w1 = __ldg(reinterpret_cast<const uint4*>(p + step));
    const uint4 w2 = __ldg(reinterpret_cast<const uint4*>(p + 2 * step));
    const uint4 w3 = __ldg(reinterpret_cast<const uint4*>(p + 3 * step));
    add8(w0); add8(w1); add8(w2); add8(w3);
  }
  for (; r < r1; r += 8, p += step) add8(__ldg(reinterpret_cast<const uint4*>(p)));
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][lane * 8 + j] = acc[j];
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
  atomicAdd(out + blockIdx.x * 256 + threadIdx.x, s);
}

inline int grid_for(long long n, int threads = 256) {
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)kNumSMs * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace dmf

using namespace dmf;

extern "C" int dmf_row_normalize_fwd(const float* X, long long ldx, int rows, int D, float eps, float* Y, long long ldy,
                                     uint16_t* Yb, long long ldyb, float* inv_norm, dmf_stream_t s) {
  DMF_REQUIRE(X && (Y || Yb) && rows >= 0 && D >= 1, "dmf_row_normalize_fwd: bad arguments");
  if (rows == 0) return 0;
  row_normalize_fwd_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)s>>>(X, ldx, rows, D, eps, Y, ldy, Yb, ldyb, inv_norm);
  return launched("dmf_row_normalize_fwd");
}
extern "C" int dmf_row_normalize_bwd(const float* Y, long long ldy, const float* inv_norm, const float* dY, long long lddy,
                                     int rows, int D, float* dX, long long lddx, int accumulate, dmf_stream_t s) {
  DMF_REQUIRE(Y && inv_norm && dY && dX && rows >= 0 && D >= 1, "dmf_row_normalize_bwd: bad arguments");
  if (rows == 0) return 0;
  row_normalize_bwd_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)s>>>(Y, ldy, inv_norm, dY, lddy, rows, D, dX, lddx, accumulate);
  return launched("dmf_row_normalize_bwd");
}
extern "C" int dmf_sumsq_f32(const float* G, long long n, float* out, dmf_stream_t s) {
  DMF_REQUIRE(G && out && n >= 0, "dmf_sumsq_f32: bad arguments");
  if (n == 0) return 0;
  sumsq_kernel<<<grid_for(n), 256, 0, (cudaStream_t)s>>>(G, n, out);
  return launched("dmf_sumsq_f32");
}

extern "C" int dmf_vmf_fwd(const float* E, long long lde, const float* nw, const float* nv, int rows, int D, float* Z,
                           long long ldz, uint16_t* Zb, long long ldzb, dmf_stream_t s) {
  DMF_REQUIRE(E && nw && nv && (Z || Zb) && rows >= 0 && D >= 2, "dmf_vmf_fwd: bad arguments");
  if (rows == 0) return 0;
  vmf_fwd_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)s>>>(E, lde, nw, nv, rows, D, Z, ldz, Zb, ldzb);
  return launched("dmf_vmf_fwd");
}
extern "C" int dmf_vmf_bwd(const float* E, long long lde, const float* nw, const float* nv, const float* dZ, long long lddz,
                           int rows, int D, float* dE, long long ldde, int accumulate, dmf_stream_t s) {
  DMF_REQUIRE(E && nw && nv && dZ && dE && rows >= 0 && D >= 2, "dmf_vmf_bwd: bad arguments");
  if (rows == 0) return 0;
  vmf_bwd_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)s>>>(E, lde, nw, nv, dZ, lddz, rows, D, dE, ldde, accumulate);
  return launched("dmf_vmf_bwd");
}
extern "C" int dmf_vmf_draw(float* nw, float* nv, int rows, int D, float kappa, unsigned long long seed,
                            unsigned long long offset, dmf_stream_t s) {
  DMF_REQUIRE(nw && nv && rows >= 0 && D >= 2 && D != 3 && kappa > 0.f, "dmf_vmf_draw: bad arguments (D=3 uses the host sampler)");
  if (rows == 0) return 0;
  vmf_draw_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)s>>>(nw, nv, rows, D, kappa, seed, offset, nullptr);
  return launched("dmf_vmf_draw");
}
extern "C" int dmf_vmf_draw_ctr(float* nw, float* nv, int rows, int D, float kappa, unsigned long long seed,
                                unsigned long long offset, const unsigned long long* counter, dmf_stream_t s) {
  DMF_REQUIRE(nw && nv && counter && rows >= 0 && D >= 2 && D != 3 && kappa > 0.f, "dmf_vmf_draw_ctr: bad arguments");
  if (rows == 0) return 0;
  vmf_draw_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)s>>>(nw, nv, rows, D, kappa, seed, offset, counter);
  return launched("dmf_vmf_draw_ctr");
}
__global__ void counter_add_kernel(unsigned long long* c, unsigned long long inc) { *c += inc; }
extern "C" int dmf_counter_add(unsigned long long* counter, unsigned long long inc, dmf_stream_t s) {
  DMF_REQUIRE(counter, "dmf_counter_add: null counter");
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)s>>>(counter, inc);
  return launched("dmf_counter_add");
}


extern "C" int dmf_augment(const float* X, long long ldx, float* Y, long long ldy, int rows, int D, float noise_scale,
                           int drop_scale, unsigned long long seed, unsigned long long offset, int* choice_out,
                           dmf_stream_t s) {
  DMF_REQUIRE(X && Y && rows >= 0 && D >= 1 && drop_scale >= 1, "dmf_augment: bad arguments");
  if (rows == 0) return 0;
  augment_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)s>>>(X, ldx, Y, ldy, rows, D, noise_scale, D / drop_scale, seed,
                                                            offset, choice_out, nullptr);
  return launched("dmf_augment");
}
extern "C" int dmf_augment_ctr(const float* X, long long ldx, float* Y, long long ldy, int rows, int D, float noise_scale,
                               int drop_scale, unsigned long long seed, unsigned long long offset, int* choice_out,
                               const unsigned long long* counter, dmf_stream_t s) {
  DMF_REQUIRE(X && Y && counter && rows >= 0 && D >= 1 && drop_scale >= 1, "dmf_augment_ctr: bad arguments");
  if (rows == 0) return 0;
  augment_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)s>>>(X, ldx, Y, ldy, rows, D, noise_scale, D / drop_scale, seed,
                                                            offset, choice_out, counter);
  return launched("dmf_augment_ctr");
}

static int fill_views(ViewPtrs& P, const float* const* in, const float* const* in2, float* const* out, int N) {
  for (int i = 0; i < N; ++i) {
    P.in[i] = in ? in[i] : nullptr;
    P.in2[i] = in2 ? in2[i] : nullptr;
    P.out[i] = out ? out[i] : nullptr;
  }
  return 0;
}

extern "C" int dmf_dmvae_head_fwd(const float* const* stats, const float* noise, int N, int B, int e, float poeT,
                                  float* const* dec_in, float* kl3, dmf_stream_t s) {
  DMF_REQUIRE(stats && noise && dec_in && kl3, "dmf_dmvae_head_fwd: null argument");
  DMF_REQUIRE(N >= 1 && N <= kMaxViews && B >= 1 && e >= 1, "dmf_dmvae_head_fwd: bad shape N=%d B=%d e=%d", N, B, e);
  ViewPtrs P;
  fill_views(P, stats, nullptr, dec_in, N);
  dmvae_head_fwd_kernel<<<grid_for((long long)B * e), 256, 0, (cudaStream_t)s>>>(P, noise, N, B, e, poeT, kl3);
  return launched("dmf_dmvae_head_fwd");
}
extern "C" int dmf_dmvae_head_bwd(const float* const* stats, const float* noise, const float* const* d_dec_in, int N,
                                  int B, int e, float poeT, const float* kl_grad3, float* const* d_stats,
                                  dmf_stream_t s) {
  DMF_REQUIRE(stats && noise && d_dec_in && d_stats && kl_grad3, "dmf_dmvae_head_bwd: null argument");
  DMF_REQUIRE(N >= 1 && N <= kMaxViews && B >= 1 && e >= 1, "dmf_dmvae_head_bwd: bad shape");
  ViewPtrs P;
  fill_views(P, stats, d_dec_in, d_stats, N);
  dmvae_head_bwd_kernel<<<grid_for((long long)B * e), 256, 0, (cudaStream_t)s>>>(P, noise, N, B, e, poeT, kl_grad3);
  return launched("dmf_dmvae_head_bwd");
}
extern "C" int dmf_dmvae_poe_mean(const float* const* stats, int N, int B, int e, float poeT, float* mu_poe, dmf_stream_t s) {
  DMF_REQUIRE(stats && mu_poe && N >= 1 && N <= kMaxViews && B >= 1 && e >= 1, "dmf_dmvae_poe_mean: bad arguments");
  ViewPtrs P;
  fill_views(P, stats, nullptr, nullptr, N);
  dmvae_poe_mean_kernel<<<grid_for((long long)B * e), 256, 0, (cudaStream_t)s>>>(P, N, B, e, poeT, mu_poe);
  return launched("dmf_dmvae_poe_mean");
}
extern "C" int dmf_dmvae_mse_fwd_bwd(const float* recon, long long ldr, const float* x, long long ldx, int N, int B, int d,
                                     int view, float w_joint, float w_cross, const float* gscale, float* out2,
                                     float* d_recon, long long lddr, dmf_stream_t s) {
  DMF_REQUIRE(recon && x && out2 && N >= 1 && B >= 1 && d >= 1, "dmf_dmvae_mse_fwd_bwd: bad arguments");
  dmvae_mse_kernel<<<grid_for((long long)N * B * d), 256, 0, (cudaStream_t)s>>>(recon, ldr, x, ldx, N, B, d, view, w_joint,
                                                                                w_cross, gscale, out2, d_recon, lddr);
  return launched("dmf_dmvae_mse_fwd_bwd");
}

extern "C" int dmf_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                             float eps, float wd, int decoupled, int step, float grad_scale, uint16_t* p_bf16,
                             dmf_stream_t s) {
  DMF_REQUIRE(p && g && m && v && n >= 0 && step >= 1, "dmf_adam_step: bad arguments");
  if (n == 0) return 0;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  adam_kernel<<<grid_for(n), 256, 0, (cudaStream_t)s>>>(p, g, m, v, n, lr, beta1, beta2, eps, wd, decoupled, step_size,
                                                        inv_bc2_sqrt, grad_scale, p_bf16);
  return launched("dmf_adam_step");
}

extern "C" int dmf_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float* state, float beta1,
                                 float beta2, float eps, float wd, int decoupled, float grad_scale, uint16_t* p_bf16,
                                 dmf_stream_t s) {
  DMF_REQUIRE(p && g && m && v && state && n >= 0, "dmf_adam_step_dev: bad arguments");
  if (n == 0) return 0;
  adam_dev_kernel<<<grid_for(n), 256, 0, (cudaStream_t)s>>>(p, g, m, v, n, state, beta1, beta2, eps, wd, decoupled,
                                                            grad_scale, p_bf16);
  int rc = launched("dmf_adam_step_dev");
  if (rc) return rc;
  adam_tick_kernel<<<1, 1, 0, (cudaStream_t)s>>>(state);
  return launched("dmf_adam_step_dev(tick)");
}
extern "C" int dmf_fill_f32(float* p, long long n, float value, dmf_stream_t s) {
  DMF_REQUIRE(p && n >= 0, "dmf_fill_f32: bad arguments");
  if (n == 0) return 0;
  fill_kernel<<<grid_for(n), 256, 0, (cudaStream_t)s>>>(p, n, value);
  return launched("dmf_fill_f32");
}

extern "C" int dmf_cast_f32_to_bf16(const float* src, long long lds, uint16_t* dst, long long ldd, int rows, int cols,
                                    dmf_stream_t s) {
  DMF_REQUIRE(src && dst && rows >= 0 && cols >= 0, "dmf_cast_f32_to_bf16: bad arguments");
  if (rows == 0 || cols == 0) return 0;
  if ((cols & 3) == 0 && (lds & 3) == 0 && (ldd & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
    cast_bf16_vec4_kernel<<<grid_for((long long)rows * (cols / 4)), 256, 0, (cudaStream_t)s>>>(src, lds, dst, ldd, rows,
                                                                                             cols / 4);
    return launched("dmf_cast_f32_to_bf16");
  }
  cast_bf16_kernel<<<grid_for((long long)rows * cols), 256, 0, (cudaStream_t)s>>>(src, lds, dst, ldd, rows, cols);
  return launched("dmf_cast_f32_to_bf16");
}
extern "C" int dmf_cast_transpose_f32_to_bf16(const float* src, long long lds, uint16_t* dst, long long ldd, int rows,
                                              int cols, dmf_stream_t s) {
  DMF_REQUIRE(src && dst && rows >= 0 && cols >= 0, "dmf_cast_transpose_f32_to_bf16: bad arguments");
  if (rows == 0 || cols == 0) return 0;
  dim3 block(32, 8), grid((cols + 31) / 32, (rows + 31) / 32);
  transpose_to_bf16_kernel<float><<<grid, block, 0, (cudaStream_t)s>>>(src, lds, dst, ldd, rows, cols);
  return launched("dmf_cast_transpose_f32_to_bf16");
}
extern "C" int dmf_transpose_bf16(const uint16_t* src, long long lds, uint16_t* dst, long long ldd, int rows, int cols,
                                  dmf_stream_t s) {
  DMF_REQUIRE(src && dst && rows >= 0 && cols >= 0, "dmf_transpose_bf16: bad arguments");
  if (rows == 0 || cols == 0) return 0;
  // full 64 x 64 tiles with 4-byte aligned rows on both sides take the wide kernel
  if ((rows & 63) == 0 && (cols & 63) == 0 && ((lds | ldd) & 1) == 0 &&
      ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 3) == 0 && rows / 64 <= 65535) {
    dim3 grid64(cols / 64, rows / 64);
    transpose_bf16_64_kernel<<<grid64, 256, 0, (cudaStream_t)s>>>(src, lds, dst, ldd, rows, cols);
    return launched("dmf_transpose_bf16");
  }
  dim3 block(32, 8), grid((cols + 31) / 32, (rows + 31) / 32);
  transpose_to_bf16_kernel<uint16_t><<<grid, block, 0, (cudaStream_t)s>>>(src, lds, dst, ldd, rows, cols);
  return launched("dmf_transpose_bf16");
}

extern "C" int dmf_cast_dual_bf16(const float* src, long long lds, uint16_t* dst, long long ldd, uint16_t* dst_t,
                                  long long ldt, float* colsum, int rows, int cols, dmf_stream_t s) {
  DMF_REQUIRE(src && (dst || dst_t || colsum) && rows >= 0 && cols >= 0, "dmf_cast_dual_bf16: bad arguments");
  if (rows == 0 || cols == 0) return 0;
  dim3 grid((cols + 63) / 64, (rows + 63) / 64);
  DMF_REQUIRE(grid.y <= 65535, "dmf_cast_dual_bf16: too many rows (%d)", rows);
  cast_dual_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(src, lds, dst, ldd, dst_t, ldt, colsum, rows, cols);
  return launched("dmf_cast_dual_bf16");
}
extern "C" int dmf_colsum_bf16(const uint16_t* X, long long ld, int M, int N, float* out, dmf_stream_t s) {
  DMF_REQUIRE(X && out && M >= 0 && N >= 0, "dmf_colsum_bf16: bad arguments");
  if (M == 0 || N == 0) return 0;
  if ((N & 255) == 0 && (ld & 7) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0) {
    // enough CTAs to fill the GPU a few times over, at least 64 rows each
    int rows = 512;
    while (rows > 64 && (long long)(N / 256) * ((M + rows - 1) / rows) < 4 * kNumSMs) rows >>= 1;
    dim3 gridw(N / 256, (M + rows - 1) / rows);
    colsum_bf16_wide_kernel<<<gridw, 256, 0, (cudaStream_t)s>>>(X, ld, M, rows, out);
    return launched("dmf_colsum_bf16");
  }
  dim3 grid((N + 63) / 64, (M + 255) / 256);
  DMF_REQUIRE(grid.y <= 65535, "dmf_colsum_bf16: too many rows (%d)", M);
  colsum_bf16_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(X, ld, M, N, out);
  return launched("dmf_colsum_bf16");
}
