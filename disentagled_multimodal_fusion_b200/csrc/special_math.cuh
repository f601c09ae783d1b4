// fp32 special functions shared by the evidential kernels.
//
// gamma3(x): lgamma, digamma and trigamma of the same argument in one evaluation, for x >= 1
// (Dirichlet parameters alpha = evidence + 1 and their sums).  Replaces torch.lgamma /
// torch.digamma calls of models/losses.py:122,193-201 and the trigamma that autograd would
// evaluate in backward.  Arguments below 16 are shifted by 6 with the recurrence written as ONE
// division:  P(x) = x(x+1)...(x+5),  sum 1/(x+k) = P'/P,  sum 1/(x+k)^2 = (P'^2 - P P'')/P^2
// (all-positive coefficients, no cancellation), then the Stirling / asymptotic series is taken at
// X = x+6 >= 7 (truncation error < 1e-9).  Accuracy vs float64 (tests/test_special_math.py):
// digamma abs 1e-6, trigamma rel 8e-7, lgamma abs 3e-6 for x<16 / rel 4e-7 above.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace dmf {

struct Gamma3 {
  float lgam, psi, psi1;
};

template <bool kNeedLgamma>
__host__ __device__ __forceinline__ Gamma3 gamma3(float x) {
  float r1 = 0.f, s2 = 0.f, lnP = 0.f, X = x;
  if (x < 16.f) {
    const float P = (((((x + 15.f) * x + 85.f) * x + 225.f) * x + 274.f) * x + 120.f) * x;
    const float P1 = ((((6.f * x + 75.f) * x + 340.f) * x + 675.f) * x + 548.f) * x + 120.f;
    float Q = 6.f;
    Q = Q * x + 150.f;
    Q = Q * x + 1635.f;
    Q = Q * x + 10200.f;
    Q = Q * x + 40208.f;
    Q = Q * x + 104370.f;
    Q = Q * x + 180455.f;
    Q = Q * x + 205800.f;
    Q = Q * x + 150152.f;
    Q = Q * x + 65760.f;
    Q = Q * x + 14400.f;
    const float rP = 1.0f / P;
    r1 = P1 * rP;
    s2 = Q * rP * rP;
    if (kNeedLgamma) lnP = logf(P);
    X = x + 6.f;
  }
  const float iX = 1.0f / X;
  const float iX2 = iX * iX;
  const float lnX = logf(X);
  Gamma3 g;
  g.psi = lnX - 0.5f * iX - iX2 * (8.3333333333e-2f - iX2 * (8.3333333333e-3f - iX2 * 3.9682539683e-3f)) - r1;
  g.psi1 = iX * (1.0f + 0.5f * iX +
                 iX2 * (1.6666666667e-1f - iX2 * (3.3333333333e-2f - iX2 * (2.3809523810e-2f - iX2 * 3.3333333333e-2f)))) + s2;
  if (kNeedLgamma) {
    g.lgam = (X - 0.5f) * lnX - X + 0.91893853320467274f +
             iX * (8.3333333333e-2f - iX2 * (2.7777777778e-3f - iX2 * 7.9365079365e-4f)) - lnP;
  } else {
    g.lgam = 0.f;
  }
  return g;
}


// Device-only fast variant used by the fused EDL kernel: shift by 4 (P = x(x+1)(x+2)(x+3), X = x+4 >= 5 keeps
// the truncation error of the same series below 1.1e-8 / 1.6e-9 / 7.6e-9), branch-free (select instead of a
// divergent branch: evidence magnitudes mix within a warp), MUFU reciprocals and logarithms (rcp.approx /
// lg2.approx, <= 1 ulp / 2^-22): ~45 FP32 instructions + 4 MUFU for all three functions.
#ifdef __CUDACC__
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ln_fast(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y * 0.6931471805599453f;
}
template <bool kNeedLgamma>
__device__ __forceinline__ Gamma3 gamma3_fast(float x) {
  const bool small = x < 16.f;
  const float xs = small ? x : 1.0f;
  const float P = (((xs + 6.f) * xs + 11.f) * xs + 6.f) * xs;
  const float P1 = ((4.f * xs + 18.f) * xs + 22.f) * xs + 6.f;
  float Q = 4.f;                       // P'^2 - P P'' (all-positive coefficients)
  Q = Q * xs + 36.f;
  Q = Q * xs + 130.f;
  Q = Q * xs + 240.f;
  Q = Q * xs + 242.f;
  Q = Q * xs + 132.f;
  Q = Q * xs + 36.f;
  const float rP = rcp_fast(P);
  const float r1 = small ? P1 * rP : 0.f;
  const float s2 = small ? Q * rP * rP : 0.f;
  const float X = small ? x + 4.f : x;
  const float iX = rcp_fast(X);
  const float iX2 = iX * iX;
  const float lnX = ln_fast(X);
  Gamma3 g;
  g.psi = lnX - 0.5f * iX - iX2 * (8.3333333333e-2f - iX2 * (8.3333333333e-3f - iX2 * 3.9682539683e-3f)) - r1;
  g.psi1 = iX * (1.0f + 0.5f * iX +
                 iX2 * (1.6666666667e-1f - iX2 * (3.3333333333e-2f - iX2 * (2.3809523810e-2f - iX2 * 3.3333333333e-2f)))) + s2;
  if (kNeedLgamma) {
    const float lnP = small ? ln_fast(P) : 0.f;
    g.lgam = (X - 0.5f) * lnX - X + 0.91893853320467274f +
             iX * (8.3333333333e-2f - iX2 * (2.7777777778e-3f - iX2 * 7.9365079365e-4f)) - lnP;
  } else {
    g.lgam = 0.f;
  }
  return g;
}
#endif

// activation_function(h, 'exp'), utils.py:46-63, same op order in fp32:
//   h <- clamp(h,-10,10); L = 13*log(10); e = exp((h+L) - logaddexp(h,L))
__host__ __device__ __forceinline__ float evidence_act(float h) {
  h = fminf(fmaxf(h, -10.f), 10.f);
  const float L = 13.f * 2.3025851f;  // fp32 log(10) times 13, as the reference computes it
  const float num = h + L;
  const float den = fmaxf(h, L) + log1pf(expf(-fabsf(h - L)));
  return expf(num - den);
}
// d e / d h (autograd of the ops above): zero outside the clamp range
__host__ __device__ __forceinline__ float evidence_act_grad(float h, float e) {
  if (!(h >= -10.f && h <= 10.f)) return 0.f;
  const float L = 13.f * 2.3025851f;
  return e * (1.0f - expf(h - fmaxf(h, L) - log1pf(expf(-fabsf(h - L)))));
}

}  // namespace dmf
