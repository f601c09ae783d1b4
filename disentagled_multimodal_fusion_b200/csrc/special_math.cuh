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

// ---------------------------------------------------------------------------------------------------------------
// Packed fp32x2 evaluation (Blackwell FFMA2 / FMUL2 / FADD2: one issue slot for two lanes of work).  The fused EDL
// kernel is issue-bound, not HBM-bound (ncu: 72 % issue slots busy at 29-40 % of the HBM peak), and ~3/4 of its
// instructions are Horner chains with constant coefficients -- exactly what f32x2 halves.  Two classes of a
// Dirichlet row travel in one 64-bit register pair.
//
// What the EDL loss needs per class is not the gamma triple itself but two combinations of it:
//     k(x) = (x-1) psi(x) - lgamma(x)        (KL term; the psi(S~) part is factored out per row)
//     h(x) = (x-1) psi1(x)                   (its gradient; h = k')
// With the shift-by-4 recurrence folded in through a float mask m = [x < 8] (no selects, every step an FMA):
//     xs = 1 + m (x-1),  P = xs(xs+1)(xs+2)(xs+3),  r1 = m P'/P,  s2 = r1^2 - m P''/P  (= sum 1/(x+k)^2),
//     X = x + 4m,  k = ln2 (m lg2 P - (4m + 1/2) lg2 X) + X - c - (x-1)(iX(1/2 + iX A) + r1) - iX B,
//     h = (x-1) (iX (1 + iX (1/2 + iX D)) + s2),   A, B, D = Stirling tails in iX^2.
// Emulated in fp32 against float64 scipy: k abs 2.5e-6 for x < 16, rel 1.5e-7 above; h rel 1.5e-6.
struct f2 {
  unsigned long long v;
};
__device__ __forceinline__ f2 pk2(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f2 splat2(float a) { return pk2(a, a); }
__device__ __forceinline__ void unpk2(f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ float lg2_fast(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ f2 rcp2(f2 a) {
  float lo, hi;
  unpk2(a, lo, hi);
  return pk2(rcp_fast(lo), rcp_fast(hi));
}
__device__ __forceinline__ f2 lg22(f2 a) {
  float lo, hi;
  unpk2(a, lo, hi);
  return pk2(lg2_fast(lo), lg2_fast(hi));
}

// shared sub-expressions of the packed evaluation
struct GammaCore2 {
  f2 u, m, r1, s2, X, iX, t, lX, lPm;
};
__device__ __forceinline__ GammaCore2 gamma_core2(f2 x) {
  GammaCore2 g;
  float x0, x1;
  unpk2(x, x0, x1);
  g.u = add2(x, splat2(-1.0f));
  g.m = pk2(x0 < 8.0f ? 1.0f : 0.0f, x1 < 8.0f ? 1.0f : 0.0f);
  const f2 xs = fma2(g.m, g.u, splat2(1.0f));
  f2 P = add2(xs, splat2(6.0f));
  P = fma2(P, xs, splat2(11.0f));
  P = fma2(P, xs, splat2(6.0f));
  P = mul2(P, xs);
  f2 P1 = fma2(splat2(4.0f), xs, splat2(18.0f));
  P1 = fma2(P1, xs, splat2(22.0f));
  P1 = fma2(P1, xs, splat2(6.0f));
  f2 P2 = fma2(splat2(12.0f), xs, splat2(36.0f));
  P2 = fma2(P2, xs, splat2(22.0f));
  const f2 rPm = mul2(rcp2(P), g.m);
  g.r1 = mul2(P1, rPm);
  g.s2 = fma2(g.r1, g.r1, mul2(mul2(P2, rPm), splat2(-1.0f)));
  g.X = fma2(g.m, splat2(4.0f), x);
  g.iX = rcp2(g.X);
  g.t = mul2(g.iX, g.iX);
  g.lX = lg22(g.X);
  g.lPm = mul2(lg22(P), g.m);
  return g;
}
// psi(x) + r1-corrected tail:  T1 = iX (1/2 + iX A) + r1   (psi = ln X - T1)
__device__ __forceinline__ f2 gamma_T1(const GammaCore2& g) {
  f2 A = fma2(g.t, splat2(3.9682539683e-3f), splat2(-8.3333333333e-3f));
  A = fma2(A, g.t, splat2(8.3333333333e-2f));
  return fma2(g.iX, fma2(g.iX, A, splat2(0.5f)), g.r1);
}
__device__ __forceinline__ f2 gamma_B(const GammaCore2& g) {
  f2 B = fma2(g.t, splat2(7.9365079365e-4f), splat2(-2.7777777778e-3f));
  return fma2(B, g.t, splat2(8.3333333333e-2f));
}
// psi1(x) = iX (1 + iX (1/2 + iX D)) + s2
__device__ __forceinline__ f2 gamma_psi1(const GammaCore2& g) {
  f2 D = fma2(g.t, splat2(2.3809523810e-2f), splat2(-3.3333333333e-2f));
  D = fma2(D, g.t, splat2(1.6666666667e-1f));
  f2 q = fma2(g.iX, D, splat2(0.5f));
  q = fma2(g.iX, q, splat2(1.0f));
  return fma2(g.iX, q, g.s2);
}
// (-k, h) of two classes: 37 packed FP ops + 8 MUFU + 2 selects for the pair (~24 issue slots per class against
// ~56 for gamma3_fast<true> plus the k / h assembly).  kn = -k so that every step is a plain FMA.
struct KH2 {
  f2 kn, h;
};
__device__ __forceinline__ KH2 gamma_kh2(f2 x) {
  float x0, x1;
  unpk2(x, x0, x1);
  const f2 u = add2(x, splat2(-1.0f));
  const f2 m = pk2(x0 < 8.0f ? 1.0f : 0.0f, x1 < 8.0f ? 1.0f : 0.0f);
  const f2 xs = fma2(m, u, splat2(1.0f));
  f2 P = add2(xs, splat2(6.0f));
  P = fma2(P, xs, splat2(11.0f));
  P = fma2(P, xs, splat2(6.0f));
  P = mul2(P, xs);
  f2 P1 = fma2(splat2(4.0f), xs, splat2(18.0f));
  P1 = fma2(P1, xs, splat2(22.0f));
  P1 = fma2(P1, xs, splat2(6.0f));
  f2 P2n = fma2(splat2(-12.0f), xs, splat2(-36.0f));
  P2n = fma2(P2n, xs, splat2(-22.0f));
  const f2 rPm = mul2(rcp2(P), m);
  const f2 r1 = mul2(P1, rPm);
  const f2 s2 = fma2(P2n, rPm, mul2(r1, r1));
  const f2 X = fma2(m, splat2(4.0f), x);
  const f2 iX = rcp2(X);
  const f2 t = mul2(iX, iX);
  const f2 lX = lg22(X);
  const f2 lPm = mul2(lg22(P), m);
  const f2 ncs = fma2(m, splat2(-4.0f), splat2(-0.5f));
  const f2 g1 = fma2(ncs, lX, lPm);
  const f2 cX = fma2(X, splat2(-1.0f), splat2(0.91893853320467274f));
  f2 kn = fma2(splat2(-0.6931471805599453f), g1, cX);
  f2 B = fma2(t, splat2(7.9365079365e-4f), splat2(-2.7777777778e-3f));
  B = fma2(B, t, splat2(8.3333333333e-2f));
  kn = fma2(iX, B, kn);
  f2 A = fma2(t, splat2(3.9682539683e-3f), splat2(-8.3333333333e-3f));
  A = fma2(A, t, splat2(8.3333333333e-2f));
  const f2 T1 = fma2(iX, fma2(iX, A, splat2(0.5f)), r1);
  KH2 r;
  r.kn = fma2(u, T1, kn);
  f2 D = fma2(t, splat2(2.3809523810e-2f), splat2(-3.3333333333e-2f));
  D = fma2(D, t, splat2(1.6666666667e-1f));
  f2 q = fma2(iX, D, splat2(0.5f));
  q = fma2(iX, q, splat2(1.0f));
  q = fma2(iX, q, s2);
  r.h = mul2(u, q);
  return r;
}
// psi of two arguments >= 1 only (aleatoric uncertainty of the evaluation pass): ~20 packed ops + 6 MUFU per pair
__device__ __forceinline__ f2 gamma_psi2(f2 x) {
  float x0, x1;
  unpk2(x, x0, x1);
  const f2 u = add2(x, splat2(-1.0f));
  const f2 m = pk2(x0 < 8.0f ? 1.0f : 0.0f, x1 < 8.0f ? 1.0f : 0.0f);
  const f2 xs = fma2(m, u, splat2(1.0f));
  f2 P = add2(xs, splat2(6.0f));
  P = fma2(P, xs, splat2(11.0f));
  P = fma2(P, xs, splat2(6.0f));
  P = mul2(P, xs);
  f2 P1n = fma2(splat2(-4.0f), xs, splat2(-18.0f));
  P1n = fma2(P1n, xs, splat2(-22.0f));
  P1n = fma2(P1n, xs, splat2(-6.0f));
  const f2 nr1 = mul2(P1n, mul2(rcp2(P), m));                 // -sum 1/(x+k)
  const f2 X = fma2(m, splat2(4.0f), x);
  const f2 iX = rcp2(X);
  const f2 t = mul2(iX, iX);
  f2 An = fma2(t, splat2(-3.9682539683e-3f), splat2(8.3333333333e-3f));
  An = fma2(An, t, splat2(-8.3333333333e-2f));
  const f2 nT = fma2(iX, fma2(iX, An, splat2(-0.5f)), nr1);   // -(iX (1/2 + iX A) + r1)
  return fma2(lg22(X), splat2(0.6931471805599453f), nT);
}

struct Gamma3x2 {
  f2 lgam, psi, psi1;
};
__device__ __forceinline__ Gamma3x2 gamma3x2(f2 x) {
  const GammaCore2 g = gamma_core2(x);
  Gamma3x2 r;
  const f2 lnX = mul2(g.lX, splat2(0.6931471805599453f));
  r.psi = add2(lnX, mul2(gamma_T1(g), splat2(-1.0f)));
  r.psi1 = gamma_psi1(g);
  f2 l = fma2(add2(g.X, splat2(-0.5f)), lnX, mul2(g.X, splat2(-1.0f)));
  l = add2(l, splat2(0.91893853320467274f));
  l = fma2(g.iX, gamma_B(g), l);
  r.lgam = fma2(splat2(-0.6931471805599453f), g.lPm, l);
  return r;
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
