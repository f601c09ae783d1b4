// K3: fused evidence fusion + EDL (AvgTrustedLoss) forward/backward + uncertainty summaries.
//
// One pass over evid [B,V,C] (fp32) replaces, per probe step, the ~60 small ATen kernels of
//   utils.py:66-116            fusion rules (cml / avg / joint / disentangled / dbf)
//   models/losses.py:117-248   edl_digamma_loss + annealed Dirichlet KL + get_dc_loss_vectorized
//   models/evidential_probe.py:139-143  u = C/S, aleatoric, argmax
// plus everything autograd would run in backward: the gradient d loss / d evid is produced in
// the same pass from closed forms:
//   d/de_c [psi(S)-psi(a_y)] = psi1(S) - [c==y] psi1(a_y)
//   dKL/de_c = (a_c-1) psi1(a_c) - (S~-C) psi1(S~)   for c != y, 0 for c == y
// so only C+2 (lgamma,digamma,trigamma) triples are evaluated per (sample, view).
//
// Memory-bound design: a CTA stages a tile of SPB samples (SPB*V*C floats, <= 12 KB) through
// shared memory with 128-bit coalesced loads, works on it in phases (row phase = one thread per
// (sample,view); element phase = one thread per (sample,view,class)), and writes grad / fused
// evidence tiles back with coalesced stores.  Algorithmic bytes per sample: 8VC + 4C + 16 + 8.
#include "common.cuh"
#include "special_math.cuh"

namespace dmf {

constexpr int kEdlThreads = 256;
constexpr int kEdlMaxTile = 3072;  // elements of evid per CTA tile (12 KB)
constexpr int kEdlEPT4 = kEdlMaxTile / (4 * kEdlThreads);   // float4 groups per thread = 3

// Layout of the dynamic shared memory (floats).  E = SPB*V*C, R = SPB*V.
struct EdlSmem {
  float* e;      // [E] evidence tile
  float* p;      // [E] projected probabilities alpha/(S+1e-8)            (dc / dbf only)
  float* S;      // [R] Dirichlet strength
  float* iT;     // [R] 1/(S+1e-8)
  float* om;     // [R] 1-u
  float* gy;     // [R] psi1(S) - psi1(alpha_y)             gradient of the label class
  float* gA;     // [R] psi1(S) - coef*(S~-C)*psi1(S~)      class-independent part for c != y
  float* psiT;   // [R] psi(S~)
  float* rowK;   // [R] (dot + gu*C) / T^2                  dc gradient, row part
  float* disc;   // [R] dbf discount
  float* pd;     // [SPB*V*V] pairwise conflict 0.5*sum_c|p_i-p_j|
  float* q;      // [SPB*V*V] q_ij = sum_c sign(p_ic-p_jc)*alpha_ic
  float* f;      // [SPB*C] fused evidence tile
  float* t;      // [SPB*C] aleatoric terms
  float* Sf;     // [SPB]
  float* psiSf;  // [SPB]
  int* y;        // [SPB]
};

__host__ __device__ inline size_t edl_smem_floats(int SPB, int V, int C) {
  const size_t E = (size_t)SPB * V * C, R = (size_t)SPB * V;
  return 2 * E + 8 * R + 2 * (size_t)SPB * V * V + 2 * (size_t)SPB * C + 3 * (size_t)SPB + 16;
}

__device__ __forceinline__ float f4_get(const float4& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }
__device__ __forceinline__ void f4_set(float4& v, int k, float x) {
  if (k == 0) v.x = x; else if (k == 1) v.y = x; else if (k == 2) v.z = x; else v.w = x;
}

// Persistent CTAs: tiles of SPB samples are dealt round-robin.  Each thread owns up to 3 float4 groups of
// the tile (loaded with one 128-bit coalesced load each, kept in registers through all phases and stored
// back as the gradient with one 128-bit store each); shared memory only carries what other threads need
// (the evidence tile for row sums / fusion, per-row statistics, the projected probabilities for the
// degree-of-conflict term).  mC / mV are 2^32/C and 2^32/V magic multipliers (exact for idx < 65536).
__global__ void __launch_bounds__(kEdlThreads, 4)
edl_fused_kernel(const float* __restrict__ evid, const long long* __restrict__ labels, dmf_edl_params prm,
                 int SPB, int ntiles, unsigned mC, unsigned mV, float lgammaC, const float* __restrict__ gscale_ptr,
                 float* __restrict__ fused_out, float* __restrict__ grad_out, float* __restrict__ u_out,
                 float* __restrict__ ale_out, int* __restrict__ pred_out, float* __restrict__ loss_parts) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[32];
  const int V = prm.V, C = prm.C, B = prm.B;
  const int tid = threadIdx.x;
  const int VC = V * C;
  const float fC = (float)C;

  EdlSmem sm;
  {
    float* p = smem;
    const size_t Ecap = ((size_t)SPB * VC + 3) & ~(size_t)3, Rcap = (size_t)SPB * V;
    sm.e = p; p += Ecap;
    sm.p = p; p += Ecap;
    sm.S = p; p += Rcap;
    sm.iT = p; p += Rcap;
    sm.om = p; p += Rcap;
    sm.gy = p; p += Rcap;
    sm.gA = p; p += Rcap;
    sm.psiT = p; p += Rcap;
    sm.rowK = p; p += Rcap;
    sm.disc = p; p += Rcap;
    sm.pd = p; p += (size_t)SPB * V * V;
    sm.q = p; p += (size_t)SPB * V * V;
    sm.f = p; p += (size_t)SPB * C;
    sm.t = p; p += (size_t)SPB * C;
    sm.Sf = p; p += SPB;
    sm.psiSf = p; p += SPB;
    sm.y = reinterpret_cast<int*>(p);
  }

  const float coef = prm.coef;
  const bool need_kl = coef != 0.f;
  const bool need_dc = prm.dc_weight != 0.f;
  const bool need_fused = fused_out || u_out || ale_out || pred_out;
  const bool need_pd = need_dc || (prm.agg == DMF_AGG_DBF && need_fused);
  const bool need_loss = grad_out || loss_parts;
  const float w_edl = prm.inv_B_global / ((float)V * (float)V);
  const float inv_vm1 = 1.0f / (float)max(1, V - 1);
  const float w_dc = prm.dc_weight * prm.inv_B_global * inv_vm1;
  const float gs = (grad_out && gscale_ptr) ? __ldg(gscale_ptr) : 1.0f;
  float acc_edl = 0.f, acc_kl = 0.f, acc_dc = 0.f;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long b0 = (long long)tile * SPB;
    const int nS = (int)min((long long)SPB, (long long)B - b0);
    const int E = nS * VC, R = nS * V;
    const float* src = evid + b0 * VC;
    const bool vec_in = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);

    // ---- P0: tile -> registers (+ smem copy for the row / fusion phases)
    float4 ev[kEdlEPT4];
#pragma unroll
    for (int k = 0; k < kEdlEPT4; ++k) {
      const int i0 = (tid + k * kEdlThreads) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i0 + 3 < E && vec_in) {
        v = __ldg(reinterpret_cast<const float4*>(src + i0));
      } else {
        if (i0 + 0 < E) v.x = __ldg(src + i0 + 0);
        if (i0 + 1 < E) v.y = __ldg(src + i0 + 1);
        if (i0 + 2 < E) v.z = __ldg(src + i0 + 2);
        if (i0 + 3 < E) v.w = __ldg(src + i0 + 3);
      }
      ev[k] = v;
      if (i0 < E) *reinterpret_cast<float4*>(sm.e + i0) = v;     // Ecap is padded to a multiple of 4
    }
    for (int i = tid; i < nS; i += kEdlThreads) {
      long long y = labels[b0 + i];
      sm.y[i] = (int)min(max(y, 0LL), (long long)(C - 1));
    }
    __syncthreads();

    // ---- P1: one thread per (sample, view) row: strength, psi / psi1 of S, alpha_y, S~
    if (need_loss || need_pd) {
      for (int r = tid; r < R; r += kEdlThreads) {
        const int b = (int)__umulhi((unsigned)r, mV);
        const float* er = sm.e + r * C;
        float S = 0.f;
        for (int c = 0; c < C; ++c) S += er[c] + 1.0f;
        sm.S[r] = S;
        const float iT = 1.0f / (S + 1e-8f);
        sm.iT[r] = iT;
        sm.om[r] = 1.0f - fC * iT;
        if (need_loss) {
          const float ay = er[sm.y[b]] + 1.0f;
          const Gamma3 gS = gamma3_fast<false>(S);
          const Gamma3 gyv = gamma3_fast<false>(ay);
          acc_edl += gS.psi - gyv.psi;
          sm.gy[r] = gS.psi1 - gyv.psi1;
          float gA = gS.psi1;
          if (need_kl) {
            const float St = S - ay + 1.0f;
            const Gamma3 gt = gamma3_fast<true>(St);
            acc_kl += gt.lgam - lgammaC;
            sm.psiT[r] = gt.psi;
            gA -= coef * (St - fC) * gt.psi1;
          }
          sm.gA[r] = gA;
        }
      }
      __syncthreads();
    }

    // ---- P2: one thread per element: class-wise KL terms + EDL gradient (registers), projected probabilities
    float4 gv[kEdlEPT4];
#pragma unroll
    for (int k = 0; k < kEdlEPT4; ++k) {
      gv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int i0 = (tid + k * kEdlThreads) * 4;
      if (i0 >= E) continue;
      float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const int idx = i0 + h;
        if (idx >= E) continue;
        const int r = (int)__umulhi((unsigned)idx, mC);
        const int c = idx - r * C;
        const int b = (int)__umulhi((unsigned)r, mV);
        const float al = f4_get(ev[k], h) + 1.0f;
        if (need_loss) {
          float g;
          if (c == sm.y[b]) {
            g = sm.gy[r];
          } else if (need_kl) {
            const float am1 = al - 1.0f;
            const Gamma3 ga = gamma3_fast<true>(al);
            acc_kl += am1 * (ga.psi - sm.psiT[r]) - ga.lgam;
            g = fmaf(coef * am1, ga.psi1, sm.gA[r]);
          } else {
            g = sm.gA[r];
          }
          f4_set(gv[k], h, g * w_edl);
        }
        if (need_pd) f4_set(pv, h, al * sm.iT[r]);
      }
      if (need_pd) *reinterpret_cast<float4*>(sm.p + i0) = pv;
    }

    // ---- P3: degree-of-conflict term (models/losses.py:161-187) and its gradient
    if (need_pd) {
      __syncthreads();
      // one thread per (sample, unordered view pair): pd_ij and the sign-weighted sums q_ij, q_ji
      const int npair = V * (V - 1) / 2;
      for (int pi = tid; pi < nS * npair; pi += kEdlThreads) {
        const int b = pi / npair;
        int rem = pi - b * npair, i = 0;
        while (rem >= V - 1 - i) { rem -= V - 1 - i; ++i; }
        const int j = i + 1 + rem;
        const float* pi_ = sm.p + (b * V + i) * C;
        const float* pj_ = sm.p + (b * V + j) * C;
        const float* ei_ = sm.e + (b * V + i) * C;
        const float* ej_ = sm.e + (b * V + j) * C;
        float pd = 0.f, qij = 0.f, qji = 0.f;
        for (int c = 0; c < C; ++c) {
          const float d = pi_[c] - pj_[c];
          pd += fabsf(d);
          const float sg = d > 0.f ? 1.0f : (d < 0.f ? -1.0f : 0.f);
          qij = fmaf(sg, ei_[c] + 1.0f, qij);
          qji = fmaf(-sg, ej_[c] + 1.0f, qji);
        }
        pd *= 0.5f;
        sm.pd[(b * V + i) * V + j] = pd;
        sm.pd[(b * V + j) * V + i] = pd;
        sm.q[(b * V + i) * V + j] = qij;
        sm.q[(b * V + j) * V + i] = qji;
      }
      for (int d = tid; d < nS * V; d += kEdlThreads) { sm.pd[d * V + (d - (int)__umulhi((unsigned)d, mV) * V)] = 0.f; }
      __syncthreads();
    }
    if (need_dc) {
      for (int r = tid; r < R; r += kEdlThreads) {
        const int b = (int)__umulhi((unsigned)r, mV), i = r - b * V;
        float gu = 0.f, dcs = 0.f, dot = 0.f;
        const float omi = sm.om[r];
        for (int j = 0; j < V; ++j) {
          if (j == i) continue;
          const float pd = sm.pd[r * V + j];
          const float omj = sm.om[b * V + j];
          gu += pd * omj;
          dcs += pd * (omi * omj);
          dot = fmaf(omj, sm.q[r * V + j], dot);
        }
        dot *= omi;
        acc_dc += dcs * inv_vm1;
        const float iT = sm.iT[r];
        sm.rowK[r] = (dot - 2.0f * gu * fC) * iT * iT;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kEdlEPT4; ++k) {
        const int i0 = (tid + k * kEdlThreads) * 4;
        if (i0 >= E) continue;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int idx = i0 + h;
          if (idx >= E) continue;
          const int r = (int)__umulhi((unsigned)idx, mC);
          const int c = idx - r * C;
          const int b = (int)__umulhi((unsigned)r, mV), kk = r - b * V;
          const float pk = sm.p[idx];
          float gp = 0.f;
          for (int j = 0; j < V; ++j) {
            if (j == kk) continue;
            const int rj = b * V + j;
            const float d = pk - sm.p[rj * C + c];
            const float sg = d > 0.f ? 1.0f : (d < 0.f ? -1.0f : 0.f);
            gp = fmaf(sg, sm.om[rj], gp);
          }
          gp *= sm.om[r];
          f4_set(gv[k], h, f4_get(gv[k], h) + w_dc * (gp * sm.iT[r] - sm.rowK[r]));
        }
      }
    }

    // ---- P4: fused evidence + summaries
    if (need_fused) {
      if (prm.agg == DMF_AGG_DBF) {  // utils.py:88-116
        for (int r = tid; r < R; r += kEdlThreads) {
          const int b = (int)__umulhi((unsigned)r, mV);
          const float ui = fC / sm.S[r];
          float agree = 1.0f;
          for (int j = 0; j < V; ++j) {
            const float uj = fC / sm.S[b * V + j];
            const float dc = sm.pd[r * V + j] * ((1.0f - ui) * (1.0f - uj));
            agree *= powf(1.0f - dc * dc * dc, 0.33333334f);
          }
          sm.disc[r] = agree;
        }
        __syncthreads();
      }
      for (int i = tid; i < nS * C; i += kEdlThreads) {
        const int b = (int)__umulhi((unsigned)i, mC), c = i - b * C;
        const float* eb = sm.e + (size_t)b * VC + c;
        float f;
        switch (prm.agg) {
          case DMF_AGG_CML: {
            f = 0.f;
            for (int v = 0; v < V; ++v) f += eb[v * C];
          } break;
          case DMF_AGG_AVG: {
            f = 0.f;
            for (int v = 0; v < V; ++v) f += eb[v * C];
            f = f / (float)V;
          } break;
          case DMF_AGG_JOINT: {
            float d = 0.f;
            for (int v = 1; v < V; ++v) d += eb[v * C];
            f = 0.5f * eb[0] + 0.5f * d;
          } break;
          case DMF_AGG_DISENTANGLED: {
            f = 0.f;
            for (int v = 1; v < V; ++v) f += eb[v * C];
          } break;
          default: {  // DBF
            f = 0.f;
            for (int v = 0; v < V; ++v) {
              const int r = b * V + v;
              const float S = sm.S[r], d = sm.disc[r];
              const float bel = (eb[v * C] / S) * d;
              const float unc = (fC / S) * d + 1.0f - d;
              f += fC * bel / (unc + 1e-6f);
            }
            f = f / (float)V;
          } break;
        }
        sm.f[i] = f;
        if (fused_out) fused_out[b0 * C + i] = f;
      }
      if (u_out || ale_out || pred_out) {
        __syncthreads();
        for (int b = tid; b < nS; b += kEdlThreads) {
          const float* fb = sm.f + b * C;
          float Sf = 0.f, best = fb[0];
          int arg = 0;
          for (int c = 0; c < C; ++c) {
            Sf += fb[c] + 1.0f;
            if (fb[c] > best) { best = fb[c]; arg = c; }
          }
          sm.Sf[b] = Sf;
          if (u_out) u_out[b0 + b] = fC / Sf;
          if (ale_out) sm.psiSf[b] = gamma3<false>(Sf + 1.0f).psi;
          if (pred_out) {
            int* po = pred_out + (b0 + b) * (V + 1);
            for (int v = 0; v < V; ++v) {
              const float* evp = sm.e + (b * V + v) * C;
              float bv = evp[0];
              int av = 0;
              for (int c = 1; c < C; ++c)
                if (evp[c] > bv) { bv = evp[c]; av = c; }
              po[v] = av;
            }
            po[V] = arg;
          }
        }
        if (ale_out) {
          __syncthreads();
          for (int i = tid; i < nS * C; i += kEdlThreads) {
            const int b = (int)__umulhi((unsigned)i, mC);
            const float al = sm.f[i] + 1.0f;
            sm.t[i] = (al / sm.Sf[b]) * (gamma3<false>(al + 1.0f).psi - sm.psiSf[b]);
          }
          __syncthreads();
          for (int b = tid; b < nS; b += kEdlThreads) {
            float a = 0.f;
            for (int c = 0; c < C; ++c) a += sm.t[b * C + c];
            ale_out[b0 + b] = -a;
          }
        }
      }
    }

    // ---- P5: gradient straight from registers (one 128-bit store per group)
    if (grad_out) {
      float* dst = grad_out + b0 * VC;
      const bool vec_out = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
#pragma unroll
      for (int k = 0; k < kEdlEPT4; ++k) {
        const int i0 = (tid + k * kEdlThreads) * 4;
        if (i0 >= E) continue;
        float4 v = gv[k];
        v.x *= gs; v.y *= gs; v.z *= gs; v.w *= gs;
        if (i0 + 3 < E && vec_out) {
          *reinterpret_cast<float4*>(dst + i0) = v;
        } else {
          if (i0 + 0 < E) dst[i0 + 0] = v.x;
          if (i0 + 1 < E) dst[i0 + 1] = v.y;
          if (i0 + 2 < E) dst[i0 + 2] = v.z;
          if (i0 + 3 < E) dst[i0 + 3] = v.w;
        }
      }
    }
    __syncthreads();     // shared tiles are reused by the next tile of this CTA
  }

  if (loss_parts) {
    const float s_edl = block_sum(acc_edl, red);
    const float s_kl = block_sum(acc_kl, red);
    const float s_dc = block_sum(acc_dc, red);
    if (tid == 0) {
      const float l0 = s_edl * w_edl, l1 = coef * s_kl * w_edl, l2 = prm.dc_weight * s_dc * prm.inv_B_global;
      atomicAdd(loss_parts + 0, l0);
      atomicAdd(loss_parts + 1, l1);
      atomicAdd(loss_parts + 2, l2);
      atomicAdd(loss_parts + 3, l0 + l1 + l2);
    }
  }
}

__global__ void evidence_fwd_kernel(const float* __restrict__ h, float* __restrict__ e, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    e[i] = evidence_act(h[i]);
}
__global__ void evidence_bwd_kernel(const float* __restrict__ h, const float* __restrict__ e,
                                    const float* __restrict__ de, float* __restrict__ dh, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dh[i] = de[i] * evidence_act_grad(h[i], e[i]);
}

}  // namespace dmf

using namespace dmf;

extern "C" int dmf_edl_fused(const float* evid, const long long* labels, const dmf_edl_params* p, const float* gscale,
                             float* fused, float* grad, float* u, float* ale, int* pred, float* loss_parts,
                             dmf_stream_t s) {
  DMF_REQUIRE(evid && labels && p, "dmf_edl_fused: null argument");
  DMF_REQUIRE(p->B >= 0 && p->V >= 1 && p->C >= 2, "dmf_edl_fused: bad shape B=%d V=%d C=%d", p->B, p->V, p->C);
  DMF_REQUIRE(p->agg >= DMF_AGG_CML && p->agg <= DMF_AGG_DBF, "dmf_edl_fused: unknown aggregation %d", p->agg);
  if (p->B == 0) return 0;
  const int VC = p->V * p->C;
  DMF_REQUIRE(VC * 4 <= kEdlMaxTile, "dmf_edl_fused: V*C=%d too large (max %d)", VC, kEdlMaxTile / 4);
  int SPB = (kEdlMaxTile / VC) & ~3;
  if (SPB > 256) SPB = 256;
  const size_t smem = edl_smem_floats(SPB, p->V, p->C) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(edl_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return fail((int)e, "dmf_edl_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  DMF_REQUIRE(smem <= 100 * 1024, "dmf_edl_fused: tile needs %zu bytes of shared memory", smem);
  const long long tiles = ((long long)p->B + SPB - 1) / SPB;
  DMF_REQUIRE(tiles < (1LL << 31), "dmf_edl_fused: batch too large");
  // persistent CTAs: as many as can be resident (shared-memory bound), in whole multiples of the SM count
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm > 4) per_sm = 4;          // register-bound residency (64 regs x 256 threads)
  if (per_sm < 1) per_sm = 1;
  const long long cap = (long long)kNumSMs * per_sm;
  const unsigned blocks = (unsigned)(tiles < cap ? tiles : cap);
  const unsigned mC = (unsigned)((1ULL << 32) / (unsigned)p->C) + 1u;
  const unsigned mV = (unsigned)((1ULL << 32) / (unsigned)p->V) + 1u;
  edl_fused_kernel<<<blocks, kEdlThreads, smem, (cudaStream_t)s>>>(evid, labels, *p, SPB, (int)tiles, mC, mV,
                                                                   lgammaf((float)p->C), gscale, fused, grad, u, ale, pred,
                                                                   loss_parts);
  return launched("dmf_edl_fused");
}

extern "C" int dmf_evidence_fwd(const float* h, float* e, long long n, dmf_stream_t s) {
  if (n <= 0) return 0;
  const int blocks = (int)min((n + 255) / 256, (long long)kNumSMs * 16);
  evidence_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)s>>>(h, e, n);
  return launched("dmf_evidence_fwd");
}
extern "C" int dmf_evidence_bwd(const float* h, const float* e, const float* de, float* dh, long long n, dmf_stream_t s) {
  if (n <= 0) return 0;
  const int blocks = (int)min((n + 255) / 256, (long long)kNumSMs * 16);
  evidence_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)s>>>(h, e, de, dh, n);
  return launched("dmf_evidence_bwd");
}
