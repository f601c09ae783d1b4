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
constexpr int kEdlMaxTile = 3072;  // elements of evid per CTA

struct EdlSmem {
  float* e;      // [E] evidence tile
  float* g;      // [E] gradient tile
  float* gp;     // [E] dc sign sums
  float* S;      // [R]
  float* rowA;   // [R] psi1(S)
  float* rowB;   // [R] psi(S~)
  float* rowC;   // [R] (S~-C) psi1(S~)
  float* rowD;   // [R] psi1(alpha_y)
  float* om;     // [R] 1-u
  float* gu;     // [R]
  float* dot;    // [R]
  float* disc;   // [R] dbf discount
  float* pd;     // [nS*V*V]
  float* f;      // [SPB*C] fused evidence tile
  float* t;      // [SPB*C] aleatoric terms
  float* Sf;     // [SPB]
  float* psiSf;  // [SPB]
  int* y;        // [SPB]
};

__host__ __device__ inline size_t edl_smem_floats(int SPB, int V, int C) {
  const size_t E = (size_t)SPB * V * C, R = (size_t)SPB * V;
  return 3 * E + 9 * R + (size_t)SPB * V * V + 2 * (size_t)SPB * C + 3 * (size_t)SPB;
}

__global__ void __launch_bounds__(kEdlThreads)
edl_fused_kernel(const float* __restrict__ evid, const long long* __restrict__ labels, dmf_edl_params prm,
                 int SPB, float lgammaC, const float* __restrict__ gscale_ptr, float* __restrict__ fused_out,
                 float* __restrict__ grad_out, float* __restrict__ u_out, float* __restrict__ ale_out,
                 int* __restrict__ pred_out, float* __restrict__ loss_parts) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[32];
  const int V = prm.V, C = prm.C, B = prm.B;
  const int tid = threadIdx.x;
  const long long b0 = (long long)blockIdx.x * SPB;
  const int nS = (int)min((long long)SPB, (long long)B - b0);
  const int VC = V * C;
  const int E = nS * VC, R = nS * V;

  EdlSmem sm;
  {
    float* p = smem;
    const size_t Ecap = (size_t)SPB * VC, Rcap = (size_t)SPB * V;
    sm.e = p; p += Ecap;
    sm.g = p; p += Ecap;
    sm.gp = p; p += Ecap;
    sm.S = p; p += Rcap;
    sm.rowA = p; p += Rcap;
    sm.rowB = p; p += Rcap;
    sm.rowC = p; p += Rcap;
    sm.rowD = p; p += Rcap;
    sm.om = p; p += Rcap;
    sm.gu = p; p += Rcap;
    sm.dot = p; p += Rcap;
    sm.disc = p; p += Rcap;
    sm.pd = p; p += (size_t)SPB * V * V;
    sm.f = p; p += (size_t)SPB * C;
    sm.t = p; p += (size_t)SPB * C;
    sm.Sf = p; p += SPB;
    sm.psiSf = p; p += SPB;
    sm.y = reinterpret_cast<int*>(p);
  }

  // ---- P0: coalesced tile load (SPB % 4 == 0 and a 16B-aligned base => tile start is 16B aligned)
  const float* src = evid + b0 * VC;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  if (vec_ok) {
    const int n4 = E >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(sm.e);
    for (int i = tid; i < n4; i += kEdlThreads) d4[i] = __ldg(s4 + i);
    for (int i = (n4 << 2) + tid; i < E; i += kEdlThreads) sm.e[i] = __ldg(src + i);
  } else {
    for (int i = tid; i < E; i += kEdlThreads) sm.e[i] = __ldg(src + i);
  }
  for (int i = tid; i < nS; i += kEdlThreads) {
    long long y = labels[b0 + i];
    sm.y[i] = (int)min(max(y, 0LL), (long long)(C - 1));
  }
  __syncthreads();

  const float coef = prm.coef;
  const bool need_kl = coef != 0.f;
  const bool need_dc = prm.dc_weight != 0.f;
  const bool need_pd = need_dc || (prm.agg == DMF_AGG_DBF && (fused_out || u_out || ale_out || pred_out));
  float acc_edl = 0.f, acc_kl = 0.f, acc_dc = 0.f;
  const float fC = (float)C;

  // ---- P1: one thread per (sample, view) row
  for (int r = tid; r < R; r += kEdlThreads) {
    const int b = r / V;
    const float* er = sm.e + r * C;
    float S = 0.f;
    for (int c = 0; c < C; ++c) S += er[c] + 1.0f;
    const int y = sm.y[b];
    const float ay = er[y] + 1.0f;
    const float St = S - ay + 1.0f;
    const Gamma3 gS = gamma3<false>(S);
    const Gamma3 gy = gamma3<false>(ay);
    acc_edl += gS.psi - gy.psi;
    sm.rowA[r] = gS.psi1;
    sm.rowD[r] = gy.psi1;
    if (need_kl) {
      const Gamma3 gt = gamma3<true>(St);
      acc_kl += gt.lgam - lgammaC;
      sm.rowB[r] = gt.psi;
      sm.rowC[r] = (St - fC) * gt.psi1;
    }
    sm.S[r] = S;
    sm.om[r] = 1.0f - fC / (S + 1e-8f);
  }
  __syncthreads();

  // ---- P2: one thread per element: class-wise KL terms and the EDL gradient
  const float w_edl = prm.inv_B_global / ((float)V * (float)V);
  for (int idx = tid; idx < E; idx += kEdlThreads) {
    const int r = idx / C;
    const int c = idx - r * C;
    const int b = r / V;
    float g;
    if (c == sm.y[b]) {
      g = sm.rowA[r] - sm.rowD[r];
    } else if (need_kl) {
      const float al = sm.e[idx] + 1.0f;
      const float am1 = al - 1.0f;
      const Gamma3 ga = gamma3<true>(al);
      acc_kl += -ga.lgam + am1 * (ga.psi - sm.rowB[r]);
      g = sm.rowA[r] + coef * (am1 * ga.psi1 - sm.rowC[r]);
    } else {
      g = sm.rowA[r];
    }
    sm.g[idx] = g * w_edl;
  }

  // ---- P3: degree-of-conflict term (models/losses.py:161-187) and its gradient
  if (need_pd) {
    const int VV = V * V;
    for (int pi = tid; pi < nS * VV; pi += kEdlThreads) {
      const int b = pi / VV;
      const int ij = pi - b * VV;
      const int i = ij / V, j = ij - i * V;
      float pd = 0.f;
      if (i != j) {
        const float* ei = sm.e + (b * V + i) * C;
        const float* ej = sm.e + (b * V + j) * C;
        const float Ti = sm.S[b * V + i] + 1e-8f, Tj = sm.S[b * V + j] + 1e-8f;
        for (int c = 0; c < C; ++c) pd += fabsf((ei[c] + 1.0f) / Ti - (ej[c] + 1.0f) / Tj);
        pd *= 0.5f;
      }
      sm.pd[pi] = pd;
    }
  }
  __syncthreads();
  if (need_dc) {
    const float inv_vm1 = 1.0f / (float)max(1, V - 1);
    for (int r = tid; r < R; r += kEdlThreads) {
      const int b = r / V, i = r - b * V;
      float gu = 0.f, dcs = 0.f;
      for (int j = 0; j < V; ++j) {
        const float pd = sm.pd[(b * V + i) * V + j];
        const float omj = sm.om[b * V + j];
        gu += pd * omj;
        dcs += pd * (sm.om[r] * omj);
      }
      sm.gu[r] = -2.0f * gu;
      acc_dc += dcs * inv_vm1;
    }
    for (int idx = tid; idx < E; idx += kEdlThreads) {
      const int r = idx / C;
      const int c = idx - r * C;
      const int b = r / V, k = r - b * V;
      const float pk = (sm.e[idx] + 1.0f) / (sm.S[r] + 1e-8f);
      float gp = 0.f;
      for (int j = 0; j < V; ++j) {
        if (j == k) continue;
        const int rj = b * V + j;
        const float pj = (sm.e[rj * C + c] + 1.0f) / (sm.S[rj] + 1e-8f);
        const float d = pk - pj;
        const float sg = d > 0.f ? 1.0f : (d < 0.f ? -1.0f : 0.f);
        gp += sg * sm.om[rj];
      }
      sm.gp[idx] = gp * sm.om[r];
    }
    __syncthreads();
    for (int r = tid; r < R; r += kEdlThreads) {
      float dot = 0.f;
      for (int c = 0; c < C; ++c) dot += sm.gp[r * C + c] * (sm.e[r * C + c] + 1.0f);
      sm.dot[r] = dot;
    }
    __syncthreads();
    const float w_dc = prm.dc_weight * prm.inv_B_global * inv_vm1;
    for (int idx = tid; idx < E; idx += kEdlThreads) {
      const int r = idx / C;
      const float T = sm.S[r] + 1e-8f;
      const float iT = 1.0f / T;
      sm.g[idx] += w_dc * (sm.gp[idx] * iT - sm.dot[r] * iT * iT - sm.gu[r] * fC * iT * iT);
    }
  }

  // ---- P4: fused evidence + summaries
  const bool need_fused = fused_out || u_out || ale_out || pred_out;
  if (need_fused) {
    if (prm.agg == DMF_AGG_DBF) {  // utils.py:88-116
      for (int r = tid; r < R; r += kEdlThreads) {
        const int b = r / V, i = r - b * V;
        const float ui = fC / sm.S[r];
        float agree = 1.0f;
        for (int j = 0; j < V; ++j) {
          const float uj = fC / sm.S[b * V + j];
          const float dc = sm.pd[(b * V + i) * V + j] * ((1.0f - ui) * (1.0f - uj));
          agree *= powf(1.0f - dc * dc * dc, 0.33333334f);
        }
        sm.disc[r] = agree;
      }
      __syncthreads();
    }
    for (int i = tid; i < nS * C; i += kEdlThreads) {
      const int b = i / C, c = i - b * C;
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
          const float* ev = sm.e + (b * V + v) * C;
          float bv = ev[0];
          int av = 0;
          for (int c = 1; c < C; ++c)
            if (ev[c] > bv) { bv = ev[c]; av = c; }
          po[v] = av;
        }
        po[V] = arg;
      }
    }
    if (ale_out) {
      __syncthreads();
      for (int i = tid; i < nS * C; i += kEdlThreads) {
        const int b = i / C;
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
  __syncthreads();

  // ---- P5: gradient tile out (coalesced), loss partial sums
  if (grad_out) {
    const float gs = gscale_ptr ? __ldg(gscale_ptr) : 1.0f;
    float* dst = grad_out + b0 * VC;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      const int n4 = E >> 2;
      const float4* s4 = reinterpret_cast<const float4*>(sm.g);
      float4* d4 = reinterpret_cast<float4*>(dst);
      for (int i = tid; i < n4; i += kEdlThreads) {
        float4 v = s4[i];
        v.x *= gs; v.y *= gs; v.z *= gs; v.w *= gs;
        d4[i] = v;
      }
      for (int i = (n4 << 2) + tid; i < E; i += kEdlThreads) dst[i] = sm.g[i] * gs;
    } else {
      for (int i = tid; i < E; i += kEdlThreads) dst[i] = sm.g[i] * gs;
    }
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
  const long long blocks = ((long long)p->B + SPB - 1) / SPB;
  edl_fused_kernel<<<(unsigned)blocks, kEdlThreads, smem, (cudaStream_t)s>>>(
      evid, labels, *p, SPB, lgammaf((float)p->C), gscale, fused, grad, u, ale, pred, loss_parts);
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
