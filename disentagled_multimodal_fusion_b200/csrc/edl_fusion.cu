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
// Memory-bound design (v3, "one thread per Dirichlet"): a thread owns one (sample, view) row and walks its C
// classes; the V rows of a sample sit in adjacent lanes of one warp and advance in LOCKSTEP over the class
// index, so everything that couples views (fused evidence, the pairwise conflict |p_i - p_j|, its sign
// pattern) is exchanged with warp shuffles -- no block barrier inside the math, no index division, no
// per-row arrays in shared memory.  Shared memory only carries the tile itself: a CTA copies its samples
// in with 128-bit coalesced loads, the row threads overwrite each evidence value by its gradient IN PLACE,
// and the tile (plus the fused-evidence tile) goes back out with 128-bit coalesced stores.  CTAs are
// persistent (tiles dealt round-robin) and 4 fit per SM.  Algorithmic bytes per sample: 8VC + 4C + 16.
#include <type_traits>
#include "common.cuh"
#include "special_math.cuh"
#include "tc_common.cuh"

namespace dmf {

// CTA shape, swept on B200 at B = 2^22, V = 4, C = 42 (tools/edl_sweep.sh; GB/s without / with the conflict term):
//   256 x 4: 3785 / 3059    256 x 3: 3880 / 2934    128 x 8: 3942 / 3135    128 x 6: 4150 / 3132
// Small CTAs desynchronise the load -> compute -> store phases of the tiles sharing an SM.
#ifndef EDL_THREADS
#define EDL_THREADS 128
#endif
#ifndef EDL_MINB
#define EDL_MINB 6
#endif
constexpr int kEdlThreads = EDL_THREADS;
constexpr int kEdlMinBlocks = EDL_MINB;          // resident CTAs per SM the register budget is sized for
constexpr int kEdlWarps = kEdlThreads / 32;
constexpr int kEdlMaxV = 8;
constexpr int kEdlTileFloats = 1344 * kEdlWarps;   // evidence tile per CTA (42 KB at 8 warps: 64 samples at V*C = 168)

// sign(d) * x for x >= 0 (alpha and 1-u are non-negative): copysign + a zero guard (3 instructions, no int->float)
__device__ __forceinline__ float sgn_mul(float d, float x) {
  const float t = __int_as_float((__float_as_int(x) & 0x7fffffff) | (__float_as_int(d) & (int)0x80000000));
  return d == 0.f ? 0.f : t;
}

// 1D bulk copies (TMA engine, no registers, no per-thread address arithmetic): global -> shared completes on an
// mbarrier, shared -> global is tracked by the issuing thread's bulk group.  Both need 16-byte aligned addresses
// and sizes; the callers fall back to cooperative loops otherwise.
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tc::smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(gdst)),
               "r"(tc::smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ bool bulk_ok(const void* g, long long bytes) {
  return ((reinterpret_cast<uintptr_t>(g) | (uintptr_t)bytes) & 15) == 0 && bytes > 0;
}

// MODE specialises the hot loops at compile time (the per-iteration uniform branches on runtime flags cost
// ~10 BRA + dead selects per element, ncu source page of v3):
//   0 = generic (any flag combination, eval outputs, DBF)      1 = train: loss + KL + conflict term, sum-type fusion
//   2 = train: loss + KL, no conflict term, sum-type fusion
//   3 = eval: forward only (fused evidence, u, aleatoric, argmax), sum-type fusion
enum { EDL_GENERIC = 0, EDL_TRAIN_DC = 1, EDL_TRAIN = 2, EDL_EVAL = 3 };

// Forward-only pass of one tile (models/evidential_probe.py:139-143,177-181, analysis.py:27-34): fused evidence of the
// sum-type rules, u = C / S, aleatoric = -sum_c (a_c/S)(psi(a_c + 1) - psi(S + 1)), per-view and fused argmax.
// Thread = one (sample, view) row as in the training path; the V lanes of a sample split the class pairs for the
// fused evidence and for the digamma terms (packed fp32x2), every lane scans its own row for the per-view argmax and
// the fused row for S / the fused argmax in class order (same summation order as the generic path: the epistemic
// ranking stays bit-exact).
template <int V, bool CE>
__device__ __forceinline__ void edl_eval_tile(const float* te, float* tf, const dmf_edl_params& prm, int C, int bs, int v, int base,
                                              bool active, long long b0, float* __restrict__ u_out,
                                              float* __restrict__ ale_out, int* __restrict__ pred_out) {
  const int VC = V * C;
  const float fC = (float)C;
  const float* sb = te + (size_t)(active ? bs : 0) * VC;
  const float* row = sb + (active ? v : 0) * C;
  float* fo = tf + (size_t)(active ? bs : 0) * C;
  // (1) fused evidence, views summed in reference order
  if (CE) {
    auto fuse_loop = [&](auto tag) {
      constexpr int AGG = decltype(tag)::value;
      for (int c = 2 * v; c < C; c += 2 * V) {
        const f2 t0 = *reinterpret_cast<const f2*>(sb + c);
        f2 sall = t0, d1 = splat2(0.f);
#pragma unroll
        for (int vv = 1; vv < V; ++vv) {
          const f2 tv = *reinterpret_cast<const f2*>(sb + vv * C + c);
          if (AGG == DMF_AGG_CML || AGG == DMF_AGG_AVG) sall = add2(sall, tv);
          if (AGG == DMF_AGG_JOINT || AGG == DMF_AGG_DISENTANGLED) d1 = vv == 1 ? tv : add2(d1, tv);
        }
        f2 f;
        if (AGG == DMF_AGG_CML) {
          f = sall;
        } else if (AGG == DMF_AGG_AVG) {
          float a0, a1;
          unpk2(sall, a0, a1);
          f = pk2(a0 / (float)V, a1 / (float)V);
        } else if (AGG == DMF_AGG_JOINT) {
          f = fma2(splat2(0.5f), t0, mul2(splat2(0.5f), d1));
        } else {
          f = d1;
        }
        if (active) *reinterpret_cast<f2*>(fo + c) = f;
      }
    };
    switch (prm.agg) {
      case DMF_AGG_CML: fuse_loop(std::integral_constant<int, DMF_AGG_CML>{}); break;
      case DMF_AGG_AVG: fuse_loop(std::integral_constant<int, DMF_AGG_AVG>{}); break;
      case DMF_AGG_JOINT: fuse_loop(std::integral_constant<int, DMF_AGG_JOINT>{}); break;
      default: fuse_loop(std::integral_constant<int, DMF_AGG_DISENTANGLED>{}); break;
    }
  } else if (active) {
    for (int c = v; c < C; c += V) {
      float t0 = sb[c], sall = t0, d1 = 0.f;
#pragma unroll
      for (int vv = 1; vv < V; ++vv) { const float tv = sb[vv * C + c]; sall += tv; d1 += tv; }
      float f;
      switch (prm.agg) {
        case DMF_AGG_CML: f = sall; break;
        case DMF_AGG_AVG: f = sall / (float)V; break;
        case DMF_AGG_JOINT: f = 0.5f * t0 + 0.5f * d1; break;
        default: f = d1; break;
      }
      fo[c] = f;
    }
  }
  __syncwarp();
  // (2) per-view argmax (first maximum wins), (3) S and argmax of the fused row in class order
  float bestv = row[0], best = fo[0], Sf = fo[0] + 1.0f;
  int argv = 0, arg = 0;
  if (pred_out) {
    for (int c = 1; c < C; ++c) {
      const float e = row[c];
      if (e > bestv) { bestv = e; argv = c; }
    }
  }
  for (int c = 1; c < C; ++c) {
    const float f = fo[c];
    Sf += f + 1.0f;
    if (f > best) { best = f; arg = c; }
  }
  if (active) {
    if (u_out && v == 0) u_out[b0 + bs] = fC / Sf;
    if (pred_out) {
      int* po = pred_out + (b0 + bs) * (V + 1);
      po[v] = argv;
      if (v == 0) po[V] = arg;
    }
  }
  // (4) aleatoric: the sample's lanes split the class pairs; psi in packed fp32x2
  if (ale_out) {
    const float iSf = 1.0f / Sf;
    float psiSf, dummy;
    unpk2(gamma_psi2(pk2(Sf + 1.0f, Sf + 1.0f)), psiSf, dummy);
    float a = 0.f;
    if (CE) {
      f2 a2 = splat2(0.f);
      const f2 npsi = splat2(-psiSf), is2 = splat2(iSf);
      for (int c = 2 * v; c < C; c += 2 * V) {
        const f2 al = add2(*reinterpret_cast<const f2*>(fo + c), splat2(1.0f));
        const f2 ps = gamma_psi2(add2(al, splat2(1.0f)));
        a2 = fma2(mul2(al, is2), add2(ps, npsi), a2);
      }
      float a0, a1;
      unpk2(a2, a0, a1);
      a = a0 + a1;
    } else {
      for (int c = v; c < C; c += V) {
        const float al = fo[c] + 1.0f;
        float ps;
        unpk2(gamma_psi2(pk2(al + 1.0f, al + 1.0f)), ps, dummy);
        a += (al * iSf) * (ps - psiSf);
      }
    }
    float tot = 0.f;
#pragma unroll
    for (int vv = 0; vv < V; ++vv) tot += __shfl_sync(0xffffffffu, a, base + vv);
    if (active && v == 0) ale_out[b0 + bs] = -tot;
  }
}

template <int VT, int MODE, bool CE>      // CE: C is even (64-bit shared-memory accesses on class pairs)
__global__ void __launch_bounds__(kEdlThreads, kEdlMinBlocks)
edl_fused_kernel(const float* __restrict__ evid, const long long* __restrict__ labels, dmf_edl_params prm,
                 int spw, int ntiles, float lgammaC, const float* __restrict__ gscale_ptr,
                 float* __restrict__ fused_out, float* __restrict__ grad_out, float* __restrict__ u_out,
                 float* __restrict__ ale_out, int* __restrict__ pred_out, float* __restrict__ loss_parts) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[32];
  __shared__ __align__(8) uint64_t ldbar;          // completion of the tile's bulk load
  constexpr int V = VT;
  const int C = prm.C, B = prm.B;
  const int VC = V * C;
  const int SPB = spw * kEdlWarps;                 // samples per tile
  float* te = smem;                                // [SPB*V*C] evidence -> gradient (in place)
  float* tf = te + (((size_t)SPB * VC + 3) & ~(size_t)3);   // [SPB*C] fused evidence
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sl = lane / V, v = lane - sl * V;      // sample slot within the warp, view
  const int base = sl * V;                         // first lane of this sample
  const bool slot_ok = sl < spw;
  const float fC = (float)C;

  const float coef = prm.coef;
  const bool need_kl = MODE != EDL_GENERIC ? true : coef != 0.f;
  const bool need_dc = MODE == EDL_TRAIN_DC ? true : (MODE == EDL_TRAIN ? false : prm.dc_weight != 0.f);
  const bool need_fused = MODE != EDL_GENERIC ? fused_out != nullptr : (fused_out || u_out || ale_out || pred_out);
  const bool dbf = MODE != EDL_GENERIC ? false : prm.agg == DMF_AGG_DBF;
  const bool need_pd = need_dc || (dbf && need_fused);
  const bool need_loss = MODE != EDL_GENERIC ? true : (grad_out || loss_parts);
  const float w_edl = prm.inv_B_global / ((float)V * (float)V);
  const float inv_vm1 = 1.0f / (float)(V > 1 ? V - 1 : 1);
  const float w_dc = prm.dc_weight * prm.inv_B_global * inv_vm1;
  const float gs = (grad_out && gscale_ptr) ? __ldg(gscale_ptr) : 1.0f;
  float acc_edl = 0.f, acc_kl = 0.f, acc_dc = 0.f;
  if (tid == 0) {
    tc::mbar_init(&ldbar, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  uint32_t ldphase = 0;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long b0 = (long long)tile * SPB;
    const int nS = (int)min((long long)SPB, (long long)B - b0);
    const int E = nS * VC;
    // ---- copy in: one bulk copy (TMA) per tile; cooperative 128-bit loads when the tile is not 16-byte aligned
    {
      const float* src = evid + b0 * VC;
      if (bulk_ok(src, (long long)E * 4)) {
        if (tid == 0) {
          bulk_wait_read();                        // the bulk stores of the previous tile have read the buffers
          tc::mbar_expect_tx(&ldbar, (uint32_t)E * 4u);
          bulk_load(te, src, (uint32_t)E * 4u, &ldbar);
        }
        tc::mbar_wait(&ldbar, ldphase);
        ldphase ^= 1u;
      } else {
        if (tid == 0) bulk_wait_read();
        __syncthreads();
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
          const int n4 = E >> 2;
          for (int i = tid; i < n4; i += kEdlThreads)
            reinterpret_cast<float4*>(te)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
          for (int i = (n4 << 2) + tid; i < E; i += kEdlThreads) te[i] = __ldg(src + i);
        } else {
          for (int i = tid; i < E; i += kEdlThreads) te[i] = __ldg(src + i);
        }
        __syncthreads();
      }
    }

    const int bs = warp * spw + sl;                // sample slot within the tile
    const bool active = slot_ok && bs < nS;
    // inactive lanes run the same instruction stream on a dummy row (row 0 of the tile) so that the
    // full-mask shuffles stay convergent; they never store
    float* row = te + (size_t)(active ? bs * V + v : 0) * C;
    if (MODE == EDL_EVAL) {
      edl_eval_tile<V, CE>(te, tf, prm, C, bs, v, base, active, b0, u_out, ale_out, pred_out);
    } else {
    int y = 0;
    if (active) {
      const long long yl = labels[b0 + bs];
      y = (int)min(max(yl, 0LL), (long long)(C - 1));
    }

    // ---- row statistics
    float S = 0.f;
    float gy = 0.f, gA = 0.f, psiT = 0.f, hy = 0.f;
    if (MODE != EDL_GENERIC) {
      // training path.  Row sum S: when the fused evidence is wanted (and C is even) ONE loop does both -- the V
      // lanes of a sample split the class pairs, each lane sums its pairs over the views in reference order (fused
      // evidence) and keeps a partial row sum per view; V shuffles per lane then complete the row sums.
      if (CE && fused_out != nullptr) {
        f2 part[V];
#pragma unroll
        for (int vv = 0; vv < V; ++vv) part[vv] = splat2(0.f);
        const float* sb = te + (size_t)(active ? bs : 0) * VC;
        float* fo = tf + (size_t)(active ? bs : 0) * C;
        auto fuse_loop = [&](auto tag) {
          constexpr int AGG = decltype(tag)::value;
          for (int c = 2 * v; c < C; c += 2 * V) {
            const f2 t0 = *reinterpret_cast<const f2*>(sb + c);
            part[0] = add2(part[0], t0);
            f2 sall = t0, d1 = splat2(0.f);
#pragma unroll
            for (int vv = 1; vv < V; ++vv) {
              const f2 tv = *reinterpret_cast<const f2*>(sb + vv * C + c);
              part[vv] = add2(part[vv], tv);
              if (AGG == DMF_AGG_CML || AGG == DMF_AGG_AVG) sall = add2(sall, tv);
              if (AGG == DMF_AGG_JOINT || AGG == DMF_AGG_DISENTANGLED) d1 = vv == 1 ? tv : add2(d1, tv);
            }
            f2 f;
            if (AGG == DMF_AGG_CML) {
              f = sall;
            } else if (AGG == DMF_AGG_AVG) {
              float a0, a1;
              unpk2(sall, a0, a1);
              f = pk2(a0 / (float)V, a1 / (float)V);
            } else if (AGG == DMF_AGG_JOINT) {
              f = fma2(splat2(0.5f), t0, mul2(splat2(0.5f), d1));
            } else {
              f = d1;
            }
            if (active) *reinterpret_cast<f2*>(fo + c) = f;
          }
        };
        switch (prm.agg) {
          case DMF_AGG_CML: fuse_loop(std::integral_constant<int, DMF_AGG_CML>{}); break;
          case DMF_AGG_AVG: fuse_loop(std::integral_constant<int, DMF_AGG_AVG>{}); break;
          case DMF_AGG_JOINT: fuse_loop(std::integral_constant<int, DMF_AGG_JOINT>{}); break;
          default: fuse_loop(std::integral_constant<int, DMF_AGG_DISENTANGLED>{}); break;
        }
        // lane v needs the sum over the sample's lanes of part[v]: in round r it receives from lane (v + r) % V the
        // partial that lane holds for view v, i.e. the source lane gives part[(own view - r) mod V]
        float ps[V];
#pragma unroll
        for (int vv = 0; vv < V; ++vv) {
          float lo, hi;
          unpk2(part[vv], lo, hi);
          ps[vv] = lo + hi;
        }
        float mine = ps[0];
#pragma unroll
        for (int vv = 1; vv < V; ++vv) mine = (v == vv) ? ps[vv] : mine;      // r = 0: own partial
#pragma unroll
        for (int r = 1; r < V; ++r) {
          int recv = v - r;
          if (recv < 0) recv += V;
          float give = ps[0];
#pragma unroll
          for (int vv = 1; vv < V; ++vv) give = (recv == vv) ? ps[vv] : give;
          int src = v + r;
          if (src >= V) src -= V;
          mine += __shfl_sync(0xffffffffu, give, base + src);
        }
        S = mine + fC;
      } else {
        f2 s2 = splat2(0.f);
        for (int c = 0; c + 1 < C; c += 2) {
          const f2 e2 = CE ? *reinterpret_cast<const f2*>(row + c) : pk2(row[c], row[c + 1]);
          s2 = add2(s2, e2);
        }
        float s_lo, s_hi;
        unpk2(s2, s_lo, s_hi);
        S = (s_lo + s_hi) + ((C & 1) ? row[C - 1] : 0.f) + fC;
      }
      const float ay = row[y] + 1.0f;
      const float St = S - ay + 1.0f;
      const Gamma3x2 ga = gamma3x2(pk2(S, St));
      const Gamma3x2 gb = gamma3x2(pk2(ay, ay));
      float psiS, psi1S, psi1T, lgT, psiY, psi1Y, lgY, dummy;
      unpk2(ga.psi, psiS, psiT);
      unpk2(ga.psi1, psi1S, psi1T);
      unpk2(ga.lgam, dummy, lgT);
      unpk2(gb.psi, psiY, dummy);
      unpk2(gb.psi1, psi1Y, dummy);
      unpk2(gb.lgam, lgY, dummy);
      const float am1y = ay - 1.0f;
      hy = am1y * psi1Y;
      gy = psi1S - psi1Y;
      gA = psi1S - coef * (St - fC) * psi1T;
      if (active) {
        acc_edl += psiS - psiY;
        // KL = lgamma(S~) - lgamma(C) - psi(S~) (S~ - C) + sum_{c != y} k(alpha_c),  k(x) = (x-1) psi(x) - lgamma(x)
        acc_kl += (lgT - lgammaC) - psiT * (St - fC) - (am1y * psiY - lgY);
      }
    } else {
      for (int c = 0; c < C; ++c) S += row[c] + 1.0f;
    }
    const float iT = 1.0f / (S + 1e-8f);
    const float om = 1.0f - fC * iT;
    if (MODE == EDL_GENERIC && need_loss) {
      const float ay = row[y] + 1.0f;
      const Gamma3 gS = gamma3_fast<false>(S);
      const Gamma3 gyv = gamma3_fast<false>(ay);
      if (active) acc_edl += gS.psi - gyv.psi;
      gy = gS.psi1 - gyv.psi1;
      gA = gS.psi1;
      if (need_kl) {
        const float St = S - ay + 1.0f;
        const Gamma3 gt = gamma3_fast<true>(St);
        if (active) acc_kl += gt.lgam - lgammaC;
        psiT = gt.psi;
        gA -= coef * (St - fC) * gt.psi1;
      }
    }
    // statistics of the other views of this sample (lane base + (v + jj) % V)
    float omj[V > 1 ? V - 1 : 1], iTj[V > 1 ? V - 1 : 1], Sj[V > 1 ? V - 1 : 1];
    int srcl[V > 1 ? V - 1 : 1];
    // power-of-two V in the training modes: partner jj = lane ^ (jj + 1) -- butterfly shuffles with immediate
    // offsets, no index registers (with 64 registers per thread ptxas rematerialised the rotation indices inside
    // the class loop: 16 of 112 instructions per class pair)
    constexpr bool kXor = MODE != EDL_GENERIC && (V & (V - 1)) == 0;
#pragma unroll
    for (int jj = 0; jj < V - 1; ++jj) {
      int vj = v + jj + 1;
      if (vj >= V) vj -= V;
      srcl[jj] = base + vj;
      if (kXor) {
        omj[jj] = __shfl_xor_sync(0xffffffffu, om, jj + 1);
        iTj[jj] = __shfl_xor_sync(0xffffffffu, iT, jj + 1);
        Sj[jj] = 0.f;
      } else {
        omj[jj] = __shfl_sync(0xffffffffu, om, srcl[jj]);
        iTj[jj] = __shfl_sync(0xffffffffu, iT, srcl[jj]);
        Sj[jj] = __shfl_sync(0xffffffffu, S, srcl[jj]);
      }
    }

    if (MODE != EDL_GENERIC) {
      // ================= training fast path (compile-time flags) =================
      // fused evidence of the sum-type rules first (the evidence tile is overwritten by the gradient below):
      // the V lanes of a sample split the classes, each value summed over the views in reference order
      if (!CE && fused_out != nullptr && active) {
        const float* sb = te + (size_t)bs * VC;
        for (int c = v; c < C; c += V) {
          float t0 = sb[c], sall = t0, d1 = 0.f;
#pragma unroll
          for (int vv = 1; vv < V; ++vv) { const float tv = sb[vv * C + c]; sall += tv; d1 += tv; }
          float f;
          switch (prm.agg) {
            case DMF_AGG_CML: f = sall; break;
            case DMF_AGG_AVG: f = sall / (float)V; break;
            case DMF_AGG_JOINT: f = 0.5f * t0 + 0.5f * d1; break;
            default: f = d1; break;   // DISENTANGLED
          }
          tf[bs * C + c] = f;
        }
      }
      __syncwarp();
      // ONE pass over the classes, two at a time in packed fp32x2: (-k, h) of the pair, KL sum, EDL gradient and
      // (MODE 1) the conflict sums.  Conflict term: with t_j = sign(p - p_j) (1 - u_j),
      //   gp_c = sum_j t_j,   sum_j (1-u_j) sum_c |p - p_j| = sum_c sum_j d_j t_j,   sum_j (1-u_j) q_j = sum_c alpha_c gp_c
      // so no per-pair accumulators are needed.  The part of the gradient that needs the completed sums (kk) and
      // the label class are patched in a light second pass.
      constexpr bool kDC = MODE == EDL_TRAIN_DC && V > 1;
      const float wg = w_edl * gs, wd = w_dc * gs * om * iT;
      const f2 one2 = splat2(1.0f), cw2 = splat2(coef * wg), ga2 = splat2(gA * wg), wd2 = splat2(wd), iT2 = splat2(iT);
      f2 niTj2[V > 1 ? V - 1 : 1];
#pragma unroll
      for (int jj = 0; jj < V - 1; ++jj) niTj2[jj] = splat2(-iTj[jj]);
      f2 acck = splat2(0.f), gu2 = splat2(0.f), dot2 = splat2(0.f);
      auto pair_step = [&](f2 e2, bool both) -> f2 {
        const f2 x2 = add2(e2, one2);
        const KH2 r = gamma_kh2(x2);
        if (both) {
          acck = add2(acck, r.kn);
        } else {
          float k0, k1;
          unpk2(r.kn, k0, k1);
          acck = add2(acck, pk2(k0, 0.f));
        }
        f2 g2 = fma2(r.h, cw2, ga2);
        if (kDC) {
          float al0, al1;
          unpk2(x2, al0, al1);
          const f2 p2 = mul2(x2, iT2);
          float gp0 = 0.f, gp1 = 0.f;
#pragma unroll
          for (int jj = 0; jj < V - 1; ++jj) {
            const float alj0 = kXor ? __shfl_xor_sync(0xffffffffu, al0, jj + 1) : __shfl_sync(0xffffffffu, al0, srcl[jj]);
            const float alj1 = kXor ? __shfl_xor_sync(0xffffffffu, al1, jj + 1) : __shfl_sync(0xffffffffu, al1, srcl[jj]);
            const f2 d2 = fma2(pk2(alj0, alj1), niTj2[jj], p2);
            float d0, d1;
            unpk2(d2, d0, d1);
            const int omb = __float_as_int(omj[jj]) & 0x7fffffff;
            const float t0 = __int_as_float(omb | (__float_as_int(d0) & (int)0x80000000));
            const float t1 = __int_as_float(omb | (__float_as_int(d1) & (int)0x80000000));
            gu2 = fma2(d2, pk2(t0, both ? t1 : 0.f), gu2);
            if (d0 != 0.f) gp0 += t0;
            if (d1 != 0.f) gp1 += t1;
          }
          if (!both) gp1 = 0.f;
          const f2 gp2 = pk2(gp0, gp1);
          dot2 = fma2(x2, gp2, dot2);
          g2 = fma2(wd2, gp2, g2);
        }
        return g2;
      };
      if (kDC) {      // two pairs in flight amortise the loop overhead; without the conflict term ptxas does worse unrolled
  #pragma unroll 2
        for (int c = 0; c + 1 < C; c += 2) {
          const f2 e2 = CE ? *reinterpret_cast<const f2*>(row + c) : pk2(row[c], row[c + 1]);
          const f2 g2 = pair_step(e2, true);
          if (active) {
            if (CE) {
              *reinterpret_cast<f2*>(row + c) = g2;
            } else {
              float g0, g1;
              unpk2(g2, g0, g1);
              row[c] = g0;
              row[c + 1] = g1;
            }
          }
        }
      } else {
  #pragma unroll 1
        for (int c = 0; c + 1 < C; c += 2) {
          const f2 e2 = CE ? *reinterpret_cast<const f2*>(row + c) : pk2(row[c], row[c + 1]);
          const f2 g2 = pair_step(e2, true);
          if (active) {
            if (CE) {
              *reinterpret_cast<f2*>(row + c) = g2;
            } else {
              float g0, g1;
              unpk2(g2, g0, g1);
              row[c] = g0;
              row[c + 1] = g1;
            }
          }
        }
      }
      if (!CE) {      // odd C: the last class rides alone (the upper lane runs on a dummy alpha = 1 and is discarded)
        const f2 g2 = pair_step(pk2(row[C - 1], 0.f), false);
        float g0, g1;
        unpk2(g2, g0, g1);
        if (active) row[C - 1] = g0;
      }
      {
        float k0, k1;
        unpk2(acck, k0, k1);
        if (active) acc_kl -= k0 + k1;         // acck holds -k
      }
      float kk = 0.f;
      if (kDC) {
        float a0, a1, b0, b1;
        unpk2(gu2, a0, a1);
        unpk2(dot2, b0, b1);
        const float gus = a0 + a1;             // sum_j (1-u_j) sum_c |p - p_j|
        const float dot = om * (b0 + b1);
        if (active) acc_dc += 0.5f * gus * om * inv_vm1;
        kk = w_dc * gs * (dot - gus * fC) * iT * iT;
      }
      if (active) {
        // label class: its EDL gradient is gy, not gA + coef h  (the conflict part is the same for every class)
        row[y] += (gy - gA - coef * hy) * wg;
        if (kDC) {
          if (CE) {
            const f2 nkk = splat2(-kk);
#pragma unroll 4
            for (int c = 0; c < C; c += 2) *reinterpret_cast<f2*>(row + c) = add2(*reinterpret_cast<const f2*>(row + c), nkk);
          } else {
#pragma unroll 4
            for (int c = 0; c < C; ++c) row[c] -= kk;
          }
        }
      }
    } else {
    // ---- loop A: pairwise conflict sums (models/losses.py:161-187): pd_kj = 0.5 sum_c |p_k - p_j|,
    //      q_kj = sum_c sign(p_kc - p_jc) alpha_kc
    float pd[V > 1 ? V - 1 : 1], rowK = 0.f, disc = 1.0f;
    if (need_pd && V > 1) {
      float q[V - 1 > 0 ? V - 1 : 1];
#pragma unroll
      for (int jj = 0; jj < V - 1; ++jj) { pd[jj] = 0.f; q[jj] = 0.f; }
#pragma unroll 2
      for (int c = 0; c < C; ++c) {
        const float al = row[c] + 1.0f;
        const float p = al * iT;
#pragma unroll
        for (int jj = 0; jj < V - 1; ++jj) {
          const float alj = __shfl_sync(0xffffffffu, al, srcl[jj]);
          const float d = p - alj * iTj[jj];
          pd[jj] += fabsf(d);
          q[jj] += sgn_mul(d, al);
        }
      }
      float gu = 0.f, dcs = 0.f, dot = 0.f;
#pragma unroll
      for (int jj = 0; jj < V - 1; ++jj) {
        pd[jj] *= 0.5f;
        gu += pd[jj] * omj[jj];
        dcs += pd[jj] * (om * omj[jj]);
        dot = fmaf(omj[jj], q[jj], dot);
      }
      dot *= om;
      if (active && need_dc) acc_dc += dcs * inv_vm1;
      rowK = (dot - 2.0f * gu * fC) * iT * iT;
      if (dbf && need_fused) {  // utils.py:88-116: discount of this view
        const float ui = fC / S;
        // multiply in absolute view order j = 0..V-1 like the reference (bit-exact uncertainty rankings);
        // the j == i factor is exactly 1 (pd_ii = 0)
        float agree = 1.0f;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          float pdj = 0.f, Sjj = S;
#pragma unroll
          for (int jj = 0; jj < V - 1; ++jj)
            if (srcl[jj] - base == j) { pdj = pd[jj]; Sjj = Sj[jj]; }
          const float uj = fC / Sjj;
          const float dc = pdj * ((1.0f - ui) * (1.0f - uj));
          agree *= powf(1.0f - dc * dc * dc, 0.33333334f);
        }
        disc = agree;
      }
    }

    // ---- loop B: class-wise KL terms, gradient (written in place), fused evidence
    float Sf = 0.f, best = 0.f, bestv = 0.f;
    int arg = 0, argv = 0;
    const float uS = fC / S;
#pragma unroll 2
    for (int c = 0; c < C; ++c) {
      const float e = row[c];
      const float al = e + 1.0f;
      if (need_loss) {
        float g;
        if (need_kl) {
          const float am1 = al - 1.0f;
          const Gamma3 ga = gamma3_fast<true>(al);
          const bool isy = c == y;
          if (active && !isy) acc_kl += am1 * (ga.psi - psiT) - ga.lgam;
          g = isy ? gy : fmaf(coef * am1, ga.psi1, gA);
        } else {
          g = (c == y) ? gy : gA;
        }
        g *= w_edl;
        if (need_dc && V > 1) {
          const float p = al * iT;
          float gp = 0.f;
#pragma unroll
          for (int jj = 0; jj < V - 1; ++jj) {
            const float alj = __shfl_sync(0xffffffffu, al, srcl[jj]);
            const float d = p - alj * iTj[jj];
            gp += sgn_mul(d, omj[jj]);
          }
          g += w_dc * (gp * om * iT - rowK);
        }
        if (active && grad_out) row[c] = g * gs;
      }
      if (need_fused) {
        // every lane of the sample forms the same fused value, summing the views in reference order
        float term = e;
        if (dbf) {
          const float bel = (e / S) * disc;
          const float unc = uS * disc + 1.0f - disc;
          term = fC * bel / (unc + 1e-6f);
        }
        // sall = ((t0 + t1) + t2) + ...   d1 = (t1 + t2) + ...   (the reference's summation orders)
        float t0 = 0.f, sall = 0.f, d1 = 0.f;
#pragma unroll
        for (int vv = 0; vv < V; ++vv) {
          const float tv = __shfl_sync(0xffffffffu, term, base + vv);
          if (vv == 0) { t0 = tv; sall = tv; } else { sall += tv; d1 += tv; }
        }
        float f;
        switch (prm.agg) {
          case DMF_AGG_CML: f = sall; break;
          case DMF_AGG_AVG: f = sall / (float)V; break;
          case DMF_AGG_JOINT: f = 0.5f * t0 + 0.5f * d1; break;
          case DMF_AGG_DISENTANGLED: f = d1; break;
          default: f = sall / (float)V; break;
        }
        if (active && v == 0) tf[bs * C + c] = f;
        if (MODE == EDL_GENERIC) {
          Sf += f + 1.0f;
          if (c == 0) { best = f; bestv = e; }
          if (f > best) { best = f; arg = c; }
          if (e > bestv) { bestv = e; argv = c; }
        }
      }
    }
    if (MODE == EDL_GENERIC && need_fused && active) {
      if (u_out && v == 0) u_out[b0 + bs] = fC / Sf;
      if (pred_out) {
        int* po = pred_out + (b0 + bs) * (V + 1);
        po[v] = argv;
        if (v == 0) po[V] = arg;
      }
    }
    if (MODE == EDL_GENERIC && ale_out) {
      // aleatoric = -sum_c (a_c/Sf) (psi(a_c + 1) - psi(Sf + 1)); the sample's lanes split the classes
      __syncwarp();
      const float psiSf = gamma3<false>(Sf + 1.0f).psi;
      float a = 0.f;
      if (active)
        for (int c = v; c < C; c += V) {
          const float al = tf[bs * C + c] + 1.0f;
          a += (al / Sf) * (gamma3<false>(al + 1.0f).psi - psiSf);
        }
      float tot = 0.f;
#pragma unroll
      for (int vv = 0; vv < V; ++vv) tot += __shfl_sync(0xffffffffu, a, base + vv);
      if (active && v == 0) ale_out[b0 + bs] = -tot;
    }
    }   // generic path
    }   // not EDL_EVAL

    // ---- copy out: gradient tile and fused-evidence tile as bulk stores (cooperative stores when misaligned)
    {
      float* gdst = grad_out ? grad_out + b0 * VC : nullptr;
      float* fdst = fused_out ? fused_out + b0 * C : nullptr;
      const int nf = nS * C;
      const bool gb = gdst && bulk_ok(gdst, (long long)E * 4), fb = fdst && bulk_ok(fdst, (long long)nf * 4);
      if (gb || fb) tc::fence_proxy_async_smem();     // this thread's generic-proxy writes -> visible to the bulk engine
      __syncthreads();
      if (tid == 0 && (gb || fb)) {
        if (gb) bulk_store(gdst, te, (uint32_t)E * 4u);
        if (fb) bulk_store(fdst, tf, (uint32_t)nf * 4u);
        bulk_commit();
      }
      if (gdst && !gb) {
        if ((reinterpret_cast<uintptr_t>(gdst) & 15) == 0) {
          const int n4 = E >> 2;
          for (int i = tid; i < n4; i += kEdlThreads) reinterpret_cast<float4*>(gdst)[i] = reinterpret_cast<const float4*>(te)[i];
          for (int i = (n4 << 2) + tid; i < E; i += kEdlThreads) gdst[i] = te[i];
        } else {
          for (int i = tid; i < E; i += kEdlThreads) gdst[i] = te[i];
        }
      }
      if (fdst && !fb) {
        if ((reinterpret_cast<uintptr_t>(fdst) & 15) == 0) {
          const int n4 = nf >> 2;
          for (int i = tid; i < n4; i += kEdlThreads) reinterpret_cast<float4*>(fdst)[i] = reinterpret_cast<const float4*>(tf)[i];
          for (int i = (n4 << 2) + tid; i < nf; i += kEdlThreads) fdst[i] = tf[i];
        } else {
          for (int i = tid; i < nf; i += kEdlThreads) fdst[i] = tf[i];
        }
      }
      if ((gdst && !gb) || (fdst && !fb)) __syncthreads();     // cooperative readers are done before the next tile lands
    }
  }
  if (tid == 0) bulk_wait_read();

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


// ------------------------------------------------------------------------------------------
// Fused evaluation reducer (analysis.py:5-399): one pass over a batch of evidences accumulates,
// per slot s (views 0..V-1, then the fused evidence as slot V), everything evaluate_subjective_model[_with_shared]
// gathers with ~40 torch kernels and ~12 .item() syncs per batch:
//   stats[s][0..7] = correct, evidence_sum, epi_sum, ale_sum, inc_N, inc_evidence_sum, inc_epi_sum, inc_ale_sum
//   class_sum[s][c] += sum_b e[b,s,c]        true_sum[s][c] += sum_{b: y_b = c} e[b,s,c]       class_counts[c] += #{y_b = c}
// One thread per (sample, slot); shared-memory accumulators per CTA, one round of global atomics per CTA.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
eval_reduce_kernel(const float* __restrict__ evid, const float* __restrict__ fused, const long long* __restrict__ labels,
                   int B, int V, int C, float* __restrict__ stats, float* __restrict__ class_sum,
                   float* __restrict__ true_sum, float* __restrict__ class_counts) {
  extern __shared__ float sm[];
  const int S1 = V + 1;
  float* s_stats = sm;                       // [S1][8]
  float* s_cls = s_stats + S1 * 8;           // [S1][C]
  float* s_true = s_cls + S1 * C;            // [S1][C]
  float* s_cnt = s_true + S1 * C;            // [C]
  const int nacc = S1 * 8 + 2 * S1 * C + C;
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const long long total = (long long)B * S1;
  const float fC = (float)C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / S1), s = (int)(idx - (long long)b * S1);
    const float* e = s < V ? evid + ((long long)b * V + s) * C : fused + (long long)b * C;
    long long yl = labels[b];
    const int y = (int)min(max(yl, 0LL), (long long)(C - 1));
    float esum = 0.f, S = 0.f, best = e[0];
    int arg = 0;
    for (int c = 0; c < C; ++c) {
      const float v = e[c];
      esum += v;
      S += v + 1.0f;
      if (v > best) { best = v; arg = c; }
      atomicAdd(&s_cls[s * C + c], v);
    }
    const float psiS = gamma3<false>(S + 1.0f).psi;
    float ale = 0.f;
    for (int c = 0; c < C; ++c) {
      const float al = e[c] + 1.0f;
      ale += (al / S) * (gamma3<false>(al + 1.0f).psi - psiS);
    }
    ale = -ale;
    const float epi = fC / S;
    const bool ok = arg == y;
    float* st = s_stats + s * 8;
    atomicAdd(st + 1, esum);
    atomicAdd(st + 2, epi);
    atomicAdd(st + 3, ale);
    if (ok) {
      atomicAdd(st + 0, 1.0f);
    } else {
      atomicAdd(st + 4, 1.0f);
      atomicAdd(st + 5, esum);
      atomicAdd(st + 6, epi);
      atomicAdd(st + 7, ale);
    }
    atomicAdd(&s_true[s * C + y], e[y]);
    if (s == V) atomicAdd(&s_cnt[y], 1.0f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S1 * 8; i += blockDim.x) if (s_stats[i] != 0.f) atomicAdd(stats + i, s_stats[i]);
  for (int i = threadIdx.x; i < S1 * C; i += blockDim.x) {
    if (s_cls[i] != 0.f) atomicAdd(class_sum + i, s_cls[i]);
    if (s_true[i] != 0.f) atomicAdd(true_sum + i, s_true[i]);
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) if (s_cnt[i] != 0.f) atomicAdd(class_counts + i, s_cnt[i]);
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

template <int VT, int MODE, bool CE>
static int launch_edl_mode(const float* evid, const long long* labels, const dmf_edl_params* p, const float* gscale, float* fused,
                      float* grad, float* u, float* ale, int* pred, float* loss_parts, cudaStream_t st) {
  const int VC = VT * p->C;
  int spw = 32 / VT;                                         // samples per warp: a sample never straddles warps
  const int cap = kEdlTileFloats / (kEdlWarps * VC);         // ... and the tile has to fit
  if (spw > cap) spw = cap;
  DMF_REQUIRE(spw >= 1, "dmf_edl_fused: V*C=%d too large (max %d)", VC, kEdlTileFloats / kEdlWarps);
  const int SPB = spw * kEdlWarps;
  const size_t smem = ((((size_t)SPB * VC + 3) & ~(size_t)3) + (size_t)SPB * p->C + 4) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(edl_fused_kernel<VT, MODE, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return fail((int)e, "dmf_edl_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  DMF_REQUIRE(smem <= 64 * 1024, "dmf_edl_fused: tile needs %zu bytes of shared memory", smem);
  const long long tiles = ((long long)p->B + SPB - 1) / SPB;
  DMF_REQUIRE(tiles < (1LL << 31), "dmf_edl_fused: batch too large");
  // persistent CTAs: whole multiples of the SM count, as many as stay resident (4 by registers, smem permitting)
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm > kEdlMinBlocks) per_sm = kEdlMinBlocks;
  if (per_sm < 1) per_sm = 1;
  const long long capb = (long long)kNumSMs * per_sm;
  const unsigned blocks = (unsigned)(tiles < capb ? tiles : capb);
  edl_fused_kernel<VT, MODE, CE><<<blocks, kEdlThreads, smem, st>>>(evid, labels, *p, spw, (int)tiles, lgammaf((float)p->C), gscale,
                                                          fused, grad, u, ale, pred, loss_parts);
  return launched("dmf_edl_fused");
}

template <int VT>
static int launch_edl(const float* evid, const long long* labels, const dmf_edl_params* p, const float* gscale, float* fused,
                      float* grad, float* u, float* ale, int* pred, float* loss_parts, cudaStream_t st) {
  // the training call (gradient and/or loss, annealed KL on, no eval outputs, sum-type fusion) gets loops
  // specialised at compile time; everything else runs the generic instantiation
  const bool train = (grad || loss_parts) && p->coef != 0.f && !u && !ale && !pred && p->agg != DMF_AGG_DBF;
  const bool ce = (p->C & 1) == 0;
  const bool eval = !grad && !loss_parts && fused && p->agg != DMF_AGG_DBF;
  if (eval)
    return ce ? launch_edl_mode<VT, EDL_EVAL, true>(evid, labels, p, gscale, fused, grad, u, ale, pred, loss_parts, st)
              : launch_edl_mode<VT, EDL_EVAL, false>(evid, labels, p, gscale, fused, grad, u, ale, pred, loss_parts, st);
  if (train && p->dc_weight != 0.f && VT > 1)
    return ce ? launch_edl_mode<VT, EDL_TRAIN_DC, true>(evid, labels, p, gscale, fused, grad, u, ale, pred, loss_parts, st)
              : launch_edl_mode<VT, EDL_TRAIN_DC, false>(evid, labels, p, gscale, fused, grad, u, ale, pred, loss_parts, st);
  if (train && (p->dc_weight == 0.f || VT == 1))
    return ce ? launch_edl_mode<VT, EDL_TRAIN, true>(evid, labels, p, gscale, fused, grad, u, ale, pred, loss_parts, st)
              : launch_edl_mode<VT, EDL_TRAIN, false>(evid, labels, p, gscale, fused, grad, u, ale, pred, loss_parts, st);
  return launch_edl_mode<VT, EDL_GENERIC, false>(evid, labels, p, gscale, fused, grad, u, ale, pred, loss_parts, st);
}

extern "C" int dmf_edl_fused(const float* evid, const long long* labels, const dmf_edl_params* p, const float* gscale,
                             float* fused, float* grad, float* u, float* ale, int* pred, float* loss_parts,
                             dmf_stream_t s) {
  DMF_REQUIRE(evid && labels && p, "dmf_edl_fused: null argument");
  DMF_REQUIRE(p->B >= 0 && p->V >= 1 && p->C >= 2, "dmf_edl_fused: bad shape B=%d V=%d C=%d", p->B, p->V, p->C);
  DMF_REQUIRE(p->V <= kEdlMaxV, "dmf_edl_fused: V=%d views; kernels are instantiated for V <= %d", p->V, kEdlMaxV);
  DMF_REQUIRE(p->agg >= DMF_AGG_CML && p->agg <= DMF_AGG_DBF, "dmf_edl_fused: unknown aggregation %d", p->agg);
  if (p->B == 0) return 0;
  cudaStream_t st = (cudaStream_t)s;
  switch (p->V) {
    case 1: return launch_edl<1>(evid, labels, p, gscale, fused, grad, u, ale, pred, loss_parts, st);
    case 2: return launch_edl<2>(evid, labels, p, gscale, fused, grad, u, ale, pred, loss_parts, st);
    case 3: return launch_edl<3>(evid, labels, p, gscale, fused, grad, u, ale, pred, loss_parts, st);
    case 4: return launch_edl<4>(evid, labels, p, gscale, fused, grad, u, ale, pred, loss_parts, st);
    case 5: return launch_edl<5>(evid, labels, p, gscale, fused, grad, u, ale, pred, loss_parts, st);
    case 6: return launch_edl<6>(evid, labels, p, gscale, fused, grad, u, ale, pred, loss_parts, st);
    case 7: return launch_edl<7>(evid, labels, p, gscale, fused, grad, u, ale, pred, loss_parts, st);
    default: return launch_edl<8>(evid, labels, p, gscale, fused, grad, u, ale, pred, loss_parts, st);
  }
}


extern "C" int dmf_eval_reduce(const float* evid, const float* fused, const long long* labels, int B, int V, int C,
                               float* stats, float* class_sum, float* true_sum, float* class_counts, dmf_stream_t s) {
  DMF_REQUIRE(evid && fused && labels && stats && class_sum && true_sum && class_counts, "dmf_eval_reduce: null argument");
  DMF_REQUIRE(B >= 0 && V >= 1 && C >= 2, "dmf_eval_reduce: bad shape B=%d V=%d C=%d", B, V, C);
  if (B == 0) return 0;
  const size_t smem = ((size_t)(V + 1) * 8 + 2 * (size_t)(V + 1) * C + C) * sizeof(float);
  DMF_REQUIRE(smem <= 48 * 1024, "dmf_eval_reduce: (V+1)*C = %d too large", (V + 1) * C);
  const long long total = (long long)B * (V + 1);
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
  eval_reduce_kernel<<<(unsigned)blocks, 256, smem, (cudaStream_t)s>>>(evid, fused, labels, B, V, C, stats, class_sum, true_sum,
                                                                    class_counts);
  return launched("dmf_eval_reduce");
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
