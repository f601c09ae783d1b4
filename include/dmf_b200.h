/*
 * dmf_b200.h -- C ABI of libdmf_b200.so: the B200 (sm_100a) kernels behind the training-step
 * hot path of Hassan-Sarwat/disentagled_multimodal_fusion.
 *
 * The reference is pure Python/PyTorch and has no FFI boundary of its own; each
 * entry point below replaces an ATen op *chain* inside one reference function, cited as
 * file:line into the reference tree.  The reference-side binding a maintainer would add (a
 * ctypes stub inside the reference module) is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns int: 0 = OK, <0 = argument/shape/alignment error, >0 = cudaError_t
 *     of the launch.  dmf_last_error() returns a thread-local message for the last failure.
 *   - the CALLER owns every buffer (device pointers unless stated "host"); the library never
 *     allocates device memory, never synchronises, and launches only on the given stream.
 *   - all matrices are row-major; "ld*" are leading dimensions in ELEMENTS.
 *   - fp32 unless a parameter is named *_bf16 (raw uint16 bfloat16 bits).
 *   - scalar results are ACCUMULATED (atomicAdd) into caller-zeroed float slots so that one
 *     device->host read per step suffices.
 */
#ifndef DMF_B200_H
#define DMF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dmf_stream_t; /* cudaStream_t */

/* ------------------------------------------------------------------ library / device guard */
int dmf_version(void);
/* copies the calling thread's last error text into buf (NUL terminated); returns its length */
int dmf_last_error(char* buf, size_t n);
/* 0 iff the current device is compute capability 10.x (B200); otherwise an error (no fallback) */
int dmf_device_check(void);
/* number of kernel launches issued by this library since load (thread-safe counter) */
long long dmf_launch_count(void);

/* ------------------------------------------------------------------ K1 grouped MLP layers
 * Replaces the per-view nn.Linear/ReLU chains of Linear.forward (models/classifiers.py:43-48) and
 * EvidentialNN.forward (:497-502), called N + N^2 times per DMVAE step (models/dmvae.py:139,154,162)
 * and 8x per DisentangledSSL step (models/disentangledssl.py:90-93,117-120): ONE launch per layer
 * over all groups (views x streams).                                                          */
enum {
  DMF_EPI_NONE = 0,      /* C = A*B                                   */
  DMF_EPI_BIAS = 1,      /* C = A*B + bias[n]                         */
  DMF_EPI_BIAS_RELU = 2, /* C = relu(A*B + bias[n])                   */
  DMF_EPI_RELU_MASK = 3, /* C = (aux[m,n] > 0) ? A*B : 0   (ReLU backward folded into dgrad) */
  DMF_EPI_BIAS_EVIDENCE = 4 /* C = evidence(A*B + bias), utils.py:46-63; aux (optional) gets the pre-activation */
};

/* One GEMM problem  C[M,N] = epi( sum_k A(m,k) * B(k,n) ).  Strides are in elements and fully
 * general, so forward (Y = X W^T), dgrad (dX = dY W) and wgrad (dW = dY^T X) are the same call. */
typedef struct {
  const void* A;  long long a_rs, a_cs;   /* A(m,k) = A[m*a_rs + k*a_cs] */
  const void* B;  long long b_rs, b_cs;   /* B(k,n) = B[k*b_rs + n*b_cs] */
  void* C;        long long ldc;          /* C[m*ldc + n], row-major      */
  const float* bias;                      /* [N] or NULL                  */
  void* aux;      long long ldaux;        /* epilogue-specific, may be NULL */
  float* rowsum_a;                        /* optional [M]: rowsum_a[m] = sum_k A(m,k)  (bias grad in wgrad form) */
  int M, N, K;
  int accumulate;                         /* 1: C += result (epilogue NONE only) */
} dmf_gemm_desc;

/* fp32 path (FFMA, 1e-5 parity path).  `groups` is a HOST array of n_groups descriptors
 * (n_groups <= 64 per call); it is passed by value to the kernel, no device copy needed.    */
int dmf_grouped_gemm_f32(const dmf_gemm_desc* groups, int n_groups, int epilogue, dmf_stream_t s);

/* bf16 tensor-core path: tcgen05.mma (kind::f16, fp32 accumulate in TMEM), operands staged by
 * TMA (128B swizzle).  Both operands K-major: A [M,K] bf16 (lda), B [N,K] bf16 (ldb):
 * C[M,N] = epi(A * B^T).  K*2 bytes and lda/ldb*2 bytes must be multiples of 16; pointers 16B
 * aligned.  out_f32 / out_bf16 / out_bf16_t may each be NULL (at least one non-NULL).
 * Groups with M >= 512 and N >= 128 run on the persistent CTA-pair kernel (cta_group::2, 256x256 tiles,
 * double-buffered TMEM accumulators); smaller ones on the single-CTA 128x128 kernel.                */
typedef struct {
  const uint16_t* A; long long lda;
  const uint16_t* B; long long ldb;
  float* out_f32;    long long ldo_f32;
  uint16_t* out_bf16; long long ldo_bf16;
  const float* bias;                 /* [N] or NULL */
  const uint16_t* mask_bf16; long long ldmask; /* DMF_EPI_RELU_MASK: zero where mask<=0 */
  int M, N, K;
  uint16_t* out_bf16_t; long long ldo_t;       /* optional TRANSPOSED bf16 copy: out_bf16_t[n*ldo_t + m] */
  int split_k;  /* 0 = auto, 1 = never, >1 = that many K splits, <0 = auto AND always accumulate (out_f32 += A*B^T:
                   wgrad straight into an existing .grad buffer).  Splitting applies only to DMF_EPI_NONE with
                   out_f32 alone (wgrad: K = batch); partial tiles are accumulated with red.add, so for
                   split_k >= 0 the CALLER must zero out_f32 first.  Honoured by the CTA-pair kernel only
                   (M >= 512 and N >= 128 in every group) */
  int mn_major; /* 1: BOTH operands are given MN-major -- A is [K, M] (lda = pitch of a k-row), B is [K, N]:
                   C[m,n] = sum_k A[k,m] * B[k,n].  This is wgrad read straight from the row-major activations
                   (dW[n_out, k_in] = sum_batch dY[batch, n_out] * X[batch, k_in]): no transposed copies.  M, N
                   multiples of 8.  CTA-pair kernel only (M >= 512 and N >= 128), else the call fails.          */
} dmf_tc_gemm_desc;
int dmf_grouped_gemm_bf16_tc(const dmf_tc_gemm_desc* groups, int n_groups, int epilogue, dmf_stream_t s);

/* The LAST encoder layer with its per-row head fused into the epilogue (bf16 operands, fp32 accumulate; one CTA pair
 * owns whole output rows, so N must be 256 or 512):  X = A W^T + b, then
 *   head 0: out = X / max(|X|_2, eps)               F.normalize, models/disentangledssl.py:139-140
 *   head 1: out = vMF reparameterised sample         models/classifiers.py:433-437 (Householder reflection of
 *           x = [w, sqrt(1 - w^2) v] from e1 onto X / |X|), noise (w [M], v [M, N-1]) from dmf_vmf_draw
 * Optional extra outputs: X itself (pre_f32 / pre_bf16: the vMF backward, the ortho term and the conditioning columns of
 * the private encoders need it) and inv_norm[M] = 1 / max(|X|, eps) (head 0; dmf_row_normalize_bwd takes it).      */
typedef struct {
  const void* A; long long lda;           /* [M, K] bf16 */
  const void* W; long long ldw;           /* [N, K] bf16 (nn.Linear layout) */
  const float* bias;                      /* [N] */
  float* pre_f32; long long ld_pre_f32;   /* X, may be NULL */
  uint16_t* pre_bf16; long long ld_pre_bf16;
  float* out_f32; long long ld_out_f32;   /* head output (at least one of out_f32 / out_bf16) */
  uint16_t* out_bf16; long long ld_out_bf16;
  float* inv_norm;                        /* [M], head 0, may be NULL */
  const float* noise_w;                   /* [M], head 1 */
  const float* noise_v;                   /* [M, N-1], head 1 */
  float eps;                              /* head 0 */
  int M, N, K;                            /* lda, ldw multiples of 8 elements; columns >= K read as zero */
} dmf_head_gemm_desc;
int dmf_head_gemm_bf16(const dmf_head_gemm_desc* groups, int n_groups, int head, dmf_stream_t s);

/* column sums  out[n] (+)= sum_m X[m*ld + n]   (bias gradients)  */
int dmf_colsum_f32(const float* X, long long ld, int M, int N, float* out, int accumulate, dmf_stream_t s);

/* dtype / layout movers for the bf16 path (memory-bound): dst[r*ldd + c] = bf16(src[r*lds + c]);
 * transpose variant writes dst[c*ldd + r].                                                    */
int dmf_cast_f32_to_bf16(const float* src, long long lds, uint16_t* dst, long long ldd, int rows, int cols, dmf_stream_t s);
int dmf_cast_transpose_f32_to_bf16(const float* src, long long lds, uint16_t* dst, long long ldd, int rows, int cols, dmf_stream_t s);
int dmf_transpose_bf16(const uint16_t* src, long long lds, uint16_t* dst, long long ldd, int rows, int cols, dmf_stream_t s);
/* one pass over fp32 src [rows, cols]: dst[r*ldd + c] (bf16, may be NULL), dst_t[c*ldt + r] (bf16 transposed,
 * may be NULL: the K-major operand of the next wgrad GEMM) and colsum[c] += sum_r src (may be NULL; caller
 * zeroes; the bias gradient).                                                                              */
int dmf_cast_dual_bf16(const float* src, long long lds, uint16_t* dst, long long ldd, uint16_t* dst_t, long long ldt,
                       float* colsum, int rows, int cols, dmf_stream_t s);
/* out[n] += sum_m X[m*ld + n] for a bf16 matrix (bias gradient from the bf16 dgrad output; caller zeroes) */
int dmf_colsum_bf16(const uint16_t* X, long long ld, int M, int N, float* out, dmf_stream_t s);

/* ------------------------------------------------------------------ K2 fused InfoNCE
 * Replaces matmul/div/max/sub/exp/sum/log/mean of SupConLoss.forward (models/losses.py:64-99)
 * without materialising the [2B,2B] logits.
 *
 * dmf_rowlse: for anchors A [Ma,D] against columns Bm [Nb,D]:  s_ij = scale * <a_i, b_j>,
 *   row_max[i] = max_j s_ij,  row_sum[i] = sum_j exp(s_ij - row_max[i]).
 *   Optional diag: if diag_offset >= 0, diag_out[i] = s_{i, diag_offset+i} (the positive / self term).
 * dtype 0: fp32 FFMA path; dtype 1: bf16 tcgen05 path (A,Bm are bf16, D % 64 == 0).           */
int dmf_rowlse(const void* A, long long lda, int Ma, const void* Bm, long long ldb, int Nb, int D,
               float scale, float* row_max, float* row_sum, long long diag_offset, float* diag_out,
               void* workspace, size_t workspace_bytes, int dtype, dmf_stream_t s);
/* scratch the bf16 path may use to split the columns across CTAs (0 is always accepted)     */
size_t dmf_rowlse_workspace_bytes(int Ma, int Nb);

/* Fixed-shift variant for L2-NORMALISED embeddings (|s| <= scale; DisentangledSSL feeds vMF samples and
 * F.normalize outputs, models/disentangledssl.py:134-140): one pass over the tiles of the [Ma x Nb] block gives
 *   row_sum[i] += sum_j exp(s_ij - shift)   and   col_sum[j] += sum_i exp(s_ij - shift)
 * (both ACCUMULATED; the caller zeroes them), so the column LSEs of S01 -- the row LSEs of the view-1 anchors --
 * come from the same tiles as the view-0 ones.  sym = 1: A is rows [row0_global, row0_global+Ma) of Bm (a
 * symmetric intra-view block): only the cyclic half window of 256-column tiles is visited and the full row sum
 * of global row g is row_sum[g-row0_global] + col_sum[g] (col_sum summed over ranks); needs row0_global % 256
 * == 0.  col_sum may be NULL when sym = 0.  diag_out[i] = s_{i, diag_offset+i} if diag_offset >= 0.
 * bf16 operands, D % 64 == 0, D <= 512.  Feed dmf_infonce_finalize with m_* = shift and l_* = these sums.  */
int dmf_infonce_rowcol_sums(const void* A, long long lda, int Ma, const void* Bm, long long ldb, int Nb, int D,
                            float scale, float shift, int sym, int row0_global, float* row_sum, float* col_sum,
                            long long diag_offset, float* diag_out, dmf_stream_t s);

/* The same pass that additionally KEEPS the probabilities (sym = 0 only): E, dmf_infonce_e_bytes(Ma, Nb) bytes,
 * receives e_ij = exp(s_ij - shift) as bf16 in blocks of [128 rows x 64 columns] (block (ib, jb) at byte offset
 * (ib * 4*ceil(Nb/256) + jb) * 16384, row-major inside; zeros outside [Ma x Nb]) -- the operand of
 * dmf_infonce_bwd_stored, which then needs no recomputation of S.  E = NULL: identical to the call above.   */
size_t dmf_infonce_e_bytes(int Ma, int Nb);
int dmf_infonce_rowcol_sums_store(const void* A, long long lda, int Ma, const void* Bm, long long ldb, int Nb, int D,
                                  float scale, float shift, int sym, int row0_global, float* row_sum, float* col_sum,
                                  long long diag_offset, float* diag_out, void* E, dmf_stream_t s);

/* Per-anchor finalisation of one SupConLoss call (models/losses.py:68-99) for BOTH anchor sets:
 *   m_full = max(m_cross, m_intra); sum_c = l_cross*exp(m_cross-m_full);
 *   loss_i = -(pos - m_full - log(sum_c + 1e-12));  lse_eff_i = m_full + log(sum_c + 1e-12)
 *   diag_i = -(self - m_intra - log(l_intra))              (no-grad loss_x / loss_y)
 * out3[0] += sum_i loss_i * inv_count ; out3[1+which] += sum_i diag_i * inv_count_diag.        */
int dmf_infonce_finalize(const float* m_cross, const float* l_cross, const float* m_intra, const float* l_intra,
                         const float* pos, const float* self, int n, float inv_count, float inv_count_diag,
                         int which, float* lse_eff, float* out3, dmf_stream_t s);

/* Backward for one anchor set (Appendix B of SURVEY, verified against reference autograd):
 *   dA[i,:] (+)= coef*g * ( sum_j [exp(s_ij - lseA_i) + exp(s_ij - lseB_j)] * b_j  - 2 * b_{pos(i)} )
 * with g read from the device scalar *gscale, pos(i) = diag_offset + i.                        */
int dmf_infonce_bwd(const void* A, long long lda, int Ma, const float* lseA,
                    const void* Bm, long long ldb, const void* BmT, long long ldbt, int Nb, const float* lseB, int D,
                    float scale, float coef, const float* gscale, long long diag_offset,
                    float* dA, long long ldda, int accumulate, int dtype, dmf_stream_t s);
/* BmT: the transposed column block [D, Nb] (dmf_transpose_bf16).  NULL for fp32, and for the bf16 path whenever
 * dmf_infonce_bwd_needs_transposed(D) returns 0 (D = 256 / 512: the kernel reads Bm as an MN-major operand). */
int dmf_infonce_bwd_needs_transposed(int D);

/* Backward of one gradient direction from the stored probabilities E of dmf_infonce_rowcol_sums_store (same Ma, Nb,
 * shift), replacing the autograd chain of models/losses.py:64-99 like dmf_infonce_bwd but with ONE product:
 *   dir 0: dOut[i,:] (+)= coef*g*( sum_j e_ij (fa_i + fb_j) Z[j,:] - 2 Z[i + diag_offset,:] ),  i < Ma, Z = columns [Nb, D]
 *   dir 1: dOut[j,:] (+)= coef*g*( sum_i e_ij (fa_i + fb_j) Z[i,:] - 2 Z[j - diag_offset,:] ),  j < Nb, Z = anchors [Ma, D]
 * fa_i = exp(shift - lseA[i]), fb_j = exp(shift - lseB[j]); Z bf16, D = 256 or 512; dOut fp32 (16-byte aligned, ldo % 4
 * == 0); work = scratch of dmf_infonce_bwd_stored_work_floats(Ma, Nb) floats (16-byte aligned).              */
size_t dmf_infonce_bwd_stored_work_floats(int Ma, int Nb);
int dmf_infonce_bwd_stored(const void* E, int Ma, int Nb, const float* lseA, const float* lseB, float shift,
                           const void* Z, long long ldz, int D, int dir, float coef, const float* gscale,
                           long long diag_offset, float* dOut, long long ldo, int accumulate, float* work, dmf_stream_t s);

/* ------------------------------------------------------------------ K4 ortho Gram pieces
 * ortho_loss (models/losses.py:104-110) = || normalize(z1)^T normalize(zs) ||_F.
 * Row normalisation fwd/bwd (also F.normalize at models/disentangledssl.py:139-140).           */
int dmf_row_normalize_fwd(const float* X, long long ldx, int rows, int D, float eps,
                          float* Y, long long ldy, uint16_t* Y_bf16, long long ldyb, float* inv_norm, dmf_stream_t s);
int dmf_row_normalize_bwd(const float* Y, long long ldy, const float* inv_norm, const float* dY, long long lddy,
                          int rows, int D, float* dX, long long lddx, int accumulate, dmf_stream_t s);
/* out[0] (+)= scale * sqrt(sum(G^2)) is NOT done here: returns sum of squares into out (atomicAdd) */
int dmf_sumsq_f32(const float* G, long long n, float* out, dmf_stream_t s);

/* ------------------------------------------------------------------ vMF reparameterised sample
 * ProbabilisticEncoder('vmf') + VonMisesFisher.rsample given explicit noise
 * (models/classifiers.py:314-335,433-437,456-466): loc=e/||e||, x=[w, sqrt(clamp(1-w^2,1e-10)) v],
 * u=(e1-loc)/(||e1-loc||+1e-5), z = x - 2<x,u>u.  noise_w [rows], noise_v [rows, D-1].          */
int dmf_vmf_fwd(const float* E, long long lde, const float* noise_w, const float* noise_v, int rows, int D,
                float* Z, long long ldz, uint16_t* Z_bf16, long long ldzb, dmf_stream_t s);
int dmf_vmf_bwd(const float* E, long long lde, const float* noise_w, const float* noise_v, const float* dZ, long long lddz,
                int rows, int D, float* dE, long long ldde, int accumulate, dmf_stream_t s);
/* Device-side draw of the vMF noise (distribution-equal to the reference sampler, not
 * stream-equal): Philox counter RNG, Wood's rejection sampler per row.           */
int dmf_vmf_draw(float* noise_w, float* noise_v, int rows, int D, float kappa, unsigned long long seed,
                 unsigned long long offset, dmf_stream_t s);

/* Device-side augment_data (utils.py:118-151): per row, with probability 1/3 each, Y = X + N(0,
 * noise_scale^2), Y = X with floor(D / drop_scale) distinct uniformly chosen columns zeroed, or Y = X.
 * Distribution-equal to the reference's host loop (numpy choice + torch.randn), not stream-equal.
 * choice_out (optional, [rows] int32) receives the per-row choice 0 / 1 / 2.                                  */
int dmf_augment(const float* X, long long ldx, float* Y, long long ldy, int rows, int D, float noise_scale,
                int drop_scale, unsigned long long seed, unsigned long long offset, int* choice_out, dmf_stream_t s);

/* The same two draws keyed additionally by a DEVICE-resident 64-bit counter (seed' = seed + *counter * odd constant),
 * so that a training step captured once into a CUDA graph draws fresh noise on every replay: capture
 * {dmf_augment_ctr, dmf_vmf_draw_ctr, ..., dmf_counter_add(counter, 1)}.  (The reference draws on the host:
 * utils.py:118-151, models/classifiers.py:314-431.)                                                             */
int dmf_vmf_draw_ctr(float* noise_w, float* noise_v, int rows, int D, float kappa, unsigned long long seed,
                     unsigned long long offset, const unsigned long long* counter, dmf_stream_t s);
int dmf_augment_ctr(const float* X, long long ldx, float* Y, long long ldy, int rows, int D, float noise_scale,
                    int drop_scale, unsigned long long seed, unsigned long long offset, int* choice_out,
                    const unsigned long long* counter, dmf_stream_t s);
int dmf_counter_add(unsigned long long* counter, unsigned long long inc, dmf_stream_t s);

/* ------------------------------------------------------------------ DMVAE head + objectives
 * Replaces chunk/exp/randn_like/PoE/KL of models/dmvae.py:74-112,142-150,170-172.
 * stats [N][B,4e] (mu_s, lv_s, mu_p, lv_p), noise [2N+1][B,e] in the reference draw order.
 * Writes decoder inputs dec_in[i] = [N*B, 2e]: block j of B rows = [z_p_i | (j==i ? z_s : z_s_uni_j)].
 * kl3[0..2] += (kl_private, kl_poe, kl_uni) already divided by B.                              */
int dmf_dmvae_head_fwd(const float* const* stats, const float* noise, int N, int B, int e, float poe_temperature,
                       float* const* dec_in, float* kl3, dmf_stream_t s);
/* backward: d_dec_in[i] [N*B,2e] and the upstream gradients of the three KL scalars
 * (kl_grad3 = d loss / d (kl_private, kl_poe, kl_uni), a DEVICE array of 3 floats; the module forms
 * loss = ... + a*(kl_private + N*kl_poe) + a*kl_uni, models/dmvae.py:174-176) -> d_stats[i] [B,4e] */
int dmf_dmvae_head_bwd(const float* const* stats, const float* noise, const float* const* d_dec_in,
                       int N, int B, int e, float poe_temperature, const float* kl_grad3,
                       float* const* d_stats, dmf_stream_t s);
/* get_embedding PoE of the means (models/dmvae.py:115-125): mu_poe [B,e]                        */
int dmf_dmvae_poe_mean(const float* const* stats, int N, int B, int e, float poe_temperature, float* mu_poe, dmf_stream_t s);
/* F.mse_loss x N^2 (models/dmvae.py:155,164) fused with its gradient: recon [N*B,d] vs x [B,d];
 * block j==view is the joint term (weight w_joint), others cross (w_cross).
 * out2[0] += w_joint*mse_joint, out2[1] += w_cross*sum_j mse_cross_j; d_recon = g*w*2*(r-x)/(B*d) */
int dmf_dmvae_mse_fwd_bwd(const float* recon, long long ldr, const float* x, long long ldx, int N, int B, int d,
                          int view, float w_joint, float w_cross, const float* gscale, float* out2,
                          float* d_recon, long long lddr, dmf_stream_t s);

/* ------------------------------------------------------------------ K3 evidence fusion + EDL
 * One pass over evid [B,V,C] replacing utils.py:66-116 (fusion rules), models/losses.py:117-248
 * (AvgTrustedLoss: EDL digamma loss + annealed KL + degree-of-conflict term) and the uncertainty
 * summaries of models/evidential_probe.py:139-143.                                             */
enum { DMF_AGG_CML = 0, DMF_AGG_AVG = 1, DMF_AGG_JOINT = 2, DMF_AGG_DISENTANGLED = 3, DMF_AGG_DBF = 4 };
typedef struct {
  int B, V, C;
  int agg;                 /* DMF_AGG_* */
  float coef;              /* min(1, annealing_step/annealing_start)            (losses.py:127-130) */
  float dc_weight;         /* gamma_t * fused                                   (losses.py:243-247) */
  float inv_B_global;      /* 1/B of the GLOBAL batch (data parallel: local sums scale by this) */
} dmf_edl_params;
/* outputs (each may be NULL): fused [B,C]; grad [B,V,C] = d loss/d evid * (*gscale, or 1 if NULL);
 * u [B] = C/S; ale [B]; pred [B,V+1] int32 (per-view argmax then fused argmax);
 * loss_parts[4] += {edl(A) term, coef*KL term, dc term, total loss}
 * The argument combination selects a compile-time specialised kernel: grad and/or loss_parts with coef != 0 and a
 * sum-type rule = training pass (packed fp32x2 math; dc_weight == 0 drops the conflict term); fused without grad /
 * loss_parts and a sum-type rule = forward-only evaluation pass; anything else (DBF, mixed outputs) = generic pass.
 * Tiles move as 1D bulk copies when evid / grad / fused tiles are 16-byte aligned (always for contiguous buffers with
 * V*C*samples-per-tile % 4 == 0), cooperative loads otherwise -- results are identical.                            */
int dmf_edl_fused(const float* evid, const long long* labels, const dmf_edl_params* p, const float* gscale,
                  float* fused, float* grad, float* u, float* ale, int* pred, float* loss_parts, dmf_stream_t s);

/* Fused evaluation reducer (analysis.py:5-399).  Slots s = 0..V-1 are the views of evid [B,V,C],
 * slot V is the fused evidence [B,C].  Everything is ACCUMULATED (caller zeroes once per evaluation):
 *   stats [V+1][8] = correct, evidence_sum, epi_sum (C/S), ale_sum, inc_N, inc_evidence_sum, inc_epi_sum, inc_ale_sum
 *   class_sum [V+1][C] = sum_b e;  true_sum [V+1][C] = sum_{b: y_b = c} e[b, c];  class_counts [C]            */
int dmf_eval_reduce(const float* evid, const float* fused, const long long* labels, int B, int V, int C,
                    float* stats, float* class_sum, float* true_sum, float* class_counts, dmf_stream_t s);

/* evidence activation alone (utils.py:46-63) and its backward (zero outside the clamp)          */
int dmf_evidence_fwd(const float* h, float* e, long long n, dmf_stream_t s);
int dmf_evidence_bwd(const float* h, const float* e, const float* de, float* dh, long long n, dmf_stream_t s);

/* ------------------------------------------------------------------ optimizer (a18)
 * Fused flat-buffer Adam / AdamW step (torch.optim semantics, models/dmvae.py:204-210,
 * models/evidential_probe.py:205-212).  decoupled=1 -> AdamW.  step is the 1-based step count.
 * Optionally refreshes a bf16 copy of the parameters for the tensor-core path.                  */
int dmf_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int decoupled, int step, float grad_scale,
                  uint16_t* p_bf16, dmf_stream_t s);
/* Same update with the step count and learning rate in DEVICE memory (state[0] = number of steps taken so far,
 * as a float; state[1] = learning rate), so that a CUDA graph holding the whole training step can be replayed:
 * the kernel uses step = state[0] + 1 and a trailing 1-thread kernel increments state[0].                 */
int dmf_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float* state, float beta1, float beta2,
                      float eps, float weight_decay, int decoupled, float grad_scale, uint16_t* p_bf16, dmf_stream_t s);
/* fill n floats with a value (buffer zeroing for the accumulate-style outputs) */
int dmf_fill_f32(float* p, long long n, float value, dmf_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* DMF_B200_H */
