/*
 * tta.h -- C ABI of libtta.so: B200 (sm_100a) kernels for the ADMM low-rank projection hot path of
 * miaoyin390/DNN-Compression-Tensor-ADMM.
 *
 * The reference has no FFI of its own (pure Python: numpy/LAPACK/tensorly/torch library calls); the
 * boundary a maintainer binds is the set of batched operators below.  Each entry cites the reference
 * call site(s) it replaces (file:line relative to the reference repo).
 *
 * Conventions
 *   - every data pointer is a DEVICE pointer owned by the caller; the library never allocates device
 *     memory and keeps no global state except `tta_last_error()`'s thread-local string;
 *   - `stream` is a `cudaStream_t` passed as `void*`; calls only enqueue work unless stated otherwise;
 *   - "task tables" are arrays of the plain structs below living in DEVICE memory (one kernel launch
 *     serves a whole table: all layers of a network step are batched);
 *   - return value: 0 on success, negative `TTA_E_*` on failure (`tta_last_error()` has the text);
 *   - matrices are row-major unless a stride says otherwise; all floating data is fp32, Gram
 *     accumulation is fp64.
 */
#ifndef TTA_H_
#define TTA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TTA_OK 0
#define TTA_E_INVALID (-1)   /* bad argument                                   */
#define TTA_E_CUDA (-2)      /* CUDA runtime error (text in tta_last_error)    */
#define TTA_E_NOCONV (-3)    /* eigensolver hit max_sweeps without converging  */
#define TTA_E_ARCH (-4)      /* device is not sm_100                           */

const char* tta_last_error(void);
int tta_version(void);
/* number of kernels this library has launched so far in this process (bench.py: `gpu_launches`) */
unsigned long long tta_launch_count(void);
/* 0 if device `dev` can run this library (compute capability 10.x), else TTA_E_ARCH. */
int tta_check_device(int dev);

/* ---------------------------------------------------------------------------------------------
 * Dual update and penalty (fused multi-tensor elementwise)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* w; /* parameter                                  */
  const float* z; /* projected copy                             */
  float* u;       /* scaled dual (read; written by dual_update) */
  float* g;       /* gradient buffer (penalty_bwd only)         */
  int64_t numel;
} tta_ew_task;

/* admm.py:71-78   U += W - Z for every listed layer; if `sqnorm_out` != NULL also
 * sqnorm_out[t] = ||W_t - Z_t||^2 (fp64; the value admm.py:76,78 logs is its square root).
 * `sqnorm_out` must be zeroed by the caller.  One launch for all tensors. */
int tta_dual_update_multi(const tta_ew_task* tasks_dev, const tta_ew_task* tasks_host, int n_tasks,
                          double* sqnorm_out, void* stream);

/* admm.py:80-85   loss_out[0] += sum_t 0.5*rho*||W_t - Z_t + U_t||^2   (fp64 accumulator, caller
 * zeroes it). */
int tta_penalty_fwd_multi(const tta_ew_task* tasks_dev, const tta_ew_task* tasks_host, int n_tasks,
                          float rho, double* loss_out, void* stream);

/* autograd of admm.py:83:  g_t (+)= grad_scale[0] * rho * (W_t - Z_t + U_t).
 * `grad_scale` is a device scalar (the incoming grad_output -- GradScaler scales the loss,
 * engines.py:315).  accumulate != 0 adds into g (existing .grad), else overwrites. */
int tta_penalty_bwd_multi(const tta_ew_task* tasks_dev, const tta_ew_task* tasks_host, int n_tasks,
                          float rho, const float* grad_scale, int accumulate, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Orthogonality regulariser of the decomposed layers' factor matrices (orthogonal.py:9-20; fine-tune loop
 * engines.py:290-291,297-298).  One problem = n vectors of length len (element t of vector i at p + i*si + t*st; one
 * of si, st is 1): a factor F with fewer rows than columns is its rows (si = cols, st = 1), otherwise its columns
 * (si = 1, st = cols).     R = X X^T - I,   loss_out[0] += rho/2 * sum R^2   (fp64, caller zeroes it),
 *                          g_i (+)= grad_scale[0] * 2 rho * sum_j R_ij x_j    (g laid out like p)
 * fwd writes R into r (n*n floats per task), bwd reads it.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* p;
  float* r;
  float* g;       /* bwd only */
  int64_t si, st;
  int32_t n, len;
} tta_orth_task;

int tta_orth_penalty_fwd_batched(const tta_orth_task* tasks_dev, const tta_orth_task* tasks_host, int n_tasks,
                                 float rho, double* loss_out, void* stream);
int tta_orth_penalty_bwd_batched(const tta_orth_task* tasks_dev, const tta_orth_task* tasks_host, int n_tasks,
                                 float rho, const float* grad_scale, int accumulate, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Unfold / fold (admm.py:45 + admm.py:96 and admm.py:99)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* w; /* (O, I, KK) contiguous                                         */
  const float* u; /* same shape, may be NULL                                       */
  float* t;       /* unfold: out (O, KK, I);  fold: in                             */
  float* z;       /* fold: out (O, I, KK); unfold: unused                          */
  int32_t O, I, KK; /* KK == 1 (linear / 1x1) degenerates to a plain add / copy    */
  int32_t pad_;
} tta_fold_task;

/* t[o,kk,i] = w[o,i,kk] + u[o,i,kk]   (V = W + U of admm.py:45 fused with the permute of :96) */
int tta_unfold_add_batched(const tta_fold_task* tasks_dev, const tta_fold_task* tasks_host,
                           int n_tasks, void* stream);
/* z[o,i,kk] = t[o,kk,i]               (admm.py:99) */
int tta_fold_store_batched(const tta_fold_task* tasks_dev, const tta_fold_task* tasks_host,
                           int n_tasks, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Gram matrices  G = sum_{b,c} a(i,b,c) a(j,b,c)   (replaces the SVD input side of ttd.py:16-17 and
 * tensorly's unfold+svd in admm.py:116,124)
 *   element a(i,b,c) lives at  a + i*si + b*sb + c*sc ;  i < k, b < nb, c < nc
 *   row Gram  A A^T of a row-major m x n matrix:  k=m, si=n, nb=1, nc=n, sc=1
 *   col Gram  A^T A:                              k=n, si=1, nb=1, nc=m, sc=n
 *   mode Gram of a (B, k, c) tensor:              si=c, sb=k*c, nc=c, sc=1, nb=B
 * Output: x (fp32) holds G as `kpad` columns of length `ld` (column j at x + j*ld), zero padded --
 * the eigensolver's in-place state.  `part` is fp64 scratch of nsplit*k*k doubles.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* a;
  double* part;
  float* x;    /* nullable when g64 is given (the fp64 eigensolver reads g64 only) */
  double* g64; /* nullable: G as a k x k row-major fp64 matrix (input of the refinement step / tta_symeig_top_batched) */
  int64_t si, sb, sc;
  int32_t k, nb, nc, nsplit;
  int32_t ld, kpad;
  const float* a2; /* nullable: second addend, same indexing -- the operand is a + a2 (V = W + U of admm.py:45 read in
                    * place: the first TT step / HOSVD mode needs no materialised sum) */
} tta_gram_task;

int tta_gram_batched(const tta_gram_task* tasks_dev, const tta_gram_task* tasks_host, int n_tasks,
                     void* stream);
/* Default on: tasks whose operand TMA can address (nb == 1, one of the two indices contiguous, 16-byte aligned base and
 * pitch, reduction length >= 32) are computed on the tensor cores -- TMA boxes of a (and a2) -> hi / lo TF32 split in
 * shared memory -> tcgen05.mma kind::tf32 (lo*hi + hi*lo + hi*hi) with a FRESH TMEM accumulator for every 32
 * reduction indices, drained into fp32 registers (round to nearest) and summed over slices in fp64.  The tensor core's
 * accumulator truncates: its error grows linearly with the accumulation length (3.5e-6 at 512, measured), 1.3e-7 for
 * one 32-index block.  The remaining tasks, and all tasks when off, use the fp64 CUDA-core kernels.  `part` is
 * reinterpreted as float[nsplit][k][k] by the tensor-core path. */
void tta_gram_enable_tc(int on);

/* ---------------------------------------------------------------------------------------------
 * Symmetric eigensolver: one-sided (Hestenes) block Jacobi on the columns of X = G
 * (replaces numpy.linalg.svd / LAPACK gesdd of ttd.py:17, admm.py:131,143 and tensorly partial_svd)
 * On return the columns of X are mutually orthogonal: x_j = lambda_j * v_j (column order arbitrary).
 * X need not be G itself: any X = G * Q with Q orthogonal is a valid start (warm start from the eigenvectors of
 * a previous, similar problem) -- the iteration converges to G * Q * J = V * Lambda all the same.
 * Synchronises `stream` once when sweeps_out is given (convergence status is read back); see below for the
 * enqueue-only mode.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  float* x;         /* kpad columns of length ld                       */
  int32_t k;        /* true order                                      */
  int32_t ld;       /* column length, multiple of 4, >= k              */
  int32_t kpad;     /* column count, multiple of bw                    */
  int32_t bw;       /* column block width: even, <= 32 (<= 16 for ld > 512) */
} tta_eig_task;

/* sweeps_out (host, nullable): number of sweeps each problem needed. */
int tta_jacobi_eigh_batched(const tta_eig_task* tasks_dev, const tta_eig_task* tasks_host,
                            int n_tasks, float tol, int max_sweeps, int32_t* scratch_dev,
                            size_t scratch_bytes, int32_t* sweeps_out, void* stream);
size_t tta_jacobi_scratch_bytes(const tta_eig_task* tasks_host, int n_tasks);
/* Asynchronous use: with sweeps_out == NULL (and every problem on a persistent cluster solver, k <= 512)
 * tta_jacobi_eigh_batched only enqueues work and returns without a host synchronisation.  After the
 * stream has been synchronised, copy the first 6*n_tasks int32 of the scratch buffer to the host and
 * pass them here: fills sweeps_out (nullable) and returns TTA_E_NOCONV if a problem did not converge. */
int tta_jacobi_read_results(const int32_t* scratch_host, const tta_eig_task* tasks_host, int n_tasks,
                            int max_sweeps, int32_t* sweeps_out);
/* Profiling aid for bench.py: when enabled, every sweep's launch sequence is bracketed by CUDA events
 * on `stream` (no host sync inside the bracket) and the elapsed device time / number of
 * jacobi_step launches are accumulated.  `tta_jacobi_profile_read` returns and clears them. */
void tta_jacobi_profile_enable(int on);
/* Test hook: != 0 routes every problem through the multi-launch solver (default: problems with
 * ld <= 512 whose geometry is kpad == 2*P*bw, P in {1,2,4,8,16}, bw even <= 32, run on the persistent
 * thread-block-cluster solver; the rest -- k up to 2048 in the Tucker sweep -- on the multi-launch one). */
void tta_jacobi_force_multilaunch(int on);
/* Test hook: 0 keeps 32 < k <= 512 problems on the column-rotation cluster kernel instead of the
 * gram-rotate-apply kernel (geometry bw == 16, kpad == 32*P, 2 <= P <= 16).  Default 1. */
void tta_jacobi_enable_gra(int on);
/* Gram-rotate-apply solver: stop after a sweep whose largest |c_pq|/sqrt(c_pp c_qq), seen before
 * rotating, is below stop_rel (the sweep leaves off-diagonals ~ stop_rel^2).  Default 3e-4;
 * 0 = stop only after a sweep without rotations. */
void tta_jacobi_set_stop_rel(float stop_rel);
void tta_jacobi_profile_read(double* step_ms, unsigned long long* step_launches);

/* ---------------------------------------------------------------------------------------------
 * Select the r dominant eigenpairs of a converged X (truncation of ttd.py:21-23, admm.py:132-134)
 *   e      (r x k) row-major: row p = p-th dominant unit eigenvector (zero row when lambda ~ 0)
 *   et     (k x r) row-major: transpose of e            (nullable)
 *   se     (r x k) row-major: sqrt(lambda_p) * e row p  (nullable; = diag(s) V^T of ttd.py:26)
 *   sigma  (r)     sqrt(lambda_p)                        (nullable)
 *   isigma (r)     1/sqrt(lambda_p), 0 when lambda ~ 0   (nullable)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* x;
  float* e;
  float* et;
  float* se;
  float* sigma;
  float* isigma;
  int32_t k, ld, r, pad_;
} tta_select_task;

int tta_select_batched(const tta_select_task* tasks_dev, const tta_select_task* tasks_host,
                       int n_tasks, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Batched fp32 GEMM  C[M,N] = (A[M,K] * B[K,N]) .* colscale[N]
 *   A(i,k) at a + i*sai + k*sak ; B(k,j) at b + k*sbk + j*sbj ; C row-major, leading dim ldc.
 * (replaces np.dot of ttd.py:26,39-40, the projection side of the SVDs, tensorly mode products)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* a;
  const float* b;
  float* c;
  const float* colscale; /* nullable */
  int64_t sai, sak, sbk, sbj, ldc;
  int32_t M, N, K;
  int32_t flags;         /* tta_gemm_f64_batched only: TTA_GEMM_STORE_F32 -- c is a float array, results are rounded
                          * on store; TTA_GEMM_GUARD -- colscale is not a scale but points to ONE double: the task
                          * is skipped when that value is 0 */
} tta_gemm_task;
#define TTA_GEMM_STORE_F32 1
#define TTA_GEMM_GUARD 2

int tta_gemm_batched(const tta_gemm_task* tasks_dev, const tta_gemm_task* tasks_host, int n_tasks,
                     void* stream);
/* Mode 1 (default): fp32 tasks of at least 0.6 GFLOP (max(M, N) >= 128, min(M, N) >= 16, K >= 64) run on the tensor
 * cores (tcgen05, 3xTF32 split: A_lo B_hi + A_hi B_lo + A_hi B_hi into a ZEROED TMEM accumulator for every 32 reduction
 * indices, running sum in fp32 registers); smaller tasks on CUDA cores, where they are faster (crossover measured in
 * scripts/bench_gemm_shapes.py: 85 against 31 TFLOP/s at 18432 x 1024 x 2048, 41 against 33 us at 105 x 4608 x 480).
 * Mode 0: CUDA cores only.  Mode 2 (tests): every task on the tensor cores.
 * Accuracy, measured on B200 (scripts/ubench/tf32_accum_error.py): the tensor core's truncating accumulation costs
 * 3.5e-6 relative error at K = 512 when the sum stays in TMEM (amplified past the 1e-4 parity bar by the next small-gap
 * SVD of a TT chain), 1.3e-7 per 32-index block -- the same grade as the CUDA-core fp32 kernel. */
void tta_gemm_enable_tc(int on);
/* Same operator in fp64: a, b, c, colscale address doubles (used by the refinement step below). */
int tta_gemm_f64_batched(const tta_gemm_task* tasks_dev, const tta_gemm_task* tasks_host, int n_tasks,
                         void* stream);

/* ---------------------------------------------------------------------------------------------
 * fp64 refinement of the fp32 Jacobi eigenvectors (one Ogita-Aishima step) fused with the dominant-r
 * selection.  Thousands of fp32 plane rotations leave an O(1e-5) error in the invariant subspace;
 * one first-order correction computed from S = Q^T G Q and T = Q^T Q in fp64 squares it away.
 * Only the r dominant vectors are corrected, so only the rows of S and T that can be selected are
 * formed: the vectors are first sorted by their fp32 eigenvalue estimate and the GEMMs compute the
 * first `wnd` rows (wnd >= r plus a safety margin against the fp32 ordering error).
 * Sequence per problem (all batched):
 *   refine_prepare : qt (k x k fp64, row p = unit vector of the p-th largest column of x) and
 *                    lam0[p] = that column's norm                                from the Jacobi state x
 *   3 x gemm_f64   : y = qt[0:wnd] * g64 ; s = y * qt^T ; t = qt[0:wnd] * qt^T   (caller enqueues these)
 *   refine_coeff   : lambda_j = s_jj / t_jj inside the window (lam0 behind it), descending rank, and for
 *                    each of the r dominant j the coefficient row c[p,:] = e_j + (first-order
 *                    correction)  (r x k fp64)
 *   1 x gemm_f64   : e64 = c * qt                                          (caller enqueues)
 *   refine_finalize: e / et / se / sigma / isigma in fp32 -- same outputs as tta_select_batched
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* x;   /* Jacobi state: kpad columns of length ld                   */
  double* qt;       /* k x k                                                     */
  const double* s;  /* wnd x k                                                   */
  const double* t;  /* wnd x k                                                   */
  double* c;        /* r x k                                                     */
  double* lam;      /* r      refined eigenvalues of the selected vectors        */
  double* lam0;     /* k      fp32 eigenvalue estimates in sorted order          */
  const double* e64;/* r x k                                                     */
  float* e;
  float* et;        /* nullable */
  float* se;        /* nullable */
  float* sigma;     /* nullable */
  float* isigma;    /* nullable */
  int32_t k, ld, r, wnd;
} tta_refine_task;

int tta_refine_prepare_batched(const tta_refine_task* tasks_dev, const tta_refine_task* tasks_host,
                               int n_tasks, void* stream);
int tta_refine_coeff_batched(const tta_refine_task* tasks_dev, const tta_refine_task* tasks_host,
                             int n_tasks, void* stream);
int tta_refine_finalize_batched(const tta_refine_task* tasks_dev, const tta_refine_task* tasks_host,
                                int n_tasks, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dominant-r symmetric eigensolver in fp64 (replaces numpy.linalg.svd / LAPACK gesdd of ttd.py:17,
 * admm.py:131,143 and tensorly's partial_svd behind admm.py:116,124 for 3 <= k <= tta_symeig_max_k()):
 * Householder tridiagonalisation of G on a thread-block cluster, multisection on the Sturm count for the r
 * largest eigenvalues, twisted factorisation for their eigenvectors, back-transformation.  Everything is
 * fp64, so no refinement step follows; tta_refine_finalize_batched (only its e64 / lam / k / r / output
 * fields are read) converts the result into the fp32 outputs of the projection.
 *   g      k x k row-major fp64, symmetric (tta_gram_task.g64); not modified
 *   work   tta_symeig_work_doubles(k, r) doubles of scratch
 *   lam    r eigenvalues, descending
 *   e64    r x k row-major: row p = eigenvector of the p-th largest eigenvalue (not normalised)
 *   status one int32 per task: 0, or 1 when two of the r eigenvalues above the zero cut are closer than 1e-9 |G|
 *          (their computed vectors may be parallel: the caller falls back to the Jacobi solver)
 * Tasks with equal cluster size should be adjacent in the table (sort by descending k): every run of
 * equal geometry is one launch.  Only enqueues work.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const double* g;
  double* work;
  double* lam;
  double* e64;
  int32_t* status;
  int32_t k, r;
} tta_symeig_task;

int tta_symeig_top_batched(const tta_symeig_task* tasks_dev, const tta_symeig_task* tasks_host, int n_tasks,
                           void* stream);
size_t tta_symeig_work_doubles(int k, int r);
int tta_symeig_max_k(void);
/* Profiling aid for bench.py: when enabled, every tridiagonalisation launch (the dominant kernel of an update) is
 * bracketed by CUDA events on the stream it is launched on; `read` synchronises those events, returns the summed
 * device time / number of launches since the last read and clears them. */
void tta_symeig_profile_enable(int on);
void tta_symeig_profile_read(double* reduce_ms, unsigned long long* reduce_launches);
/* tta_symeig_profile_enable(2): five events per tta_symeig_top_batched call on the caller's stream (before the reduction,
 * after it, after the eigenvalues, after the vectors, after the back-transformation).  Readback, in call order, 7 doubles
 * per call: largest k, number of tasks, and the five times in ms after `origin` (a cudaEvent_t the caller recorded). */
int tta_symeig_stage_profile_read(void* origin, double* out, int max_records);

/* out[t] = sum of squares of n floats (fp64): tensorly `tl.norm(core, 2)**2` in the HOOI stopping rule. */
typedef struct {
  const float* x;
  int64_t n;
} tta_sqnorm_task;
int tta_sqnorm_batched(const tta_sqnorm_task* tasks_dev, const tta_sqnorm_task* tasks_host,
                       int n_tasks, double* out_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Decomposed-layer forward building blocks (TTLinear.py:75-93, TTConv.py:130-153): bf16 operands,
 * fp32 accumulation.  Single-problem calls (one layer at a time, as the modules' forward() runs).
 * ------------------------------------------------------------------------------------------- */
/* C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) on the tcgen05 tensor cores.  A, B bf16 K-major (leading
 * dimensions lda, ldb in elements), C bf16 (out_fp32 == 0) or fp32, row-major with leading dim ldc.
 * K, lda, ldb multiples of 8; A, B 16-byte aligned. */
int tta_gemm_bf16_tc(const void* a, int64_t lda, const void* b, int64_t ldb, void* c, int64_t ldc, int M,
                     int N, int K, const float* bias, int out_fp32, void* stream);
/* Skinny contraction (K, N <= 96), one thread per row: C(i, j) = sum_k arow_i[k] * B[j*K + k] with
 * arow_i = a + (i / a_inner) * a_outer + (i % a_inner) * K, stored at
 * c + (i / m_inner) * s_outer + (i % m_inner) * s_inner + j * s_col, plus
 * bias[(i % m_inner) * bias_inner + j * bias_col] when bias != NULL.  A is fp32 or bf16, B bf16. */
int tta_small_gemm(const void* a, int a_is_f32, const void* b_bf16, void* c, int c_is_f32, const float* bias,
                   int64_t M, int N, int K, int64_t a_inner, int64_t a_outer, int64_t m_inner, int64_t s_outer,
                   int64_t s_inner, int64_t s_col, int64_t bias_inner, int64_t bias_col, void* stream);
/* y (bf16) = x (fp32), n elements */
int tta_cast_bf16(const float* x, void* y, int64_t n, void* stream);
/* Activation layouts of the conv forwards (TTConv.py:132-149, TKConv.py:205-222).
 * x (B, C, HW) fp32 -> y (B, HW, ldc) bf16; and back (bf16 or fp32 source) with the bias add fused. */
int tta_nchw_to_nhwc_bf16(const float* x, void* y, int B, int C, int HW, int ldc, void* stream);
int tta_nhwc_to_nchw_f32(const void* x, int x_is_f32, float* y, const float* bias, int B, int C, int HW, int ldc,
                         void* stream);
/* im2col rows of a (B, H, W, ldx) bf16 activation for the k x k core convolution (TTConv.py:139,
 * TKConv.py:95): out is (B*Ho*Wo) x ldo, column (kh*KW + kw)*C + c, zero padded. */
int tta_im2col_bf16(const void* x, void* out, int B, int H, int W, int C, int ldx, int KH, int KW, int sh, int sw,
                    int ph, int pw, int dh, int dw, int Ho, int Wo, int ldo, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused forward of a factorised convolution:  1x1 (C_in -> r_a)  ->  k x k (r_a -> r_b, stride, zero
 * padding)  ->  1x1 (r_b -> C_out) + bias in ONE kernel, intermediates in shared memory, fp32.
 * Replaces the op chains of TTConv2dM.forward (TTConv.py:130-153; the host folds the in-core chain into
 * a_in and the out-core chain into a_out) and TKConv2dC/M.forward (TKConv.py:93-98, 205-222: first
 * factor, core, last factor).
 *   x (B, C_in, H, W) NCHW fp32;  a_in (r_a x C_in);  kern (r_b, r_a, k, k);  a_out (C_out x r_b);
 *   bias (C_out) nullable;  y (B, C_out, Ho, Wo) NCHW fp32.   k in {1, 3}, stride in {1, 2}, dilation 1.
 * ------------------------------------------------------------------------------------------- */
int tta_ttconv_fused_fwd(const float* x, const float* a_in, const float* kern, const float* a_out,
                         const float* bias, float* y, int B, int Cin, int H, int W, int Ra, int Rb,
                         int Cout, int KS, int stride, int pad, void* stream);

/* Same contraction on the tensor cores (bf16 tcgen05.mma, fp32 TMEM accumulators, both intermediates in shared memory
 * as bf16; csrc/ttconv_tc.cu): pixel-major implicit GEMM over the zero-padded position space, the nine taps of the
 * 3 x 3 stage are row-shifted views of the stage-1 result (un-swizzled K-major operands: a shift is a descriptor start
 * address).  Serves k = 3, stride 1 or 2 (the stride-1 result is formed at every position and the odd ones are dropped:
 * the three contractions are cheap next to the activation traffic), pad 1, channel counts and ranks <= 64
 * (tta_ttconv_tc_supported); other geometries
 * return TTA_E_INVALID -- the callers fall back to tta_ttconv_fused_fwd.  The weights are packed once per weight change
 * into the shared-memory image of the kernel (tta_ttconv_tc_pack: bf16 planes of a_in, the nine taps of kern, a_out,
 * fp32 bias; tta_ttconv_tc_blob_bytes bytes) and fetched by every CTA with one bulk copy.  Output within 1e-2 relative
 * of the fp32 contraction (three bf16 roundings), the bar of the decomposed-layer forwards. */
int tta_ttconv_tc_supported(int Cin, int Ra, int Rb, int Cout, int KS, int stride, int pad);
int64_t tta_ttconv_tc_blob_bytes(int Cin, int Ra, int Rb, int Cout);
int tta_ttconv_tc_pack(const float* a_in, const float* kern, const float* a_out, const float* bias, void* blob,
                       int Cin, int Ra, int Rb, int Cout, void* stream);
int tta_ttconv_tc_fwd(const float* x, const void* blob, float* y, int B, int Cin, int H, int W, int Ra, int Rb,
                      int Cout, int KS, int stride, int pad, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused two-factor linear forward on the tcgen05 tensor cores (TMA-fed, persistent, the rank-N1
 * intermediate stays in TMEM / shared memory):
 *     y[M, N2] = bf16( x[M, K1] * w1[N1, K1]^T ) * w2[N2, N1]^T + bias[N2]
 * Replaces the op chain of TTLinearM.forward (TTLinear.py:75-93; the host folds the input-side cores
 * into w1 and the output-side cores into w2) and of TKLinearM.forward (TKLinear.py:60-75).
 *   x, w1, w2 bf16 row-major with leading dimensions ldx, ld1, ld2 (multiples of 8, 16-byte aligned
 *   bases);  bias fp32 nullable;  y fp32 (out_fp32 != 0, ldy % 4 == 0) or bf16 (ldy % 8 == 0);
 *   N1 <= 384 (TMEM budget: N1 + 2 output chunks <= 512 columns).  y is written by TMA stores in 16-byte
 *   granules: when N2 is not a multiple of 4 (fp32) / 8 (bf16) the pad columns up to that multiple (they lie
 *   inside ldy) receive zeros.
 * ------------------------------------------------------------------------------------------- */
int tta_lowrank2_fwd(const void* x, int64_t ldx, const void* w1, int64_t ld1, const void* w2, int64_t ld2,
                     const float* bias, void* y, int64_t ldy, int out_fp32, int64_t M, int K1, int N1, int N2,
                     void* stream);

/* ---------------------------------------------------------------------------------------------
 * Weight-gradient GEMM of the fused training path (fwd_common.LowRank2Fn.backward; replaces torch.mm / cuBLAS):
 *     c[M, N] (fp32, row pitch ldc) = a^T b,   a (K x M) and b (K x N) bf16 row-major (pitches lda, ldb: multiples of 8,
 *     16-byte aligned bases) -- the reduction index K (tokens) is the slow index of both operands.
 * TMA boxes land as MN-major SWIZZLE_128B operands of tcgen05.mma (no transposition), split-K over CTAs, partial
 * tiles in `workspace` (tta_gemm_bf16_tn_workspace_bytes), summed in a fixed order.
 * ------------------------------------------------------------------------------------------- */
int64_t tta_gemm_bf16_tn_workspace_bytes(int M, int N, int K);
int tta_gemm_bf16_tn(const void* a, int64_t lda, const void* b, int64_t ldb, float* c, int64_t ldc, int M, int N, int K,
                     void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TTA_H_ */
