/*
 * cdr_b200.h -- C ABI of the B200 (sm_100a) implementation of the
 * convex_dim_red alternating-update hot path.
 *
 * The reference (azedarach/matrix-factorization-case-studies) is pure Python:
 * it has no FFI / plugin interface, so the drop-in boundary is its Python API
 * (src/convex_dim_red/__init__.py:5-11).  This header is the thin C layer the
 * Python host code (matrix-factorization-case-studies_b200/convex_dim_red)
 * binds with ctypes; each entry point names the reference code it replaces.
 * Citations are relative to /root/reference/src/convex_dim_red/.
 *
 * Conventions
 *   - all matrix pointers are DEVICE pointers to IEEE fp64, row-major;
 *   - every function is asynchronous on `stream` and returns 0 on success,
 *     a positive cudaError_t if a launch failed, or a negative CDR_ERR_* code;
 *   - nothing is allocated behind the caller's back: workspaces are passed in
 *     and sized by the `*_workspace_bytes` twins;
 *   - "padded" matrices have a leading dimension that is a multiple of
 *     CDR_LD_ALIGN doubles with the padding columns zero-filled;
 *   - `flags` (may be NULL) points at a device-resident cdr_loop_state; when
 *     its `done` field is non-zero every kernel returns without touching
 *     memory, so an iteration captured in a CUDA graph can be replayed past
 *     convergence without changing the result.
 */
#ifndef CDR_B200_H
#define CDR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDR_LD_ALIGN 32          /* doubles */
#define CDR_MAX_COMPONENTS 64    /* k <= 64 */
#define CDR_MAX_MEMORY 8         /* non-monotone line-search memory */

#define CDR_ERR_INVALID_ARGUMENT (-1)
#define CDR_ERR_UNSUPPORTED (-2)
#define CDR_ERR_WORKSPACE (-3)
#define CDR_ERR_NOT_APPLICABLE (-4) /* fused path does not cover this shape: use the unfused calls */

typedef void* cdr_stream_t; /* cudaStream_t */

/* Options of the simplex-constrained QP solver (defaults: spg.py:287-291). */
typedef struct cdr_spg_params {
    double gamma;
    double sigma_one;
    double sigma_two;
    double lambda_min;
    double alpha0;
    double alpha_min;
    double alpha_max;
    double epsilon_one;
    double epsilon_two;
    int memory;
    int max_iterations;
    int max_feval;
    int use_infinity_norm;
} cdr_spg_params;

/* Device-resident control block shared by the kernels of one fit.  The
 * `done` field must stay first (cdr_flags aliases it). */
typedef struct cdr_loop_state {
    int done;             /* set on convergence, failure or iteration limit */
    int error_stage;      /* 0 none, 1 scale factors, 2 dictionary, 3 weights */
    int n_iter;           /* completed outer iterations */
    int converged;        /* stopping rule fired */
    int max_iterations;
    int stopping_rule;    /* 0 abs_delta_f, 1 rel_delta_f (archetypal_analysis.py:177-197) */
    int require_monotone; /* archetypal_analysis.py:167-174 */
    int spg_iter;         /* inner (dictionary) SPG iteration counter */
    int spg_feval;
    int spg_active;       /* inner SPG still iterating */
    int spg_alpha_set;    /* 0 until the first step length is initialised (spg.py:178-189) */
    int spg_warnings;     /* sticky: bit 0 step below lambda_min, bit 1 feval limit, bit 2 iteration limit */
    double tolerance;
    double trace_data;    /* trace(K) or ||X||_F^2 */
    double cost;          /* current cost */
    double old_cost;      /* cost at the start of the outer iteration */
    double f_old;
    double f_new;
    double lam;
    double alpha;         /* spectral step length */
    double delta;         /* <d, g> */
    double dd;            /* <d, d> */
    double a0;            /* linear trace term at x */
    double a1;            /* linear trace term along d */
    double beta;
    double res2;
    double resinf;
    double penalty;       /* GPNH: lambda_W * phi(W) */
    double f_mem[CDR_MAX_MEMORY];
    unsigned int tickets[4]; /* "last CTA" counters of the fused kernels; zero between launches */
} cdr_loop_state;

typedef cdr_loop_state cdr_flags;

const char* cdr_version(void);
int cdr_device_check(void); /* 0 if the current device is sm_100 */
unsigned long long cdr_launch_count(void); /* kernels launched by this library so far */
/* host-only: kernel choice and tile geometry of the two streaming passes for a shape
 * (12 ints, see stream_gemm.cu); lets the dispatch logic be tested without a GPU */
int cdr_debug_stream_plan(int T, int d, int k, int with_epilogue, int* out);
/* fp64 tensor-pipe (DMMA.8x8x4) throughput probe: `blocks` CTAs x 8 warps x 8 independent
 * chains x `iters` MMAs on registers; out: blocks * 256 doubles.  The caller times the launch;
 * flops = 2 * 256 * 8 * iters * 8 * blocks.  The result is the roofline denominator of the
 * tensor-bound shapes (Gram, k = 64), which MEASURED_PEAKS.json does not carry. */
int cdr_debug_dmma_probe(double* out, int blocks, int iters, cdr_stream_t stream);
/* debugging aids: CDR_DEBUG_SYNC=1 synchronises after every launch and reports the first
 * failing one; CDR_TIME_LAUNCHES=1 records an event after every (eager, default-stream) launch
 * and this call prints the time between consecutive events per launch site. */
int cdr_debug_timing_report(int skip);

/* ------------------------------------------------------------------ simplex
 * Euclidean projection of each row / column of A onto the probability simplex.
 * Replaces simplex_projection.py:13-47 (simplex_project_vector / _rows /
 * _columns).  out may alias A. */
int cdr_simplex_project_rows(const double* A, double* out, int m, int n, long lda, long ldo,
                             const cdr_flags* flags, cdr_stream_t stream);
int cdr_simplex_project_columns(const double* A, double* out, int m, int n, long lda, long ldo,
                                const cdr_flags* flags, cdr_stream_t stream);

/* ------------------------------------------------------------------ per-sample QP
 * For every sample t solve  min 1/2 z'A'z + b_t'z  over the simplex with the
 * SPG of spg.py:286-398, where A' = diag(alpha) A diag(alpha) and
 * b_t[c] = -alpha[c] * B[t*sb_t + c*sb_c]  (alpha may be NULL = ones).
 * Replaces _gu_update_kernel_aa_weights / _update_kernel_aa_weights
 * (archetypal_analysis.py:344-396) and _gu_update_gpnh_weights
 * (gpnh_convex_coding.py:229-251).  Z (T x k, row-major) is updated in place.
 * n_iter / n_feval (int[T], may be NULL) receive per-sample counts. */
int cdr_quad_simplex_spg_batched(const double* A, const double* alpha, const double* B,
                                 long sb_t, long sb_c, double* Z, int T, int k,
                                 const cdr_spg_params* params, int* n_iter, int* n_feval,
                                 const cdr_flags* flags, cdr_stream_t stream);

/* ------------------------------------------------------------------ streaming contractions
 * X is T x d with padded leading dimension ldx.
 *
 * reduce over samples:   out[k x ldo] = E (L X),  L[i][t] = Lp[i*sLi + t*sLt],
 *   E (k x k, may be NULL = identity).  Replaces dictionary.dot(X),
 *   X.T.dot(weights) (archetypal_analysis.py:545-549,618,641) and
 *   weights.T.dot(X) + lstsq back-substitution (gpnh_convex_coding.py:219-224).
 * reduce over features:  out[k x ldo] (k x T) = M X',  M is k x ldm padded.
 *   Replaces CX.dot(X.T), X.dot(XtZ) (archetypal_analysis.py:546,549,619,642)
 *   and X.dot(dictionary) (gpnh_convex_coding.py:271,352).
 * gram:                  K[T x ldk] = X X' (archetypal_analysis.py:1032).
 */
size_t cdr_reduce_samples_workspace_bytes(int T, int d, int k);
int cdr_reduce_samples(const double* Lp, long sLi, long sLt, const double* X, long ldx, int T,
                       int d, int k, const double* E, double* out, long ldo, void* workspace,
                       size_t workspace_bytes, const cdr_flags* flags, cdr_stream_t stream);
size_t cdr_reduce_features_workspace_bytes(int T, int d, int k);
int cdr_reduce_features(const double* M, long ldm, const double* X, long ldx, int T, int d,
                        int k, double* out, long ldo, void* workspace, size_t workspace_bytes,
                        const cdr_flags* flags, cdr_stream_t stream);
size_t cdr_gram_workspace_bytes(int T, int d);
int cdr_gram(const double* X, long ldx, int T, int d, double* K, long ldk, void* workspace,
             size_t workspace_bytes, cdr_stream_t stream);
/* K = X X' as a SYRK (tiles on and above the diagonal, mirrored; T^2 d flops; csrc/syrk.cu).
 * part_index / part_count (0 / 1 for everything): only the upper-triangle tiles
 * t = part_index (mod part_count) and their mirror images are written -- ranks of a process
 * group deal the tiles this way and sum the zero-initialised matrices. */
size_t cdr_syrk_workspace_bytes(int T, int d);
int cdr_syrk(const double* X, long ldx, int T, int d, double* K, long ldk, int part_index,
             int part_count, void* workspace, size_t workspace_bytes, cdr_stream_t stream);
/* sum of squares of a T x d matrix (trace of the Gram, archetypal_analysis.py:552) */
int cdr_frobenius_sq(const double* X, long ldx, int T, int d, double* out, void* workspace,
                     size_t workspace_bytes, cdr_stream_t stream);
size_t cdr_frobenius_workspace_bytes(void);
/* fixed-order sum of n doubles (k-means inertia) */
int cdr_sum_vector(const double* v, int n, double* out, cdr_stream_t stream);
/* ||X - Z A||_F^2 with direct differences (gpnh_convex_coding.py:199-210,
 * archetypal_analysis.py:1196-1197); Z is T x k, A is k x lda; part: T doubles */
int cdr_residual_sq(const double* X, long ldx, int T, int d, const double* Z, int k,
                    const double* A, long lda, double* out, double* part, cdr_stream_t stream);

/* ------------------------------------------------------------------ small products
 * out[i][j] = scale * sum_n A[i*sAi + n*sAn] * B[j*sBj + n*sBn]   (mode 0)
 * out[i][j] = scale * sum_n (A[i,n] - B[j,n])^2                    (mode 1)
 * Up to CDR_GRAM_BATCH independent products per launch. */
#define CDR_GRAM_BATCH 4
typedef struct cdr_small_gram_desc {
    const double* A;
    const double* B;
    double* out;      /* ka x kb, row-major, leading dimension kb */
    long sAi, sAn, sBj, sBn;
    int ka, kb, n, mode;
    double scale;
} cdr_small_gram_desc;
size_t cdr_small_gram_workspace_bytes(void);
int cdr_small_gram(const cdr_small_gram_desc* descs, int count, void* workspace,
                   size_t workspace_bytes, const cdr_flags* flags, cdr_stream_t stream);

/* P = pinv(S / n_samples + lambda * G) / n_samples for the symmetric k x k
 * matrix S, G = 4/(d k (k-1)) (k I - 1); minimum-norm with the
 * numpy.linalg.lstsq(rcond=None) cut-off.  Replaces the k x k solve of
 * gpnh_convex_coding.py:221-226.  workspace: cdr_sym_pinv_workspace_bytes(k). */
size_t cdr_sym_pinv_workspace_bytes(int k);
int cdr_gpnh_solve_matrix(const double* ZtZ, int k, int n_samples, int n_features,
                          double lambda_W, double* P, void* workspace, size_t workspace_bytes,
                          const cdr_flags* flags, cdr_stream_t stream);

/* ------------------------------------------------------------------ AA dictionary SPG steps
 * Device-side pieces of spg() (spg.py:46-283) specialised to the AA dictionary
 * problem (archetypal_analysis.py:261-341); see DESIGN.md for the sequence.
 * Shapes: C, G, D, CK, DK, KZt are k x T with leading dimension ldt. */
typedef struct cdr_aa_buffers {
    double* C;          /* k x T dictionary (iterate x) */
    double* G;          /* gradient */
    double* D;          /* search direction */
    double* CK;         /* C K  (or (C X) X') */
    double* DK;         /* D K */
    double* KZt;        /* (K Z)' */
    double* alpha;      /* k scale factors */
    double* ZtZ;        /* k x k */
    double* CKCt;       /* k x k */
    double* CKZ;        /* k x k */
    double* G01;        /* k x k: CK D' */
    double* G11;        /* k x k: DK D' */
    double* row_scratch;/* 8 * k doubles of per-row partial results */
    cdr_loop_state* state;
    double* cost_deltas;/* max_iterations doubles */
    int k, T;
    long ldt;
    double grad_scale;  /* 1/T (archetypal_analysis.py:297) or 1/k (:288) */
    double cost_scale;  /* 1/k (archetypal_analysis.py:265,277) */
} cdr_aa_buffers;

/* P = pinv(S) for a symmetric k x k matrix (same cut-off) */
int cdr_sym_pinv(const double* S, int k, double* P, const cdr_flags* flags, cdr_stream_t stream);
/* start of an outer iteration: old_cost = cost, iteration-limit test */
int cdr_loop_begin(cdr_loop_state* state, cdr_stream_t stream);
/* spg.py:146-189: x = P(C) in place, f(x), df(x) -> G, first step length.
 * CK / CKCt must correspond to the projected C. */
int cdr_aa_spg_begin(const cdr_aa_buffers* b, const cdr_spg_params* p, cdr_stream_t stream);
/* spg.py:191-194: D = P(x - alpha G) - x and the row partials of <D,G>, <D,D> */
int cdr_aa_spg_direction(const cdr_aa_buffers* b, const cdr_spg_params* p, cdr_stream_t stream);
/* spg.py:196-229: needs DK = D K and G01 = CK D', G11 = DK D'; runs the whole
 * non-monotone Armijo search on the device and applies x += lam D, CK += lam DK */
int cdr_aa_spg_linesearch(const cdr_aa_buffers* b, const cdr_spg_params* p, cdr_stream_t stream);
/* spg.py:231-281: new gradient, Barzilai-Borwein step, residual test, counters */
int cdr_aa_spg_update(const cdr_aa_buffers* b, const cdr_spg_params* p, int compute_residual,
                      cdr_stream_t stream);
/* cost after a sub-step and the monotonicity / convergence tests of
 * archetypal_analysis.py:167-197, 623-663.  stage: 2 dictionary, 3 weights;
 * end_of_iteration != 0 also applies the stopping rule and advances n_iter. */
int cdr_aa_cost_check(const cdr_aa_buffers* b, int stage, int end_of_iteration,
                      cdr_stream_t stream);
/* delta != 0: _update_kernel_aa_scale_factors (archetypal_analysis.py:220-258), the generic
 * spg() on the k-vector b->alpha over the box [1 - delta, 1 + delta]^k, from b->ZtZ, b->CKCt,
 * b->CKZ and state->trace_data; one warp, no host round trip.  params: defaults of spg()
 * (spg.py:46-51). */
int cdr_aa_scale_factors_step(const cdr_aa_buffers* b, const cdr_spg_params* p, double delta,
                              cdr_stream_t stream);
/* stand-alone df(C) -> b->G and f(C) -> *out (archetypal_analysis.py:261-301) */
int cdr_aa_gradient(const cdr_aa_buffers* b, cdr_stream_t stream);
int cdr_aa_dictionary_cost(const cdr_aa_buffers* b, double trace_data, double* out,
                           cdr_stream_t stream);

/* ------------------------------------------------------------------ GPNH cost bookkeeping
 * gpnh_convex_coding.py:352-399.  tr(W'X'Z) = trace of XWtZ (k x k),
 * tr(Z'Z W'W) from the two k x k matrices; reg_pairs (k x k, may be NULL)
 * holds ||w_i - w_j||^2. */
int cdr_gpnh_cost_check(cdr_loop_state* state, double* cost_deltas, const double* XWtZ,
                        const double* ZtZ, const double* WtW, const double* reg_pairs, int k,
                        int n_samples, int n_features, double lambda_W, int stage,
                        int end_of_iteration, cdr_stream_t stream);

/* ------------------------------------------------------------------ furthest sum
 * Greedy furthest-sum selection (furthest_sum.py:23-127) on a T x T
 * dissimilarity matrix; also builds the dissimilarities from a Gram matrix
 * (archetypal_analysis.py:96-100).  selected: int64[k] (device). */
int cdr_dissimilarity_from_gram(const double* K, long ldk, int T, double* D, long ldd,
                                cdr_stream_t stream);
size_t cdr_furthest_sum_workspace_bytes(int T, int k, int extra_steps);
int cdr_furthest_sum(const double* D, long ldd, int T, int k, int start_index,
                     const int64_t* exclude, int n_exclude, int extra_steps, int64_t* selected,
                     void* workspace, size_t workspace_bytes, cdr_stream_t stream);

/* ------------------------------------------------------------------ k-means (Lloyd)
 * Pieces of scikit-learn 1.9.0's KMeans(algorithm='lloyd') around the two
 * streaming passes (the reference calls sklearn.cluster.KMeans at
 * bin/run_hadisst_kmeans.py:128-131):
 *   xct = centres X'        -> cdr_reduce_features
 *   labels: first argmin_j ||c_j||^2 - 2 xct[j][t]   (_k_means_lloyd.pyx:193-213)
 *   sums  = one_hot(labels) X -> cdr_reduce_samples
 *   centres = sums / counts, squared shift            (_k_means_common.pyx:215-262)
 * labels int32[T] is read (previous labels) and written; counts int[k] must be
 * zeroed by the caller; changed is OR-ed with 1 if any label moved. */
int cdr_row_sqnorms(const double* C, long ldc, int k, int d, double* out, cdr_stream_t stream);
int cdr_kmeans_labels(const double* xct, long ldt, const double* cnorm, int T, int k,
                      int32_t* labels, double* onehot, long ldo, int* counts, int* changed,
                      cdr_stream_t stream);
int cdr_kmeans_sqdist(const double* X, long ldx, int T, int d, const double* C, long ldc,
                      const int32_t* labels, double* out, cdr_stream_t stream);
int cdr_kmeans_update(const double* sums, long lds, const double* counts, double* centres,
                      long ldc, int k, int d, double* shift, cdr_stream_t stream);
/* column means / variances in NumPy's row-sequential order (sklearn _kmeans.py:285-294,
 * 1486-1493) and in-place centring X += sign * mean */
int cdr_column_moments(const double* X, long ldx, int T, int d, double* mean, double* var,
                       cdr_stream_t stream);
int cdr_center_columns(double* X, long ldx, int T, int d, const double* mean, double sign,
                       cdr_stream_t stream);

/* One Lloyd iteration (scikit-learn `_kmeans_single_lloyd`, _kmeans.py:620-760) behind one
 * call, stopping rule on the device: E step from the per-strip partials of centres . X', the
 * per-cluster sums as one_hot' X, centre update (with the next ||c_j||^2) and the tests "labels
 * unchanged" / "total shift <= tol" / iteration limit by the last CTA (csrc/kmeans_iter.cu):
 * four kernels.  X is the centred
 * data; `counts` is int[k]; the state block starts zeroed except max_iter and tol_abs.  An
 * empty cluster stops the loop before the centre update with needs_relocation set (the caller
 * relocates and finishes that iteration with the stand-alone calls above).  Returns
 * CDR_ERR_NOT_APPLICABLE for shapes the strip kernels do not cover or k > 16. */
typedef struct cdr_kmeans_state {
    int done;              /* must stay first (the streaming kernels read it as cdr_flags) */
    int n_iter;            /* completed Lloyd iterations */
    int max_iter;
    int strict;            /* labels did not change: strict convergence */
    int needs_relocation;  /* an empty cluster was found */
    int changed;           /* a label changed in the current E step */
    unsigned int ticket;
    int reserved_;
    double tol_abs;        /* tol * mean per-feature variance (_kmeans.py:285-294) */
    double shift_total;
} cdr_kmeans_state;

typedef struct cdr_kmeans_problem {
    const double* X;       /* T x d centred data, leading dimension ldx */
    long ldx;
    int T, d, k;
    double* centres;       /* k x ldx, updated in place */
    int32_t* labels;       /* T, read (previous labels) and written */
    double* onehot;        /* k x ldt */
    long ldt;
    double* sums;          /* k x ldx */
    double* cnorm;         /* k */
    double* shift;         /* k */
    int* counts;           /* k */
    cdr_kmeans_state* state;
    void* workspace;       /* cdr_kmeans_workspace_bytes(T, d, k) */
    size_t workspace_bytes;
} cdr_kmeans_problem;

size_t cdr_kmeans_workspace_bytes(int T, int d, int k);
int cdr_kmeans_fused_applicable(int T, int d, int k);
/* ||c_j||^2 of the current centres, counters cleared: before the first iteration and after the
 * caller changed the centres itself */
int cdr_kmeans_prepare_enqueue(const cdr_kmeans_problem* problem, cdr_stream_t stream);
int cdr_kmeans_iterate_enqueue(const cdr_kmeans_problem* problem, cdr_stream_t stream);

/* ------------------------------------------------------------------ peer-memory collectives
 * The exchange steps of the sample-sharded fit as kernels of this
 * library over NVLink / NVSwitch peer memory, one process per GPU.  They stand where the
 * reference would call nothing at all (it is single-process): sum over ranks of the k x d
 * partials of dictionary.dot(X) / weights.T.dot(X) (archetypal_analysis.py:545-549,
 * gpnh_convex_coding.py:219-224), of the k x k statistics (archetypal_analysis.py:541-556,
 * gpnh_convex_coding.py:292-314), and the gathering of the k x T_local column blocks.
 *
 * Each rank allocates one symmetric region (same size everywhere), exports it, and imports
 * the regions of the other ranks (the 64-byte handles travel through the host-side process
 * group).  The first CDR_PEER_HEADER_BYTES of a region hold flags and counters and must not
 * be used for data.  A buffer that takes part in a collective sits at the same byte offset in
 * every region.  All ranks must issue the same sequence of collective calls with the same
 * sizes.  STATUS: opt-in (CDR_PEER_COLLECTIVES=1); see DESIGN.md section 7.
 */
#define CDR_MAX_PEERS 8
#define CDR_PEER_HEADER_BYTES (1u << 20)
#define CDR_PEER_MAX_CTAS 256    /* grid limit of the collective kernels */
#define CDR_PEER_MAX_STRIPS 2048 /* strips of the fused reduce-over-samples kernel */
#define CDR_PEER_SMALL_MAX 800   /* doubles of a one-shot exchange between kernel tails */

typedef struct cdr_peer_group {
    int world;
    int rank;
    void* region[CDR_MAX_PEERS]; /* region[rank] is the local allocation */
    size_t region_bytes;
    size_t inbox_offset;         /* world + 1 slots of inbox_slot_bytes: push area of the fused
                                    kernel (one slot per sending rank, one for the results) */
    size_t inbox_slot_bytes;     /* >= 4 x the bytes of the exchanged k x ldo matrix: two
                                    alternating sets of 16-byte tagged words per double */
} cdr_peer_group;

int cdr_peer_region_alloc(size_t bytes, void** region);  /* cudaMalloc + clear */
int cdr_peer_region_free(void* region);
int cdr_peer_export(void* region, unsigned char handle[64]);
int cdr_peer_import(const unsigned char handle[64], void** region);
int cdr_peer_release(void* imported_region);
/* synchronising read (and clear) of the local time-out indicator: 0 = no wait timed out */
int cdr_peer_error(const cdr_peer_group* group, int* error_out, cdr_stream_t stream);

/* In-place sum over ranks of the n doubles at `offset` of every region (n even, offset a
 * multiple of 16).  Two-shot: chunk c is summed in rank order by rank c % world, which
 * pulls the peers' partials and pushes the sum to every rank. */
int cdr_peer_allreduce(const cdr_peer_group* group, size_t offset, size_t n,
                       const cdr_flags* flags, cdr_stream_t stream);

/* Every rank pushes its k x ncols block (src: local device pointer, leading dimension lds)
 * into columns [col0, col0 + ncols) of the k x ldd matrix at dst_offset of every region.
 * max_cols = the largest ncols of any rank (sizes the grid identically everywhere). */
int cdr_peer_allgather_columns(const cdr_peer_group* group, const double* src, long lds,
                               size_t dst_offset, long ldd, int k, int col0, int ncols,
                               int max_cols, const cdr_flags* flags, cdr_stream_t stream);

/* cdr_reduce_samples fused with the sum over ranks: out (k x ldo, at out_offset of every
 * region) = sum_r E (L_r X_r).  Each CTA pushes the tile of its feature strip straight from
 * the epilogue into the inbox of the strip's owner rank -- every double as two 8-byte words
 * carrying 32 data bits and a 32-bit epoch tag, so that no fence and no flag message is
 * needed --, the owner sums the world tiles in rank order as they arrive and pushes the
 * result, tagged, to every rank; the kernel ends when the local `out` is complete.  T_min = the smallest local T of any rank (keeps the kernel choice identical
 * on all ranks).  Returns CDR_ERR_NOT_APPLICABLE for shapes the strip kernel does not cover
 * (then: cdr_reduce_samples followed by cdr_peer_allreduce). */
int cdr_reduce_samples_allreduce(const cdr_peer_group* group, const double* Lp, long sLi,
                                 long sLt, const double* X, long ldx, int T, int T_min, int d,
                                 int k, const double* E, size_t out_offset, long ldo,
                                 const cdr_flags* flags, cdr_stream_t stream);

/* ------------------------------------------------------------------ whole outer iterations
 * One call enqueues a complete outer (alternating) iteration on `stream`; the convergence and
 * monotonicity tests run on the device (cdr_loop_state), so a host loop -- or a CUDA graph
 * replayed N times -- only has to look at `state->done` every now and then.  These are the
 * entry points a non-Python binder needs in place of the reference's loops
 *   _iterate_gpnh_convex_coding  (gpnh_convex_coding.py:282-402)
 *   _iterate_aa                  (archetypal_analysis.py:534-670).
 * All pointers are device pointers; X, WT, C ... are padded (leading dimension a multiple of
 * CDR_LD_ALIGN, padding zero).  The `state` block must have been initialised by the caller
 * (tolerance, max_iterations, stopping_rule, require_monotone, trace_data; everything else
 * zero).  Protocol:   prepare_enqueue once, then iterate_enqueue until state->done.
 */
typedef struct cdr_gpnh_problem {
    const double* X;      /* T x d data (this rank's rows), leading dimension ldx */
    long ldx;
    int T, d, k;
    int T_total;          /* = T on a single GPU; all ranks' rows in a sample-sharded fit */
    double lambda_W;
    double* Z;            /* T x k weights, dense, updated in place */
    double* WT;           /* k x ldx: the dictionary transposed (gpnh_convex_coding.py:226), updated in place */
    double* XWt;          /* k x ldt scratch: (X W)' */
    long ldt;             /* >= T, multiple of CDR_LD_ALIGN */
    double* ZtZ;          /* k x k statistics, maintained across iterations */
    double* XWtZ;
    double* WtW;
    double* REG;          /* k x k pairwise ||w_i - w_j||^2 (lambda_W != 0) */
    double* P;            /* k x k solve matrix of the next dictionary step */
    cdr_loop_state* state;
    double* cost_deltas;  /* max_iterations doubles */
    cdr_spg_params weights_params;
    void* workspace;      /* cdr_gpnh_workspace_bytes(T, d, k) */
    size_t workspace_bytes;
    /* sample-sharded fit over peer memory (NULL / world 1: single GPU): WT must live in the
     * symmetric region (same offset on every rank), with an inbox slot of k * ldx doubles;
     * T_min = the smallest local T of any rank.  Only cdr_gpnh_iterate_enqueue on the
     * three-kernel path is offered; the statistics of the first iteration (Z'Z summed over
     * ranks, cost, old_cost, P) are the caller's to form. */
    const cdr_peer_group* peers;
    int T_min;
} cdr_gpnh_problem;

size_t cdr_gpnh_workspace_bytes(int T, int d, int k);
/* gpnh_convex_coding.py:292-314: products and cost of the initial factors, then the start of
 * the first iteration (old_cost, solve matrix P). */
int cdr_gpnh_prepare_enqueue(const cdr_gpnh_problem* problem, cdr_stream_t stream);
/* gpnh_convex_coding.py:339-399: dictionary update, weights update, both cost checks, stopping
 * rule, start of the next iteration.  For k <= 16 at streaming shapes
 * (cdr_gpnh_fused_applicable) this is three kernels: W' = P Z'X (reduce over samples, solve in
 * the epilogue), X W as per-strip partials with W'W as a by-product of the same pass, and the
 * per-sample QPs with the statistics Z'Z, tr(W'X'Z), both cost checks and the next solve
 * matrix formed by the last CTA to finish; XWt / XWtZ are then scratch of
 * cdr_gpnh_prepare_enqueue only.  Other shapes run the general kernel sequence. */
int cdr_gpnh_iterate_enqueue(const cdr_gpnh_problem* problem, cdr_stream_t stream);
int cdr_gpnh_fused_applicable(int T, int d, int k); /* host-only: 1 = three-kernel path */

/* Archetypal analysis in feature space (archetypal_analysis.py:534-670), delta = 0 (no scale
 * factor update), dictionary SPG with 1 <= max_iterations <= 8 inner iterations (longer inner
 * loops need the host to look at the state between bursts: use the cdr_aa_spg_* pieces). */
typedef struct cdr_aa_problem {
    const double* X;      /* T x d data (this rank's rows), leading dimension ldx */
    long ldx;
    int T, d;
    cdr_aa_buffers buf;   /* C, G, D, CK, DK, KZt (k x ldt), alpha, k x k statistics, state;
                             buf.T = all samples (= T on a single GPU) */
    double* Z;            /* T x k weights, dense, updated in place */
    double* tmp_kd;       /* k x ldx scratch: D X, Z'X */
    cdr_spg_params dictionary_params; /* defaults spg.py:46-51 */
    cdr_spg_params weights_params;    /* defaults spg.py:287-291 */
    void* workspace;      /* cdr_aa_workspace_bytes(T, d, k) */
    size_t workspace_bytes;
    /* sample-sharded fit over peer memory (NULL / world 1: single GPU, row0 = 0): this rank
     * owns samples row0 .. row0 + T; the k x buf.T matrices are replicated; DK, KZt and tmp_kd
     * must live in the symmetric region (same offsets on every rank) with an inbox slot of
     * k * ldx doubles; T_min = the smallest local T of any rank.  Only cdr_aa_iterate_enqueue
     * on the eight-kernel path is offered; the products of the first iteration are the
     * caller's to form. */
    int row0;
    int T_min;
    const cdr_peer_group* peers;
} cdr_aa_problem;

size_t cdr_aa_workspace_bytes(int T, int d, int k);
/* archetypal_analysis.py:541-556: C X X', X X' Z, Z'Z, C X X' C', the initial cost; then the
 * projection of the start that spg() applies first (spg.py:146-148), old_cost, and the
 * gradient of the first dictionary step in buf.G. */
int cdr_aa_prepare_enqueue(const cdr_aa_problem* problem, cdr_stream_t stream);
/* archetypal_analysis.py:586-663: dictionary SPG step(s), weights update, both cost checks,
 * stopping rule.  With one inner iteration and k <= 16 at streaming shapes
 * (cdr_aa_fused_applicable) this is eight kernels, see csrc/iterate_aa.cu.  The buffers carry
 * the state from call to call (C, CK, KZt, the k x k statistics and, on the eight-kernel path
 * of a single GPU, the gradient of the next dictionary step in buf.G): call it on what
 * cdr_aa_prepare_enqueue or the previous call left. */
int cdr_aa_iterate_enqueue(const cdr_aa_problem* problem, cdr_stream_t stream);
int cdr_aa_fused_applicable(int T, int d, int k, int dictionary_max_iterations);

#ifdef __cplusplus
}
#endif

#endif /* CDR_B200_H */
