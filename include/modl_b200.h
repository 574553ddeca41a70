/*
 * modl_b200.h -- C ABI of the B200-native MODL minibatch hot path.
 *
 * Every entry point replaces one piece of the reference's per-minibatch inner loop
 * (DictFact._single_batch_fit, modl/decomposition/dict_fact.py:495-526) or one of the
 * compiled helpers that loop imports (dict_fact.py:13-18).  The reference interface each
 * function stands in for is cited as  [ref: file:line].  Paths are relative to the
 * reference checkout (arthurmensch/modl).
 *
 * Conventions
 *   - plain C: pointers, sizes, scalars.  No torch / C++ types.
 *   - `_f32` / `_f64` suffix = element type of every floating array of that call
 *     (the reference's Cython `floating` fused type).
 *   - all array arguments of the device entry points are DEVICE pointers unless the
 *     parameter name starts with `h_` (host).  Matrices are row-major ("C order") with
 *     the leading dimension passed explicitly (`ld*`, in elements).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Calls are asynchronous with respect to the host unless stated otherwise.
 *   - return value: 0 = MODL_OK, otherwise a MODL_E* code; modl_last_error() gives
 *     the text.  The library never throws and never falls back to a CPU path.
 *   - buffers are caller-owned and updated in place, exactly like the ndarray
 *     attributes the reference mutates.
 */
#ifndef MODL_B200_H_
#define MODL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MODL_OK 0
#define MODL_EINVAL 1      /* bad shape / null pointer / unsupported option   */
#define MODL_ECUDA 2       /* a CUDA runtime call or kernel launch failed      */
#define MODL_ENOTSPD 3     /* Cholesky hit a non-positive pivot (posv info>0)  */
#define MODL_ENOMEM 4

typedef struct modl_ctx modl_ctx;         /* per-device context: workspace, SM count     */
typedef struct modl_rng modl_rng;         /* host MT19937 stream                          */
typedef struct modl_sampler modl_sampler; /* host feature-subset sampler                  */

int modl_version(void);
const char *modl_last_error(void);
/* sizeof(modl_step_params), sizeof(modl_fit_params), sizeof(modl_fit_batches) as this library was compiled: a binding
 * that mirrors the structs (ctypes, cgo, JNA ...) checks its own layout against them before the first call. */
void modl_struct_sizes(int64_t *h_out3);

/* ------------------------------------------------------------------------------------
 * Host-side bookkeeping: bit-exact with the reference (integer streams).
 * ---------------------------------------------------------------------------------- */

/* [ref: modl/utils/randomkit/random_fast.pyx:49-74  RandomState(seed), .seed()] */
modl_rng *modl_rs_create(uint64_t seed);
void modl_rs_destroy(modl_rng *rs);
void modl_rs_seed(modl_rng *rs, uint64_t seed);
/* [ref: random_fast.pyx:76-77  RandomState.randint(high) -> uniform integer in [0, high]] */
int64_t modl_rs_randint(modl_rng *rs, uint64_t high);
/* [ref: random_fast.pyx:146-147  RandomState.binomial(n, p)] */
int64_t modl_rs_binomial(modl_rng *rs, int64_t n, double p);
/* [ref: random_fast.pyx:79-85  RandomState.permutation(size)] */
void modl_rs_permutation(modl_rng *rs, int64_t *h_out, int64_t size);
/* [ref: random_fast.pyx:87-125  RandomState.shuffle(x) for a 1-D integer array] */
void modl_rs_shuffle(modl_rng *rs, int64_t *h_x, int64_t n);
/* [ref: random_fast.pyx:127-144  RandomState.shuffle_with_trace(list)]: draws the swap
 * list once; h_trace[i] = original position of the row that ends at position i, so
 * every array of the list becomes x[h_trace].  h_swap may be NULL. */
void modl_rs_shuffle_with_trace(modl_rng *rs, int64_t n, int64_t *h_swap, int64_t *h_trace);

/* [ref: modl/utils/randomkit/sampler.pyx:10-39  Sampler(range, rand_size, replacement, seed)] */
modl_sampler *modl_sampler_create(int64_t range, int rand_size, int replacement, uint64_t seed);
void modl_sampler_destroy(modl_sampler *s);
/* [ref: sampler.pyx:41-69  Sampler.yield_subset(reduction)]: writes the (unsorted) subset
 * into h_out (capacity >= range) and returns its length. */
int64_t modl_sampler_yield_subset(modl_sampler *s, double reduction, int64_t *h_out);

/* [ref: modl/decomposition/dict_fact_fast.pyx:115-122  _batch_weight] */
double modl_batch_weight(int64_t count, int64_t batch_size, double learning_rate, double offset);

/* ------------------------------------------------------------------------------------
 * Device context
 * ---------------------------------------------------------------------------------- */
int modl_ctx_create(int device, modl_ctx **out);
void modl_ctx_destroy(modl_ctx *ctx);
int modl_ctx_sm_count(const modl_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t modl_ctx_launch_count(const modl_ctx *ctx);
/* tunables: "bcd_cluster" (largest thread-block cluster the dictionary update may use, 0 = none),
 * "cd_warps" (warps per CTA of the CD kernel, 0 = auto), "force_global_gram" (debug),
 * "tc_gemm" (1 = float32 contractions on tcgen05 with the 3xTF32 split, 0 = CUDA-core FFMA GEMM),
 * "bcd_pilot" (0 = plain per-atom cluster kernel instead of the look-ahead pilot kernel), "bcd_coop_min_cols" (fewest
 * columns per CTA of the grid-wide dictionary update, default 128), "bcd_flag_barrier" (1 = per-CTA epoch flags instead
 * of one atomic counter at its grid barrier; measured slower, off), "bcd_timing" (debug stamps). */
int modl_ctx_set_option(modl_ctx *ctx, const char *name, int value);
/* Synchronises `stream`, then returns MODL_ENOTSPD if a Cholesky pivot was non-positive since
 * the last check (LAPACK posv info > 0, which the reference ignores:
 * dict_fact_fast.pyx:88-92, 188-192), else MODL_OK.  Clears the flag. */
int modl_ctx_check_info(modl_ctx *ctx, void *stream);

/* Per-phase device timing of modl_batch_fit_* (CUDA events on the caller's stream; while
 * enabled every step ends with a host synchronisation, so use it in a dedicated pass).
 * modl_ctx_profile(ctx, 1) resets the counters; modl_ctx_profile_read copies the accumulated
 * milliseconds per phase (h_ms[MODL_PROF_PHASES]) and the number of steps measured. */
enum {
    MODL_PROF_GATHER = 0,   /* subset upload, column gathers, row norms          */
    MODL_PROF_GRAM = 1,     /* G and Dx products                                  */
    MODL_PROF_AVERAGE = 2,  /* Dx_average_/G_average_ running averages            */
    MODL_PROF_CODE = 3,     /* coordinate descent / Cholesky code solve           */
    MODL_PROF_STATS = 4,    /* C_ and B_ accumulation                             */
    MODL_PROF_DICT_PREP = 5,/* B_[:, subset] gather, G_ downdate, order upload    */
    MODL_PROF_DICT_BCD = 6, /* sequential atom update kernel                      */
    MODL_PROF_DICT_POST = 7,/* scatter into components_, G_ update                */
    MODL_PROF_PHASES = 8
};
int modl_ctx_profile(modl_ctx *ctx, int enable);
int modl_ctx_profile_read(modl_ctx *ctx, double *h_ms, int64_t *h_steps);

/* ------------------------------------------------------------------------------------
 * Elastic-net ball helpers   [ref: modl/utils/math/enet.pxd:10-16]
 * ---------------------------------------------------------------------------------- */
/* enet_norm(v, l1_ratio) for `rows` vectors of length n (row r at v + r*ld); out[rows].
 * [ref: modl/utils/math/enet.pyx:125-148] */
int modl_enet_norm_f32(modl_ctx *, const float *v, int64_t rows, int64_t n, int64_t ld,
                       float l1_ratio, float *out, void *stream);
int modl_enet_norm_f64(modl_ctx *, const double *v, int64_t rows, int64_t n, int64_t ld,
                       double l1_ratio, double *out, void *stream);
/* enet_projection(v, out, radius, l1_ratio) per row; radius[rows] is a device array.
 * `out` may alias `v`.  [ref: enet.pyx:38-122] */
int modl_enet_projection_f32(modl_ctx *, const float *v, float *out, int64_t rows, int64_t n,
                             int64_t ld, const float *radius, float l1_ratio, void *stream);
int modl_enet_projection_f64(modl_ctx *, const double *v, double *out, int64_t rows, int64_t n,
                             int64_t ld, const double *radius, double l1_ratio, void *stream);
/* enet_scale(X[r], l1_ratio, radius) in place for every row.  [ref: enet.pyx:150-168] */
int modl_enet_scale_f32(modl_ctx *, float *X, int64_t rows, int64_t n, int64_t ld,
                        float l1_ratio, float radius, void *stream);
int modl_enet_scale_f64(modl_ctx *, double *X, int64_t rows, int64_t n, int64_t ld,
                        double l1_ratio, double radius, void *stream);

/* ------------------------------------------------------------------------------------
 * Code computation   [ref: DictFact._compute_code, dict_fact.py:577-648]
 * ---------------------------------------------------------------------------------- */
/* Gathered Gram / correlation products [ref: dict_fact.py:589-604, and :67-71 for the
 * un-subsampled transform case]:
 *     G  (k x k, ld k) = scale * D[:, subset] . D[:, subset]^T      (if G  != NULL)
 *     Dx (b x k, ld k) = scale * X[:, subset] . D[:, subset]^T      (if Dx != NULL)
 *     xnorm2[b]        = squared L2 norm of every FULL row of X     (if xnorm2 != NULL)
 * subset: device int64[s]; NULL means "all p features" (s is ignored).
 * D is k x p (ldd), X is b x p (ldx). */
int modl_gram_dx_f32(modl_ctx *, const float *D, int64_t ldd, const float *X, int64_t ldx,
                     const int64_t *subset, int64_t s, int64_t k, int64_t b, int64_t p,
                     float scale, float *G, float *Dx, float *xnorm2, void *stream);
int modl_gram_dx_f64(modl_ctx *, const double *D, int64_t ldd, const double *X, int64_t ldx,
                     const int64_t *subset, int64_t s, int64_t k, int64_t b, int64_t p,
                     double scale, double *G, double *Dx, double *xnorm2, void *stream);

/* Batched elastic-net regression over ONE shared Gram matrix
 * [ref: _enet_regression_single_gram, dict_fact_fast.pyx:125-215].
 *   G k x k (symmetric; its lower triangle is read), Dx b x k (OVERWRITTEN by the ridge
 *   branch, like the reference :188-197), code n x k (rows indices[ii] are read as the warm
 *   start and written), indices int64[b] (NULL = arange(b)).
 *   The CD stop test needs |x_i|^2 of the full data row [ref: :334-336]: pass either the
 *   rows X (b x p, ldx) or precomputed xnorm2[b]; X may be NULL when xnorm2 is given.
 *   l1_ratio == 0 -> Cholesky ridge branch [ref: :174-197]; else cyclic coordinate descent
 *   [ref: :198-214 -> enet_coordinate_descent_gram :270-427].
 *   sweeps (int32[b], optional): sweeps executed per sample (diagnostic; 0 for ridge). */
int modl_enet_regression_single_gram_f32(modl_ctx *, const float *G, float *Dx, const float *X,
                                         int64_t ldx, int64_t p, const float *xnorm2, float *code,
                                         const int64_t *indices, int64_t b, int64_t k,
                                         float l1_ratio, float alpha, int positive, float tol,
                                         int max_iter, int32_t *sweeps, void *stream);
int modl_enet_regression_single_gram_f64(modl_ctx *, const double *G, double *Dx, const double *X,
                                         int64_t ldx, int64_t p, const double *xnorm2, double *code,
                                         const int64_t *indices, int64_t b, int64_t k,
                                         double l1_ratio, double alpha, int positive, double tol,
                                         int max_iter, int32_t *sweeps, void *stream);

/* Same with one Gram matrix per sample, G is b x k x k
 * [ref: _enet_regression_multi_gram, dict_fact_fast.pyx:33-113].  The ridge branch works on
 * a copy (the reference destroys the upper triangle of G[ii], :82-94; we leave G intact). */
int modl_enet_regression_multi_gram_f32(modl_ctx *, const float *G, float *Dx, const float *X,
                                        int64_t ldx, int64_t p, const float *xnorm2, float *code,
                                        const int64_t *indices, int64_t b, int64_t k,
                                        float l1_ratio, float alpha, int positive, float tol,
                                        int max_iter, int32_t *sweeps, void *stream);
int modl_enet_regression_multi_gram_f64(modl_ctx *, const double *G, double *Dx, const double *X,
                                        int64_t ldx, int64_t p, const double *xnorm2, double *code,
                                        const int64_t *indices, int64_t b, int64_t k,
                                        double l1_ratio, double alpha, int positive, double tol,
                                        int max_iter, int32_t *sweeps, void *stream);

/* G_average[ii] = (1 - w[ii]) * G_average[ii] + w[ii] * G for the rows `indices` of the
 * n x k x k array [ref: _update_G_average, dict_fact_fast.pyx:217-228 and its caller
 * dict_fact.py:605-618 which gathers/scatters the rows; indices NULL = rows 0..b-1]. */
int modl_update_G_average_f32(modl_ctx *, float *G_average, const float *G, const float *w_sample,
                              const int64_t *indices, int64_t b, int64_t k, void *stream);
int modl_update_G_average_f64(modl_ctx *, double *G_average, const double *G, const double *w_sample,
                              const int64_t *indices, int64_t b, int64_t k, void *stream);
/* Dx_average[idx] = (1-w)*Dx_average[idx] + w*Dx ; Dx <- Dx_average[idx]
 * [ref: dict_fact.py:596-601]. */
int modl_update_Dx_average_f32(modl_ctx *, float *Dx_average, float *Dx, const float *w_sample,
                               const int64_t *indices, int64_t b, int64_t k, void *stream);
int modl_update_Dx_average_f64(modl_ctx *, double *Dx_average, double *Dx, const double *w_sample,
                               const int64_t *indices, int64_t b, int64_t k, void *stream);

/* ------------------------------------------------------------------------------------
 * Surrogate statistics   [ref: DictFact._update_C / _update_B, dict_fact.py:559-575]
 *   C_ = (1-w) C_ + (w/b) code^T code   (k x k)
 *   B_ = (1-w) B_ + (w/b) code^T X      (k x p, all p columns)
 * `code` is the b x k batch code (rows gathered from code_[indices] when indices != NULL,
 * code then being the full n x k array).  overwrite != 0 selects the 'sgd' branch
 * (C_ = code^T code / b, :574-575, :565-566).
 * ---------------------------------------------------------------------------------- */
int modl_update_stats_f32(modl_ctx *, const float *code, const int64_t *indices, const float *X,
                          int64_t ldx, float *C, float *B, int64_t ldb, double w, int64_t b,
                          int64_t k, int64_t p, int overwrite, void *stream);
int modl_update_stats_f64(modl_ctx *, const double *code, const int64_t *indices, const double *X,
                          int64_t ldx, double *C, double *B, int64_t ldb, double w, int64_t b,
                          int64_t k, int64_t p, int overwrite, void *stream);

/* ------------------------------------------------------------------------------------
 * Dictionary update   [ref: DictFact._update_dict, dict_fact.py:650-715]
 * Block coordinate descent over the atoms in `h_order` (host int64[k], the permutation the
 * caller drew from its NumPy RandomState, :672) restricted to the feature `subset`
 * (device int64[s]):  gathers components_[:, subset] and B_[:, subset] (the reference's
 * gradient_[:, subset] = B_[:, subset], :532), runs the sequential atom updates with the
 * elastic-net-ball projection and comp_norm_ slack accounting (:675-694), scatters the
 * panel back (:709) and, when G_full != NULL (G_agg == 'full'), down/up-dates G_ (:667-668,
 * :711-715).
 *   mode 0 = 'variational' (:675-694); mode 1 = 'sgd' (:695-708) using w and step_size.
 * ---------------------------------------------------------------------------------- */
int modl_update_dict_f32(modl_ctx *, float *components, int64_t ldd, const float *B, int64_t ldb,
                         const float *C, float *comp_norm, float *G_full,
                         const int64_t *subset, int64_t s, const int64_t *h_order,
                         int64_t k, int64_t p, float comp_l1_ratio, int comp_pos,
                         int mode, double w, double step_size, void *stream);
int modl_update_dict_f64(modl_ctx *, double *components, int64_t ldd, const double *B, int64_t ldb,
                         const double *C, double *comp_norm, double *G_full,
                         const int64_t *subset, int64_t s, const int64_t *h_order,
                         int64_t k, int64_t p, double comp_l1_ratio, int comp_pos,
                         int mode, double w, double step_size, void *stream);

/* ------------------------------------------------------------------------------------
 * The fused minibatch step   [ref: DictFact._single_batch_fit, dict_fact.py:507-522]
 * One call = subset upload, gathers + Gram products, optional running averages, code
 * solve, statistics update, dictionary update.  Host bookkeeping (sampler draw, n_iter_,
 * sample_n_iter_, w, w_sample, atom order) is done by the caller with the host entry
 * points above and passed in.
 * ---------------------------------------------------------------------------------- */
enum { MODL_AGG_MASKED = 0, MODL_AGG_FULL = 1, MODL_AGG_AVERAGE = 2 };

typedef struct modl_step_params {
    /* shapes */
    int64_t n_samples, n_features, n_components, batch_size;
    /* batch inputs */
    const void *X;             /* device b x p */
    int64_t ldx;
    const int64_t *indices;    /* device int64[b] (rows of code_/averages), NULL = arange */
    const int64_t *h_subset;   /* HOST int64[s]: sampler output for this batch           */
    int64_t subset_len;
    const int64_t *h_order;    /* HOST int64[k]: atom permutation                        */
    const void *w_sample;      /* device real[b] (only read by the 'average' modes)      */
    double w;                  /* batch weight                                            */
    /* estimator state (device, in place) */
    void *components;          /* k x p */
    void *code;                /* n x k */
    void *C;                   /* k x k */
    void *B;                   /* k x p */
    void *comp_norm;           /* k     */
    void *G_full;              /* k x k, G_agg == 'full' only                            */
    void *Dx_average;          /* n x k, Dx_agg == 'average' only                        */
    void *G_average;           /* n x k x k, G_agg == 'average' only                     */
    /* hyper-parameters (DictFact.__init__, dict_fact.py:128-153) */
    double reduction, code_alpha, code_l1_ratio, comp_l1_ratio, tol, step_size;
    int max_iter, code_pos, comp_pos, Dx_agg, G_agg, optimizer_sgd;
    /* diagnostics (optional, device int32[b]) */
    int32_t *sweeps;
    /* Sample-sharded data parallelism (one process per GPU).  `phases` selects which parts of
     * the step this call runs (0 = all four, the single-GPU case).  With stats_inc != NULL the
     * STATS phase does not touch C_/B_ but writes this rank's share of the increments
     *     stats_inc[0 : k*k]      = (w / global_batch) code^T code
     *     stats_inc[k*k : k*k+k*p] = (w / global_batch) code^T X
     * which the caller sums over ranks (NCCL all-reduce over NVLink) before the APPLY phase
     * folds them in:  C_ = (1-w) C_ + sum,  B_ = (1-w) B_ + sum.  batch_size is the number of
     * rows THIS rank holds; global_batch (0 = batch_size) the rows of the whole minibatch. */
    int phases;
    int64_t global_batch;
    void *stats_inc;           /* device real[k*k + k*p]                                  */
    /* Overlapped exchange (optional).  The dictionary update only needs C_ and the subset columns of
     * B_, so with inc_sub != NULL the STATS phase also writes the compact buffer
     *     inc_sub[0 : k*k]           = stats_inc[0 : k*k]
     *     inc_sub[k*k : k*k + k*lds] = stats_inc B-part restricted to the subset columns (row pitch
     *                                  lds = subset_len rounded up to 4)
     * which the caller all-reduces FIRST (1.5 MB instead of 10.5 MB at the benchmark shape).
     * MODL_PHASE_APPLY_SUB folds it in:  C_ = (1-w) C_ + sum,  and leaves the panel
     * (1-w) B_[:, subset] + sum in the workspace for the next MODL_PHASE_DICT of this context;
     * B_ itself is updated by MODL_PHASE_APPLY_B from the full all-reduced stats_inc, which the
     * caller may run on another stream while the dictionary update is in flight. */
    void *inc_sub;             /* device real[k*k + k*lds]                                */
    /* cudaEvent_t (or NULL) recorded on the call's stream right after MODL_PHASE_APPLY_SUB: "B_[:, subset]
     * has been read", so that APPLY_SUB and DICT can share one call while another stream waits for exactly
     * that point before it rewrites B_. */
    void *ev_after_apply_sub;
    /* Two-stream schedule of modl_partial_fit_* (below).  `slot` (0/1) selects one of two copies of the per-step
     * device inputs (feature subset, atom order, B_[:, subset] panel) so that MODL_PHASE_PREFETCH of step t+1 may
     * run while the dictionary update of step t still reads its own.  sm_avail (0 = all): SMs the full-width
     * product may count on while the dictionary update's cluster is resident.  start_flag (device uint32, or
     * NULL): the dictionary-update kernel stores start_serial there as soon as its CTAs are resident, which is
     * what the second stream waits for before it launches anything that could take those SMs. */
    int slot;
    int sm_avail;
    void *start_flag;
    uint32_t start_serial;
    /* 1: h_subset / h_order sit in PINNED (mapped) host memory that stays untouched until the step has run; the
     * device then fetches them with a small kernel that reads the host memory directly instead of two host->device
     * copies -- copies would queue on the copy engine behind a 20 MB batch upload in flight (0.4 ms). */
    int h_inputs_mapped;
} modl_step_params;

/* MODL_PHASE_STATS_SUB (same call as MODL_PHASE_CODE) computes ONLY what the dictionary update waits
 * for, straight into inc_sub:  (w/batch) code^T code  and  (w/batch) code^T X[:, subset]  -- a k x s
 * product instead of k x p.  MODL_PHASE_STATS_B (a call of its own, any stream) then produces the
 * full-width statistic: B_ = (1-w) B_ + (w/batch) code^T X when stats_inc == NULL, or the raw
 * increment into stats_inc (for the caller's all-reduce + MODL_PHASE_APPLY_B) otherwise.  Run on a
 * second stream it hides the 2.6 GFLOP / 60 MB product behind the sequential dictionary update. */
enum { MODL_PHASE_CODE = 1, MODL_PHASE_STATS = 2, MODL_PHASE_APPLY = 4, MODL_PHASE_DICT = 8,
       MODL_PHASE_APPLY_SUB = 16, MODL_PHASE_APPLY_B = 32, MODL_PHASE_STATS_SUB = 64, MODL_PHASE_STATS_B = 128,
       /* the feature subset uploaded by the previous call of this context belongs to the same step: skip the upload */
       MODL_PHASE_REUSE_SUBSET = 256,
       /* A call of its own, on any stream, ahead of the step: everything that depends on the batch, the subset and
        * the atom order but not on the dictionary -- index uploads, the X side of the subset gather (packed rows, row
        * norms), the packed X[:, subset]^T operand and (with FUSED_APPLY) the B_[:, subset] panel. */
       MODL_PHASE_PREFETCH = 512,
       /* flag: this step's MODL_PHASE_PREFETCH has completed (ordered by the caller's events) */
       MODL_PHASE_INPUTS_READY = 1024,
       /* flag, one GPU: STATS_SUB folds its product straight into C_ and into the B_[:, subset] panel
        * ([panel | C_] = keep [panel | C_] + (w/b) code^T [X_sub | code], one tensor-core launch); no APPLY_SUB */
       MODL_PHASE_FUSED_APPLY = 2048,
       /* flag, with FUSED_APPLY and INPUTS_READY: the B_[:, subset] panel is gathered by THIS call (the step's PREFETCH ran
        * without FUSED_APPLY, while the previous step's full-width product could still be writing B_) */
       MODL_PHASE_GATHER_B = 4096 };

int modl_batch_fit_f32(modl_ctx *, const modl_step_params *prm, void *stream);
int modl_batch_fit_f64(modl_ctx *, const modl_step_params *prm, void *stream);

/* ------------------------------------------------------------------------------------
 * The minibatch loop   [ref: DictFact.partial_fit, dict_fact.py:313-337, over _single_batch_fit :495-526]
 *
 * One call = every `batch_size` block of the rows passed to partial_fit: per block the host statements of the
 * reference in their order (feature subset from the sampler :507, n_iter_ :509, sample_n_iter_ :510, batch
 * weight :515, the atom order the caller drew :672) and the device phases of modl_batch_fit_* on two streams:
 * the dictionary update (a 16-CTA cluster, sequential over atoms) only needs C_ and the SUBSET columns of B_
 * (:532), so the full-width product B_ = (1-w) B_ + w/b code^T X (:559-566) and the next block's input
 * preparation run on a second stream on the SMs the dictionary update leaves idle.  Results are those of
 * calling modl_batch_fit_* block by block (B_[:, subset] panel and B_ agree to rounding of two equivalent
 * products).  Rows may come from device memory or from host memory (staged through three device slots on a
 * copy stream, so that the copy of block i+1 overlaps the kernels of block i, within a call and across calls).
 * ---------------------------------------------------------------------------------- */
typedef struct modl_fit modl_fit;     /* streams, events, staging slots and (optionally) NCCL communicators of one estimator */

typedef struct modl_fit_params {       /* the estimator: state arrays (device, updated in place) and hyper-parameters */
    int64_t n_samples, n_features, n_components, batch_size;
    void *components, *code, *C, *B, *comp_norm, *G_full, *Dx_average, *G_average;
    double reduction, learning_rate, code_alpha, code_l1_ratio, comp_l1_ratio, tol, step_size;
    int max_iter, code_pos, comp_pos, Dx_agg, G_agg, optimizer_sgd;
} modl_fit_params;

typedef struct modl_fit_batches {      /* one partial_fit call */
    const void *X;                     /* n_rows x n_features, leading dimension ldx */
    int64_t ldx, n_rows;
    int x_location;                    /* 0 = device, 1 = pinned host, 2 = pageable host */
    const int64_t *h_sample_indices;   /* HOST int64[n_rows]: rows of code_ / the running averages (NULL = 0..n_rows-1) */
    modl_sampler *sampler;             /* feature_sampler_ of the estimator */
    int64_t *h_n_iter;                 /* HOST: n_iter_, updated */
    int64_t *h_sample_n_iter;          /* HOST int64[n_samples]: sample_n_iter_, updated (may be NULL) */
    int update_counters;               /* 0: the caller has already bumped both counters for these rows */
    const int64_t *h_orders;           /* HOST int64[n_batches x k]: one atom permutation per block, drawn by the caller from
                                          its NumPy RandomState in block order [ref: :672] */
    const void *h_w_sample;            /* HOST real[n_rows]: sample_n_iter_ ** -sample_learning_rate [ref: :513]; 'average' modes */
    int32_t *sweeps;                   /* device int32[batch_size] or NULL: CD sweeps per sample of the LAST block */
    int64_t *h_last_subset;            /* HOST int64[n_features] or NULL: receives the last block's feature subset */
    int64_t *h_last_subset_len;
    void *h_code_out;                  /* HOST (pinned) real[n_rows x k] or NULL: every block's code, copied out on a stream
                                          of its own as soon as the block's solve has finished */
    int wait_host;                     /* 1: return once X has been consumed and h_code_out is complete; 0: fully asynchronous
                                          (the caller keeps X untouched until modl_fit_synchronize) */
    int fence;                         /* do the loop's own streams wait for what the caller enqueued on `stream` before this
                                          call?  0 = automatic: yes for device rows (they may be the output of a pending kernel),
                                          no for host rows; 1 = yes (the caller wrote a state array or X with device work since the
                                          last call); 2 = no (device rows that are final: nothing pending on `stream` writes them).
                                          The first call of a loop object always waits.  The wait serialises the next block's
                                          input preparation behind the previous call's dictionary update (both are on `stream`). */
} modl_fit_batches;

int modl_fit_create(modl_ctx *ctx, modl_fit **out);
void modl_fit_destroy(modl_fit *fit);
/* "overlap" (1 = the two-stream schedule above, 0 = one fused call per block on the caller's stream),
 * "gate" (1 = the second stream starts only once the dictionary-update kernel is resident),
 * "graph" (on one GPU, after the first four blocks, the launches of a block are stream-captured, folded into the previous
 * block's executable CUDA graph with cudaGraphExecUpdate and launched as a graph -- same kernels, same results:
 * 1 (default) = the three calls of the two-stream schedule as three graphs; 2 = one graph per block on the caller's
 * stream, the full-width product forked inside it; 0 = plain stream launches.  Blocks whose dictionary update runs on the
 * cooperative grid -- panels larger than one thread-block cluster -- keep plain launches whatever the option). */
int modl_fit_set_option(modl_fit *fit, const char *name, int value);
/* h_out[0] = graph launches so far, h_out[1] = executable graphs rebuilt because the block's topology changed,
 * h_out[2] = calls run with plain launches because they could not be captured (a workspace slot had to grow),
 * h_out[3] = times the full-width product gave up waiting (20 ms) for the dictionary kernel to become resident
 * (always 0 unless a dictionary kernel failed to start).  Synchronises the device. */
int modl_fit_graph_stats(modl_fit *fit, int64_t *h_out4);
/* Blocks until the loop's own streams (copies, second stream, code read-back) are idle. */
int modl_fit_synchronize(modl_fit *fit);
/* Debug: timeline of the next `steps` steps of the two-stream schedule (0 = off), then
 * modl_fit_trace_read: h_ms[step][10] = milliseconds since arming at which each point was reached on its stream
 * (NaN = not recorded): 0 H2D start, 1 H2D end (copy stream); 2 PREFETCH start, 3 PREFETCH end (second stream);
 * 4 critical path start, 5 codes + subset statistics done, 6 dictionary update done (caller's stream); 7 full-width
 * product start, 8 full-width product end (second stream); 9 code read-back done. */
int modl_fit_trace(modl_fit *fit, int steps);
int modl_fit_trace_read(modl_fit *fit, double *h_ms, int capacity_steps, int *h_steps);
int modl_partial_fit_f32(modl_fit *fit, const modl_fit_params *est, const modl_fit_batches *io, void *stream);
int modl_partial_fit_f64(modl_fit *fit, const modl_fit_params *est, const modl_fit_batches *io, void *stream);

/* Sample-sharded data parallelism, one process per GPU (the reference's analogue is its thread pool over
 * slices of a batch, dict_fact.py:584-586, 621-635).  Every rank passes ITS rows of each block (the same count on
 * every rank); the loop sums the statistics increments over ranks with ncclAllReduce -- the k x (k + s) part the
 * dictionary update waits for on the caller's stream, the k x p part of B_ on the second stream behind the
 * dictionary update -- and every rank applies the same deterministic dictionary update to bit-identical inputs.
 * modl_nccl_unique_id: rank 0 draws two ids (128 bytes each) and ships them to the other ranks by any means
 * (torch.distributed broadcast); modl_fit_set_comm is collective. */
int modl_nccl_unique_id(void *h_out, int64_t capacity);
int modl_fit_set_comm(modl_fit *fit, int world, int rank, const void *h_id_main, const void *h_id_side);

/* ------------------------------------------------------------------------------------
 * Recsys: the missing-value path   [ref: RecsysDictFact, modl/decomposition/recsys.py:147-213,
 * :254-265; _predict, modl/decomposition/recsys_fast.pyx:10-37]   
 *
 * X is a CSR matrix resident on the device: indptr int64[n+1], indices int32[nnz], data[nnz].
 * The dictionary is held twice: components (k x p, ldd) for modl_update_dict_*, and its transpose
 * Dt (p x k, ldt) for the column gathers below; modl_recsys_sync_transposed_* refreshes the rows
 * `subset` of Dt (all p rows when subset == NULL) after a dictionary update.  1 <= k <= 128.
 * ---------------------------------------------------------------------------------- */

/* Per-row Gram and correlation over the row's observed columns [ref: _single_sample_update
 * recsys.py:169-177 and _refit :254-264]:  G[ii] = D_sub D_sub^T + (alpha / reduction) I,
 * Dx[ii] = D_sub x_sub with reduction = p / len(observed).  Rows: rows[ii] (device int64[b]), or
 * row0 + ii when rows == NULL.  The solve itself (recsys.py:178, :265) is
 * modl_enet_regression_multi_gram_* with l1_ratio = 0, alpha = 0 on (G, Dx). */
int modl_recsys_gram_dx_f32(modl_ctx *, const float *Dt, int64_t ldt, const int64_t *indptr,
                            const int32_t *indices, const float *data, const int64_t *rows, int64_t row0,
                            int64_t b, int64_t k, int64_t p, double alpha, float *G, float *Dx, void *stream);
int modl_recsys_gram_dx_f64(modl_ctx *, const double *Dt, int64_t ldt, const int64_t *indptr,
                            const int32_t *indices, const double *data, const int64_t *rows, int64_t row0,
                            int64_t b, int64_t k, int64_t p, double alpha, double *G, double *Dx, void *stream);

/* The B_ update of one minibatch [ref: recsys.py:168, :182-185].  The batch's stored entries are given
 * ordered by (column, position of the row in the batch): subset int64[s] = the sorted distinct columns
 * (recsys.py:159-161), col_ptr int64[s+1] their entry ranges, entry_row int64[] the row of code (n x k)
 * each entry belongs to, entry_val its value.  Per column, in that order:
 *   feature_n_iter[j] += 1;  w_B = min(1, w * n_iter / feature_n_iter[j]);
 *   B[:, j] = (1 - w_B) B[:, j] + code[row] * (x * w_B)        (float64 arithmetic, rounded like NumPy's) */
int modl_recsys_update_B_f32(modl_ctx *, float *B, int64_t ldb, const float *code, const int64_t *subset,
                             const int64_t *col_ptr, const int64_t *entry_row, const float *entry_val,
                             int64_t *feature_n_iter, int64_t s, int64_t k, double w, int64_t n_iter, void *stream);
int modl_recsys_update_B_f64(modl_ctx *, double *B, int64_t ldb, const double *code, const int64_t *subset,
                             const int64_t *col_ptr, const int64_t *entry_row, const double *entry_val,
                             int64_t *feature_n_iter, int64_t s, int64_t k, double w, int64_t n_iter, void *stream);

/* C_ = (1 - w) C_ + (w / b) code[rows]^T code[rows]   [ref: recsys.py:156-157]; rows NULL = rows 0..b-1 */
int modl_recsys_update_C_f32(modl_ctx *, float *C, const float *code, const int64_t *rows, int64_t b, int64_t k,
                             double w, void *stream);
int modl_recsys_update_C_f64(modl_ctx *, double *C, const double *code, const int64_t *rows, int64_t b, int64_t k,
                             double w, void *stream);

/* Dt[subset[j], :] = components[:, subset[j]] (the scatter of recsys.py:213 seen from the transposed copy) */
int modl_recsys_sync_transposed_f32(modl_ctx *, const float *components, int64_t ldd, float *Dt, int64_t ldt,
                                    const int64_t *subset, int64_t s, int64_t k, void *stream);
int modl_recsys_sync_transposed_f64(modl_ctx *, const double *components, int64_t ldd, double *Dt, int64_t ldt,
                                    const int64_t *subset, int64_t s, int64_t k, void *stream);

/* out[e] = <code[u], components[:, indices[e]]> for the stored entries e of every row u
 * [ref: _predict, recsys_fast.pyx:10-37]; out is float64 like the reference's X_data. */
int modl_recsys_predict_f32(modl_ctx *, const float *code, const float *Dt, int64_t ldt, const int64_t *indptr,
                            const int32_t *indices, int64_t n_rows, int64_t k, double *out, void *stream);
int modl_recsys_predict_f64(modl_ctx *, const double *code, const double *Dt, int64_t ldt, const int64_t *indptr,
                            const int32_t *indices, int64_t n_rows, int64_t k, double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MODL_B200_H_ */
