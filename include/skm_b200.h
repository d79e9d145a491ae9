/*
 * skm_b200.h -- C ABI of libskm_b200.so, the B200 (sm_100a) engine for the
 * sparsified K-means hot path of stephenbeckr/SparsifiedKMeans.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / MATLAB
 * types.  Citations are relative to the reference repository root.
 *
 * Two levels:
 *
 *  Level 1 -- stateless replacements for the reference's five MEX gateways.
 *    Host pointers in, host pointers out, IEEE double arithmetic in the
 *    reference's operation order (bit-identical results).  A MEX shim unpacks
 *    its mxArrays and calls exactly one of these (see mex/ and INTEGRATION.md).
 *
 *  Level 2 -- resident handles.  The sparsified matrix is uploaded once
 *    (skm_dataset_*), a Lloyd state is bound to it (skm_lloyd_*), and every
 *    iteration runs on the GPU: masked distance + argmin (replaces
 *    private/findClusterAssignments.m:60-83,168-171 +
 *    private/SparseMatrixMinusCluster.c:117-183), per-cluster sums and support
 *    counts (replaces kmeans_sparsified.m:430-453), centre finalisation
 *    (kmeans_sparsified.m:448,470-471).  Multi-GPU: each process owns a
 *    contiguous block of columns; skm_lloyd_partials() exposes the device
 *    buffer [S | N | counts | sumsq] that the host all-reduces (NCCL) between
 *    skm_lloyd_accumulate() and skm_lloyd_finalize().
 *
 * Conventions: all matrices column-major (MATLAB); points are COLUMNS of the
 * p x n sparse matrix X (as inside kmeans_sparsified.m after :213-218); CSC
 * row indices are 0-based and sorted ascending within a column (what MATLAB
 * stores); assignments returned to the host are 1-based like MATLAB's.
 *
 * Every function returns SKM_OK (0) or an error code; the message is available
 * from skm_last_error().  Nothing here long-jumps or calls mex*; no host
 * pointer is retained after a call returns.  Calls on one context are
 * synchronous with respect to the host unless stated otherwise.
 */
#ifndef SKM_B200_H
#define SKM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKM_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------ */
#define SKM_OK               0
#define SKM_ERR_INVALID      1   /* bad argument; maps to the reference's mexErrMsgTxt usage errors */
#define SKM_ERR_CUDA         2   /* CUDA runtime / launch failure */
#define SKM_ERR_NOMEM        3
#define SKM_ERR_UNSUPPORTED  4
#define SKM_ERR_STATE        5   /* call order violated (e.g. accumulate before assign) */

/* ---- element types ----------------------------------------------------- */
#define SKM_F32 0
#define SKM_F64 1
#define SKM_I32 2
#define SKM_I64 3   /* also used for MATLAB mwIndex (uint64 < 2^63) */
#define SKM_U16 4   /* row indices only (p <= 65536): 2 bytes per entry over PCIe */

typedef struct skm_ctx     skm_ctx;
typedef struct skm_dataset skm_dataset;
typedef struct skm_lloyd   skm_lloyd;

/* ---- context ----------------------------------------------------------- */

int         skm_abi_version(void);
/* Bind to CUDA device `device`.  `cuda_stream` may be NULL (the library creates
 * its own non-blocking stream) or a cudaStream_t owned by the caller. */
int         skm_ctx_create(int device, void *cuda_stream, skm_ctx **out);
void        skm_ctx_destroy(skm_ctx *ctx);
/* Message of the last failure on this thread ("" if none). ctx may be NULL. */
const char *skm_last_error(const skm_ctx *ctx);
void       *skm_ctx_stream(skm_ctx *ctx);          /* the cudaStream_t in use */
int         skm_ctx_device(const skm_ctx *ctx);
int         skm_ctx_sync(skm_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t     skm_ctx_launch_count(const skm_ctx *ctx);

/* Chunks of skm_second_pass the tensor-core filter (tcgen05 tf32 product + exact evaluation of its candidates)
 * handled so far on this context, and chunks after which it was switched off for the rest of a call because more
 * than an eighth of the points stayed uncertain (data without cluster structure). */
int         skm_ctx_tc_chunks(const skm_ctx *ctx, int64_t *kept, int64_t *dropped);

/* Per-kernel device timing with CUDA events on the context stream (off by default).
 * skm_ctx_timing_read synchronises, writes the summed milliseconds and launch-group counts
 * of the 8 slots {assign, recheck, accumulate, finalize, prep, fwht, kpp, upload} since the
 * previous read into ms[8] / counts[8], and resets them. */
int         skm_ctx_timing_enable(skm_ctx *ctx, int on);
int         skm_ctx_timing_read(skm_ctx *ctx, double *ms, int64_t *counts);

/* ---- Level 1: stateless, exact fp64, host buffers ----------------------- */

/* dist = SparseMatrixMinusCluster(X, c [,beta])   (private/SparseMatrixMinusCluster.c:2-11,44-47)
 *   dist[k + K*j] = sqrt( sum_{t in col j} (x[t] - c[ir[t] + p*k])^2 ), terms added in stored
 *   order (:169-182).  has_beta != 0 selects the K==1-only variant of :118-129.
 *   jc[n+1], ir[nnz] are MATLAB mwIndex (uint64). */
int skm_sparse_matrix_minus_cluster(skm_ctx *ctx, int64_t p, int64_t n, int64_t K,
                                    const uint64_t *jc, const uint64_t *ir, const double *pr,
                                    const double *centers, int has_beta, double beta,
                                    double *dist);

/* [innerProd, normX2] = SparseMatrixInnerProduct(X, c)  (private/SparseMatrixInnerProduct.c:2-9,87-100)
 *   normsq may be NULL.  c has length p (the reference checks it against n, :71-77; we check p). */
int skm_sparse_matrix_inner_product(skm_ctx *ctx, int64_t p, int64_t n,
                                    const uint64_t *jc, const uint64_t *ir, const double *pr,
                                    const double *c, double *inner, double *normsq);

/* normX2 = SparseMatrixColumnNormSq(X)  (private/SparseMatrixColumnNormSq.c:2-8,71-77) */
int skm_sparse_matrix_column_normsq(skm_ctx *ctx, int64_t p, int64_t n,
                                    const uint64_t *jc, const double *pr, double *normsq);

/* w = hadamard(x) and w = hadamard_pthreads(x): unnormalised Sylvester-ordered
 * Walsh-Hadamard transform of each column of the m x n matrix x, m a power of
 * two >= 2 (private/hadamard.c:57-92,97-111; private/hadamard_pthreads.c:69-90). */
int skm_hadamard(skm_ctx *ctx, int64_t m, int64_t n, const double *x, double *w);

/* ---- Level 2: resident data -------------------------------------------- */

typedef struct skm_dataset_info {
    int64_t p, n, nnz;
    int64_t max_col_nnz;     /* largest number of stored entries in a column */
    int32_t store_dtype;     /* SKM_F32 (fast path) or SKM_F64 (exact path) */
    int32_t reserved;
    int64_t device_bytes;    /* HBM held by this dataset */
    int64_t stream_bytes;    /* bytes the assignment kernel streams per pass */
} skm_dataset_info;

/* Upload a p x n CSC matrix.  jc has n+1 entries of type jc_type (SKM_I32/SKM_I64),
 * ir has nnz entries of type ir_type (SKM_I32/SKM_I64/SKM_U16), val has nnz entries of type val_type
 * (SKM_F32/SKM_F64).  on_device != 0: the three pointers are device pointers on
 * ctx's device (used when the data was produced on the GPU).
 * store_dtype SKM_F64 keeps doubles and every operation is bit-exact with the
 * reference; SKM_F32 rounds values to float (the "identical input" for parity
 * is then that rounded value) and enables the fast assignment kernel, whose
 * assignments are still bit-identical to the reference's on those inputs
 * (uncertified columns are re-evaluated in fp64 in the reference order). */
int  skm_dataset_create_csc(skm_ctx *ctx, int64_t p, int64_t n,
                            const void *jc, int jc_type, const void *ir, int ir_type,
                            const void *val, int val_type, int store_dtype, int on_device,
                            skm_dataset **out);
/* Same, with the number of centres the caller is going to use (0: unknown).  For large host matrices the upload is
 * pipelined in column chunks (conversion, validation and the counting pass of the row-major image follow the upload
 * chunk by chunk); with K_hint > 0 the entry order of the kernel family that serves K centres is built chunk by chunk
 * while the upload is still in flight instead of at the first skm_lloyd_assign. */
int  skm_dataset_create_csc_hint(skm_ctx *ctx, int64_t p, int64_t n,
                                 const void *jc, int jc_type, const void *ir, int ir_type,
                                 const void *val, int val_type, int store_dtype, int on_device,
                                 int64_t K_hint, skm_dataset **out);
/* In-place production of an SKM_F32 dataset: allocate the device CSC arrays (int64 colptr[n+1], int32 rowidx[nnz],
 * float val[nnz]), let the caller's own kernels fill them (skm_dataset_csc_ptrs returns the device pointers), then
 * skm_dataset_commit validates them (sorted rows in range, colptr[n] == nnz) and builds the streamed images.  A
 * shard that fills most of the HBM never exists twice this way (skm_dataset_create_csc with on_device copies). */
int  skm_dataset_alloc_csc(skm_ctx *ctx, int64_t p, int64_t n, int64_t nnz, skm_dataset **out);
int  skm_dataset_csc_ptrs(skm_dataset *ds, void **colptr, void **rowidx, void **val);
int  skm_dataset_commit(skm_dataset *ds);
/* min(X(:)) and max(X(:)) over all p*n elements, implicit zeros included (Start='uniform', kmeans_sparsified.m:388-390). */
int  skm_dataset_minmax(skm_dataset *ds, double *mn, double *mx);
void skm_dataset_destroy(skm_dataset *ds);
int  skm_dataset_get_info(const skm_dataset *ds, skm_dataset_info *info);
/* Verification / statistics of the streamed (SELL-32) image the fast assignment kernel reads.
 * layout: -1 = check the current image, 0 / 1 / 2 = first re-order it for kernel family 0 (16-byte
 * gathers, single table, greedy quarter-warp order), 1 (8-byte gathers, dual table, exact
 * edge-coloured half-warp order) or 2 (16-byte gathers, dual table, edge-coloured quarter-warp
 * order); 1 and 2 need columns of at most 254 entries.
 * out[0] = columns whose image differs from the CSC matrix (must be 0), out[1] = gather steps,
 * out[2] = shared-memory wavefronts those steps cost (== out[1] when conflict-free),
 * out[3] = layout of the image. */
int  skm_dataset_layout_check(skm_dataset *ds, int layout, int64_t out[4]);
/* Densify column j (0-based) into out[p] (kmeans_sparsified.m:436, X(:,iMax)). */
int  skm_dataset_get_column(skm_dataset *ds, int64_t j, double *out);

/* Operator: [assignments, distances] = findClusterAssignments(X, centers, [], gamma)
 * for sparse X and dense centers (private/findClusterAssignments.m:77-80,168-171).
 * centers: host, p x K.  has_gamma != 0: centres are divided by gamma first (:78).
 * assign_out (1-based) / dist_out (Euclidean, not squared) may each be NULL. */
int  skm_assign(skm_dataset *ds, const double *centers, int64_t K, int has_gamma, double gamma,
                int32_t *assign_out, double *dist_out);

/* Same operator, sparse-centres branch (private/findClusterAssignments.m:63-75):
 * structural zeros of `centers` (exact 0.0) are outside each centre's support;
 * per centre k the sum runs over rows in supp(X(:,j)) ∩ supp(c_k), the values are
 * divided by nnz(c_k)/p and the centre by gamma when has_gamma != 0. Exact fp64. */
int  skm_assign_sparse_centers(skm_dataset *ds, const double *centers, int64_t K,
                               int has_gamma, double gamma,
                               int32_t *assign_out, double *dist_out);

/* Full K x n distance matrix for a resident dataset (exact fp64); dist is a HOST buffer. */
int  skm_masked_distances(skm_dataset *ds, const double *centers, int64_t K, double *dist);

typedef struct skm_iter_stats {
    double  dff;            /* ||C_old - C_new||_F                 kmeans_sparsified.m:470 */
    double  sumsq;          /* sum_j distance_j^2  (obj = sqrt)    kmeans_sparsified.m:471 */
    int64_t n_empty;        /* clusters with no members (left untouched by finalize) */
    int64_t n_rechecked;    /* columns the fast kernel could not certify (re-done in fp64) */
    int64_t n_points;       /* points accumulated (global after the all-reduce) */
    int32_t has_nan;        /* any NaN in the new centres           kmeans_sparsified.m:480 */
    int32_t reserved;
} skm_iter_stats;

int   skm_lloyd_create(skm_dataset *ds, int64_t K, skm_lloyd **out);
void  skm_lloyd_destroy(skm_lloyd *L);
int   skm_lloyd_set_centers(skm_lloyd *L, const double *centers /* host p x K */);
int   skm_lloyd_get_centers(skm_lloyd *L, double *centers /* host p x K */);
/* centres as they were before the last skm_lloyd_finalize (centersOld, kmeans_sparsified.m:428) */
int   skm_lloyd_get_centers_old(skm_lloyd *L, double *centers /* host p x K */);
int   skm_lloyd_set_center_column(skm_lloyd *L, int64_t k, const double *col /* host p */);
/* K1: masked distance + argmin of every local column against the current centres
 * (divided by gamma when has_gamma).  Asynchronous on the context stream. */
int   skm_lloyd_assign(skm_lloyd *L, int has_gamma, double gamma);
/* Same, sparse-centres branch (private/findClusterAssignments.m:63-75): exact zeros of the
 * current centres are structural zeros.  Exact fp64.  Synchronous. */
int   skm_lloyd_assign_sparse(skm_lloyd *L, int has_gamma, double gamma);
/* K2: zero the partials and add this shard's per-cluster row sums S (p x K),
 * support counts N (p x K), member counts (K) and sum of squared distances (1).
 * Asynchronous on the context stream. */
int   skm_lloyd_accumulate(skm_lloyd *L);
/* How skm_lloyd_assign finds the nearest centre.  0 (default): every centre is evaluated for every column in
 * every call.  1: bounded (SKM_F32 datasets) -- a lower bound on the distance to every centre but the
 * assigned one is carried from call to call and lowered by the movement of the centres (the masked distance
 * is a seminorm of the centre, so it changes by at most ||c_new - c_old||); a column whose distance to its
 * own centre stays below that bound, with the fast kernel's rounding guard and a 1e-6 relative margin, keeps
 * its assignment after ONE centre evaluation; the others are re-evaluated against every centre (lists of 2048
 * columns and more by an fp32 list kernel with the fast kernel's guard first; what that cannot certify, and short
 * lists, in fp64 in the reference's order).  Assignments and distances are the same as in mode 0.  The bounds are dropped whenever
 * the mode is set. */
int   skm_lloyd_set_assign_mode(skm_lloyd *L, int mode);
/* Columns the last bounded skm_lloyd_assign had to re-evaluate (-1: that call evaluated every column). */
int   skm_lloyd_last_assign(skm_lloyd *L, int64_t *n_flagged);
/* Partial-distance pruning of the full assignment pass (plans that need several launches, K > 16).  -1 (default):
 * automatic, 0: off, 1: always try.  A pass over the first ~15 % of every column's stored entries for all K centres
 * gives a candidate winner and -- every term of the masked distance being non-negative -- a lower bound on the distance
 * to every other centre; the candidate is then evaluated exactly on all entries and kept iff it stays below that bound
 * (rounding guards on both sides).  The prefix runs on a half-precision centre table (all K <= 64 centres in one launch)
 * whose rounding is part of the bound.  Columns that cannot be kept are evaluated against every centre (fp32 list kernel
 * with the usual guard, then fp64 in the reference's order for what it cannot certify) when they are fewer than n/4
 * (n/16 for K > 128); otherwise the ordinary full pass runs and the pruned pass sits out 1, 2, 4, ... 32 calls.
 * Assignments and distances are the same as without it. */
int   skm_lloyd_set_prune(skm_lloyd *L, int mode);
/* Columns the last pruned pass could not keep (-1: the last pass was not pruned) and the entry pairs it read per column. */
int   skm_lloyd_last_prune(skm_lloyd *L, int64_t *not_kept, int64_t *pairs);
/* Which kernels the full assignment pass of skm_lloyd_assign runs.  -1 (default): automatic.  1: the tensor-core
 * plan (SKM_F32 datasets, 2 <= K <= 128, p <= 4096; never chosen automatically: it measured 4.9 ms against 5.35 ms
 * per pass at K = 64 and needs a fourth image of X): every column is
 * densified on the fly into an fp16 operand tile and the K scores  sum_{r in supp} (c_rk^2 - 2 x_r c_rk)  come
 * out of tcgen05.mma (private/SparseMatrixMinusCluster.c:169-182 expanded); that product only FILTERS -- the best
 * centre is then evaluated exactly (same fp32 sum and rounding guard as the gather kernels), a rigorous bound on
 * the filter's rounding excludes the others, near-ties are settled among the best three in fp32 and whatever is
 * left in fp64 in the reference's order.  Assignments are the same as with 0 (the gather kernels of mode 0
 * evaluate every centre in fp32 for every column). */
int   skm_lloyd_set_tc_filter(skm_lloyd *L, int mode);
/* After skm_lloyd_finalize / refresh_diff: columns the exact evaluation of the filter's winner could not keep, and
 * columns that went on to the fp64 kernel, in the last tensor-core pass (-1: the last pass was not one). */
int   skm_lloyd_last_tc(skm_lloyd *L, int64_t *not_kept, int64_t *not_resolved);
/* Test hook: runs the tensor-core pass for the current centres and returns the raw filter scores, scaled back,
 * as [n][bn] floats (bn = 32, 64 or 128 by K; scores_host must hold n * 128 floats to be safe). */
int   skm_debug_tc_scores(skm_lloyd *L, int has_gamma, double gamma, float *scores_host, int64_t *bn_out);
/* How skm_lloyd_accumulate obtains the sums.  0 (default): recompute from all columns every iteration, as
 * the reference does (kmeans_sparsified.m:430-453).  1: incremental -- the per-shard sums of the previous
 * iteration are kept and only the columns whose assignment changed move their entries between clusters
 * (the sums depend on nothing but the assignments, so the centres are the same up to fp64 rounding; every
 * 64th iteration, or when more than n/16 columns moved, the sums are recomputed in full).  The sum of squared
 * distances and the member counts are always exact for the current assignment. */
int   skm_lloyd_set_update_mode(skm_lloyd *L, int mode);
/* What the last skm_lloyd_accumulate did: kind 0 = full recompute, 1 = moved n_changed columns,
 * 2 = no assignment changed (n_changed = 0).  n_changed = -1 after a full recompute. */
int   skm_lloyd_last_update(skm_lloyd *L, int *kind, int64_t *n_changed);
/* Device pointer / length (in doubles) of [S | N | counts | sumsq] for the collective. */
void *skm_lloyd_partials(skm_lloyd *L, int64_t *n_doubles);
/* K3: C(:,k) = gamma*S(:,k)./(N(:,k)+1e-16) (ml_correction) or S(:,k)/count_k, for
 * non-empty clusters (kmeans_sparsified.m:448,450); fills stats (synchronises). */
int   skm_lloyd_finalize(skm_lloyd *L, double gamma, int ml_correction, skm_iter_stats *stats);
/* Recompute dff / has_nan against the centres as they are now, without an update: used
 * after the host patched columns for EmptyAction (kmeans_sparsified.m:432-445,470). */
int   skm_lloyd_refresh_diff(skm_lloyd *L, skm_iter_stats *stats);
int   skm_lloyd_get_counts(skm_lloyd *L, int64_t *counts /* host K, after finalize */);
/* Local results of the last skm_lloyd_assign (1-based assignments; either may be NULL). */
int   skm_lloyd_get_assignments(skm_lloyd *L, int32_t *assign_out, double *dist_out);
/* First local column attaining the largest distance (kmeans_sparsified.m:435). */
int   skm_lloyd_argmax_distance(skm_lloyd *L, double *maxdist, int64_t *j);
/* Name of the kernel skm_lloyd_assign runs for this state (reporting only; thread-local string). */
const char *skm_lloyd_kernel_name(skm_lloyd *L);
/* Device pointers to the local results (int32 0-based assignments, float/double distances). */
void *skm_lloyd_assign_ptr(skm_lloyd *L);
void *skm_lloyd_dist_ptr(skm_lloyd *L, int *dtype);

/* One Lloyd iteration over a matrix that lives in HOST memory (the stateless form of the path,
 * and the out-of-core form for matrices larger than HBM): columns are streamed in chunks of
 * `chunk_cols` (0 = automatic) over PCIe on a copy stream while the previous chunk is laid out,
 * assigned (K1 + fp64 re-evaluation) and accumulated (K2) on the compute stream; K3 runs at the
 * end.  Same arithmetic and exactness guarantee as skm_lloyd_assign/accumulate/finalize on an
 * SKM_F32 dataset.  Pinned host buffers overlap best.  assign_out (1-based) / dist_out may be
 * NULL; centers_out receives the updated centres (empty clusters keep their input column). */
/* `reduce` (may be NULL) is called once, after the last chunk, with the device buffer
 * [S | N | counts | sumsq] and the compute stream: a multi-GPU caller all-reduces it there
 * (asynchronously on that stream) and returns 0. */
typedef int (*skm_reduce_fn)(void *partials_dev, int64_t n_doubles, void *cuda_stream, void *user);
int   skm_lloyd_step_host(skm_ctx *ctx, int64_t p, int64_t n, const void *jc, int jc_type,
                          const void *ir, int ir_type, const void *val, int val_type,
                          const double *centers, int64_t K, int has_gamma, double gamma_dist,
                          double gamma_update, int ml_correction, int64_t chunk_cols,
                          double *centers_out, int32_t *assign_out, double *dist_out,
                          skm_iter_stats *stats, skm_reduce_fn reduce, void *reduce_user);

/* Second pass over the ORIGINAL dense data (kmeans_sparsified.m:542-560 in core;
 * private/recalculateAssignmentLargeFile.m:85-113 out of core -- the same two computations chunk by chunk):
 *   centers_out(:,k) = mean of the columns j with assign_in[j] == k+1 (zero column for an empty cluster;
 *                      label 0 = unassigned is skipped); counts_out[k] = members          (:545-551)
 *   [assign_out, dist_out] = nearest column of `centers` in plain Euclidean distance: the dense
 *                      branch of private/findClusterAssignments.m:124-171 (gamma is not used there) (:558)
 * x: dense p x n column-major (points are columns), SKM_F32/SKM_F64, in HOST memory (streamed over
 * PCIe in column chunks on a copy stream; a memory-mapped file works) or, x_on_device != 0, in device
 * memory.  Every value is multiplied by `scale` first (the reference's X*(1+2*eps), :292; pass 1.0 for none).
 * Either half may be skipped: centers_out == NULL (then assign_in may be NULL) or assign_out == dist_out
 * == NULL (then centers may be NULL).  Distances are evaluated in fp32 as sum (x-c)^2 with the same
 * rounding guard as the sparsified kernel; columns whose winner cannot be certified are re-evaluated in
 * fp64 (n_rechecked, may be NULL).  assign_out is 1-based. */
int   skm_second_pass(skm_ctx *ctx, int64_t p, int64_t n, const void *x, int x_type, int x_on_device,
                      double scale, const double *centers, int64_t K, const int32_t *assign_in,
                      double *centers_out, int64_t *counts_out, int32_t *assign_out, double *dist_out,
                      int64_t chunk_cols, int64_t *n_rechecked);

/* k-means++ support (private/Arthur_initialization.m:39-53): fold the masked
 * distance to ONE new centre into the running minimum kept on the device.
 * first != 0 resets the running minimum.  sum_d2 receives sum_j mind_j^2 over
 * the local columns. */
int   skm_kpp_update(skm_dataset *ds, const double *center /* host p */, int has_gamma,
                     double gamma, int first, double *sum_d2);
/* Same for the gamma-less call of private/Arthur_initialization.m:26 (unbiasedInitialization = false): the
 * chosen centres stay SPARSE there, so findClusterAssignments takes its sparse-centres branch
 * (private/findClusterAssignments.m:70-74): the sum runs over supp(X(:,j)) /\ supp(center) only -- exact zeros
 * of `center` are structural zeros. */
int   skm_kpp_update_sparse(skm_dataset *ds, const double *center /* host p */, int first, double *sum_d2);
/* Smallest local column index j with cumsum(mind^2)[j] > target (clamped to n-1). */
int   skm_kpp_pick(skm_dataset *ds, double target, int64_t *j);
int   skm_kpp_get_mindist(skm_dataset *ds, double *mind /* host n */);

/* ---- several GPUs driven by ONE host process -----------------------------------
 * The reference is one MATLAB process calling its MEX gateway on the main thread; a MATLAB session cannot be
 * torchrun.  skm_multi owns one context and one host worker thread per device, shards the columns of X in
 * contiguous blocks (device g owns [g*n/G, (g+1)*n/G)) and replaces the per-iteration all-reduce + K3 by one
 * kernel per device that reads every peer's partials [S | N | counts | sumsq] through NVLink peer memory, sums
 * them in device order (bit-identical centres on every device) and finalises the centres in the same pass.
 * All calls are synchronous for the caller and may be made from any single host thread.  (One process per GPU
 * with an NCCL all-reduce of skm_lloyd_partials() remains available: sparsifiedkmeans_b200/distributed.py.) */
typedef struct skm_multi         skm_multi;
typedef struct skm_multi_dataset skm_multi_dataset;
typedef struct skm_multi_lloyd   skm_multi_lloyd;

/* ndev <= 0: every visible device; devices == NULL: devices 0..ndev-1. */
int      skm_multi_create(int ndev, const int *devices, skm_multi **out);
void     skm_multi_destroy(skm_multi *m);
int      skm_multi_ndev(const skm_multi *m);
skm_ctx *skm_multi_ctx(skm_multi *m, int i);
int      skm_multi_peer_access(const skm_multi *m);   /* 1: partials are read through peer memory; 0: staged copies */

/* skm_dataset_create_csc for a host matrix, column blocks uploaded concurrently (one worker per device). */
int  skm_multi_dataset_create_csc(skm_multi *m, int64_t p, int64_t n, const void *jc, int jc_type,
                                  const void *ir, int ir_type, const void *val, int val_type,
                                  int store_dtype, skm_multi_dataset **out);
/* Adopt ndev datasets built any other way (shards[g] on skm_multi_ctx(m, g), in column order, e.g. from
 * skm_dataset_from_dense_host with col0 = the block's first column); ownership moves to the result. */
int  skm_multi_dataset_from_shards(skm_multi *m, skm_dataset *const *shards, skm_multi_dataset **out);
void skm_multi_dataset_destroy(skm_multi_dataset *md);
skm_dataset *skm_multi_dataset_shard(skm_multi_dataset *md, int i, int64_t *col0);
int  skm_multi_dataset_get_info(const skm_multi_dataset *md, skm_dataset_info *info);   /* totals over the shards */
int  skm_multi_dataset_get_column(skm_multi_dataset *md, int64_t j /* global */, double *out);
/* k-means++ rounds over the shards (private/Arthur_initialization.m:39-53); sparse_center selects
 * skm_kpp_update_sparse.  skm_multi_kpp_pick returns the GLOBAL column whose cumulative D^2 passes target. */
int  skm_multi_kpp_update(skm_multi_dataset *md, const double *center, int has_gamma, double gamma, int first,
                          int sparse_center, double *sum_d2);
int  skm_multi_kpp_pick(skm_multi_dataset *md, double target, int64_t *j);

int  skm_multi_lloyd_create(skm_multi_dataset *md, int64_t K, skm_multi_lloyd **out);
void skm_multi_lloyd_destroy(skm_multi_lloyd *L);
int  skm_multi_lloyd_set_modes(skm_multi_lloyd *L, int update_mode, int assign_mode);   /* skm_lloyd_set_*_mode per shard */
int  skm_multi_lloyd_set_centers(skm_multi_lloyd *L, const double *centers /* host p x K */);
int  skm_multi_lloyd_set_center_column(skm_multi_lloyd *L, int64_t k, const double *col);
int  skm_multi_lloyd_get_centers(skm_multi_lloyd *L, double *centers);
int  skm_multi_lloyd_get_centers_old(skm_multi_lloyd *L, double *centers);
int  skm_multi_lloyd_get_centers_of(skm_multi_lloyd *L, int device_index, double *centers);
/* One Lloyd iteration (kmeans_sparsified.m:420-471): K1 + K2 on every shard concurrently, then the fused
 * peer-memory reduction + K3 on every device.  sparse_centers != 0 takes the sparse-centres branch of the
 * operator (private/findClusterAssignments.m:63-75).  stats are global. */
int  skm_multi_lloyd_step(skm_multi_lloyd *L, int has_gamma, double gamma_dist, double gamma_update,
                          int ml_correction, int sparse_centers, skm_iter_stats *stats);
int  skm_multi_lloyd_refresh_diff(skm_multi_lloyd *L, skm_iter_stats *stats);
int  skm_multi_lloyd_get_counts(skm_multi_lloyd *L, int64_t *counts /* host K, global */);
/* assignments (1-based) / distances of all n columns, in column order */
int  skm_multi_lloyd_get_assignments(skm_multi_lloyd *L, int32_t *assign_out, double *dist_out);
int  skm_multi_lloyd_argmax_distance(skm_multi_lloyd *L, double *maxdist, int64_t *j /* global, first occurrence */);
int  skm_multi_lloyd_launch_count(skm_multi_lloyd *L, int64_t *launches);

/* ---- preconditioning on device (kmeans_sparsified.m:238-248,286-295) ------ */

/* Y = hadamard(D * [X; 0]) / sqrt(p2) for a dense p x n host matrix (fp64 in,
 * computed in `compute_dtype`, fp64 out, p2 x n).  signs has p2 entries of +-1. */
int   skm_mix_hadamard(skm_ctx *ctx, int64_t p, int64_t p2, int64_t n, const double *x,
                       const double *signs, int compute_dtype, double *y);

/* Fused precondition + row sample, all on device, producing a resident dataset:
 * column j of the result keeps m distinct rows of hadamard(D*x_j)/sqrt(p2), each divided by
 * (m/p2) (private/randsample_fixedNumberEntries.m:30-31,62).  x_dev is a device pointer to a
 * dense p2 x n float matrix (column-major).  rows_dev: device int32[m*n], column j keeps
 * rows[m*j .. m*j+m) (0-based, distinct, any order) -- or NULL, and the rows are drawn on the
 * device: uniform without replacement (the contract of private/randsample_block.m:44-84) from
 * Philox4x32-10 keyed by `seed` and counted by the GLOBAL column index col0 + j, so a column's
 * sample does not depend on the sharding.  Exact zeros are kept as stored entries with value 0. */
int   skm_fwht_sample_f32(skm_ctx *ctx, int64_t p2, int64_t n, int64_t m, const float *x_dev,
                          const float *signs_dev, const int32_t *rows_dev, uint64_t seed, int64_t col0,
                          skm_dataset **out);
/* The row sets skm_fwht_sample_f32 draws for (seed, col0): rows_dev int32[m*n], ascending per column. */
int   skm_sample_rows(skm_ctx *ctx, int64_t p2, int64_t n, int64_t m, uint64_t seed, int64_t col0,
                      int32_t *rows_dev);
/* The whole precondition + sample stage of kmeans_sparsified.m:286-334 from a dense HOST matrix
 * x (p x n column-major, x_type SKM_F32/SKM_F64, points are columns): column chunks cross PCIe
 * on a copy stream, are zero-padded to p2 rows, scaled by (1+2*eps) (:292), cast to fp32 and run
 * through the fused sign-flip + FWHT + /sqrt(p2) + on-device row sample (m rows, divided by
 * m/p2) straight into a resident SKM_F32 dataset.  signs: host double[p2] of +-1. */
int   skm_dataset_from_dense_host(skm_ctx *ctx, int64_t p, int64_t p2, int64_t n, const void *x, int x_type,
                                  const double *signs, int64_t m, uint64_t seed, int64_t col0,
                                  int64_t chunk_cols, skm_dataset **out);
/* ---- DCT sketch (kmeans_sparsified.m:226-231,256-258: chosen when p is not a power of two) ---- */

/* y = dct(D .* x) (inverse == 0: mix, :295) or y = D .* idct(x) (inverse != 0: unmix, :296) for a dense
 * p x n HOST matrix, column-major, fp64; dct is MATLAB's orthonormal DCT-II along columns.  signs
 * (+-1, length p) may be NULL.  Applied as one dense fp64 product on the GPU (hand-written tiled kernel). */
int   skm_dct_mix(skm_ctx *ctx, int64_t p, int64_t n, const double *x, const double *signs, int inverse, double *y);
/* skm_dataset_from_dense_host for the DCT sketch (no zero-padding: p2 = p): chunks cross PCIe, are
 * multiplied by T*diag(signs)*(1+2eps) on the tensor cores (tcgen05 tf32, both operands split in tf32-exact halves:
 * 3xTF32, fp32 accuracy) and m rows per column are kept, divided by m/p.
 * rows_host: int32[m*n] explicit 0-based rows per column (any order, distinct) or NULL = drawn on the
 * device (Philox4x32-10 keyed by seed, counted by the global column col0+j). */
int   skm_dataset_from_dense_host_dct(skm_ctx *ctx, int64_t p, int64_t n, const void *x, int x_type,
                                      const double *signs, int64_t m, uint64_t seed, int64_t col0,
                                      const int32_t *rows_host, int64_t chunk_cols, skm_dataset **out);
/* The rows skm_dataset_from_dense_host_dct draws: rows_dev int32[m*n] (device), ascending per column;
 * any 1 <= m <= p <= 32768 (skm_sample_rows is the power-of-two variant fused into the FWHT). */
int   skm_sample_rows_general(skm_ctx *ctx, int64_t p, int64_t n, int64_t m, uint64_t seed, int64_t col0,
                              int32_t *rows_dev);
/* In-place device FWHT of a dense p2 x n float matrix with sign flip and 1/sqrt(p2). */
int   skm_fwht_f32_inplace(skm_ctx *ctx, int64_t p2, int64_t n, float *x_dev,
                           const float *signs_dev);

#ifdef __cplusplus
}
#endif
#endif /* SKM_B200_H */
