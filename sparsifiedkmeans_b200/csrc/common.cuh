// common.cuh -- shared declarations of libskm_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/skm_b200.h"

#define SKM_SLICE 32          // columns per SELL slice == warp width

void skm_set_error(const char *fmt, ...);
struct skm_ctx;
// Big device arrays (the images of a dataset): cudaMalloc / cudaFree of gigabytes cost ~5 ms per GB and synchronise the
// device; the stream-ordered pool keeps freed blocks for the next dataset (release threshold SKM_POOL_RETAIN_GB, default
// 48) and frees without a device-wide stall.  Falls back to cudaMalloc when pools are unavailable or SKM_NO_POOL is set.
int  skm_big_alloc(skm_ctx *ctx, void **ptr, size_t bytes, const char *what);
void skm_big_free(skm_ctx *ctx, void *ptr);

#define SKM_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            skm_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),  \
                          __FILE__, __LINE__);                                      \
            return SKM_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

#define SKM_CHECK_LAUNCH(ctx)                                                       \
    do {                                                                            \
        (ctx)->launches++;                                                          \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) {                                                   \
            skm_set_error("kernel launch failed: %s (%s:%d)",                       \
                          cudaGetErrorString(e__), __FILE__, __LINE__);             \
            return SKM_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

#define SKM_TRY(expr)                                                               \
    do {                                                                            \
        int rc__ = (expr);                                                          \
        if (rc__ != SKM_OK) return rc__;                                            \
    } while (0)

#define SKM_REQUIRE(cond, ...)                                                      \
    do {                                                                            \
        if (!(cond)) {                                                              \
            skm_set_error(__VA_ARGS__);                                             \
            return SKM_ERR_INVALID;                                                 \
        }                                                                           \
    } while (0)

// per-kernel device timing (CUDA events on the launching stream), off by default
enum { SKM_T_ASSIGN = 0, SKM_T_RECHECK = 1, SKM_T_ACCUM = 2, SKM_T_FINAL = 3, SKM_T_PREP = 4,
       SKM_T_FWHT = 5, SKM_T_KPP = 6, SKM_T_UPLOAD = 7, SKM_T_SLOTS = 8 };
#define SKM_T_RING 512

#define SKM_RED_BLOCKS 4096
struct skm_ctx {
    int          device;
    cudaStream_t stream;
    bool         own_stream;
    int64_t      launches;
    int          sm_count;
    int          smem_optin;       // max dynamic shared memory per block (bytes)
    int         *d_flag;           // device int[4] scratch (validation / counters)
    int         *h_flag;           // pinned host mirror
    double      *red_scratch;      // [SKM_RED_BLOCKS] per-block partials of the order-independent reductions (update.cu)
    unsigned int *red_ticket;      // their ticket counter (zero between launches)
    bool         timing;
    cudaEvent_t (*ev)[SKM_T_RING][2];   // [SKM_T_SLOTS][SKM_T_RING][2], created lazily
    int          ev_count[SKM_T_SLOTS];
    void        *stream_cache;              // staging buffers of skm_lloyd_step_host, reused across calls
    void       (*stream_cache_free)(void *);
    int64_t      tc_chunks;                 // chunks the tensor-core filter of the second pass ran on (and kept)
    int64_t      tc_chunks_dropped;         // chunks after which it was switched off (too many uncertain points)
    bool         use_pool;                  // big arrays come from the device's stream-ordered memory pool
    void        *bounce[2];                 // pinned staging for read-backs into pageable memory (lazily allocated)
    cudaEvent_t  bounce_ev[2];
};
// device -> pageable host memory through two pinned staging buffers (a direct cudaMemcpy into fresh pageable memory
// runs at ~4 GB/s); synchronises the stream
int skm_d2h_pageable(skm_ctx *ctx, void *dst, const void *src_dev, size_t bytes);

// RAII: records a start event now and a stop event at scope exit when timing is enabled
struct SkmTimed {
    skm_ctx *ctx; int slot; int idx;
    SkmTimed(skm_ctx *c, int s) : ctx(c), slot(s), idx(-1) {
        if (!c->timing || !c->ev || c->ev_count[s] >= SKM_T_RING) return;
        idx = c->ev_count[s]++;
        cudaEventRecord(c->ev[s][idx][0], c->stream);
    }
    ~SkmTimed() { if (idx >= 0) cudaEventRecord(ctx->ev[slot][idx][1], ctx->stream); }
};

// RAII device buffer used for temporaries inside one API call.
struct DevBuf {
    void  *ptr = nullptr;
    size_t bytes = 0;
    DevBuf() {}
    ~DevBuf() { if (ptr) cudaFree(ptr); }
    int alloc(size_t n) {
        if (ptr) { cudaFree(ptr); ptr = nullptr; }
        bytes = n;
        if (n == 0) return SKM_OK;
        cudaError_t e = cudaMalloc(&ptr, n);
        if (e != cudaSuccess) {
            ptr = nullptr;
            skm_set_error("cudaMalloc(%zu bytes) failed: %s", n, cudaGetErrorString(e));
            cudaGetLastError();
            return SKM_ERR_NOMEM;
        }
        return SKM_OK;
    }
    template <class T> T *as() const { return (T *)ptr; }
    void *release() { void *p = ptr; ptr = nullptr; return p; }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

struct skm_dataset {
    skm_ctx *ctx;
    int64_t  p, n, nnz, max_col_nnz;
    int      store_dtype;               // SKM_F32 | SKM_F64
    // CSC in stored order (always present)
    int64_t *colptr;                    // [n+1]
    int32_t *rowidx;                    // [nnz]
    void    *val;                       // float[nnz] or double[nnz]
    // SELL-32 image for the fast kernel (SKM_F32 only):
    //   slice s holds columns [32s, 32s+32); its entries are interleaved so that lane l's
    //   pair t2 = (row_{2t2}, val_{2t2}, row_{2t2+1}, val_{2t2+1}) sits at
    //   sell[slice_ptr[s] + t2*32 + l]; pad entries are (row = p, val = 0).
    int4    *sell;
    int64_t *slice_ptr;                 // [nslices+1], in int4 units
    int64_t  nslices, sell_elems;       // sell_elems in int4 units
    int      sell_mode;                 // 0: single table, quarter-warp (row mod 8) greedy order
                                        // 1: dual table, half-warp (row mod 16) edge-coloured order (8-byte gathers)
                                        // 2: dual table, quarter-warp (row mod 8) edge-coloured order (16-byte gathers)
    bool     sell_plain;                // stored order with true rows (streamed views)
    int      sell_wmax;                 // largest slice width in entries
    bool     uniform_width;             // every slice has the same width
    int      sell_width2;               // pairs per column if uniform_width
    // row-major image for K2 (SKM_F32 only, see csr.cu): (column, value bits) pairs per row
    int2    *csr;
    int64_t *rowptr;                    // [p+1] device
    int64_t *h_rowptr;                  // [p+1] host copy
    // K2 work list: unit u covers csr[unit_start[u] .. unit_start[u+1]) of row unit_row[u]
    int32_t *unit_row;
    int64_t *unit_start;                // [2*nunits]: starts then ends; a unit never crosses a row
    int64_t  nunits;
    unsigned long long *unit_counter;   // device counter K2 pulls work units from
    // k-means++ running minimum distance (allocated on first use)
    double  *kpp_mind;                  // [n]
    double  *kpp_cum;                   // [n] inclusive scan of mind^2
    int32_t *kpp_flag;                  // [n] columns the fp32 filter of a k-means++ round could not skip; then the counter
    double  *kpp_c;                     // [p] scaled centre of the current round, then float[p + 2] (its fp32 copy, max |c|)
    int64_t  kpp_last_exact;            // columns the last round evaluated exactly (-1: all of them)
    // tile/stripe block image for the tensor-core filter (tcsparse.cu), built on first use
    uint2   *tsb;                       // [nnz] (key, value bits), blocks of (128 columns x 64 rows)
    int64_t *tsb_ptr;                   // [ntiles * stripes + 1]
    int64_t  tsb_ntiles;
    int      tsb_stripes;
    float    tsb_sigma;                 // power of two the stored values are multiplied by
    int      tsb_max_block;             // entries of the largest block
    float   *colnorm2;                  // [n] fp32 |x_j|^2 of the stored values
    int     *xmax_bits;                 // device int[4]: bits of max |x|, largest block
    int64_t  device_bytes;
    bool     uncommitted;               // skm_dataset_alloc_csc without skm_dataset_commit yet
};

struct skm_lloyd {
    skm_dataset *ds;
    skm_ctx     *ctx;        // kept separately so destroy never touches a dataset that is gone
    int64_t  K;
    double  *centers;        // [p*K] column-major, current centres (unscaled)
    double  *centers_old;    // [p*K]
    double  *cscaled_t;      // [(p+1)*K] row-major c' = centers/gamma (double), row p = 0
    float   *table;          // fast-path table, fp32, row-major with padded stride
    float   *cmax;           // [1] max |c'|
    int32_t *assign;         // [n] 0-based
    void    *assign_c;       // [n] compact copy for K2's gather (uint8 if K<=256 else uint16/int32)
    float   *dist_f32;       // [n] (SKM_F32 datasets)
    double  *dist_f64;       // [n] (SKM_F64 datasets, and rechecked columns mirror)
    float   *best2;          // [2n] running best/second-best for K-chunked launches
    int32_t *flagged;        // [n] uncertified columns
    int     *nflag;          // device counter
    double  *partials;       // [2*p*K + K + 1]
    double  *stats;          // device [8]: dff^2, has_nan, n_empty, ...
    double  *h_stats;        // pinned
    int64_t *h_counts;       // pinned [K]
    // incremental K2 (opt-in, skm_lloyd_set_update_mode): sums of the previous iteration kept per shard
    int      update_mode;    // 0 = recompute every iteration (default), 1 = incremental
    double  *acc_local;      // [2*p*K + K + 1] this shard's [S | N | counts | -] (partials is what the all-reduce sums)
    int32_t *assign_prev;    // [n] assignments acc_local corresponds to
    int32_t *changed;        // [n] columns whose assignment changed
    int     *nchanged;       // device counter
    bool     acc_valid;
    int      incr_run;       // incremental iterations since the last full recompute
    int      last_update_kind; // 0 full, 1 incremental, 2 nothing changed
    int64_t  last_changed;
    // bounded assignment (opt-in, skm_lloyd_set_assign_mode): Hamerly-style bounds carried across iterations
    int      assign_mode;    // 0 = evaluate every centre for every column (default), 1 = bounded
    float   *lb;             // [n] lower bound on the distance to every centre but the assigned one
    bool     lb_valid;
    double  *centers_prev;   // [p*K] the centres (unscaled) the bounds refer to
    double   gamma_prev;     // and the scaling they were divided by (NaN: none)
    float   *table_t;        // [K][p+1] fp32 centres, one row per centre (bounded kernel)
    float   *shift;          // [K + 4] per-centre movement; then max, second max, argmax
    void    *nchanged_pred;  // device counter of the a-priori keep test
    int64_t  last_predicted_keep;   // columns the a-priori test proved to keep their centre (-1: not run)
    int64_t  last_bounded_flagged;   // columns the bounds could not keep (-1: the pass evaluated everything)
    int      bounded_skip, bounded_backoff;   // passes to sit out after a bounded pass that kept too few columns
    // partial-distance pruning of the full pass (multi-launch plans, K > 16): -1 = automatic, 0 = off, 1 = always try
    int      prune_mode;
    int      prune_skip, prune_backoff;   // full passes to run unpruned after a pruned pass that kept too few columns
    float   *table_rm;       // [p + 1][kpad] fp32 centres, row-major (assign_cols.cu)
    void    *prune_table16;  // half-precision centre table of the one-launch prefix (prefix16.cu)
    float   *prune_scale;    // [4] its scale, 1/scale and rounding bound
    int64_t  last_prune[2];               // columns the pruned pass could not keep (-1: not tried), entry pairs it read
    bool     last_pruned;                 // the last assignment pass was the pruned plan (it kept enough columns)
    // tensor-core filter plan (tcsparse.cu): -1 = automatic, 0 = off, 1 = on
    int      tc_filter;
    void    *tc_bimg;        // swizzled fp16 centre image
    float   *tc_scale;       // [4] sigma, 1/sigma^2, off flag
    uint32_t *tc_cand;       // [n] second | third << 16
    float   *tc_lb4;         // [n]
    float   *tc_zshift;      // [K + 4] zeros (the bounded kernel's movement input)
    int32_t *flagged2;       // [n]
    int64_t  last_tc[3];     // columns: not kept by the candidate pass, not resolved among the best three, (unused)
    bool     assigned, accumulated;
    bool     dist_is_f64;    // which of dist_f64 / dist_f32 the last assignment wrote
    int64_t  last_rechecked;
};

// ---- launchers (implemented across the .cu files) -------------------------

// convert.cu
int skm_launch_convert_index(skm_ctx *ctx, const void *src, int src_type, int64_t count,
                             void *dst, int dst_is_i64);
int skm_launch_convert_value(skm_ctx *ctx, const void *src, int src_type, int64_t count,
                             void *dst, int dst_type);
int skm_validate_csc(skm_ctx *ctx, int64_t p, int64_t n, int64_t nnz, const int64_t *colptr,
                     const int32_t *rowidx, int64_t *max_col_nnz);
int skm_build_sell(skm_dataset *ds);
// re-order the SELL image in place for the kernel family `mode` (0 / 1, see skm_dataset); no-op if current
int skm_sell_ensure_layout(skm_dataset *ds, int mode);
// debug/verification: out[0] = columns whose SELL entries differ from the CSC image, out[1] = steps
// (per quarter- or half-warp), out[2] = shared-memory wavefronts those steps cost
int skm_sell_check(skm_dataset *ds, int64_t out[3]);
int skm_build_csr(skm_dataset *ds);      // csr.cu
// asynchronous pieces used by the streamed path (stream.cu); no host synchronisation inside
int skm_launch_validate_async(skm_ctx *ctx, int64_t p, int64_t n, int64_t nnz, const int64_t *colptr,
                              const int32_t *rowidx, int *flags_dev /* [0]=error bits, [1]=max col nnz */);
int skm_launch_validate_cols_async(skm_ctx *ctx, int64_t n, int64_t nnz, const int64_t *colptr, int *flags_dev);
int skm_launch_validate_rows_async(skm_ctx *ctx, int64_t p, int64_t count, const int32_t *rowidx, int *flags_dev);
int skm_launch_rebase_colptr(skm_ctx *ctx, const void *src, int src_type, int64_t count, int64_t base, int64_t *dst);
int skm_build_sell_async(skm_ctx *ctx, skm_dataset *view, int32_t *width2_scratch, int64_t *elems_scratch,
                         void *cub_tmp, size_t cub_tmp_bytes);
size_t skm_sell_scan_tmp_bytes(int64_t nslices);

// exact.cu
struct ExactArgs {
    int64_t p, n, K;
    const int64_t *colptr;
    const int32_t *rowidx;
    const void    *val;
    int            val_type;        // SKM_F32 / SKM_F64
    const double  *ct;              // row-major [p][K] scaled centres
    const uint8_t *mask;            // optional row-major [p][K] support mask (sparse centres)
    const double  *xdiv;            // optional [K] divisor applied to x (sparse centres)
};
int skm_launch_exact_dist(skm_ctx *ctx, const ExactArgs &a, int64_t j0, int64_t j1, double *dist);
int skm_launch_exact_dist_beta(skm_ctx *ctx, const ExactArgs &a, double beta, double *dist);
int skm_launch_exact_assign(skm_ctx *ctx, const ExactArgs &a, int32_t *assign, double *dist64,
                            float *dist32, const int32_t *subset, const int *subset_count_dev,
                            int64_t subset_max, float *lb = nullptr);
int skm_launch_inner_product(skm_ctx *ctx, int64_t n, const int64_t *colptr, const int32_t *rowidx,
                             const double *val, const double *c, double *inner, double *normsq);
int skm_launch_prep_centers(skm_ctx *ctx, int64_t p, int64_t K, const double *centers,
                            int has_gamma, double gamma, double *ct /* [(p+1)*K] */,
                            uint8_t *mask /* nullable */, double *xdiv /* nullable */);

// assign_fast.cu
struct FastPlan {
    int kc;           // centres per launch (compile-time chunk)
    int ks;           // table row stride in floats
    int nchunks;      // launches per assignment pass
    size_t smem;      // dynamic shared memory per block
    int threads;
    bool global_table; // table gathered from global memory (too large for shared memory)
    bool mode64;       // LDS.64 kernel on a dual table (needs the SELL image in layout mode 1)
    bool dual8;        // 16-byte kernel on a dual table (needs the SELL image in layout mode 2)
    int  boff;         // first row of the second table copy (mode64 / dual8), else 0
    int  layout;       // SELL layout mode this plan reads: 0, 1 or 2
    int64_t rows;      // table rows per chunk (p + 1, or boff + p for a dual table)
};
// max_col_nnz < 0: the caller's SELL image is in stored order / cannot be re-laid out -> never mode64
bool skm_fast_plan(const skm_ctx *ctx, int64_t p, int64_t K, FastPlan *plan, int64_t max_col_nnz = -1);
size_t skm_fast_table_floats(int64_t p, const FastPlan &pl);
int64_t skm_dual_boff(int64_t p);
int skm_fast_stride(int kc);     // padded table row stride (floats) of the 16-byte kernels
int  skm_launch_build_table(skm_ctx *ctx, int64_t p, int64_t K, const double *ct, const FastPlan &pl,
                            float *table, float *cmax);
int  skm_launch_assign_fast(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const FastPlan &pl,
                            const float *table, const float *cmax, int32_t *assign, float *dist,
                            float *best2, int32_t *flagged, int *nflag, const int *m_dev = nullptr,
                            float *lb = nullptr, int max_pairs = 0);

// prefix16.cu: the prefix launch of the pruned pass on a half-precision table (all K <= 64 centres in one launch)
struct Prefix16Plan {
    int kc;           // centres per launch: 32 or 64
    int row_bytes;    // table row: kc halves + 16 bytes of padding
    int nchunks;      // launches (K > 64: several)
    size_t smem;
};
bool   skm_prefix16_plan(const skm_ctx *ctx, int64_t p, int64_t K, Prefix16Plan *pl);
size_t skm_prefix16_table_bytes(int64_t p, const Prefix16Plan &pl);
int    skm_launch_build_table16(skm_ctx *ctx, int64_t p, int64_t K, const double *ct, const Prefix16Plan &pl, const float *cmax,
                                void *table, float *scale);
int    skm_launch_prefix16(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const Prefix16Plan &pl, const void *table,
                           const float *scale, int32_t *assign, float *best2, float *lb, int max_pairs);

// assign_cols.cu: every centre for a list of columns, fp32 with the K1 guard (warp per column, column-major image)
int  skm_assign_cols_kpad(int64_t K);
bool skm_assign_cols_supported(const skm_dataset *ds, int64_t K);
int  skm_launch_build_table_rm(skm_ctx *ctx, int64_t p, int64_t K, const double *ct, float *table, float *cmax);
int  skm_launch_assign_cols(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const float *table_rm, const float *cmax,
                            const int32_t *list, int64_t nlist, int32_t *assign, float *dist, float *lb,
                            int32_t *flagged_out, int *nflag_out);

// bounded.cu
int skm_launch_center_shift(skm_ctx *ctx, int64_t p, int64_t K, const double *centers, double *centers_prev,
                            int has_gamma, double gamma, float *shift /* [K+4] */);
int skm_launch_bound_predict(skm_ctx *ctx, int64_t n, int64_t K, const float *lb, const float *dist, const int32_t *assign,
                             const float *shift, unsigned long long *count_dev);
int skm_launch_build_table_t(skm_ctx *ctx, int64_t p, int64_t K, const double *ct, float *table_t, float *cmax);
int skm_launch_assign_bounded(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const float *table_t, const float *cmax,
                              const float *shift, const int32_t *assign, float *lb, float *dist,
                              int32_t *flagged, int *nflag);

// update.cu
int skm_launch_accumulate(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const int32_t *assign,
                          void *assign_c, const float *dist32, const double *dist64, double *partials,
                          bool zero_first = true);
int skm_launch_finalize(skm_ctx *ctx, int64_t p, int64_t K, const double *partials, double gamma,
                        int ml_correction, double *centers, double *centers_old, double *stats);
int skm_launch_diff_assign(skm_ctx *ctx, int64_t n, int64_t K, const int32_t *assign, const int32_t *prev,
                           const float *dist32, const double *dist64, int32_t *changed, int *nchanged, double *sumsq);
int skm_launch_move_changed(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const int32_t *assign, int32_t *prev,
                            const int32_t *changed, const int *nchanged, int64_t nchanged_host, double *acc);
// a_out[j] = a[j] + 1 (MATLAB indices), d_out[j] = (double)d[j]; either pair may be NULL
int skm_launch_export(skm_ctx *ctx, int64_t n, const int32_t *a, int32_t *a_out, const float *d, double *d_out);
int skm_launch_argmax(skm_ctx *ctx, int64_t n, const float *dist32, const double *dist64,
                      double *out_val, int64_t *out_idx);

// tcsparse.cu: tensor-core (tcgen05, fp16) filter for the sparsified K1 at many centres
bool   skm_tcs_supported(const skm_ctx *ctx, const skm_dataset *ds, int64_t K);
int    skm_tcs_bn(int64_t K);
size_t skm_tcs_bimg_bytes(int64_t p, int64_t K);
int    skm_tsb_build(skm_dataset *ds);
void   skm_tsb_free(skm_dataset *ds);
int    skm_launch_tcs_centres(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const double *ct, const float *cmax,
                              float *scale, void *bimg);
int    skm_launch_tcs_filter(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const void *bimg, const float *scale,
                             const float *cmax, int32_t *assign, float *lb, uint32_t *cand, float *lb4, float *dbg_scores);
int    skm_launch_tcs_resolve(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const float *table_t, const float *cmax,
                              const int32_t *flagged_in, const int *nflag_in, int64_t nflag_host, const uint32_t *cand,
                              const float *lb4, int32_t *assign, float *dist, float *lb, int32_t *flagged_out, int *nflag_out);
int    skm_sell_ensure_any(skm_dataset *ds);
int    skm_sell_layout_range(skm_dataset *ds, int mode, int64_t slice0, int64_t nsl, unsigned long long *ovf_dev);
// csr.cu in three steps, so the row-major image can follow the upload chunk by chunk (chunks of whole 512-column tiles)
struct SkmCsrBuild { void *counts, *offsets, *scan_tmp; size_t scan_bytes; int64_t max_tiles, nchunks; bool active; };
int    skm_csr_begin(skm_dataset *ds, SkmCsrBuild *b, int64_t max_chunk_cols, int64_t nchunks);
int    skm_csr_chunk(skm_dataset *ds, SkmCsrBuild *b, int64_t c, int64_t j0, int64_t j1, int64_t e0, int64_t e1);
int    skm_csr_finish(skm_dataset *ds, SkmCsrBuild *b);
void   skm_csr_abort(SkmCsrBuild *b);

// tcgemm.cu: tensor-core (tcgen05, tf32) filter + exact evaluation of the candidates for the dense second pass
bool   skm_tc_dense_usable(int64_t p, int64_t K);
size_t skm_tc_scratch_bytes(int64_t p, int64_t K, int64_t nc);
int    skm_launch_tc_prep(skm_ctx *ctx, int64_t p, int64_t K, const double *ct /* [p+1][K] */, void *scratch);
int    skm_launch_dense_assign_tc(skm_ctx *ctx, int64_t p, int64_t nc, int64_t K, const float *x32, void *scratch,
                                  const float *cmax, int32_t *assign, float *dist, int32_t *flagged, int *nflag);

int    skm_launch_cast_split(skm_ctx *ctx, int64_t nc, int64_t p, int64_t p_pad, const void *x, int x_type, double scale,
                             float *hi, float *lo);
int    skm_launch_tc_dct(skm_ctx *ctx, int64_t p, int64_t p_pad, int64_t nc, const float *xh, const float *xl,
                         const float *mh, const float *ml, float *y, int64_t ldy);

// fwht.cu
int skm_launch_fwht_f64(skm_ctx *ctx, int64_t m, int64_t n, double *x_inplace,
                        const double *signs /* nullable, length m */, double divide_by /* 0 = none */);
int skm_launch_fwht_f32(skm_ctx *ctx, int64_t m, int64_t n, float *x_inplace,
                        const float *signs /* nullable */, float divide_by /* 0 = none */);
int skm_launch_fwht_sample_f32(skm_ctx *ctx, int64_t p2, int64_t n, int64_t m, const float *x,
                               const float *signs, const int32_t *rows /* nullable: sample on device */,
                               uint64_t seed, int64_t col0, int64_t *colptr, int32_t *rowidx, float *val);
int skm_launch_sample_rows(skm_ctx *ctx, int64_t p2, int64_t n, int64_t m, uint64_t seed, int64_t col0,
                           int32_t *rows_out);

// kpp.cu
int skm_launch_kpp_update(skm_ctx *ctx, const skm_dataset *ds, const double *c_scaled /* dev [p] */,
                          int first, double *mind, int masked = 0);
int skm_launch_scan_sq(skm_ctx *ctx, int64_t n, const double *mind, double *cum);
bool skm_kpp_filter_usable(const skm_ctx *ctx, const skm_dataset *ds);
int skm_launch_kpp_update_filtered(skm_ctx *ctx, const skm_dataset *ds, const double *c_scaled, double *mind,
                                   float *c32, int32_t *flagged, int *nflag);
