// api.cu -- the extern "C" surface declared in include/skm_b200.h: contexts, level-1 stateless
// MEX replacements, resident datasets, the Lloyd state machine and k-means++ support.
#include "common.cuh"
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <new>
#include <vector>

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void skm_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" const char *skm_last_error(const skm_ctx *) { return g_err; }
extern "C" int skm_abi_version(void) { return SKM_ABI_VERSION; }

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
extern "C" int skm_ctx_create(int device, void *cuda_stream, skm_ctx **out)
{
    if (!out) { skm_set_error("skm_ctx_create: out is NULL"); return SKM_ERR_INVALID; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        skm_set_error("no CUDA device available (%s); libskm_b200 has no CPU fallback",
                      e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return SKM_ERR_CUDA;
    }
    SKM_REQUIRE(device >= 0 && device < ndev, "device %d out of range (have %d)", device, ndev);
    SKM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SKM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        skm_set_error("device %d is sm_%d%d; libskm_b200 is built for sm_100a only", device, prop.major, prop.minor);
        return SKM_ERR_UNSUPPORTED;
    }
    skm_ctx *ctx = new (std::nothrow) skm_ctx();
    if (!ctx) { skm_set_error("out of host memory"); return SKM_ERR_NOMEM; }
    ctx->device = device;
    ctx->launches = 0;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (cuda_stream) { ctx->stream = (cudaStream_t)cuda_stream; ctx->own_stream = false; }
    else {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete ctx; skm_set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); return SKM_ERR_CUDA; }
        ctx->own_stream = true;
    }
    ctx->d_flag = nullptr;
    ctx->h_flag = nullptr;
    ctx->red_scratch = nullptr;
    ctx->red_ticket = nullptr;
    ctx->timing = false;
    ctx->stream_cache = nullptr;
    ctx->stream_cache_free = nullptr;
    ctx->tc_chunks = 0;
    ctx->tc_chunks_dropped = 0;
    ctx->use_pool = false;
    ctx->bounce[0] = ctx->bounce[1] = nullptr;
    ctx->bounce_ev[0] = ctx->bounce_ev[1] = nullptr;
    {
        int pools = 0;
        cudaMemPool_t pool;
        if (!getenv("SKM_NO_POOL") && cudaDeviceGetAttribute(&pools, cudaDevAttrMemoryPoolsSupported, device) == cudaSuccess && pools &&
            cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            const char *g = getenv("SKM_POOL_RETAIN_GB");
            uint64_t keep = (uint64_t)((g && *g) ? atof(g) : 48.0) << 30;
            if (cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess) ctx->use_pool = true;
        }
        cudaGetLastError();
    }
    ctx->ev = nullptr;
    memset(ctx->ev_count, 0, sizeof ctx->ev_count);
    if (cudaMalloc((void **)&ctx->d_flag, 16 * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&ctx->red_scratch, SKM_RED_BLOCKS * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&ctx->red_ticket, sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(ctx->red_ticket, 0, sizeof(unsigned int)) != cudaSuccess ||
        cudaMallocHost((void **)&ctx->h_flag, 16 * sizeof(int)) != cudaSuccess) {
        skm_set_error("context scratch allocation failed");
        skm_ctx_destroy(ctx);
        return SKM_ERR_NOMEM;
    }
    *out = ctx;
    return SKM_OK;
}

extern "C" void skm_ctx_destroy(skm_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->stream_cache && ctx->stream_cache_free) ctx->stream_cache_free(ctx->stream_cache);
    if (ctx->d_flag) cudaFree(ctx->d_flag);
    if (ctx->red_scratch) cudaFree(ctx->red_scratch);
    if (ctx->red_ticket) cudaFree(ctx->red_ticket);
    if (ctx->h_flag) cudaFreeHost(ctx->h_flag);
    for (int i = 0; i < 2; ++i) { if (ctx->bounce[i]) cudaFreeHost(ctx->bounce[i]); if (ctx->bounce_ev[i]) cudaEventDestroy(ctx->bounce_ev[i]); }
    if (ctx->ev) {
        for (int s = 0; s < SKM_T_SLOTS; ++s)
            for (int i = 0; i < SKM_T_RING; ++i) { cudaEventDestroy(ctx->ev[s][i][0]); cudaEventDestroy(ctx->ev[s][i][1]); }
        delete[] ctx->ev;
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" void *skm_ctx_stream(skm_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int skm_ctx_device(const skm_ctx *ctx) { return ctx ? ctx->device : -1; }
extern "C" int64_t skm_ctx_launch_count(const skm_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int skm_ctx_tc_chunks(const skm_ctx *ctx, int64_t *kept, int64_t *dropped)
{
    SKM_REQUIRE(ctx, "ctx is NULL");
    if (kept) *kept = ctx->tc_chunks;
    if (dropped) *dropped = ctx->tc_chunks_dropped;
    return SKM_OK;
}
extern "C" int skm_ctx_sync(skm_ctx *ctx)
{
    SKM_REQUIRE(ctx, "ctx is NULL");
    SKM_CUDA(cudaSetDevice(ctx->device));
    SKM_CUDA(cudaStreamSynchronize(ctx->stream));
    return SKM_OK;
}

extern "C" int skm_ctx_timing_enable(skm_ctx *ctx, int on)
{
    SKM_REQUIRE(ctx, "ctx is NULL");
    SKM_CUDA(cudaSetDevice(ctx->device));
    if (on && !ctx->ev) {
        ctx->ev = new (std::nothrow) cudaEvent_t[SKM_T_SLOTS][SKM_T_RING][2];
        if (!ctx->ev) { skm_set_error("out of host memory"); return SKM_ERR_NOMEM; }
        for (int s = 0; s < SKM_T_SLOTS; ++s)
            for (int i = 0; i < SKM_T_RING; ++i) {
                SKM_CUDA(cudaEventCreate(&ctx->ev[s][i][0]));
                SKM_CUDA(cudaEventCreate(&ctx->ev[s][i][1]));
            }
    }
    ctx->timing = on != 0;
    return SKM_OK;
}

extern "C" int skm_ctx_timing_read(skm_ctx *ctx, double *ms, int64_t *counts)
{
    SKM_REQUIRE(ctx && ms && counts, "NULL argument");
    SKM_CUDA(cudaSetDevice(ctx->device));
    SKM_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int s = 0; s < SKM_T_SLOTS; ++s) {
        double tot = 0.0;
        for (int i = 0; i < ctx->ev_count[s]; ++i) {
            float t = 0.f;
            SKM_CUDA(cudaEventElapsedTime(&t, ctx->ev[s][i][0], ctx->ev[s][i][1]));
            tot += t;
        }
        ms[s] = tot;
        counts[s] = ctx->ev_count[s];
        ctx->ev_count[s] = 0;
    }
    return SKM_OK;
}

static int enter(skm_ctx *ctx)
{
    SKM_REQUIRE(ctx, "ctx is NULL");
    SKM_CUDA(cudaSetDevice(ctx->device));
    return SKM_OK;
}

static int dev_alloc(void **ptr, size_t bytes, const char *what)
{
    *ptr = nullptr;
    cudaError_t e = cudaMalloc(ptr, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        cudaGetLastError();
        skm_set_error("cudaMalloc(%s, %zu bytes) failed: %s", what, bytes, cudaGetErrorString(e));
        return SKM_ERR_NOMEM;
    }
    return SKM_OK;
}

int skm_big_alloc(skm_ctx *ctx, void **ptr, size_t bytes, const char *what)
{
    *ptr = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e;
    if (ctx->use_pool) {
        e = cudaMallocAsync(ptr, bytes, ctx->stream);
        if (e != cudaSuccess) {                                  // give back what the pool holds in reserve and retry once
            cudaGetLastError();
            cudaMemPool_t pool;
            cudaStreamSynchronize(ctx->stream);
            if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
            e = cudaMallocAsync(ptr, bytes, ctx->stream);
        }
    } else e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *ptr = nullptr;
        skm_set_error("device allocation (%s, %zu bytes) failed: %s", what, bytes, cudaGetErrorString(e));
        return SKM_ERR_NOMEM;
    }
    return SKM_OK;
}

void skm_big_free(skm_ctx *ctx, void *ptr)
{
    if (!ptr) return;
    if (ctx->use_pool) cudaFreeAsync(ptr, ctx->stream);
    else cudaFree(ptr);
}

int skm_d2h_pageable(skm_ctx *ctx, void *dst, const void *src_dev, size_t bytes)
{
    const size_t CH = (size_t)8 << 20;
    if (bytes < 4 * CH) {                                        // small: not worth the staging
        if (bytes) SKM_CUDA(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        SKM_CUDA(cudaStreamSynchronize(ctx->stream));
        return SKM_OK;
    }
    for (int i = 0; i < 2; ++i) {
        if (!ctx->bounce[i]) SKM_CUDA(cudaMallocHost(&ctx->bounce[i], CH));
        if (!ctx->bounce_ev[i]) SKM_CUDA(cudaEventCreateWithFlags(&ctx->bounce_ev[i], cudaEventDisableTiming));
    }
    const size_t nch = (bytes + CH - 1) / CH;
    for (size_t c = 0; c < nch + 1; ++c) {
        if (c < nch) {                                           // chunk c crosses PCIe ...
            const size_t off = c * CH, sz = std::min(CH, bytes - off);
            SKM_CUDA(cudaMemcpyAsync(ctx->bounce[c & 1], (const char *)src_dev + off, sz, cudaMemcpyDeviceToHost, ctx->stream));
            SKM_CUDA(cudaEventRecord(ctx->bounce_ev[c & 1], ctx->stream));
        }
        if (c > 0) {                                             // ... while chunk c-1 is copied out of the staging buffer
            const size_t off = (c - 1) * CH, sz = std::min(CH, bytes - off);
            SKM_CUDA(cudaEventSynchronize(ctx->bounce_ev[(c - 1) & 1]));
            memcpy((char *)dst + off, ctx->bounce[(c - 1) & 1], sz);
        }
    }
    return SKM_OK;
}

static int h2d(skm_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (bytes == 0) return SKM_OK;
    SKM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return SKM_OK;
}

static int d2h_sync(skm_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (bytes) SKM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SKM_CUDA(cudaStreamSynchronize(ctx->stream));
    return SKM_OK;
}

// ---------------------------------------------------------------------------
// datasets
// ---------------------------------------------------------------------------
static size_t type_size(int t) { return t == SKM_U16 ? 2 : ((t == SKM_F32 || t == SKM_I32) ? 4 : 8); }

#include <chrono>
static bool skm_trace_on() { static const bool on = getenv("SKM_TRACE") != nullptr; return on; }
static double skm_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define SKM_TRACE_POINT(label, t0)                                                             \
    do { if (skm_trace_on()) { cudaDeviceSynchronize(); const double t1__ = skm_now();         \
         fprintf(stderr, "[skm trace] %-28s %8.2f ms\n", label, 1e3 * (t1__ - (t0))); (t0) = t1__; } } while (0)

extern "C" void skm_dataset_destroy(skm_dataset *ds)
{
    if (!ds) return;
    cudaSetDevice(ds->ctx->device);
    cudaStreamSynchronize(ds->ctx->stream);
    const double td0 = skm_trace_on() ? skm_now() : 0.0;
    struct Tr { double t0; ~Tr() { if (skm_trace_on()) fprintf(stderr, "[skm trace] %-28s %8.2f ms\n", "dataset_destroy", 1e3 * (skm_now() - t0)); } } tr{td0};
    cudaDeviceSynchronize();                       // other streams (a caller's views of the arrays) are done as well
    double tdd = skm_now();
    skm_big_free(ds->ctx, ds->colptr);
    skm_big_free(ds->ctx, ds->rowidx);
    skm_big_free(ds->ctx, ds->val);
    skm_big_free(ds->ctx, ds->sell);
    cudaFree(ds->slice_ptr);
    cudaFree(ds->kpp_mind);
    cudaFree(ds->kpp_cum);
    skm_big_free(ds->ctx, ds->kpp_flag);
    cudaFree(ds->kpp_c);
    skm_big_free(ds->ctx, ds->csr);
    SKM_TRACE_POINT("destroy: big frees", tdd);
    cudaFree(ds->rowptr);
    cudaFree(ds->unit_row);
    cudaFree(ds->unit_start);
    cudaFree(ds->unit_counter);
    skm_tsb_free(ds);
    SKM_TRACE_POINT("destroy: small frees", tdd);
    free(ds->h_rowptr);
    delete ds;
}

// takes ownership of colptr/rowidx/val (device, final types)
static int dataset_finish(skm_dataset *ds)
{
    skm_ctx *ctx = ds->ctx;
    double t0 = skm_now();
    SKM_TRY(skm_validate_csc(ctx, ds->p, ds->n, ds->nnz, ds->colptr, ds->rowidx, &ds->max_col_nnz));
    SKM_TRACE_POINT("validate", t0);
    ds->device_bytes = (int64_t)sizeof(int64_t) * (ds->n + 1) + (int64_t)sizeof(int32_t) * ds->nnz +
                       (int64_t)type_size(ds->store_dtype) * ds->nnz;
    if (ds->store_dtype == SKM_F32) {
        SKM_TRY(skm_build_sell(ds));
        SKM_TRACE_POINT("build_sell (widths, alloc)", t0);
        SKM_TRY(skm_build_csr(ds));
        SKM_TRACE_POINT("build_csr", t0);
    }
    return SKM_OK;
}

// ---- pipelined creation from host arrays ----
// The upload of a large matrix is PCIe-bound (88 ms for config 2) and used to be followed by ~85 ms of image building
// plus, at the first assignment, ~55 ms of entry-order scheduling -- with the GPU idle during the upload.  Here the
// columns cross PCIe in chunks on a copy stream while the previous chunk is converted, validated, counted for the
// row-major image and (when the caller says how many centres it will use) laid out in the entry order of that kernel
// family.  Only the offsets scan and the scatter of the row-major image are left when the last chunk has landed.
static int64_t host_jc(const void *jc, int jc_type, int64_t j)
{
    return jc_type == SKM_I32 ? (int64_t)((const int32_t *)jc)[j] : ((const int64_t *)jc)[j];
}

static int dataset_fill_pipelined(skm_dataset *ds, const void *jc, int jc_type, const void *ir, int ir_type,
                                  const void *val, int val_type, int64_t K_hint)
{
    skm_ctx *ctx = ds->ctx;
    const int64_t p = ds->p, n = ds->n, nnz = ds->nnz;
    double t0 = skm_now();
    // 1. column pointers, validated before anything walks them
    {
        DevBuf sj;
        if (jc_type != SKM_I64) {
            SKM_TRY(sj.alloc(type_size(jc_type) * (n + 1)));
            SKM_TRY(h2d(ctx, sj.ptr, jc, type_size(jc_type) * (n + 1)));
            SKM_TRY(skm_launch_convert_index(ctx, sj.ptr, jc_type, n + 1, ds->colptr, 1));
        } else SKM_TRY(h2d(ctx, ds->colptr, jc, sizeof(int64_t) * (n + 1)));
        SKM_CUDA(cudaMemsetAsync(ctx->d_flag, 0, 4 * sizeof(int), ctx->stream));
        SKM_TRY(skm_launch_validate_cols_async(ctx, n, nnz, ds->colptr, ctx->d_flag));
        SKM_CUDA(cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SKM_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->h_flag[0] & 1) { skm_set_error("invalid CSC column pointers (must start at 0, be non-decreasing, end at nnz)"); return SKM_ERR_INVALID; }
        ds->max_col_nnz = ctx->h_flag[1];
    }
    ds->device_bytes = (int64_t)sizeof(int64_t) * (n + 1) + (int64_t)sizeof(int32_t) * nnz + (int64_t)sizeof(float) * nnz;
    // 2. skeletons of the two images (sizes depend on the column pointers only)
    SKM_TRY(skm_build_sell(ds));
    int layout = -1;
    if (K_hint > 0 && ds->sell && ds->sell_wmax > 0 && ds->sell_wmax <= 254) {
        FastPlan pl;
        if (skm_fast_plan(ctx, p, K_hint, &pl, ds->max_col_nnz) && (pl.layout == 1 || pl.layout == 2)) layout = pl.layout;
    }
    SKM_TRACE_POINT("pipelined: colptr + skeletons", t0);
    // 3. chunks of whole 512-column tiles, about 24 M entries each
    const int64_t avg = nnz / n > 0 ? nnz / n : 1;
    const char *ce = getenv("SKM_PIPELINE_CHUNK");                  // entries per chunk (tests shrink it)
    const int64_t target = (ce && atoll(ce) > 0) ? atoll(ce) : 24000000;
    int64_t cols = (target / avg + 511) / 512 * 512;
    if (cols < 512) cols = 512;
    const int64_t nchunks = (n + cols - 1) / cols;
    int64_t max_e = 0;
    for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t j0 = c * cols, j1 = std::min(n, j0 + cols);
        max_e = std::max(max_e, host_jc(jc, jc_type, j1) - host_jc(jc, jc_type, j0));
    }
    SkmCsrBuild cb;
    SKM_TRY(skm_csr_begin(ds, &cb, cols, nchunks));
    const bool stage_i = ir_type != SKM_I32, stage_v = val_type != SKM_F32;
    cudaStream_t cs = nullptr;
    cudaEvent_t up[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
    DevBuf si[2], sv[2], ovf;
    int rc = SKM_OK;
    auto cleanup = [&]() {
        if (cs) { cudaStreamSynchronize(cs); cudaStreamDestroy(cs); }
        cudaStreamSynchronize(ctx->stream);
        for (int i = 0; i < 2; ++i) { if (up[i]) cudaEventDestroy(up[i]); if (done[i]) cudaEventDestroy(done[i]); }
    };
    do {
        if (cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) { skm_set_error("cudaStreamCreate failed"); rc = SKM_ERR_CUDA; break; }
        for (int i = 0; i < 2 && rc == SKM_OK; ++i) {
            if (cudaEventCreateWithFlags(&up[i], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess) { skm_set_error("cudaEventCreate failed"); rc = SKM_ERR_CUDA; }
            if (rc == SKM_OK && stage_i) rc = si[i].alloc(type_size(ir_type) * (size_t)max_e);
            if (rc == SKM_OK && stage_v) rc = sv[i].alloc(type_size(val_type) * (size_t)max_e);
        }
        if (rc != SKM_OK) break;
        if ((rc = ovf.alloc(sizeof(unsigned long long)))) break;
        if (cudaMemsetAsync(ovf.ptr, 0, sizeof(unsigned long long), ctx->stream) != cudaSuccess) { rc = SKM_ERR_CUDA; break; }
        // the copy stream may only start once the stream-ordered allocations and the memsets above have happened
        if (cudaEventRecord(done[0], ctx->stream) != cudaSuccess || cudaStreamWaitEvent(cs, done[0], 0) != cudaSuccess) { rc = SKM_ERR_CUDA; break; }
        for (int64_t c = 0; c < nchunks && rc == SKM_OK; ++c) {
            const int slot = (int)(c & 1);
            const int64_t j0 = c * cols, j1 = std::min(n, j0 + cols);
            const int64_t e0 = host_jc(jc, jc_type, j0), e1 = host_jc(jc, jc_type, j1), ne = e1 - e0;
            if (ne < 0 || e0 < 0 || e1 > nnz) { skm_set_error("invalid CSC column pointers"); rc = SKM_ERR_INVALID; break; }
            cudaError_t e = cudaSuccess;
            if (c >= 2) e = cudaStreamWaitEvent(cs, done[slot], 0);              // the staging slot has been converted
            if (e == cudaSuccess && ne > 0) {
                void *di = stage_i ? si[slot].ptr : (void *)(ds->rowidx + e0);
                void *dv = stage_v ? sv[slot].ptr : (void *)((float *)ds->val + e0);
                e = cudaMemcpyAsync(di, (const char *)ir + (size_t)e0 * type_size(ir_type), (size_t)ne * type_size(ir_type), cudaMemcpyHostToDevice, cs);
                if (e == cudaSuccess)
                    e = cudaMemcpyAsync(dv, (const char *)val + (size_t)e0 * type_size(val_type), (size_t)ne * type_size(val_type), cudaMemcpyHostToDevice, cs);
            }
            if (e == cudaSuccess) e = cudaEventRecord(up[slot], cs);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, up[slot], 0);
            if (e != cudaSuccess) { skm_set_error("pipelined upload failed: %s", cudaGetErrorString(e)); rc = SKM_ERR_CUDA; break; }
            if (ne > 0) {
                if (stage_i && (rc = skm_launch_convert_index(ctx, si[slot].ptr, ir_type, ne, ds->rowidx + e0, 0))) break;
                if (stage_v && (rc = skm_launch_convert_value(ctx, sv[slot].ptr, val_type, ne, (float *)ds->val + e0, SKM_F32))) break;
            }
            if (cudaEventRecord(done[slot], ctx->stream) != cudaSuccess) { rc = SKM_ERR_CUDA; break; }
            if ((rc = skm_launch_validate_rows_async(ctx, p, ne, ds->rowidx + e0, ctx->d_flag))) break;
            if ((rc = skm_csr_chunk(ds, &cb, c, j0, j1, e0, e1))) break;
            if (layout > 0 && (rc = skm_sell_layout_range(ds, layout, j0 / SKM_SLICE, (j1 - j0 + SKM_SLICE - 1) / SKM_SLICE, ovf.as<unsigned long long>()))) break;
        }
        if (rc != SKM_OK) break;
        if (cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) { skm_set_error("pipelined upload failed: %s", cudaGetErrorString(cudaGetLastError())); rc = SKM_ERR_CUDA; break; }
        if (ctx->h_flag[0] & 2) { skm_set_error("invalid CSC row index (must lie in [0,p))"); rc = SKM_ERR_INVALID; break; }
        SKM_TRACE_POINT("pipelined: upload + per-chunk work", t0);
        if (layout > 0) { ds->sell_mode = layout; ds->sell_plain = false; }
    } while (0);
    cleanup();
    if (rc != SKM_OK) { skm_csr_abort(&cb); return rc; }
    rc = skm_csr_finish(ds, &cb);
    SKM_TRACE_POINT("pipelined: csr finish", t0);
    return rc;
}

extern "C" int skm_dataset_create_csc(skm_ctx *ctx, int64_t p, int64_t n, const void *jc, int jc_type,
                                      const void *ir, int ir_type, const void *val, int val_type,
                                      int store_dtype, int on_device, skm_dataset **out)
{
    return skm_dataset_create_csc_hint(ctx, p, n, jc, jc_type, ir, ir_type, val, val_type, store_dtype, on_device, 0, out);
}

extern "C" int skm_dataset_create_csc_hint(skm_ctx *ctx, int64_t p, int64_t n, const void *jc, int jc_type,
                                           const void *ir, int ir_type, const void *val, int val_type,
                                           int store_dtype, int on_device, int64_t K_hint, skm_dataset **out)
{
    SKM_TRY(enter(ctx));
    SKM_REQUIRE(out, "out is NULL");
    *out = nullptr;
    SKM_REQUIRE(p >= 0 && n >= 0, "negative dimensions");
    SKM_REQUIRE(p < 2147483647LL, "p must be below 2^31-1");
    SKM_REQUIRE(jc, "jc is NULL");
    SKM_REQUIRE(jc_type == SKM_I32 || jc_type == SKM_I64, "jc_type must be SKM_I32 or SKM_I64");
    SKM_REQUIRE(ir_type == SKM_I32 || ir_type == SKM_I64 || ir_type == SKM_U16, "ir_type must be SKM_I32, SKM_I64 or SKM_U16");
    SKM_REQUIRE(ir_type != SKM_U16 || p <= 65536, "SKM_U16 row indices need p <= 65536");
    SKM_REQUIRE(val_type == SKM_F32 || val_type == SKM_F64, "val_type must be SKM_F32 or SKM_F64");
    SKM_REQUIRE(store_dtype == SKM_F32 || store_dtype == SKM_F64, "store_dtype must be SKM_F32 or SKM_F64");

    // nnz = jc[n]
    int64_t nnz = 0;
    {
        const char *last = (const char *)jc + (size_t)n * type_size(jc_type);
        char buf[8] = {0};
        if (on_device) {
            SKM_CUDA(cudaMemcpyAsync(buf, last, type_size(jc_type), cudaMemcpyDeviceToHost, ctx->stream));
            SKM_CUDA(cudaStreamSynchronize(ctx->stream));
        } else memcpy(buf, last, type_size(jc_type));
        nnz = (jc_type == SKM_I32) ? (int64_t) * (int32_t *)buf : *(int64_t *)buf;
    }
    SKM_REQUIRE(nnz >= 0, "jc[n] is negative");
    SKM_REQUIRE(nnz == 0 || (ir && val), "ir/val are NULL but nnz > 0");

    skm_dataset *ds = new (std::nothrow) skm_dataset();
    if (!ds) { skm_set_error("out of host memory"); return SKM_ERR_NOMEM; }
    memset(ds, 0, sizeof *ds);
    ds->ctx = ctx;
    ds->p = p; ds->n = n; ds->nnz = nnz;
    ds->store_dtype = store_dtype;
    int rc = SKM_OK;
    double tt0 = skm_now();
    do {
        if ((rc = skm_big_alloc(ds->ctx, (void **)&ds->colptr, sizeof(int64_t) * (n + 1), "colptr"))) break;
        if ((rc = skm_big_alloc(ds->ctx, (void **)&ds->rowidx, sizeof(int32_t) * nnz, "rowidx"))) break;
        if ((rc = skm_big_alloc(ds->ctx, &ds->val, type_size(store_dtype) * nnz, "val"))) break;
        SKM_TRACE_POINT("alloc csc", tt0);
        if (!on_device && store_dtype == SKM_F32 && n > 0 && nnz >= (1 << 22) && !getenv("SKM_NO_PIPELINE")) {
            rc = dataset_fill_pipelined(ds, jc, jc_type, ir, ir_type, val, val_type, K_hint);
            break;
        }
        // stage raw arrays (host -> device) unless they already live on the device
        DevBuf sj, si, sv;
        const void *dj = jc, *di = ir, *dv = val;
        if (!on_device) {
            if (jc_type != SKM_I64) {
                if ((rc = sj.alloc(type_size(jc_type) * (n + 1)))) break;
                if ((rc = h2d(ctx, sj.ptr, jc, type_size(jc_type) * (n + 1)))) break;
                dj = sj.ptr;
            } else {
                if ((rc = h2d(ctx, ds->colptr, jc, sizeof(int64_t) * (n + 1)))) break;
                dj = nullptr;
            }
            if (nnz) {
                if (ir_type != SKM_I32) {
                    if ((rc = si.alloc(type_size(ir_type) * nnz))) break;
                    if ((rc = h2d(ctx, si.ptr, ir, type_size(ir_type) * nnz))) break;
                    di = si.ptr;
                } else {
                    if ((rc = h2d(ctx, ds->rowidx, ir, sizeof(int32_t) * nnz))) break;
                    di = nullptr;
                }
                if (val_type != store_dtype) {
                    if ((rc = sv.alloc(type_size(val_type) * nnz))) break;
                    if ((rc = h2d(ctx, sv.ptr, val, type_size(val_type) * nnz))) break;
                    dv = sv.ptr;
                } else {
                    if ((rc = h2d(ctx, ds->val, val, type_size(val_type) * nnz))) break;
                    dv = nullptr;
                }
            } else { di = nullptr; dv = nullptr; }
        }
        if (dj && (rc = skm_launch_convert_index(ctx, dj, jc_type, n + 1, ds->colptr, 1))) break;
        if (di && nnz && (rc = skm_launch_convert_index(ctx, di, ir_type, nnz, ds->rowidx, 0))) break;
        if (dv && nnz && (rc = skm_launch_convert_value(ctx, dv, val_type, nnz, ds->val, store_dtype))) break;
        cudaError_t e = cudaStreamSynchronize(ctx->stream);      // staging buffers die here
        if (e != cudaSuccess) { skm_set_error("upload failed: %s", cudaGetErrorString(e)); rc = SKM_ERR_CUDA; break; }
        SKM_TRACE_POINT("h2d + convert", tt0);
        rc = dataset_finish(ds);
    } while (0);
    if (rc != SKM_OK) { skm_dataset_destroy(ds); return rc; }
    *out = ds;
    return SKM_OK;
}

// ---- in-place production: the caller (a generator or a preconditioning kernel of its own) writes the CSC arrays
// straight into the dataset's buffers, so a shard that fills most of the HBM never exists twice ----
extern "C" int skm_dataset_alloc_csc(skm_ctx *ctx, int64_t p, int64_t n, int64_t nnz, skm_dataset **out)
{
    SKM_TRY(enter(ctx));
    SKM_REQUIRE(out, "out is NULL");
    *out = nullptr;
    SKM_REQUIRE(p >= 0 && n >= 0 && nnz >= 0, "negative dimensions");
    SKM_REQUIRE(p < 2147483647LL, "p must be below 2^31-1");
    skm_dataset *ds = new (std::nothrow) skm_dataset();
    if (!ds) { skm_set_error("out of host memory"); return SKM_ERR_NOMEM; }
    memset(ds, 0, sizeof *ds);
    ds->ctx = ctx; ds->p = p; ds->n = n; ds->nnz = nnz; ds->store_dtype = SKM_F32;
    ds->uncommitted = true;
    int rc = SKM_OK;
    do {
        if ((rc = skm_big_alloc(ds->ctx, (void **)&ds->colptr, sizeof(int64_t) * (n + 1), "colptr"))) break;
        if ((rc = skm_big_alloc(ds->ctx, (void **)&ds->rowidx, sizeof(int32_t) * nnz, "rowidx"))) break;
        if ((rc = skm_big_alloc(ds->ctx, &ds->val, sizeof(float) * nnz, "val"))) break;
        // the caller fills the arrays on streams of its own: the (stream-ordered) allocations must have happened
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { skm_set_error("allocation failed: %s", cudaGetErrorString(cudaGetLastError())); rc = SKM_ERR_CUDA; }
    } while (0);
    if (rc != SKM_OK) { skm_dataset_destroy(ds); return rc; }
    *out = ds;
    return SKM_OK;
}

extern "C" int skm_dataset_csc_ptrs(skm_dataset *ds, void **colptr, void **rowidx, void **val)
{
    SKM_REQUIRE(ds, "NULL argument");
    if (colptr) *colptr = ds->colptr;
    if (rowidx) *rowidx = ds->rowidx;
    if (val) *val = ds->val;
    return SKM_OK;
}

extern "C" int skm_dataset_commit(skm_dataset *ds)
{
    SKM_REQUIRE(ds, "NULL argument");
    SKM_TRY(enter(ds->ctx));
    if (!ds->uncommitted) { skm_set_error("skm_dataset_commit: the dataset is already committed"); return SKM_ERR_STATE; }
    SKM_CUDA(cudaDeviceSynchronize());                     // the producer may have written on any stream
    int64_t last = 0;
    SKM_TRY(d2h_sync(ds->ctx, &last, ds->colptr + ds->n, sizeof last));
    SKM_REQUIRE(last == ds->nnz, "skm_dataset_commit: colptr[n] = %lld but the dataset was allocated for %lld entries",
                (long long)last, (long long)ds->nnz);
    SKM_TRY(dataset_finish(ds));
    ds->uncommitted = false;
    return SKM_OK;
}

// min / max over ALL p*n elements of the sparse matrix (implicit zeros included), the min(X(:)) / max(X(:)) of
// Start = 'uniform' (kmeans_sparsified.m:388-390)
namespace {
template <typename VT>
__global__ void k_minmax(int64_t nnz, const VT *__restrict__ val, double *__restrict__ out /* [2] init +inf,-inf */)
{
    double mn = __longlong_as_double(0x7ff0000000000000LL), mx = -mn;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) {
        const double v = (double)val[i];
        if (v < mn) mn = v;
        if (v > mx) mx = v;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        // doubles of one sign order like their integer images; do the two signs separately
        unsigned long long *o0 = (unsigned long long *)out, *o1 = o0 + 1;
        unsigned long long old = *o0, assumed;
        do { assumed = old; if (!(mn < __longlong_as_double((long long)assumed))) break;
             old = atomicCAS(o0, assumed, (unsigned long long)__double_as_longlong(mn)); } while (old != assumed);
        old = *o1;
        do { assumed = old; if (!(mx > __longlong_as_double((long long)assumed))) break;
             old = atomicCAS(o1, assumed, (unsigned long long)__double_as_longlong(mx)); } while (old != assumed);
    }
}
}  // namespace

extern "C" int skm_dataset_minmax(skm_dataset *ds, double *mn, double *mx)
{
    SKM_REQUIRE(ds && mn && mx, "NULL argument");
    skm_ctx *ctx = ds->ctx;
    SKM_TRY(enter(ctx));
    const double inf = INFINITY;
    double h[2] = {inf, -inf};
    if (ds->nnz > 0) {
        DevBuf d;
        SKM_TRY(d.alloc(sizeof h));
        SKM_TRY(h2d(ctx, d.ptr, h, sizeof h));
        const int64_t blocks = std::min<int64_t>((ds->nnz + 255) / 256, (int64_t)ctx->sm_count * 8);
        if (ds->store_dtype == SKM_F32) k_minmax<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(ds->nnz, (const float *)ds->val, d.as<double>());
        else k_minmax<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(ds->nnz, (const double *)ds->val, d.as<double>());
        SKM_CHECK_LAUNCH(ctx);
        SKM_TRY(d2h_sync(ctx, h, d.ptr, sizeof h));
    }
    if (ds->nnz < ds->p * ds->n) { if (!(h[0] < 0.0)) h[0] = fmin(h[0], 0.0); if (!(h[1] > 0.0)) h[1] = fmax(h[1], 0.0); }
    *mn = h[0]; *mx = h[1];
    return SKM_OK;
}

extern "C" int skm_dataset_get_info(const skm_dataset *ds, skm_dataset_info *info)
{
    SKM_REQUIRE(ds && info, "NULL argument");
    info->p = ds->p; info->n = ds->n; info->nnz = ds->nnz;
    info->max_col_nnz = ds->max_col_nnz;
    info->store_dtype = ds->store_dtype;
    info->reserved = 0;
    info->device_bytes = ds->device_bytes;
    info->stream_bytes = ds->store_dtype == SKM_F32 ? ds->sell_elems * 16
                                                    : ds->nnz * 12 + (ds->n + 1) * 8;
    return SKM_OK;
}

extern "C" int skm_dataset_layout_check(skm_dataset *ds, int layout, int64_t *out)
{
    SKM_REQUIRE(ds && out, "NULL argument");
    SKM_TRY(enter(ds->ctx));
    SKM_REQUIRE(layout >= -1 && layout <= 2, "layout must be -1 (current), 0, 1 or 2");
    if (layout >= 0 && ds->store_dtype == SKM_F32) SKM_TRY(skm_sell_ensure_layout(ds, layout));
    SKM_TRY(skm_sell_check(ds, out));
    out[3] = ds->sell_mode;
    return SKM_OK;
}

extern "C" int skm_dataset_get_column(skm_dataset *ds, int64_t j, double *out)
{
    SKM_REQUIRE(ds && out, "NULL argument");
    SKM_TRY(enter(ds->ctx));
    SKM_REQUIRE(j >= 0 && j < ds->n, "column %lld out of range", (long long)j);
    skm_ctx *ctx = ds->ctx;
    int64_t range[2];
    SKM_TRY(d2h_sync(ctx, range, ds->colptr + j, sizeof range));
    int64_t cnt = range[1] - range[0];
    std::vector<int32_t> rows(cnt);
    std::vector<double> vals(cnt);
    for (int64_t i = 0; i < ds->p; ++i) out[i] = 0.0;
    if (cnt == 0) return SKM_OK;
    SKM_CUDA(cudaMemcpyAsync(rows.data(), ds->rowidx + range[0], sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
    if (ds->store_dtype == SKM_F64) {
        SKM_TRY(d2h_sync(ctx, vals.data(), (const double *)ds->val + range[0], sizeof(double) * cnt));
    } else {
        std::vector<float> vf(cnt);
        SKM_TRY(d2h_sync(ctx, vf.data(), (const float *)ds->val + range[0], sizeof(float) * cnt));
        for (int64_t i = 0; i < cnt; ++i) vals[i] = (double)vf[i];
    }
    for (int64_t i = 0; i < cnt; ++i) out[rows[i]] = vals[i];
    return SKM_OK;
}

static ExactArgs exact_args(const skm_dataset *ds, int64_t K, const double *ct)
{
    ExactArgs a;
    a.p = ds->p; a.n = ds->n; a.K = K;
    a.colptr = ds->colptr; a.rowidx = ds->rowidx; a.val = ds->val;
    a.val_type = ds->store_dtype;
    a.ct = ct; a.mask = nullptr; a.xdiv = nullptr;
    return a;
}

// ---------------------------------------------------------------------------
// Lloyd state
// ---------------------------------------------------------------------------
extern "C" void skm_lloyd_destroy(skm_lloyd *L)
{
    if (!L) return;
    cudaSetDevice(L->ctx->device);
    cudaStreamSynchronize(L->ctx->stream);
    cudaFree(L->centers); cudaFree(L->centers_old); cudaFree(L->cscaled_t); cudaFree(L->table);
    skm_big_free(L->ctx, L->assign_c);
    cudaFree(L->cmax); skm_big_free(L->ctx, L->assign); skm_big_free(L->ctx, L->dist_f32); skm_big_free(L->ctx, L->dist_f64);
    skm_big_free(L->ctx, L->best2); skm_big_free(L->ctx, L->flagged); cudaFree(L->nflag); cudaFree(L->partials);
    cudaFree(L->stats);
    cudaFree(L->acc_local); skm_big_free(L->ctx, L->assign_prev); skm_big_free(L->ctx, L->changed); cudaFree(L->nchanged);
    skm_big_free(L->ctx, L->lb); cudaFree(L->centers_prev); cudaFree(L->table_t); cudaFree(L->shift); cudaFree(L->nchanged_pred);
    cudaFree(L->prune_table16); cudaFree(L->prune_scale); cudaFree(L->table_rm);
    cudaFree(L->tc_bimg); cudaFree(L->tc_scale); skm_big_free(L->ctx, L->tc_cand); skm_big_free(L->ctx, L->tc_lb4); cudaFree(L->tc_zshift); skm_big_free(L->ctx, L->flagged2);
    if (L->h_stats) cudaFreeHost(L->h_stats);
    if (L->h_counts) cudaFreeHost(L->h_counts);
    delete L;
}

static int skm_lloyd_create_ex(skm_dataset *ds, int64_t K, int want_f64_dist, skm_lloyd **out);
extern "C" int skm_lloyd_create(skm_dataset *ds, int64_t K, skm_lloyd **out)
{
    return skm_lloyd_create_ex(ds, K, 1, out);
}

static int skm_lloyd_create_ex(skm_dataset *ds, int64_t K, int want_f64_dist, skm_lloyd **out)
{
    SKM_REQUIRE(ds && out, "NULL argument");
    *out = nullptr;
    SKM_TRY(enter(ds->ctx));
    SKM_REQUIRE(K >= 1, "K must be >= 1");
    SKM_REQUIRE(K < (1 << 24), "K too large");
    if (ds->uncommitted) { skm_set_error("the dataset was allocated with skm_dataset_alloc_csc but never committed"); return SKM_ERR_STATE; }
    skm_lloyd *L = new (std::nothrow) skm_lloyd();
    if (!L) { skm_set_error("out of host memory"); return SKM_ERR_NOMEM; }
    memset(L, 0, sizeof *L);
    L->ds = ds; L->ctx = ds->ctx; L->K = K;
    L->tc_filter = -1;
    L->prune_mode = -1;
    L->last_prune[0] = L->last_prune[1] = -1;
    L->last_tc[0] = L->last_tc[1] = L->last_tc[2] = -1;
    const int64_t p = ds->p, n = ds->n;
    int rc = SKM_OK;
    do {
        if ((rc = dev_alloc((void **)&L->centers, sizeof(double) * p * K, "centers"))) break;
        if ((rc = dev_alloc((void **)&L->centers_old, sizeof(double) * p * K, "centers_old"))) break;
        if ((rc = dev_alloc((void **)&L->cscaled_t, sizeof(double) * (p + 1) * K, "cscaled"))) break;
        if ((rc = skm_big_alloc(L->ctx, (void **)&L->assign, sizeof(int32_t) * n, "assign"))) break;
        if ((rc = skm_big_alloc(L->ctx, &L->assign_c, sizeof(int32_t) * n, "assign_c"))) break;
        if ((rc = dev_alloc((void **)&L->partials, sizeof(double) * (2 * p * K + K + 1), "partials"))) break;
        if ((rc = dev_alloc((void **)&L->stats, sizeof(double) * 8, "stats"))) break;
        if ((rc = dev_alloc((void **)&L->nflag, sizeof(int) * 4, "nflag"))) break;
        if (cudaMemsetAsync(L->nflag, 0, sizeof(int) * 4, ds->ctx->stream) != cudaSuccess) { rc = SKM_ERR_CUDA; break; }   // read_stats copies all four
        if ((rc = dev_alloc((void **)&L->cmax, sizeof(float) * 4, "cmax"))) break;
        if (ds->store_dtype == SKM_F32) {
            if ((rc = skm_big_alloc(L->ctx, (void **)&L->dist_f32, sizeof(float) * n, "dist"))) break;
            if ((rc = skm_big_alloc(L->ctx, (void **)&L->flagged, sizeof(int32_t) * n, "flagged"))) break;
            FastPlan pl;
            if (skm_fast_plan(ds->ctx, p, K, &pl, ds->max_col_nnz)) {
                if ((rc = dev_alloc((void **)&L->table, sizeof(float) * skm_fast_table_floats(p, pl), "table"))) break;
                if (pl.nchunks > 1 && (rc = skm_big_alloc(L->ctx, (void **)&L->best2, sizeof(float) * 2 * n, "best2"))) break;
            }
        }
        if (ds->store_dtype == SKM_F64 || want_f64_dist) {
            if ((rc = skm_big_alloc(L->ctx, (void **)&L->dist_f64, sizeof(double) * n, "dist64"))) break;
        }
        if (cudaMallocHost((void **)&L->h_stats, sizeof(double) * 8) != cudaSuccess ||
            cudaMallocHost((void **)&L->h_counts, sizeof(int64_t) * (K + 2)) != cudaSuccess) {
            skm_set_error("pinned host allocation failed");
            rc = SKM_ERR_NOMEM;
            break;
        }
        cudaError_t e = cudaMemsetAsync(L->centers, 0, sizeof(double) * p * K, ds->ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(L->centers_old, 0, sizeof(double) * p * K, ds->ctx->stream);
        if (e != cudaSuccess) { skm_set_error("memset failed: %s", cudaGetErrorString(e)); rc = SKM_ERR_CUDA; break; }
    } while (0);
    if (rc != SKM_OK) { skm_lloyd_destroy(L); return rc; }
    *out = L;
    return SKM_OK;
}

extern "C" int skm_lloyd_set_centers(skm_lloyd *L, const double *centers)
{
    SKM_REQUIRE(L && centers, "NULL argument");
    SKM_TRY(enter(L->ds->ctx));
    SKM_TRY(h2d(L->ds->ctx, L->centers, centers, sizeof(double) * L->ds->p * L->K));
    SKM_CUDA(cudaStreamSynchronize(L->ds->ctx->stream));
    return SKM_OK;
}

extern "C" int skm_lloyd_get_centers(skm_lloyd *L, double *centers)
{
    SKM_REQUIRE(L && centers, "NULL argument");
    SKM_TRY(enter(L->ds->ctx));
    return d2h_sync(L->ds->ctx, centers, L->centers, sizeof(double) * L->ds->p * L->K);
}

extern "C" int skm_lloyd_get_centers_old(skm_lloyd *L, double *centers)
{
    SKM_REQUIRE(L && centers, "NULL argument");
    SKM_TRY(enter(L->ds->ctx));
    return d2h_sync(L->ds->ctx, centers, L->centers_old, sizeof(double) * L->ds->p * L->K);
}

extern "C" int skm_lloyd_set_center_column(skm_lloyd *L, int64_t k, const double *col)
{
    SKM_REQUIRE(L && col, "NULL argument");
    SKM_REQUIRE(k >= 0 && k < L->K, "centre index out of range");
    SKM_TRY(enter(L->ds->ctx));
    SKM_TRY(h2d(L->ds->ctx, L->centers + k * L->ds->p, col, sizeof(double) * L->ds->p));
    SKM_CUDA(cudaStreamSynchronize(L->ds->ctx->stream));
    return SKM_OK;
}

extern "C" int skm_lloyd_set_assign_mode(skm_lloyd *L, int mode)
{
    SKM_REQUIRE(L, "NULL argument");
    SKM_REQUIRE(mode == 0 || mode == 1, "assign mode must be 0 (evaluate every centre) or 1 (bounded)");
    SKM_TRY(enter(L->ds->ctx));
    if (mode == 1 && L->ds->store_dtype != SKM_F32) {
        skm_set_error("the bounded assignment needs an SKM_F32 dataset");
        return SKM_ERR_UNSUPPORTED;
    }
    if (mode == 1) {
        // each buffer on its own: the pruned / tensor-core passes may already have allocated some of them
        const int64_t p = L->ds->p, n = L->ds->n, K = L->K;
        if (!L->lb) SKM_TRY(skm_big_alloc(L->ctx, (void **)&L->lb, sizeof(float) * n, "lb"));
        if (!L->centers_prev) {
            SKM_TRY(dev_alloc((void **)&L->centers_prev, sizeof(double) * p * K, "centers_prev"));
            // the first movement is measured against zeros and never used (no bounds exist yet), but it is read (initcheck)
            SKM_CUDA(cudaMemsetAsync(L->centers_prev, 0, sizeof(double) * p * K, L->ds->ctx->stream));
        }
        if (!L->table_t) SKM_TRY(dev_alloc((void **)&L->table_t, sizeof(float) * K * (p + 1), "table_t"));
        if (!L->shift) SKM_TRY(dev_alloc((void **)&L->shift, sizeof(float) * (K + 4), "shift"));
        if (!L->nchanged_pred) SKM_TRY(dev_alloc(&L->nchanged_pred, 16, "predict counter"));
    }
    L->assign_mode = mode;
    L->lb_valid = false;
    L->bounded_skip = 0;
    L->bounded_backoff = 0;
    return SKM_OK;
}

// ---- partial-distance pruning of the full pass (plans with several launches, K > 16) ----
// Every term of the masked distance is non-negative, so the sum over a PREFIX of a column's entries is a lower bound of
// its distance to that centre.  A pass over the first ~15 % of every column's entries for all K centres (the same
// kernels, `max_pairs`) yields a candidate winner and, from the second-smallest partial sum minus its rounding guard, a
// lower bound on the distance to every other centre; `k_assign_bounded` then evaluates the candidate exactly on all
// entries and keeps it iff it stays below that bound.  Columns it cannot keep are evaluated against every centre (fp64
// when fewer than n/16; the ordinary full pass when many, after which the pruned pass sits out 1, 2, 4, ... 32 calls).  Exact: a kept
// winner beats a rigorous lower bound of every other centre.  It pays when clusters are separated (mixture at K = 64:
// 5.4 -> ~2 ms per pass) and costs one wasted attempt in 33 on data without structure.
// prefix launches on the half-precision table unless it does not fit or SKM_PRUNE_F32 asks for the fp32 kernels
static bool prune_half_table(const skm_ctx *ctx, int64_t p, int64_t K, Prefix16Plan *hp)
{
    const bool f32 = getenv("SKM_PRUNE_F32") != nullptr;          // read every call: the tests switch it
    return !f32 && skm_prefix16_plan(ctx, p, K, hp);
}

static bool prune_wanted(skm_lloyd *L)
{
    if (L->prune_mode == 0) return false;
    if (L->prune_mode == 1) return true;
    static const char *e = getenv("SKM_PRUNE");
    if (e && *e && atoi(e) == 0) return false;
    if (L->prune_skip > 0) { L->prune_skip -= 1; return false; }
    return true;
}

extern "C" int skm_lloyd_set_prune(skm_lloyd *L, int mode)
{
    SKM_REQUIRE(L, "NULL argument");
    SKM_REQUIRE(mode >= -1 && mode <= 1, "prune mode must be -1 (automatic), 0 (off) or 1 (always try)");
    L->prune_mode = mode;
    L->prune_skip = L->prune_backoff = 0;
    return SKM_OK;
}

extern "C" int skm_lloyd_last_prune(skm_lloyd *L, int64_t *not_kept, int64_t *pairs)
{
    SKM_REQUIRE(L, "NULL argument");
    if (not_kept) *not_kept = L->last_prune[0];
    if (pairs) *pairs = L->last_prune[1];
    return SKM_OK;
}

// ---- tensor-core filter plan (tcsparse.cu) ----
static bool tc_wanted(const skm_lloyd *L)
{
    const skm_dataset *ds = L->ds;
    if (!skm_tcs_supported(ds->ctx, ds, L->K)) return false;
    if (L->tc_filter == 0) return false;
    if (L->tc_filter == 1) return true;
    // automatic = off: measured 4.9 ms per config-3 shard pass (filter 4.2 + candidate pass) against 5.35 ms for the
    // gather kernels -- not worth a fourth image of X (profiles/r2_tcsparse.md); SKM_TC_FILTER=1 turns it on
    static const char *e = getenv("SKM_TC_FILTER");
    return e && *e && atoi(e) != 0;
}

static int tc_prepare(skm_lloyd *L)
{
    skm_dataset *ds = L->ds;
    const int64_t p = ds->p, n = ds->n, K = L->K;
    if (!L->lb) SKM_TRY(skm_big_alloc(L->ctx, (void **)&L->lb, sizeof(float) * n, "lb"));
    if (!L->table_t) SKM_TRY(dev_alloc((void **)&L->table_t, sizeof(float) * K * (p + 1), "table_t"));
    if (!L->tc_bimg) SKM_TRY(dev_alloc(&L->tc_bimg, skm_tcs_bimg_bytes(p, K), "centre image"));
    if (!L->tc_scale) SKM_TRY(dev_alloc((void **)&L->tc_scale, sizeof(float) * 4, "tc scale"));
    if (!L->tc_cand) SKM_TRY(skm_big_alloc(L->ctx, (void **)&L->tc_cand, sizeof(uint32_t) * n, "tc candidates"));
    if (!L->tc_lb4) SKM_TRY(skm_big_alloc(L->ctx, (void **)&L->tc_lb4, sizeof(float) * n, "tc bound"));
    if (!L->flagged2) SKM_TRY(skm_big_alloc(L->ctx, (void **)&L->flagged2, sizeof(int32_t) * n, "flagged2"));
    if (!L->tc_zshift) {
        SKM_TRY(dev_alloc((void **)&L->tc_zshift, sizeof(float) * (K + 4), "zero shift"));
        SKM_CUDA(cudaMemsetAsync(L->tc_zshift, 0, sizeof(float) * (K + 4), ds->ctx->stream));
    }
    SKM_TRY(skm_tsb_build(ds));
    SKM_TRY(skm_sell_ensure_any(ds));
    return SKM_OK;
}

// the full assignment pass on the tensor cores: filter -> exact evaluation of the winner (every column) -> exact
// evaluation of the best three (columns whose runner-up is within the filter's error) -> fp64 (what is left)
static int tc_assign(skm_lloyd *L, const ExactArgs &ea, float *dbg_scores)
{
    skm_dataset *ds = L->ds;
    skm_ctx *ctx = ds->ctx;
    {
        SkmTimed t(ctx, SKM_T_ASSIGN);
        SKM_TRY(skm_launch_build_table_t(ctx, ds->p, L->K, L->cscaled_t, L->table_t, L->cmax));
        SKM_TRY(skm_launch_tcs_centres(ctx, ds, L->K, L->cscaled_t, L->cmax, L->tc_scale, L->tc_bimg));
        SKM_TRY(skm_launch_tcs_filter(ctx, ds, L->K, L->tc_bimg, L->tc_scale, L->cmax, L->assign, L->lb, L->tc_cand, L->tc_lb4,
                                      dbg_scores));
        SKM_TRY(skm_launch_assign_bounded(ctx, ds, L->K, L->table_t, L->cmax, L->tc_zshift, L->assign, L->lb, L->dist_f32,
                                          L->flagged, L->nflag));
    }
    {
        SkmTimed t(ctx, SKM_T_RECHECK);
        SKM_CUDA(cudaMemcpyAsync(L->nflag + 2, L->nflag, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
        SKM_TRY(skm_launch_tcs_resolve(ctx, ds, L->K, L->table_t, L->cmax, L->flagged, L->nflag, ds->n, L->tc_cand, L->tc_lb4,
                                       L->assign, L->dist_f32, L->lb, L->flagged2, L->nflag + 1));
        SKM_TRY(skm_launch_exact_assign(ctx, ea, L->assign, nullptr, L->dist_f32, L->flagged2, L->nflag + 1, ds->n, L->lb));
        // n_rechecked keeps its meaning: columns that went to the fp64 kernel
        SKM_CUDA(cudaMemcpyAsync(L->nflag, L->nflag + 1, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    L->last_tc[0] = L->last_tc[1] = -2;               // on the device until read_stats
    return SKM_OK;
}

extern "C" int skm_lloyd_set_tc_filter(skm_lloyd *L, int mode)
{
    SKM_REQUIRE(L, "NULL argument");
    SKM_REQUIRE(mode >= -1 && mode <= 1, "tc filter mode must be -1 (automatic), 0 (off) or 1 (on)");
    if (mode == 1 && !skm_tcs_supported(L->ds->ctx, L->ds, L->K)) {
        skm_set_error("the tensor-core filter needs an SKM_F32 dataset, 2 <= K <= 128 and p <= 4096");
        return SKM_ERR_UNSUPPORTED;
    }
    L->tc_filter = mode;
    return SKM_OK;
}

extern "C" int skm_lloyd_last_tc(skm_lloyd *L, int64_t *not_kept, int64_t *not_resolved)
{
    SKM_REQUIRE(L, "NULL argument");
    if (not_kept) *not_kept = L->last_tc[0];
    if (not_resolved) *not_resolved = L->last_tc[1];
    return SKM_OK;
}

// tests: the raw scores of the filter, scaled back, [n][BN] floats (BN = 32 / 64 / 128 by K), after an assignment pass
extern "C" int skm_debug_tc_scores(skm_lloyd *L, int has_gamma, double gamma, float *scores_host, int64_t *bn_out)
{
    SKM_REQUIRE(L && scores_host, "NULL argument");
    skm_dataset *ds = L->ds;
    skm_ctx *ctx = ds->ctx;
    SKM_TRY(enter(ctx));
    if (!skm_tcs_supported(ctx, ds, L->K)) { skm_set_error("tensor-core filter not supported for this dataset / K"); return SKM_ERR_UNSUPPORTED; }
    SKM_TRY(tc_prepare(L));
    const int64_t bn = skm_tcs_bn(L->K);
    DevBuf sc;
    SKM_TRY(sc.alloc(sizeof(float) * (size_t)ds->n * bn));
    SKM_TRY(skm_launch_prep_centers(ctx, ds->p, L->K, L->centers, has_gamma, gamma, L->cscaled_t, nullptr, nullptr));
    ExactArgs ea = exact_args(ds, L->K, L->cscaled_t);
    SKM_TRY(tc_assign(L, ea, sc.as<float>()));
    L->dist_is_f64 = false; L->assigned = true; L->accumulated = false;
    if (bn_out) *bn_out = bn;
    return d2h_sync(ctx, scores_host, sc.ptr, sizeof(float) * (size_t)ds->n * bn);
}

extern "C" int skm_lloyd_last_assign(skm_lloyd *L, int64_t *n_flagged)
{
    SKM_REQUIRE(L, "NULL argument");
    if (n_flagged) *n_flagged = L->last_bounded_flagged;
    return SKM_OK;
}

// Every centre for the columns a bounded / pruned pass could not keep (L->flagged, nfl of them, counted on the host).
// Lists of 2048 columns and more first take the K1 arithmetic in fp32 restricted to the list (assign_cols.cu: a warp per
// column on the column-major image, certified with the K1 guard; ~1.5 us per 1000 columns at K = 64 whatever the list
// length), and only what that cannot certify goes to the fp64 kernel (reference order, ~3.4 us per 1000 columns); shorter
// lists go to fp64 directly.  Both refresh lb for the columns they touch.  (A lane-per-column walk of the SELL image was
// tried first and dropped: its 16-byte loads 512 bytes apart cost 64-byte DRAM bursts, profiles/r2_prune.md.)
static int reevaluate_flagged(skm_lloyd *L, const ExactArgs &ea, int64_t nfl)
{
    skm_dataset *ds = L->ds;
    skm_ctx *ctx = ds->ctx;
    const char *le = getenv("SKM_LIST_MIN");                          // knob (read every call: tests switch it): shortest list for the fp32 kernel
    const char *how = getenv("SKM_REEVAL");                           // knob: "fp64" switches the fp32 list kernel off
    const bool cols_ok = !(how && !strcmp(how, "fp64")) && skm_assign_cols_supported(ds, L->K);
    if (cols_ok && nfl >= (le ? atoll(le) : 2048)) {
        if (!L->flagged2) SKM_TRY(skm_big_alloc(L->ctx, (void **)&L->flagged2, sizeof(int32_t) * ds->n, "flagged2"));
        if (!L->table_rm)
            SKM_TRY(dev_alloc((void **)&L->table_rm, sizeof(float) * (size_t)(ds->p + 1) * skm_assign_cols_kpad(L->K), "row-major table"));
        SKM_TRY(skm_launch_build_table_rm(ctx, ds->p, L->K, L->cscaled_t, L->table_rm, L->cmax));
        SKM_TRY(skm_launch_assign_cols(ctx, ds, L->K, L->table_rm, L->cmax, L->flagged, nfl, L->assign, L->dist_f32, L->lb,
                                       L->flagged2, L->nflag + 1));
        SKM_TRY(skm_launch_exact_assign(ctx, ea, L->assign, nullptr, L->dist_f32, L->flagged2, L->nflag + 1, ds->n, L->lb));
        SKM_CUDA(cudaMemcpyAsync(L->nflag, L->nflag + 1, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));   // statistics: fp64 columns
        return SKM_OK;
    }
    SKM_TRY(skm_launch_exact_assign(ctx, ea, L->assign, nullptr, L->dist_f32, L->flagged, L->nflag, ds->n, L->lb));
    return SKM_OK;
}

extern "C" int skm_lloyd_assign(skm_lloyd *L, int has_gamma, double gamma)
{
    SKM_REQUIRE(L, "NULL argument");
    skm_dataset *ds = L->ds;
    skm_ctx *ctx = ds->ctx;
    SKM_TRY(enter(ctx));
    SKM_REQUIRE(!has_gamma || gamma == gamma, "gamma is NaN");
    FastPlan pl;
    const bool fast = ds->store_dtype == SKM_F32 && L->table && skm_fast_plan(ctx, ds->p, L->K, &pl, ds->max_col_nnz);
    bool use_tc = fast && tc_wanted(L);
    if (use_tc) {
        const int rc = tc_prepare(L);                                    // one-off image build; falls back if it does not fit
        if ((rc == SKM_ERR_NOMEM || rc == SKM_ERR_UNSUPPORTED) && L->tc_filter != 1) use_tc = false;
        else SKM_TRY(rc);
    }
    if (fast && !use_tc) SKM_TRY(skm_sell_ensure_layout(ds, pl.layout)); // entry order of the kernel family (one-off)
    L->last_tc[0] = L->last_tc[1] = -1;
    {
        SkmTimed t(ctx, SKM_T_PREP);
        SKM_TRY(skm_launch_prep_centers(ctx, ds->p, L->K, L->centers, has_gamma, gamma, L->cscaled_t, nullptr, nullptr));
    }
    ExactArgs ea = exact_args(ds, L->K, L->cscaled_t);
    L->last_rechecked = -1;
    L->last_bounded_flagged = -1;
    L->last_prune[0] = L->last_prune[1] = -1;
    L->last_pruned = false;
    const bool want_bounds = fast && L->assign_mode == 1;
    float *lb = want_bounds ? L->lb : nullptr;
    if (want_bounds) {
        // bounded pass: valid when the bounds exist and refer to the same scaling of the centres
        const double gnow = has_gamma ? gamma : nan("");
        const bool same_scale = (gnow == L->gamma_prev) || (gnow != gnow && L->gamma_prev != L->gamma_prev);
        bool usable = L->lb_valid && same_scale && L->assigned;
        // back off after a pass that kept too few columns: while the centres still move a lot the bounded pass
        // is pure overhead (the full pass rewrites every bound anyway, so skipping costs nothing)
        if (usable && L->bounded_skip > 0) { L->bounded_skip -= 1; usable = false; }
        {
            SkmTimed t(ctx, SKM_T_PREP);
            SKM_TRY(skm_launch_center_shift(ctx, ds->p, L->K, L->centers, L->centers_prev, has_gamma, gamma, L->shift));
        }
        L->gamma_prev = gnow;
        if (usable && !L->dist_is_f64) {
            // a-priori test (12 bytes per column, no entry read): how many columns are CERTAIN to keep their centre
            // after this move?  While the centres still move a lot that is next to none, and the bounded pass
            // would only be an extra pass in front of the full one.
            SkmTimed t(ctx, SKM_T_PREP);
            unsigned long long *cnt = reinterpret_cast<unsigned long long *>(L->nchanged_pred);
            SKM_TRY(skm_launch_bound_predict(ctx, ds->n, L->K, L->lb, L->dist_f32, L->assign, L->shift, cnt));
            unsigned long long hcnt = 0;
            SKM_CUDA(cudaMemcpyAsync(&hcnt, cnt, sizeof hcnt, cudaMemcpyDeviceToHost, ctx->stream));
            SKM_CUDA(cudaStreamSynchronize(ctx->stream));
            L->last_predicted_keep = (int64_t)hcnt;
            if ((int64_t)hcnt < ds->n - ds->n / 4) usable = false;      // the pass pays off only if few columns need the full evaluation
        }
        if (usable) {
            int64_t nfl = 0;
            {
                SkmTimed t(ctx, SKM_T_ASSIGN);
                SKM_TRY(skm_launch_build_table_t(ctx, ds->p, L->K, L->cscaled_t, L->table_t, L->cmax));
                SKM_TRY(skm_launch_assign_bounded(ctx, ds, L->K, L->table_t, L->cmax, L->shift, L->assign, L->lb, L->dist_f32,
                                                  L->flagged, L->nflag));
                SKM_CUDA(cudaMemcpyAsync(ctx->h_flag + 13, L->nflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                SKM_CUDA(cudaStreamSynchronize(ctx->stream));
                nfl = ctx->h_flag[13];
            }
            L->last_bounded_flagged = nfl;
            if (nfl <= ds->n / 8) {
                L->bounded_backoff = 0;
                // few columns left their bound: every centre for those (fp32 list pass + fp64, or fp64 alone); refreshes their lb
                SkmTimed t(ctx, SKM_T_RECHECK);
                SKM_TRY(reevaluate_flagged(L, ea, nfl));
                L->dist_is_f64 = false;
                L->assigned = true;
                L->accumulated = false;
                return SKM_OK;
            }
            // many movers: the full pass below re-evaluates everything and rewrites every bound
            L->bounded_backoff = L->bounded_backoff ? (L->bounded_backoff < 8 ? 2 * L->bounded_backoff : 8) : 1;
            L->bounded_skip = L->bounded_backoff;
        }
    }
    bool pruned = false;
    {
        const int64_t w2max = ds->uniform_width ? ds->sell_width2 : ds->sell_wmax / 2;
        if (fast && !use_tc && pl.nchunks > 1 && !pl.global_table && w2max >= 8 && ds->n > 0 && prune_wanted(L)) {
            const int64_t p = ds->p, n = ds->n, K = L->K;
            if (!L->lb) SKM_TRY(skm_big_alloc(ctx, (void **)&L->lb, sizeof(float) * n, "lb"));
            if (!L->table_t) SKM_TRY(dev_alloc((void **)&L->table_t, sizeof(float) * K * (p + 1), "table_t"));
            if (!L->tc_zshift) {
                SKM_TRY(dev_alloc((void **)&L->tc_zshift, sizeof(float) * (K + 4), "zero shift"));
                SKM_CUDA(cudaMemsetAsync(L->tc_zshift, 0, sizeof(float) * (K + 4), ctx->stream));
            }
            int pairs = (int)((w2max * 15 + 99) / 100);
            if (pairs < 4) pairs = 4;          // 8 entries: fewer leave too many columns to the fallback (profiles/r2_prune.md)
            {
                static const char *pe = getenv("SKM_PRUNE_PAIRS");          // tuning knob
                if (pe && atoi(pe) > 0) pairs = atoi(pe);
            }
            int64_t nfl = 0;
            {
                SkmTimed t(ctx, SKM_T_ASSIGN);
                Prefix16Plan hp;
                if (prune_half_table(ctx, p, K, &hp)) {
                    // one launch over a half-precision table (prefix16.cu); cmax comes from the fp32 row table
                    if (!L->prune_table16) {
                        SKM_TRY(dev_alloc(&L->prune_table16, skm_prefix16_table_bytes(p, hp), "half-precision prefix table"));
                        SKM_TRY(dev_alloc((void **)&L->prune_scale, 4 * sizeof(float), "prefix scale"));
                    }
                    SKM_TRY(skm_launch_build_table_t(ctx, p, K, L->cscaled_t, L->table_t, L->cmax));
                    SKM_TRY(skm_launch_build_table16(ctx, p, K, L->cscaled_t, hp, L->cmax, L->prune_table16, L->prune_scale));
                    SKM_TRY(skm_launch_prefix16(ctx, ds, K, hp, L->prune_table16, L->prune_scale, L->assign, L->best2, L->lb, pairs));
                } else {
                    SKM_TRY(skm_launch_build_table(ctx, p, K, L->cscaled_t, pl, L->table, L->cmax));
                    SKM_TRY(skm_launch_assign_fast(ctx, ds, K, pl, L->table, L->cmax, L->assign, L->dist_f32, L->best2,
                                                   L->flagged, L->nflag, nullptr, L->lb, pairs));
                    SKM_TRY(skm_launch_build_table_t(ctx, p, K, L->cscaled_t, L->table_t, L->cmax));
                }
                SKM_TRY(skm_launch_assign_bounded(ctx, ds, K, L->table_t, L->cmax, L->tc_zshift, L->assign, L->lb, L->dist_f32,
                                                  L->flagged, L->nflag));
                SKM_CUDA(cudaMemcpyAsync(ctx->h_flag + 10, L->nflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                SKM_CUDA(cudaStreamSynchronize(ctx->stream));
                nfl = ctx->h_flag[10];
            }
            L->last_prune[0] = nfl; L->last_prune[1] = pairs;
            // re-evaluating a column against every centre costs ~3.4 us per 1000 columns in fp64 and ~1.5 us with the fp32 list
            // kernel in front: up to n/16 resp. n/4 columns the pruned pass still beats the full one
            if (nfl <= (skm_assign_cols_supported(ds, K) ? n / 4 : n / 16)) {
                SkmTimed t(ctx, SKM_T_RECHECK);
                SKM_TRY(reevaluate_flagged(L, ea, nfl));
                L->prune_backoff = 0;
                pruned = true;
            } else {
                // no structure to exploit (or centres far from converged): the full pass below redoes every column
                L->prune_backoff = L->prune_backoff ? (L->prune_backoff < 32 ? 2 * L->prune_backoff : 32) : 1;
                L->prune_skip = L->prune_backoff;
            }
        }
    }
    if (pruned) {
        L->last_pruned = true;
        L->dist_is_f64 = false;
        if (want_bounds) L->lb_valid = true;
    } else if (use_tc) {
        SKM_TRY(tc_assign(L, ea, nullptr));
        L->dist_is_f64 = false;
        if (want_bounds) L->lb_valid = true;
    } else if (fast) {
        {
            SkmTimed t(ctx, SKM_T_ASSIGN);
            SKM_TRY(skm_launch_build_table(ctx, ds->p, L->K, L->cscaled_t, pl, L->table, L->cmax));
            SKM_TRY(skm_launch_assign_fast(ctx, ds, L->K, pl, L->table, L->cmax, L->assign, L->dist_f32, L->best2,
                                           L->flagged, L->nflag, nullptr, lb));
        }
        // columns the guard could not certify: fp64, reference order
        SkmTimed t(ctx, SKM_T_RECHECK);
        SKM_TRY(skm_launch_exact_assign(ctx, ea, L->assign, nullptr, L->dist_f32, L->flagged, L->nflag, ds->n, lb));
        L->dist_is_f64 = false;
        if (want_bounds) L->lb_valid = true;
    } else {
        SkmTimed t(ctx, SKM_T_ASSIGN);
        SKM_TRY(skm_launch_exact_assign(ctx, ea, L->assign, L->dist_f64, L->dist_f32, nullptr, nullptr, 0));
        SKM_CUDA(cudaMemsetAsync(L->nflag, 0, sizeof(int), ctx->stream));
        L->dist_is_f64 = true;
    }
    L->assigned = true;
    L->accumulated = false;
    return SKM_OK;
}

extern "C" int skm_lloyd_set_update_mode(skm_lloyd *L, int mode)
{
    SKM_REQUIRE(L, "NULL argument");
    SKM_REQUIRE(mode == 0 || mode == 1, "update mode must be 0 (recompute) or 1 (incremental)");
    SKM_TRY(enter(L->ds->ctx));
    if (mode == 1 && !L->acc_local) {
        const int64_t p = L->ds->p, n = L->ds->n, K = L->K;
        SKM_TRY(dev_alloc((void **)&L->acc_local, sizeof(double) * (2 * p * K + K + 1), "acc_local"));
        SKM_TRY(skm_big_alloc(L->ctx, (void **)&L->assign_prev, sizeof(int32_t) * n, "assign_prev"));
        SKM_TRY(skm_big_alloc(L->ctx, (void **)&L->changed, sizeof(int32_t) * n, "changed"));
        SKM_TRY(dev_alloc((void **)&L->nchanged, sizeof(int) * 4, "nchanged"));
    }
    L->update_mode = mode;
    L->acc_valid = false;
    return SKM_OK;
}

extern "C" int skm_lloyd_accumulate(skm_lloyd *L)
{
    SKM_REQUIRE(L, "NULL argument");
    skm_ctx *ctx = L->ds->ctx;
    SKM_TRY(enter(ctx));
    if (!L->assigned) { skm_set_error("skm_lloyd_accumulate called before skm_lloyd_assign"); return SKM_ERR_STATE; }
    SkmTimed t(ctx, SKM_T_ACCUM);
    const int64_t p = L->ds->p, n = L->ds->n, K = L->K;
    const double *d64 = L->dist_is_f64 ? L->dist_f64 : nullptr;
    if (L->update_mode == 0) {
        SKM_TRY(skm_launch_accumulate(ctx, L->ds, K, L->assign, L->assign_c, L->dist_f32, d64, L->partials));
        L->last_update_kind = 0; L->last_changed = -1;
        L->accumulated = true;
        return SKM_OK;
    }
    // incremental: acc_local holds this shard's [S | N | counts] for assign_prev.  The sums only depend on the
    // assignments, so columns that kept theirs contribute nothing new; a full recompute every 64 incremental
    // iterations bounds the drift of the +/- updates (each is one fp64 rounding).
    const int64_t nacc = 2 * p * K + K;
    bool full = !L->acc_valid || L->incr_run >= 64;
    int64_t nch = 0;
    if (!full) {
        SKM_TRY(skm_launch_diff_assign(ctx, n, K, L->assign, L->assign_prev, L->dist_f32, d64, L->changed, L->nchanged,
                                       L->acc_local + nacc));
        SKM_CUDA(cudaMemcpyAsync(ctx->h_flag + 12, L->nchanged, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SKM_CUDA(cudaStreamSynchronize(ctx->stream));
        nch = ctx->h_flag[12];
        if (nch > n / 16) full = true;                      // many movers: the row-major pass is cheaper
    }
    if (full) {
        SKM_TRY(skm_launch_accumulate(ctx, L->ds, K, L->assign, L->assign_c, L->dist_f32, d64, L->acc_local));
        SKM_CUDA(cudaMemcpyAsync(L->assign_prev, L->assign, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        L->acc_valid = true; L->incr_run = 0; L->last_update_kind = 0; L->last_changed = -1;
    } else {
        SKM_TRY(skm_launch_move_changed(ctx, L->ds, K, L->assign, L->assign_prev, L->changed, L->nchanged, nch, L->acc_local));
        L->incr_run += 1; L->last_update_kind = nch ? 1 : 2; L->last_changed = nch;
    }
    SKM_CUDA(cudaMemcpyAsync(L->partials, L->acc_local, sizeof(double) * (nacc + 1), cudaMemcpyDeviceToDevice, ctx->stream));
    L->accumulated = true;
    return SKM_OK;
}

extern "C" int skm_lloyd_last_update(skm_lloyd *L, int *kind, int64_t *n_changed)
{
    SKM_REQUIRE(L, "NULL argument");
    if (kind) *kind = L->last_update_kind;
    if (n_changed) *n_changed = L->last_changed;
    return SKM_OK;
}

extern "C" void *skm_lloyd_partials(skm_lloyd *L, int64_t *n_doubles)
{
    if (!L) return nullptr;
    if (n_doubles) *n_doubles = 2 * L->ds->p * L->K + L->K + 1;
    return L->partials;
}

static int read_stats(skm_lloyd *L, skm_iter_stats *stats)
{
    skm_ctx *ctx = L->ds->ctx;
    const int64_t p = L->ds->p, K = L->K;
    std::vector<double> tail(K + 1);
    SKM_CUDA(cudaMemcpyAsync(L->h_stats, L->stats, sizeof(double) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SKM_CUDA(cudaMemcpyAsync(ctx->h_flag, L->nflag, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SKM_TRY(d2h_sync(ctx, tail.data(), L->partials + 2 * p * K, sizeof(double) * (K + 1)));
    if (L->last_tc[0] == -2) { L->last_tc[0] = ctx->h_flag[2]; L->last_tc[1] = ctx->h_flag[1]; }
    int64_t n_empty = 0, npts = 0;
    for (int64_t k = 0; k < K; ++k) {
        L->h_counts[k] = (int64_t)llround(tail[k]);
        npts += L->h_counts[k];
        if (L->h_counts[k] == 0) ++n_empty;
    }
    L->last_rechecked = ctx->h_flag[0];
    if (stats) {
        stats->dff = sqrt(L->h_stats[0]);
        stats->sumsq = tail[K];
        stats->n_empty = n_empty;
        stats->n_rechecked = L->last_rechecked;
        stats->n_points = npts;
        stats->has_nan = L->h_stats[1] != 0.0;
        stats->reserved = 0;
    }
    return SKM_OK;
}

extern "C" int skm_lloyd_finalize(skm_lloyd *L, double gamma, int ml_correction, skm_iter_stats *stats)
{
    SKM_REQUIRE(L, "NULL argument");
    SKM_TRY(enter(L->ds->ctx));
    if (!L->accumulated) { skm_set_error("skm_lloyd_finalize called before skm_lloyd_accumulate"); return SKM_ERR_STATE; }
    {
        SkmTimed t(L->ds->ctx, SKM_T_FINAL);
        SKM_TRY(skm_launch_finalize(L->ds->ctx, L->ds->p, L->K, L->partials, gamma, ml_correction, L->centers,
                                    L->centers_old, L->stats));
    }
    return read_stats(L, stats);
}

// recompute dff / has_nan after the host patched centre columns (EmptyAction = singleton)
extern "C" int skm_lloyd_refresh_diff(skm_lloyd *L, skm_iter_stats *stats)
{
    SKM_REQUIRE(L, "NULL argument");
    SKM_TRY(enter(L->ds->ctx));
    SKM_TRY(skm_launch_finalize(L->ds->ctx, L->ds->p, L->K, nullptr, 0.0, 0, L->centers, L->centers_old, L->stats));
    return read_stats(L, stats);
}

extern "C" int skm_lloyd_get_counts(skm_lloyd *L, int64_t *counts)
{
    SKM_REQUIRE(L && counts, "NULL argument");
    for (int64_t k = 0; k < L->K; ++k) counts[k] = L->h_counts[k];
    return SKM_OK;
}

extern "C" int skm_lloyd_get_assignments(skm_lloyd *L, int32_t *assign_out, double *dist_out)
{
    SKM_REQUIRE(L, "NULL argument");
    skm_ctx *ctx = L->ds->ctx;
    SKM_TRY(enter(ctx));
    if (!L->assigned) { skm_set_error("no assignments yet"); return SKM_ERR_STATE; }
    const int64_t n = L->ds->n;
    if (n == 0) return SKM_OK;
    // 1-based indices (MATLAB) and doubles are produced on the device: a host loop over 1e7 columns cost 50 ms
    void *sa = nullptr, *sd = nullptr;
    int rc = SKM_OK;
    do {
        if (assign_out && (rc = skm_big_alloc(ctx, &sa, sizeof(int32_t) * n, "assignment export"))) break;
        if (dist_out && !L->dist_is_f64 && (rc = skm_big_alloc(ctx, &sd, sizeof(double) * n, "distance export"))) break;
        if ((rc = skm_launch_export(ctx, n, assign_out ? L->assign : nullptr, (int32_t *)sa,
                                    (dist_out && !L->dist_is_f64) ? L->dist_f32 : nullptr, (double *)sd))) break;
        if (assign_out && (rc = skm_d2h_pageable(ctx, assign_out, sa, sizeof(int32_t) * n))) break;
        if (dist_out && (rc = skm_d2h_pageable(ctx, dist_out, L->dist_is_f64 ? (const void *)L->dist_f64 : sd, sizeof(double) * n))) break;
    } while (0);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    skm_big_free(ctx, sa);
    skm_big_free(ctx, sd);
    if (rc == SKM_OK && e != cudaSuccess) { skm_set_error("read-back failed: %s", cudaGetErrorString(e)); rc = SKM_ERR_CUDA; }
    return rc;
}

extern "C" int skm_lloyd_argmax_distance(skm_lloyd *L, double *maxdist, int64_t *j)
{
    SKM_REQUIRE(L && maxdist && j, "NULL argument");
    skm_ctx *ctx = L->ds->ctx;
    SKM_TRY(enter(ctx));
    if (!L->assigned) { skm_set_error("no assignments yet"); return SKM_ERR_STATE; }
    if (L->ds->n == 0) { *maxdist = -1.0; *j = -1; return SKM_OK; }
    DevBuf v, i;
    SKM_TRY(v.alloc(sizeof(double)));
    SKM_TRY(i.alloc(sizeof(int64_t)));
    SKM_TRY(skm_launch_argmax(ctx, L->ds->n, L->dist_f32, L->dist_is_f64 ? L->dist_f64 : nullptr, v.as<double>(), i.as<int64_t>()));
    SKM_CUDA(cudaMemcpyAsync(maxdist, v.ptr, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return d2h_sync(ctx, j, i.ptr, sizeof(int64_t));
}

extern "C" const char *skm_lloyd_kernel_name(skm_lloyd *L)
{
    static thread_local char name[96];
    name[0] = 0;
    if (!L) return name;
    FastPlan pl;
    const skm_dataset *ds = L->ds;
    if (ds->store_dtype == SKM_F32 && L->table && skm_fast_plan(ds->ctx, ds->p, L->K, &pl, ds->max_col_nnz)) {
        if (tc_wanted(L) && ds->tsb) snprintf(name, sizeof name, "k_tcs_filter<%d> + k_assign_bounded", skm_tcs_bn(L->K));
        else if (L->last_prune[0] >= 0 && L->last_pruned) {
            Prefix16Plan hp;
            if (prune_half_table(ds->ctx, ds->p, L->K, &hp))
                snprintf(name, sizeof name, "k_prefix16<%d> x%d on %lld of the entry pairs + k_assign_bounded", hp.kc, hp.nchunks,
                         (long long)L->last_prune[1]);
            else
                snprintf(name, sizeof name, "k_assign_fast<%d>%s x%d on %lld of the entry pairs + k_assign_bounded", pl.kc,
                         pl.dual8 ? " (dual table)" : "", pl.nchunks, (long long)L->last_prune[1]);
        }
        else if (pl.mode64) snprintf(name, sizeof name, "k_assign_fast64<%d>", pl.kc);
        else snprintf(name, sizeof name, "k_assign_fast<%d>%s%s x%d", pl.kc, pl.global_table ? " (global table)" : "",
                      pl.dual8 ? " (dual table)" : "", pl.nchunks);
    } else snprintf(name, sizeof name, "k_exact_assign");
    return name;
}

extern "C" void *skm_lloyd_assign_ptr(skm_lloyd *L) { return L ? L->assign : nullptr; }
extern "C" void *skm_lloyd_dist_ptr(skm_lloyd *L, int *dtype)
{
    if (!L) return nullptr;
    if (L->dist_is_f64) { if (dtype) *dtype = SKM_F64; return L->dist_f64; }
    if (dtype) *dtype = SKM_F32;
    return L->dist_f32;
}

// ---------------------------------------------------------------------------
// operator-level calls on a resident dataset
// ---------------------------------------------------------------------------
extern "C" int skm_assign(skm_dataset *ds, const double *centers, int64_t K, int has_gamma, double gamma,
                          int32_t *assign_out, double *dist_out)
{
    SKM_REQUIRE(ds && centers, "NULL argument");
    skm_lloyd *L = nullptr;
    SKM_TRY(skm_lloyd_create(ds, K, &L));
    int rc = skm_lloyd_set_centers(L, centers);
    if (rc == SKM_OK) rc = skm_lloyd_assign(L, has_gamma, gamma);
    if (rc == SKM_OK) rc = skm_lloyd_get_assignments(L, assign_out, dist_out);
    skm_lloyd_destroy(L);
    return rc;
}

// sparse-centres branch on the Lloyd state (findClusterAssignments.m:63-75): exact fp64
extern "C" int skm_lloyd_assign_sparse(skm_lloyd *L, int has_gamma, double gamma)
{
    SKM_REQUIRE(L, "NULL argument");
    skm_dataset *ds = L->ds;
    skm_ctx *ctx = ds->ctx;
    SKM_TRY(enter(ctx));
    L->lb_valid = false;                                   // the bounds refer to dense-centre distances
    const int64_t p = ds->p, K = L->K;
    DevBuf mask, xdiv;
    SKM_TRY(mask.alloc((size_t)(p + 1) * K));
    SKM_TRY(xdiv.alloc(sizeof(double) * K));
    SKM_TRY(skm_launch_prep_centers(ctx, p, K, L->centers, has_gamma, gamma, L->cscaled_t, mask.as<uint8_t>(),
                                    xdiv.as<double>()));
    ExactArgs ea = exact_args(ds, K, L->cscaled_t);
    ea.mask = mask.as<uint8_t>();
    ea.xdiv = xdiv.as<double>();
    SKM_TRY(skm_launch_exact_assign(ctx, ea, L->assign, L->dist_f64, L->dist_f32, nullptr, nullptr, 0));
    SKM_CUDA(cudaMemsetAsync(L->nflag, 0, sizeof(int), ctx->stream));
    SKM_CUDA(cudaStreamSynchronize(ctx->stream));        // mask/xdiv are freed on return
    L->assigned = true;
    L->accumulated = false;
    L->dist_is_f64 = true;
    return SKM_OK;
}

extern "C" int skm_assign_sparse_centers(skm_dataset *ds, const double *centers, int64_t K, int has_gamma,
                                         double gamma, int32_t *assign_out, double *dist_out)
{
    SKM_REQUIRE(ds && centers, "NULL argument");
    skm_lloyd *L = nullptr;
    SKM_TRY(skm_lloyd_create_ex(ds, K, 1, &L));
    int rc = skm_lloyd_set_centers(L, centers);
    if (rc == SKM_OK) rc = skm_lloyd_assign_sparse(L, has_gamma, gamma);
    if (rc == SKM_OK) rc = skm_lloyd_get_assignments(L, assign_out, dist_out);
    skm_lloyd_destroy(L);
    return rc;
}

extern "C" int skm_masked_distances(skm_dataset *ds, const double *centers, int64_t K, double *dist)
{
    SKM_REQUIRE(ds && centers && dist, "NULL argument");
    skm_ctx *ctx = ds->ctx;
    SKM_TRY(enter(ctx));
    SKM_REQUIRE(K >= 1, "K must be >= 1");
    const int64_t p = ds->p, n = ds->n;
    DevBuf dc, ct, tmp;
    SKM_TRY(dc.alloc(sizeof(double) * p * K));
    SKM_TRY(ct.alloc(sizeof(double) * (p + 1) * K));
    SKM_TRY(h2d(ctx, dc.ptr, centers, sizeof(double) * p * K));
    SKM_TRY(skm_launch_prep_centers(ctx, p, K, dc.as<double>(), 0, 0.0, ct.as<double>(), nullptr, nullptr));
    ExactArgs ea = exact_args(ds, K, ct.as<double>());
    // K x n in column chunks so the device temporary stays bounded
    const int64_t chunk = std::max<int64_t>(1, (int64_t)(256LL << 20) / (8 * K));
    SKM_TRY(tmp.alloc(sizeof(double) * std::min(chunk, std::max<int64_t>(n, 1)) * K));
    for (int64_t j0 = 0; j0 < n; j0 += chunk) {
        const int64_t j1 = std::min(n, j0 + chunk);
        SKM_TRY(skm_launch_exact_dist(ctx, ea, j0, j1, tmp.as<double>()));
        SKM_TRY(d2h_sync(ctx, dist + j0 * K, tmp.ptr, sizeof(double) * (j1 - j0) * K));
    }
    SKM_CUDA(cudaStreamSynchronize(ctx->stream));
    return SKM_OK;
}

// ---------------------------------------------------------------------------
// Level 1
// ---------------------------------------------------------------------------
extern "C" int skm_sparse_matrix_minus_cluster(skm_ctx *ctx, int64_t p, int64_t n, int64_t K,
                                               const uint64_t *jc, const uint64_t *ir, const double *pr,
                                               const double *centers, int has_beta, double beta, double *dist)
{
    SKM_TRY(enter(ctx));
    SKM_REQUIRE(jc && centers && dist, "NULL argument");
    SKM_REQUIRE(K >= 1, "Center must have at least one column");
    if (has_beta && K != 1) {
        // SparseMatrixMinusCluster.c:119-120
        skm_set_error("Have not yet implemented case for using 'beta' with p x k (k!=1) centers");
        return SKM_ERR_INVALID;
    }
    skm_dataset *ds = nullptr;
    SKM_TRY(skm_dataset_create_csc(ctx, p, n, jc, SKM_I64, ir, SKM_I64, pr, SKM_F64, SKM_F64, 0, &ds));
    int rc;
    if (!has_beta) rc = skm_masked_distances(ds, centers, K, dist);
    else {
        rc = SKM_OK;
        DevBuf dc, dd;
        do {
            if ((rc = dc.alloc(sizeof(double) * p))) break;
            if ((rc = dd.alloc(sizeof(double) * std::max<int64_t>(n, 1)))) break;
            if ((rc = h2d(ctx, dc.ptr, centers, sizeof(double) * p))) break;
            ExactArgs ea = exact_args(ds, 1, dc.as<double>());
            if ((rc = skm_launch_exact_dist_beta(ctx, ea, beta, dd.as<double>()))) break;
            rc = d2h_sync(ctx, dist, dd.ptr, sizeof(double) * n);
        } while (0);
    }
    skm_dataset_destroy(ds);
    return rc;
}

static int inner_common(skm_ctx *ctx, int64_t p, int64_t n, const uint64_t *jc, const uint64_t *ir,
                        const double *pr, const double *c, double *inner, double *normsq)
{
    SKM_TRY(enter(ctx));
    skm_dataset *ds = nullptr;
    // ColumnNormSq never looks at the row indices: fabricate zeros when ir is NULL
    std::vector<uint64_t> zeros;
    if (!ir) { zeros.assign((size_t)jc[n], 0); ir = zeros.data(); }
    SKM_TRY(skm_dataset_create_csc(ctx, p, n, jc, SKM_I64, ir, SKM_I64, pr, SKM_F64, SKM_F64, 0, &ds));
    int rc = SKM_OK;
    DevBuf dc, di, dn;
    do {
        if (c) {
            if ((rc = dc.alloc(sizeof(double) * std::max<int64_t>(p, 1)))) break;
            if ((rc = h2d(ctx, dc.ptr, c, sizeof(double) * p))) break;
        }
        if (inner && (rc = di.alloc(sizeof(double) * std::max<int64_t>(n, 1)))) break;
        if (normsq && (rc = dn.alloc(sizeof(double) * std::max<int64_t>(n, 1)))) break;
        if ((rc = skm_launch_inner_product(ctx, n, ds->colptr, ds->rowidx, (const double *)ds->val,
                                           c ? dc.as<double>() : nullptr, inner ? di.as<double>() : nullptr,
                                           normsq ? dn.as<double>() : nullptr))) break;
        if (inner && (rc = d2h_sync(ctx, inner, di.ptr, sizeof(double) * n))) break;
        if (normsq && (rc = d2h_sync(ctx, normsq, dn.ptr, sizeof(double) * n))) break;
        rc = d2h_sync(ctx, nullptr, nullptr, 0);
    } while (0);
    skm_dataset_destroy(ds);
    return rc;
}

extern "C" int skm_sparse_matrix_inner_product(skm_ctx *ctx, int64_t p, int64_t n, const uint64_t *jc,
                                               const uint64_t *ir, const double *pr, const double *c,
                                               double *inner, double *normsq)
{
    SKM_REQUIRE(jc && c && inner, "NULL argument");
    SKM_REQUIRE(jc[n] == 0 || ir, "ir is NULL");
    return inner_common(ctx, p, n, jc, ir, pr, c, inner, normsq);
}

extern "C" int skm_sparse_matrix_column_normsq(skm_ctx *ctx, int64_t p, int64_t n, const uint64_t *jc,
                                               const double *pr, double *normsq)
{
    SKM_REQUIRE(jc && normsq, "NULL argument");
    return inner_common(ctx, p, n, jc, nullptr, pr, nullptr, nullptr, normsq);
}

extern "C" int skm_hadamard(skm_ctx *ctx, int64_t m, int64_t n, const double *x, double *w)
{
    SKM_TRY(enter(ctx));
    SKM_REQUIRE(x && w, "NULL argument");
    // hadamard.c:97-111,134-140
    SKM_REQUIRE(m >= 2 && (m & (m - 1)) == 0, "hadamard: number of rows must be a power of two >= 2");
    if (n == 0) return SKM_OK;
    // column chunks bound the device footprint
    const int64_t chunk = std::max<int64_t>(1, (int64_t)(1LL << 30) / (8 * m));
    DevBuf d;
    SKM_TRY(d.alloc(sizeof(double) * m * std::min(chunk, n)));
    for (int64_t j0 = 0; j0 < n; j0 += chunk) {
        const int64_t j1 = std::min(n, j0 + chunk);
        SKM_TRY(h2d(ctx, d.ptr, x + j0 * m, sizeof(double) * m * (j1 - j0)));
        SKM_TRY(skm_launch_fwht_f64(ctx, m, j1 - j0, d.as<double>(), nullptr, 0.0));
        SKM_TRY(d2h_sync(ctx, w + j0 * m, d.ptr, sizeof(double) * m * (j1 - j0)));
    }
    return SKM_OK;
}

// ---------------------------------------------------------------------------
// preconditioning
// ---------------------------------------------------------------------------
namespace {
__global__ void k_pad_cast(int64_t p, int64_t p2, int64_t n, const double *__restrict__ x, float *__restrict__ y32,
                           double *__restrict__ y64)
{
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; idx < p2 * n; idx += stride) {
        const int64_t col = idx / p2, r = idx % p2;
        const double v = r < p ? x[col * p + r] : 0.0;
        if (y32) y32[idx] = (float)v;
        else y64[idx] = v;
    }
}
__global__ void k_f32_to_f64(int64_t total, const float *__restrict__ a, double *__restrict__ b)
{
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; idx < total; idx += stride) b[idx] = (double)a[idx];
}
}  // namespace

extern "C" int skm_mix_hadamard(skm_ctx *ctx, int64_t p, int64_t p2, int64_t n, const double *x,
                                const double *signs, int compute_dtype, double *y)
{
    SKM_TRY(enter(ctx));
    SKM_REQUIRE(x && signs && y, "NULL argument");
    SKM_REQUIRE(p2 >= 2 && (p2 & (p2 - 1)) == 0 && p <= p2, "p2 must be a power of two >= max(2,p)");
    SKM_REQUIRE(compute_dtype == SKM_F32 || compute_dtype == SKM_F64, "bad compute dtype");
    if (n == 0) return SKM_OK;
    const int64_t chunk = std::max<int64_t>(1, (int64_t)(1LL << 29) / (8 * p2));
    const int64_t cn = std::min(chunk, n);
    DevBuf din, dwork, dout, dsign;
    SKM_TRY(din.alloc(sizeof(double) * p * cn));
    SKM_TRY(dout.alloc(sizeof(double) * p2 * cn));
    SKM_TRY(dsign.alloc(sizeof(double) * p2));
    std::vector<float> s32;
    if (compute_dtype == SKM_F32) {
        SKM_TRY(dwork.alloc(sizeof(float) * p2 * cn));
        s32.resize(p2);
        for (int64_t i = 0; i < p2; ++i) s32[i] = (float)signs[i];
        SKM_TRY(h2d(ctx, dsign.ptr, s32.data(), sizeof(float) * p2));
    } else SKM_TRY(h2d(ctx, dsign.ptr, signs, sizeof(double) * p2));
    const int64_t cap = (int64_t)ctx->sm_count * 32;
    for (int64_t j0 = 0; j0 < n; j0 += chunk) {
        const int64_t j1 = std::min(n, j0 + chunk), c = j1 - j0;
        SKM_TRY(h2d(ctx, din.ptr, x + j0 * p, sizeof(double) * p * c));
        int64_t blocks = std::min(cap, (p2 * c + 255) / 256);
        if (compute_dtype == SKM_F32) {
            k_pad_cast<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, p2, c, din.as<double>(), dwork.as<float>(), nullptr);
            SKM_CHECK_LAUNCH(ctx);
            SKM_TRY(skm_launch_fwht_f32(ctx, p2, c, dwork.as<float>(), dsign.as<float>(), sqrtf((float)p2)));
            k_f32_to_f64<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p2 * c, dwork.as<float>(), dout.as<double>());
            SKM_CHECK_LAUNCH(ctx);
        } else {
            k_pad_cast<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, p2, c, din.as<double>(), nullptr, dout.as<double>());
            SKM_CHECK_LAUNCH(ctx);
            SKM_TRY(skm_launch_fwht_f64(ctx, p2, c, dout.as<double>(), dsign.as<double>(), sqrt((double)p2)));
        }
        SKM_TRY(d2h_sync(ctx, y + j0 * p2, dout.ptr, sizeof(double) * p2 * c));
    }
    return SKM_OK;
}

extern "C" int skm_fwht_f32_inplace(skm_ctx *ctx, int64_t p2, int64_t n, float *x_dev, const float *signs_dev)
{
    SKM_TRY(enter(ctx));
    SKM_REQUIRE(x_dev, "NULL argument");
    return skm_launch_fwht_f32(ctx, p2, n, x_dev, signs_dev, sqrtf((float)p2));
}

extern "C" int skm_sample_rows(skm_ctx *ctx, int64_t p2, int64_t n, int64_t m, uint64_t seed, int64_t col0,
                               int32_t *rows_dev)
{
    SKM_TRY(enter(ctx));
    SKM_REQUIRE(rows_dev, "NULL argument");
    SKM_TRY(skm_launch_sample_rows(ctx, p2, n, m, seed, col0, rows_dev));
    SKM_CUDA(cudaStreamSynchronize(ctx->stream));
    return SKM_OK;
}

extern "C" int skm_fwht_sample_f32(skm_ctx *ctx, int64_t p2, int64_t n, int64_t m, const float *x_dev,
                                   const float *signs_dev, const int32_t *rows_dev, uint64_t seed, int64_t col0,
                                   skm_dataset **out)
{
    SKM_TRY(enter(ctx));
    SKM_REQUIRE(x_dev && signs_dev && out, "NULL argument");
    *out = nullptr;
    skm_dataset *ds = new (std::nothrow) skm_dataset();
    if (!ds) { skm_set_error("out of host memory"); return SKM_ERR_NOMEM; }
    memset(ds, 0, sizeof *ds);
    ds->ctx = ctx;
    ds->p = p2; ds->n = n; ds->nnz = n * m;
    ds->store_dtype = SKM_F32;
    int rc = SKM_OK;
    do {
        if ((rc = skm_big_alloc(ds->ctx, (void **)&ds->colptr, sizeof(int64_t) * (n + 1), "colptr"))) break;
        if ((rc = skm_big_alloc(ds->ctx, (void **)&ds->rowidx, sizeof(int32_t) * ds->nnz, "rowidx"))) break;
        if ((rc = skm_big_alloc(ds->ctx, &ds->val, sizeof(float) * ds->nnz, "val"))) break;
        if (n == 0) { if ((rc = (cudaMemsetAsync(ds->colptr, 0, sizeof(int64_t), ctx->stream) == cudaSuccess) ? SKM_OK : SKM_ERR_CUDA)) break; }
        if ((rc = skm_launch_fwht_sample_f32(ctx, p2, n, m, x_dev, signs_dev, rows_dev, seed, col0, ds->colptr, ds->rowidx,
                                             (float *)ds->val))) break;
        rc = dataset_finish(ds);
    } while (0);
    if (rc != SKM_OK) { skm_dataset_destroy(ds); return rc; }
    *out = ds;
    return SKM_OK;
}


namespace {
template <typename T>
__global__ void k_pad_cast_scaled(int64_t p, int64_t p2, int64_t n, const T *__restrict__ x, double scale, float *__restrict__ y)
{
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; idx < p2 * n; idx += stride) {
        const int64_t col = idx / p2, r = idx % p2;
        y[idx] = r < p ? (float)((double)x[col * p + r] * scale) : 0.f;
    }
}
__global__ void k_fill_colptr(int64_t n, int64_t m, int64_t *__restrict__ colptr)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; j <= n; j += stride) colptr[j] = j * m;
}
}  // namespace

// Whole precondition + sample pipeline from a dense HOST matrix (p x n column-major, points are
// columns): chunks cross PCIe, are zero-padded to p2 rows, multiplied by (1+2*eps)
// (kmeans_sparsified.m:292) and cast to fp32, then run through the fused sign-flip + FWHT +
// on-device row sample kernel straight into the resident dataset.
extern "C" int skm_dataset_from_dense_host(skm_ctx *ctx, int64_t p, int64_t p2, int64_t n, const void *x, int x_type,
                                           const double *signs, int64_t m, uint64_t seed, int64_t col0,
                                           int64_t chunk_cols, skm_dataset **out)
{
    SKM_TRY(enter(ctx));
    SKM_REQUIRE(x && signs && out, "NULL argument");
    *out = nullptr;
    SKM_REQUIRE(x_type == SKM_F32 || x_type == SKM_F64, "x_type must be SKM_F32 or SKM_F64");
    SKM_REQUIRE(p2 >= 32 && p2 <= 32768 && (p2 & (p2 - 1)) == 0 && p >= 1 && p <= p2, "need a power of two 32 <= p2 <= 32768 and p <= p2");
    SKM_REQUIRE(m >= 1 && m <= p2 && n >= 0, "need 1 <= m <= p2");
    const size_t xs = x_type == SKM_F32 ? 4 : 8;
    if (chunk_cols <= 0) chunk_cols = std::max<int64_t>(1, (int64_t)(512LL << 20) / (int64_t)(p2 * 4));
    chunk_cols = std::min<int64_t>(chunk_cols, std::max<int64_t>(n, 1));
    skm_dataset *ds = new (std::nothrow) skm_dataset();
    if (!ds) { skm_set_error("out of host memory"); return SKM_ERR_NOMEM; }
    memset(ds, 0, sizeof *ds);
    ds->ctx = ctx; ds->p = p2; ds->n = n; ds->nnz = n * m; ds->store_dtype = SKM_F32;
    int rc = SKM_OK;
    do {
        if ((rc = skm_big_alloc(ds->ctx, (void **)&ds->colptr, sizeof(int64_t) * (n + 1), "colptr"))) break;
        if ((rc = skm_big_alloc(ds->ctx, (void **)&ds->rowidx, sizeof(int32_t) * ds->nnz, "rowidx"))) break;
        if ((rc = skm_big_alloc(ds->ctx, &ds->val, sizeof(float) * ds->nnz, "val"))) break;
        DevBuf raw[2], dense, dsign, scratch_colptr;
        std::vector<float> s32(p2);
        for (int64_t i = 0; i < p2; ++i) s32[i] = (float)signs[i];
        if ((rc = dsign.alloc(sizeof(float) * p2))) break;
        if ((rc = h2d(ctx, dsign.ptr, s32.data(), sizeof(float) * p2))) break;
        if ((rc = raw[0].alloc(xs * p * chunk_cols)) || (rc = raw[1].alloc(xs * p * chunk_cols))) break;
        if ((rc = dense.alloc(sizeof(float) * p2 * chunk_cols))) break;
        if ((rc = scratch_colptr.alloc(sizeof(int64_t) * (chunk_cols + 1)))) break;
        cudaStream_t copy_stream;
        if (cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking) != cudaSuccess) { rc = SKM_ERR_CUDA; skm_set_error("cudaStreamCreate failed"); break; }
        cudaEvent_t up[2], freed[2];
        for (int i = 0; i < 2; ++i) { cudaEventCreateWithFlags(&up[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&freed[i], cudaEventDisableTiming); }
        const double scale = 1.0 + 2.0 * 2.220446049250313e-16;
        const int64_t nchunks = (n + chunk_cols - 1) / chunk_cols;
        auto issue = [&](int64_t c) {
            const int64_t j0 = c * chunk_cols, nc = std::min(chunk_cols, n - j0);
            if (c >= 2) cudaStreamWaitEvent(copy_stream, freed[c & 1], 0);
            cudaMemcpyAsync(raw[c & 1].ptr, (const char *)x + (size_t)j0 * p * xs, xs * p * nc, cudaMemcpyHostToDevice, copy_stream);
            cudaEventRecord(up[c & 1], copy_stream);
        };
        if (nchunks > 0) issue(0);
        for (int64_t c = 0; c < nchunks && rc == SKM_OK; ++c) {
            if (c + 1 < nchunks) issue(c + 1);
            const int64_t j0 = c * chunk_cols, nc = std::min(chunk_cols, n - j0);
            cudaStreamWaitEvent(ctx->stream, up[c & 1], 0);
            const int64_t blocks = std::min<int64_t>((p2 * nc + 255) / 256, (int64_t)ctx->sm_count * 32);
            if (x_type == SKM_F32)
                k_pad_cast_scaled<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, p2, nc, raw[c & 1].as<float>(), scale, dense.as<float>());
            else
                k_pad_cast_scaled<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, p2, nc, raw[c & 1].as<double>(), scale, dense.as<float>());
            ctx->launches++;
            cudaEventRecord(freed[c & 1], ctx->stream);
            rc = skm_launch_fwht_sample_f32(ctx, p2, nc, m, dense.as<float>(), dsign.as<float>(), nullptr, seed, col0 + j0,
                                            scratch_colptr.as<int64_t>(), ds->rowidx + j0 * m, (float *)ds->val + j0 * m);
        }
        cudaStreamSynchronize(copy_stream);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(up[i]); cudaEventDestroy(freed[i]); }
        cudaStreamDestroy(copy_stream);
        if (rc != SKM_OK) break;
        k_fill_colptr<<<(unsigned)std::min<int64_t>((n + 256) / 256, (int64_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(n, m, ds->colptr);
        ctx->launches++;
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { skm_set_error("precondition pipeline failed: %s", cudaGetErrorString(e)); rc = SKM_ERR_CUDA; break; }
        rc = dataset_finish(ds);
    } while (0);
    if (rc != SKM_OK) { skm_dataset_destroy(ds); return rc; }
    *out = ds;
    return SKM_OK;
}

// ---------------------------------------------------------------------------
// k-means++
// ---------------------------------------------------------------------------
static int kpp_update_common(skm_dataset *ds, const double *center, int has_gamma, double gamma, int first,
                             int masked, double *sum_d2);
extern "C" int skm_kpp_update(skm_dataset *ds, const double *center, int has_gamma, double gamma, int first,
                              double *sum_d2)
{
    return kpp_update_common(ds, center, has_gamma, gamma, first, 0, sum_d2);
}
extern "C" int skm_kpp_update_sparse(skm_dataset *ds, const double *center, int first, double *sum_d2)
{
    return kpp_update_common(ds, center, 0, 0.0, first, 1, sum_d2);
}
static int kpp_update_common(skm_dataset *ds, const double *center, int has_gamma, double gamma, int first,
                             int masked, double *sum_d2)
{
    SKM_REQUIRE(ds && center, "NULL argument");
    skm_ctx *ctx = ds->ctx;
    SKM_TRY(enter(ctx));
    const int64_t p = ds->p, n = ds->n;
    const int64_t nb = (n + 1023) / 1024;
    if (!ds->kpp_mind) {
        SKM_TRY(dev_alloc((void **)&ds->kpp_mind, sizeof(double) * n, "kpp_mind"));
        SKM_TRY(dev_alloc((void **)&ds->kpp_cum, sizeof(double) * (nb + 1), "kpp_bsum"));
        first = 1;
    }
    std::vector<double> c(center, center + p);
    if (has_gamma) for (int64_t i = 0; i < p; ++i) c[i] = center[i] / gamma;    // full(centers)/gamma, IEEE division
    // scratch kept with the dataset: a cudaMalloc / cudaFree pair per round costs more than the round's kernels at 8 GPUs
    if (!ds->kpp_c) SKM_TRY(dev_alloc((void **)&ds->kpp_c, sizeof(double) * (p + 1) + sizeof(float) * (p + 4), "kpp centre"));
    double *dcp = ds->kpp_c;
    float *c32p = reinterpret_cast<float *>(ds->kpp_c + p + 1);
    SKM_CUDA(cudaStreamSynchronize(ctx->stream));                     // `c` of the previous call is no longer being copied
    SKM_TRY(h2d(ctx, dcp, c.data(), sizeof(double) * p));
    ds->kpp_last_exact = -1;
    if (!first && !masked && skm_kpp_filter_usable(ctx, ds)) {
        // rounds after the first: an fp32 pass proves for most columns that the new centre is farther than their
        // current minimum; only the rest is evaluated in fp64 (values stay bit-identical to the reference's)
        if (!ds->kpp_flag) SKM_TRY(skm_big_alloc(ctx, (void **)&ds->kpp_flag, sizeof(int32_t) * (n + 4), "kpp filter list"));
        SKM_TRY(skm_sell_ensure_any(ds));
        int *nflag = reinterpret_cast<int *>(ds->kpp_flag + n);
        SKM_TRY(skm_launch_kpp_update_filtered(ctx, ds, dcp, ds->kpp_mind, c32p, ds->kpp_flag, nflag));
        SKM_CUDA(cudaMemcpyAsync(ctx->h_flag + 11, nflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        ds->kpp_last_exact = -2;                                      // in h_flag[11] once the block sums have been read
    } else SKM_TRY(skm_launch_kpp_update(ctx, ds, dcp, first, ds->kpp_mind, masked));
    SKM_TRY(skm_launch_scan_sq(ctx, n, ds->kpp_mind, ds->kpp_cum));
    std::vector<double> bs(nb);
    SKM_TRY(d2h_sync(ctx, bs.data(), ds->kpp_cum, sizeof(double) * nb));
    if (ds->kpp_last_exact == -2) ds->kpp_last_exact = ctx->h_flag[11];
    double tot = 0.0;
    for (int64_t b = 0; b < nb; ++b) tot += bs[b];
    if (sum_d2) *sum_d2 = tot;
    return SKM_OK;
}

extern "C" int skm_kpp_pick(skm_dataset *ds, double target, int64_t *j)
{
    SKM_REQUIRE(ds && j, "NULL argument");
    skm_ctx *ctx = ds->ctx;
    SKM_TRY(enter(ctx));
    if (!ds->kpp_mind) { skm_set_error("skm_kpp_pick before skm_kpp_update"); return SKM_ERR_STATE; }
    const int64_t n = ds->n, nb = (n + 1023) / 1024;
    SKM_REQUIRE(n > 0, "empty dataset");
    std::vector<double> bs(nb);
    SKM_TRY(d2h_sync(ctx, bs.data(), ds->kpp_cum, sizeof(double) * nb));
    double acc = 0.0;
    int64_t b = 0;
    for (; b < nb; ++b) {
        if (acc + bs[b] > target) break;
        acc += bs[b];
    }
    if (b >= nb) { *j = n - 1; return SKM_OK; }
    const int64_t i0 = b * 1024, i1 = std::min(n, i0 + 1024);
    std::vector<double> md(i1 - i0);
    SKM_TRY(d2h_sync(ctx, md.data(), ds->kpp_mind + i0, sizeof(double) * (i1 - i0)));
    for (int64_t i = i0; i < i1; ++i) {
        acc += md[i - i0] * md[i - i0];
        if (acc > target) { *j = i; return SKM_OK; }
    }
    *j = i1 - 1;     // rounding between the tree sum and the sequential sum
    return SKM_OK;
}

extern "C" int skm_kpp_get_mindist(skm_dataset *ds, double *mind)
{
    SKM_REQUIRE(ds && mind, "NULL argument");
    SKM_TRY(enter(ds->ctx));
    if (!ds->kpp_mind) { skm_set_error("skm_kpp_get_mindist before skm_kpp_update"); return SKM_ERR_STATE; }
    return d2h_sync(ds->ctx, mind, ds->kpp_mind, sizeof(double) * ds->n);
}
