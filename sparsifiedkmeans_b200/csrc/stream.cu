// stream.cu -- one Lloyd iteration over a sparsified matrix that lives in HOST memory.
//
// The stateless form of the hot path (what a MEX call with X in MATLAB's memory amounts to) and
// the out-of-core form for matrices larger than HBM: columns are cut into chunks; while chunk c
// is converted, laid out (stored-order SELL), assigned (K1 + fp64 re-evaluation) and accumulated
// (K2, atomic variant) on the compute stream, chunk c+1 crosses PCIe on a copy stream into the
// other staging slot.  The bank-aware SELL order and the row-major image are NOT built here:
// they pay for themselves only over several iterations on resident data.  K3 runs once at the
// end.  Results are identical to the resident path's (same kernels, same exactness guard).
#include "common.cuh"
#include <algorithm>
#include <new>
#include <vector>

namespace {

struct Slot {
    DevBuf raw_jc, raw_ir, raw_val;            // staging for host types that need conversion
    DevBuf colptr, rowidx, val, sell, slice_ptr, w2, elems, cub_tmp;
    DevBuf assign, dist, flagged, nflag, flags, out_assign, out_dist;
    cudaEvent_t h2d_done = nullptr, free_ev = nullptr;
    bool used = false;
    ~Slot() { if (h2d_done) cudaEventDestroy(h2d_done); if (free_ev) cudaEventDestroy(free_ev); }
};

__global__ void k_publish(int64_t n, const int32_t *__restrict__ a, const float *__restrict__ d,
                          int32_t *__restrict__ a1, double *__restrict__ d64)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; j < n; j += stride) {
        if (a1) a1[j] = a[j] + 1;               // MATLAB is 1-based
        if (d64) d64[j] = (double)d[j];
    }
}

// everything skm_lloyd_step_host allocates, kept in the context between calls (cudaMalloc /
// cudaFree of gigabytes per call is slow and synchronises the device)
struct StreamCache {
    int64_t p = -1, K = -1, chunk_cols = -1, cap_nnz = -1, cap_sell = -1;
    int jc_type = -1, ir_type = -1, val_type = -1, want_assign = -1, want_dist = -1, table_floats = -1;
    DevBuf dcen, dold, ct, table, cmax, partials, dstats, best2;
    Slot slots[2];
    cudaStream_t copy_stream = nullptr;
    ~StreamCache() { if (copy_stream) cudaStreamDestroy(copy_stream); }
};

void free_stream_cache(void *p) { delete (StreamCache *)p; }

size_t tsize(int t) { return t == SKM_U16 ? 2 : ((t == SKM_F32 || t == SKM_I32) ? 4 : 8); }

int64_t host_index(const void *a, int type, int64_t i)
{
    return type == SKM_I32 ? (int64_t)((const int32_t *)a)[i] : ((const int64_t *)a)[i];
}

}  // namespace

static int step_host_impl(skm_ctx *ctx, int64_t p, int64_t n, const void *jc, int jc_type,
                          const void *ir, int ir_type, const void *val, int val_type,
                          const double *centers, int64_t K, int has_gamma, double gamma_dist,
                          double gamma_update, int ml_correction, int64_t chunk_cols,
                          double *centers_out, int32_t *assign_out, double *dist_out,
                          skm_iter_stats *stats, skm_reduce_fn reduce, void *reduce_user);

extern "C" int skm_lloyd_step_host(skm_ctx *ctx, int64_t p, int64_t n, const void *jc, int jc_type,
                                   const void *ir, int ir_type, const void *val, int val_type,
                                   const double *centers, int64_t K, int has_gamma, double gamma_dist,
                                   double gamma_update, int ml_correction, int64_t chunk_cols,
                                   double *centers_out, int32_t *assign_out, double *dist_out,
                                   skm_iter_stats *stats, skm_reduce_fn reduce, void *reduce_user)
{
    const int rc = step_host_impl(ctx, p, n, jc, jc_type, ir, ir_type, val, val_type, centers, K, has_gamma, gamma_dist,
                                  gamma_update, ml_correction, chunk_cols, centers_out, assign_out, dist_out, stats, reduce,
                                  reduce_user);
    if (rc != SKM_OK && ctx) {
        // an error exit may leave the H2D copy of the next chunk in flight on the copy stream, reading the
        // caller's host buffers, and result copies pending on the compute stream: drain both before returning
        // so that "no host pointer is retained after a call returns" also holds on failure
        StreamCache *sc = (StreamCache *)ctx->stream_cache;
        if (sc && sc->copy_stream) cudaStreamSynchronize(sc->copy_stream);
        cudaStreamSynchronize(ctx->stream);
        cudaGetLastError();
    }
    return rc;
}

static int step_host_impl(skm_ctx *ctx, int64_t p, int64_t n, const void *jc, int jc_type,
                          const void *ir, int ir_type, const void *val, int val_type,
                          const double *centers, int64_t K, int has_gamma, double gamma_dist,
                          double gamma_update, int ml_correction, int64_t chunk_cols,
                          double *centers_out, int32_t *assign_out, double *dist_out,
                          skm_iter_stats *stats, skm_reduce_fn reduce, void *reduce_user)
{
    SKM_REQUIRE(ctx && jc && centers && centers_out, "NULL argument");
    SKM_CUDA(cudaSetDevice(ctx->device));
    SKM_REQUIRE(p >= 1 && n >= 0 && K >= 1, "bad dimensions");
    SKM_REQUIRE(jc_type == SKM_I32 || jc_type == SKM_I64, "jc_type must be SKM_I32 or SKM_I64");
    SKM_REQUIRE(ir_type == SKM_I32 || ir_type == SKM_I64 || ir_type == SKM_U16, "ir_type must be SKM_I32, SKM_I64 or SKM_U16");
    SKM_REQUIRE(ir_type != SKM_U16 || p <= 65536, "SKM_U16 row indices need p <= 65536");
    SKM_REQUIRE(val_type == SKM_F32 || val_type == SKM_F64, "val_type must be SKM_F32 or SKM_F64");
    const int64_t nnz = host_index(jc, jc_type, n);
    SKM_REQUIRE(nnz >= 0 && host_index(jc, jc_type, 0) == 0, "invalid CSC column pointers");
    SKM_REQUIRE(nnz == 0 || (ir && val), "ir/val are NULL but nnz > 0");
    FastPlan pl;
    if (!skm_fast_plan(ctx, p, K, &pl)) {
        skm_set_error("skm_lloyd_step_host: no assignment plan for p=%lld K=%lld", (long long)p, (long long)K);
        return SKM_ERR_UNSUPPORTED;
    }

    // ---- chunking: ~32 M stored entries per chunk unless the caller fixed the width ----
    const double avg = n > 0 ? (double)nnz / (double)n : 0.0;
    if (chunk_cols <= 0) chunk_cols = (int64_t)std::max(1024.0, 33554432.0 / std::max(avg, 1.0));
    chunk_cols = std::min<int64_t>(std::max<int64_t>(chunk_cols, 32), std::max<int64_t>(n, 32));
    chunk_cols = (chunk_cols + 31) & ~(int64_t)31;
    const int64_t cap_nnz = (int64_t)(avg * chunk_cols * 1.5) + 4096;
    const int64_t cap_sell = cap_nnz / 2 + 32 * (chunk_cols / 32 + 1) * 2;      // int4 units
    (void)0;

    // ---- device state: reused from the previous call when the shapes match ----
    const int64_t npart = 2 * p * K + K + 1;
    const int table_floats = (int)((p + 1) * pl.ks * pl.nchunks);
    StreamCache *sc = (StreamCache *)ctx->stream_cache;
    const bool reuse = sc && sc->p == p && sc->K == K && sc->chunk_cols == chunk_cols && sc->cap_nnz >= cap_nnz &&
                       sc->cap_sell >= cap_sell && sc->jc_type == jc_type && sc->ir_type == ir_type &&
                       sc->val_type == val_type && sc->want_assign >= (assign_out != nullptr) &&
                       sc->want_dist >= (dist_out != nullptr) && sc->table_floats == table_floats;
    if (!reuse) {
        if (sc) { SKM_CUDA(cudaStreamSynchronize(ctx->stream)); delete sc; ctx->stream_cache = nullptr; }
        sc = new (std::nothrow) StreamCache();
        if (!sc) { skm_set_error("out of host memory"); return SKM_ERR_NOMEM; }
        ctx->stream_cache = sc;
        ctx->stream_cache_free = free_stream_cache;
        int rc = SKM_OK;
        do {
            if ((rc = sc->dcen.alloc(sizeof(double) * p * K))) break;
            if ((rc = sc->dold.alloc(sizeof(double) * p * K))) break;
            if ((rc = sc->ct.alloc(sizeof(double) * (p + 1) * K))) break;
            if ((rc = sc->table.alloc(sizeof(float) * (size_t)table_floats))) break;
            if ((rc = sc->cmax.alloc(16))) break;
            if ((rc = sc->partials.alloc(sizeof(double) * npart))) break;
            if ((rc = sc->dstats.alloc(sizeof(double) * 8))) break;
            if (pl.nchunks > 1 && (rc = sc->best2.alloc(sizeof(float) * 2 * chunk_cols))) break;
            if (cudaStreamCreateWithFlags(&sc->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { rc = SKM_ERR_CUDA; skm_set_error("cudaStreamCreate failed"); break; }
            const int64_t nsl = chunk_cols / 32 + 1;
            for (Slot &s : sc->slots) {
                if ((rc = s.raw_jc.alloc(tsize(jc_type) * (chunk_cols + 1)))) break;
                if (ir_type != SKM_I32 && (rc = s.raw_ir.alloc(tsize(ir_type) * cap_nnz))) break;
                if (val_type != SKM_F32 && (rc = s.raw_val.alloc(tsize(val_type) * cap_nnz))) break;
                if ((rc = s.colptr.alloc(sizeof(int64_t) * (chunk_cols + 1)))) break;
                if ((rc = s.rowidx.alloc(sizeof(int32_t) * cap_nnz))) break;
                if ((rc = s.val.alloc(sizeof(float) * cap_nnz))) break;
                if ((rc = s.sell.alloc(sizeof(int4) * cap_sell))) break;
                if ((rc = s.slice_ptr.alloc(sizeof(int64_t) * (nsl + 1)))) break;
                if ((rc = s.w2.alloc(sizeof(int32_t) * (nsl + 1)))) break;
                if ((rc = s.elems.alloc(sizeof(int64_t) * (nsl + 1)))) break;
                if ((rc = s.cub_tmp.alloc(skm_sell_scan_tmp_bytes(nsl) + 256))) break;
                if ((rc = s.assign.alloc(sizeof(int32_t) * chunk_cols))) break;
                if ((rc = s.dist.alloc(sizeof(float) * chunk_cols))) break;
                if ((rc = s.flagged.alloc(sizeof(int32_t) * chunk_cols))) break;
                if ((rc = s.nflag.alloc(16))) break;
                if ((rc = s.flags.alloc(16))) break;
                if (assign_out && (rc = s.out_assign.alloc(sizeof(int32_t) * chunk_cols))) break;
                if (dist_out && (rc = s.out_dist.alloc(sizeof(double) * chunk_cols))) break;
                if (cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming) != cudaSuccess ||
                    cudaEventCreateWithFlags(&s.free_ev, cudaEventDisableTiming) != cudaSuccess) { rc = SKM_ERR_CUDA; skm_set_error("cudaEventCreate failed"); break; }
            }
        } while (0);
        if (rc != SKM_OK) { delete sc; ctx->stream_cache = nullptr; return rc; }
        sc->p = p; sc->K = K; sc->chunk_cols = chunk_cols; sc->cap_nnz = cap_nnz; sc->cap_sell = cap_sell;
        sc->jc_type = jc_type; sc->ir_type = ir_type; sc->val_type = val_type;
        sc->want_assign = assign_out != nullptr; sc->want_dist = dist_out != nullptr; sc->table_floats = table_floats;
    }
    DevBuf &dcen = sc->dcen, &dold = sc->dold, &ct = sc->ct, &table = sc->table, &cmax = sc->cmax,
           &partials = sc->partials, &dstats = sc->dstats, &best2 = sc->best2;
    Slot *slots = sc->slots;
    cudaStream_t copy_stream = sc->copy_stream;
    slots[0].used = slots[1].used = false;

    cudaStream_t cs = ctx->stream;
    SKM_CUDA(cudaMemcpyAsync(dcen.ptr, centers, sizeof(double) * p * K, cudaMemcpyHostToDevice, cs));
    SKM_CUDA(cudaMemsetAsync(partials.ptr, 0, sizeof(double) * npart, cs));
    SKM_CUDA(cudaMemsetAsync(slots[0].flags.ptr, 0, 16, cs));
    SKM_CUDA(cudaMemsetAsync(slots[1].flags.ptr, 0, 16, cs));
    SKM_TRY(skm_launch_prep_centers(ctx, p, K, dcen.as<double>(), has_gamma, gamma_dist, ct.as<double>(), nullptr, nullptr));
    SKM_TRY(skm_launch_build_table(ctx, p, K, ct.as<double>(), pl, table.as<float>(), cmax.as<float>()));

    // ---- chunk list (a chunk never exceeds cap_nnz entries) ----
    std::vector<int64_t> bounds{0};
    while (bounds.back() < n) {
        int64_t j0 = bounds.back(), j1 = std::min(n, j0 + chunk_cols);
        const int64_t t0 = host_index(jc, jc_type, j0);
        while (j1 > j0 + 1 && host_index(jc, jc_type, j1) - t0 > cap_nnz) j1 = j0 + (j1 - j0) / 2;
        if (host_index(jc, jc_type, j1) - t0 > cap_nnz) {
            skm_set_error("skm_lloyd_step_host: column %lld holds more entries than a chunk can stage", (long long)j0);
            return SKM_ERR_UNSUPPORTED;
        }
        bounds.push_back(j1);
    }
    const int nchunks = (int)bounds.size() - 1;
    int64_t rechecked = 0;

    auto issue_h2d = [&](int c) -> int {
        Slot &s = slots[c & 1];
        const int64_t j0 = bounds[c], j1 = bounds[c + 1], nc = j1 - j0;
        const int64_t t0 = host_index(jc, jc_type, j0), t1 = host_index(jc, jc_type, j1), nz = t1 - t0;
        if (s.used) SKM_CUDA(cudaStreamWaitEvent(copy_stream, s.free_ev, 0));
        SKM_CUDA(cudaMemcpyAsync(s.raw_jc.ptr, (const char *)jc + (size_t)j0 * tsize(jc_type), tsize(jc_type) * (nc + 1),
                                 cudaMemcpyHostToDevice, copy_stream));
        if (nz > 0) {
            void *dir = ir_type == SKM_I32 ? s.rowidx.ptr : s.raw_ir.ptr;
            void *dvl = val_type == SKM_F32 ? s.val.ptr : s.raw_val.ptr;
            SKM_CUDA(cudaMemcpyAsync(dir, (const char *)ir + (size_t)t0 * tsize(ir_type), tsize(ir_type) * nz,
                                     cudaMemcpyHostToDevice, copy_stream));
            SKM_CUDA(cudaMemcpyAsync(dvl, (const char *)val + (size_t)t0 * tsize(val_type), tsize(val_type) * nz,
                                     cudaMemcpyHostToDevice, copy_stream));
        }
        SKM_CUDA(cudaEventRecord(s.h2d_done, copy_stream));
        s.used = true;
        return SKM_OK;
    };

    if (nchunks > 0) SKM_TRY(issue_h2d(0));
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) SKM_TRY(issue_h2d(c + 1));          // next chunk crosses PCIe meanwhile
        Slot &s = slots[c & 1];
        const int64_t j0 = bounds[c], j1 = bounds[c + 1], nc = j1 - j0;
        const int64_t t0 = host_index(jc, jc_type, j0), t1 = host_index(jc, jc_type, j1), nz = t1 - t0;
        SKM_CUDA(cudaStreamWaitEvent(cs, s.h2d_done, 0));
        SkmTimed tt(ctx, SKM_T_UPLOAD);
        SKM_TRY(skm_launch_rebase_colptr(ctx, s.raw_jc.ptr, jc_type, nc + 1, t0, s.colptr.as<int64_t>()));
        if (nz > 0 && ir_type != SKM_I32) SKM_TRY(skm_launch_convert_index(ctx, s.raw_ir.ptr, ir_type, nz, s.rowidx.ptr, 0));
        if (nz > 0 && val_type != SKM_F32) SKM_TRY(skm_launch_convert_value(ctx, s.raw_val.ptr, val_type, nz, s.val.ptr, SKM_F32));
        SKM_TRY(skm_launch_validate_async(ctx, p, nc, nz, s.colptr.as<int64_t>(), s.rowidx.as<int32_t>(), s.flags.as<int>()));

        skm_dataset view;
        memset(&view, 0, sizeof view);
        view.ctx = ctx; view.p = p; view.n = nc; view.nnz = nz; view.store_dtype = SKM_F32;
        view.colptr = s.colptr.as<int64_t>(); view.rowidx = s.rowidx.as<int32_t>(); view.val = s.val.ptr;
        view.sell = s.sell.as<int4>(); view.slice_ptr = s.slice_ptr.as<int64_t>();
        view.max_col_nnz = 1;
        SKM_TRY(skm_build_sell_async(ctx, &view, s.w2.as<int32_t>(), s.elems.as<int64_t>(), s.cub_tmp.ptr, s.cub_tmp.bytes));
        // capacity / validity check of this chunk (one tiny read-back)
        int64_t sell_elems = 0;
        SKM_CUDA(cudaMemcpyAsync(&sell_elems, view.slice_ptr + view.nslices, sizeof(int64_t), cudaMemcpyDeviceToHost, cs));
        SKM_CUDA(cudaMemcpyAsync(ctx->h_flag, s.flags.ptr, 2 * sizeof(int), cudaMemcpyDeviceToHost, cs));
        SKM_CUDA(cudaStreamSynchronize(cs));
        if (ctx->h_flag[0]) {
            skm_set_error("invalid CSC data in columns [%lld,%lld) (row index out of range or bad column pointers)",
                          (long long)j0, (long long)j1);
            return SKM_ERR_INVALID;
        }
        ExactArgs ea;
        ea.p = p; ea.n = nc; ea.K = K; ea.colptr = view.colptr; ea.rowidx = view.rowidx; ea.val = view.val;
        ea.val_type = SKM_F32; ea.ct = ct.as<double>(); ea.mask = nullptr; ea.xdiv = nullptr;
        if (sell_elems <= cap_sell) {
            view.sell_elems = sell_elems;
            SKM_TRY(skm_launch_assign_fast(ctx, &view, K, pl, table.as<float>(), cmax.as<float>(), s.assign.as<int32_t>(),
                                           s.dist.as<float>(), best2.as<float>(), s.flagged.as<int32_t>(), s.nflag.as<int>(),
                                           s.flags.as<int>() + 1));
            SKM_TRY(skm_launch_exact_assign(ctx, ea, s.assign.as<int32_t>(), nullptr, s.dist.as<float>(),
                                            s.flagged.as<int32_t>(), s.nflag.as<int>(), nc));
        } else {    // very ragged chunk: the padded image would not fit its slot
            SKM_CUDA(cudaMemsetAsync(s.nflag.ptr, 0, sizeof(int), cs));
            SKM_TRY(skm_launch_exact_assign(ctx, ea, s.assign.as<int32_t>(), nullptr, s.dist.as<float>(), nullptr, nullptr, 0));
        }
        SKM_TRY(skm_launch_accumulate(ctx, &view, K, s.assign.as<int32_t>(), nullptr, s.dist.as<float>(), nullptr,
                                      partials.as<double>(), false));
        if (assign_out || dist_out) {
            int64_t blocks = std::min<int64_t>((nc + 255) / 256, (int64_t)ctx->sm_count * 8);
            k_publish<<<(unsigned)blocks, 256, 0, cs>>>(nc, s.assign.as<int32_t>(), s.dist.as<float>(),
                                                       assign_out ? s.out_assign.as<int32_t>() : nullptr,
                                                       dist_out ? s.out_dist.as<double>() : nullptr);
            SKM_CHECK_LAUNCH(ctx);
            if (assign_out) SKM_CUDA(cudaMemcpyAsync(assign_out + j0, s.out_assign.ptr, sizeof(int32_t) * nc, cudaMemcpyDeviceToHost, cs));
            if (dist_out) SKM_CUDA(cudaMemcpyAsync(dist_out + j0, s.out_dist.ptr, sizeof(double) * nc, cudaMemcpyDeviceToHost, cs));
        }
        int nf = 0;
        SKM_CUDA(cudaMemcpyAsync(&nf, s.nflag.ptr, sizeof(int), cudaMemcpyDeviceToHost, cs));
        SKM_CUDA(cudaEventRecord(s.free_ev, cs));
        SKM_CUDA(cudaStreamSynchronize(cs));       // the next H2D is already in flight on copy_stream
        rechecked += nf;
    }

    // ---- multi-GPU: the caller sums the partials over ranks on the compute stream ----
    if (reduce) {
        const int rrc = reduce(partials.ptr, npart, (void *)cs, reduce_user);
        if (rrc != 0) { skm_set_error("skm_lloyd_step_host: the reduce callback failed (%d)", rrc); return SKM_ERR_CUDA; }
    }
    // ---- K3 ----
    SKM_TRY(skm_launch_finalize(ctx, p, K, partials.as<double>(), gamma_update, ml_correction, dcen.as<double>(),
                                dold.as<double>(), dstats.as<double>()));
    std::vector<double> tail(K + 1);
    double hs[8];
    SKM_CUDA(cudaMemcpyAsync(hs, dstats.ptr, sizeof hs, cudaMemcpyDeviceToHost, cs));
    SKM_CUDA(cudaMemcpyAsync(tail.data(), partials.as<double>() + 2 * p * K, sizeof(double) * (K + 1), cudaMemcpyDeviceToHost, cs));
    SKM_CUDA(cudaMemcpyAsync(centers_out, dcen.ptr, sizeof(double) * p * K, cudaMemcpyDeviceToHost, cs));
    SKM_CUDA(cudaStreamSynchronize(cs));
    SKM_CUDA(cudaStreamSynchronize(copy_stream));
    if (stats) {
        int64_t n_empty = 0, npts = 0;
        for (int64_t k = 0; k < K; ++k) { const int64_t c = (int64_t)llround(tail[k]); npts += c; n_empty += (c == 0); }
        stats->dff = sqrt(hs[0]);
        stats->sumsq = tail[K];
        stats->n_empty = n_empty;
        stats->n_rechecked = rechecked;
        stats->n_points = npts;
        stats->has_nan = hs[1] != 0.0;
        stats->reserved = 0;
    }
    return SKM_OK;
}
