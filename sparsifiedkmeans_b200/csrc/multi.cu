// multi.cu -- several GPUs of one box driven by ONE host process (a MATLAB session cannot be torchrun).
//
// The reference is a single MATLAB process; its MEX gateway is called on MATLAB's main thread.  To let that
// one caller use all GPUs of the box, skm_multi owns one skm_ctx (and one host worker thread) per device,
// shards the columns in contiguous blocks (GPU g owns [g*n/G, (g+1)*n/G)), runs
// K1/K2 of every shard concurrently, and replaces the per-iteration all-reduce + K3 by ONE kernel per device
// that reads every peer's partials [S | N | counts | sumsq] straight through NVLink peer memory, sums them
// in device order (so every device gets bit-identical sums, whatever the timing) and finalises the centres
// (kmeans_sparsified.m:448,470-471) in the same pass: the collective is fused into its consumer.  The
// ordering between devices is carried by CUDA events recorded on the owning stream and waited on by the
// peers' streams; no host thread blocks inside an iteration except to read the statistics back.
#include "common.cuh"
#include <math.h>
#include <string.h>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

int skm_launch_finalize_peers(skm_ctx *ctx, int64_t p, int64_t K, int ndev, const double *const *parts_dev,
                              double gamma, int ml_correction, double *centers, double *centers_old,
                              double *stats, double *tail);

namespace {

struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> task;
    bool has_task = false, done = false, quit = false;
    int rc = SKM_OK;
    std::string err;
};

void worker_main(Worker *w, int device)
{
    cudaSetDevice(device);
    for (;;) {
        std::function<int()> t;
        {
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv.wait(lk, [&] { return w->has_task || w->quit; });
            if (w->quit) return;
            t = w->task;
        }
        const int rc = t();
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->rc = rc;
            w->err = rc == SKM_OK ? "" : skm_last_error(nullptr);     // the message is thread-local: carry it over
            w->has_task = false;
            w->done = true;
        }
        w->cv.notify_all();
    }
}

}  // namespace

struct skm_multi {
    int ndev = 0;
    bool peer = false;                      // every pair has direct peer access (else: staged copies)
    std::vector<skm_ctx *> ctx;
    std::vector<Worker *> w;
};

struct skm_multi_dataset {
    skm_multi *m = nullptr;
    int64_t p = 0, n = 0;
    std::vector<skm_dataset *> shard;
    std::vector<int64_t> col0;              // [ndev + 1]
    std::vector<double> kpp_sums;           // local sums of D^2 of the last skm_multi_kpp_update
};

struct skm_multi_lloyd {
    skm_multi_dataset *md = nullptr;
    int64_t K = 0;
    std::vector<skm_lloyd *> L;
    std::vector<cudaEvent_t> ev_acc, ev_fin;
    std::vector<double **> d_parts;         // per device: device array of ndev pointers to the partials it reads
    std::vector<double *> tail;             // per device: reduced [counts (K) | sumsq]
    std::vector<double *> stage;            // per device: staging copies of the peers' partials (no peer access)
    bool fin_recorded = false;
    std::vector<int64_t> counts;
    std::vector<int64_t> rechecked;
};

// run fn(g) on every device's worker thread; returns the first failure (its message becomes this thread's)
static int run_all(skm_multi *m, const std::function<int(int)> &fn)
{
    for (int g = 0; g < m->ndev; ++g) {
        Worker *w = m->w[g];
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->task = [fn, g] { return fn(g); };
            w->done = false;
            w->has_task = true;
        }
        w->cv.notify_all();
    }
    int rc = SKM_OK;
    for (int g = 0; g < m->ndev; ++g) {
        Worker *w = m->w[g];
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return w->done; });
        if (w->rc != SKM_OK && rc == SKM_OK) { rc = w->rc; skm_set_error("device %d: %s", m->ctx[g]->device, w->err.c_str()); }
    }
    return rc;
}

extern "C" void skm_multi_destroy(skm_multi *m)
{
    if (!m) return;
    for (Worker *w : m->w) {
        if (!w) continue;
        { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; }
        w->cv.notify_all();
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    for (skm_ctx *c : m->ctx) skm_ctx_destroy(c);
    delete m;
}

extern "C" int skm_multi_create(int ndev, const int *devices, skm_multi **out)
{
    SKM_REQUIRE(out, "out is NULL");
    *out = nullptr;
    int have = 0;
    cudaError_t e = cudaGetDeviceCount(&have);
    if (e != cudaSuccess || have == 0) {
        skm_set_error("no CUDA device available (%s); libskm_b200 has no CPU fallback",
                      e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return SKM_ERR_CUDA;
    }
    if (ndev <= 0) ndev = have;                                   // 0 = every visible device
    SKM_REQUIRE(ndev <= have, "asked for %d devices, %d visible", ndev, have);
    skm_multi *m = new (std::nothrow) skm_multi();
    if (!m) { skm_set_error("out of host memory"); return SKM_ERR_NOMEM; }
    m->ndev = ndev;
    for (int g = 0; g < ndev; ++g) {
        const int dev = devices ? devices[g] : g;
        for (int h = 0; h < g; ++h)
            if (m->ctx[h]->device == dev) { skm_set_error("device %d listed twice", dev); skm_multi_destroy(m); return SKM_ERR_INVALID; }
        skm_ctx *c = nullptr;
        const int rc = skm_ctx_create(dev, nullptr, &c);
        if (rc != SKM_OK) { skm_multi_destroy(m); return rc; }
        m->ctx.push_back(c);
    }
    // peer access between every pair (NVLink / NVSwitch on a B200 box)
    m->peer = true;
    for (int g = 0; g < ndev && m->peer; ++g)
        for (int h = 0; h < ndev; ++h) {
            if (g == h) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, m->ctx[g]->device, m->ctx[h]->device);
            if (!can) { m->peer = false; break; }
        }
    if (m->peer)
        for (int g = 0; g < ndev; ++g) {
            cudaSetDevice(m->ctx[g]->device);
            for (int h = 0; h < ndev; ++h) {
                if (g == h) continue;
                e = cudaDeviceEnablePeerAccess(m->ctx[h]->device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess) { cudaGetLastError(); m->peer = false; }
            }
        }
    for (int g = 0; g < ndev; ++g) {
        Worker *w = new (std::nothrow) Worker();
        if (!w) { skm_set_error("out of host memory"); skm_multi_destroy(m); return SKM_ERR_NOMEM; }
        m->w.push_back(w);
        w->th = std::thread(worker_main, w, m->ctx[g]->device);
    }
    *out = m;
    return SKM_OK;
}

extern "C" int skm_multi_ndev(const skm_multi *m) { return m ? m->ndev : 0; }
extern "C" skm_ctx *skm_multi_ctx(skm_multi *m, int i) { return (m && i >= 0 && i < m->ndev) ? m->ctx[i] : nullptr; }
extern "C" int skm_multi_peer_access(const skm_multi *m) { return (m && m->peer) ? 1 : 0; }

// ---------------------------------------------------------------------------------------------- datasets
extern "C" void skm_multi_dataset_destroy(skm_multi_dataset *md)
{
    if (!md) return;
    for (skm_dataset *d : md->shard) skm_dataset_destroy(d);
    delete md;
}

extern "C" int skm_multi_dataset_from_shards(skm_multi *m, skm_dataset *const *shards, skm_multi_dataset **out)
{
    SKM_REQUIRE(m && shards && out, "NULL argument");
    *out = nullptr;
    skm_multi_dataset *md = new (std::nothrow) skm_multi_dataset();
    if (!md) { skm_set_error("out of host memory"); return SKM_ERR_NOMEM; }
    md->m = m;
    md->col0.assign(m->ndev + 1, 0);
    for (int g = 0; g < m->ndev; ++g) {
        skm_dataset *d = shards[g];
        if (!d || d->ctx != m->ctx[g]) { delete md; skm_set_error("shard %d is not a dataset of device %d's context", g, g); return SKM_ERR_INVALID; }
        if (g == 0) md->p = d->p;
        if (d->p != md->p) { delete md; skm_set_error("shard %d has %lld rows, shard 0 has %lld", g, (long long)d->p, (long long)md->p); return SKM_ERR_INVALID; }
        md->col0[g + 1] = md->col0[g] + d->n;
    }
    md->n = md->col0[m->ndev];
    md->shard.assign(shards, shards + m->ndev);                   // ownership moves to the multi dataset
    *out = md;
    return SKM_OK;
}

static size_t tsz(int t) { return t == SKM_U16 ? 2 : ((t == SKM_F32 || t == SKM_I32) ? 4 : 8); }
static int64_t hidx(const void *a, int type, int64_t i) { return type == SKM_I32 ? (int64_t)((const int32_t *)a)[i] : ((const int64_t *)a)[i]; }

extern "C" int skm_multi_dataset_create_csc(skm_multi *m, int64_t p, int64_t n, const void *jc, int jc_type,
                                            const void *ir, int ir_type, const void *val, int val_type,
                                            int store_dtype, skm_multi_dataset **out)
{
    SKM_REQUIRE(m && jc && out, "NULL argument");
    *out = nullptr;
    SKM_REQUIRE(p >= 0 && n >= 0, "negative dimensions");
    SKM_REQUIRE(jc_type == SKM_I32 || jc_type == SKM_I64, "jc_type must be SKM_I32 or SKM_I64");
    SKM_REQUIRE(ir_type == SKM_I32 || ir_type == SKM_I64 || ir_type == SKM_U16, "ir_type must be SKM_I32, SKM_I64 or SKM_U16");
    SKM_REQUIRE(val_type == SKM_F32 || val_type == SKM_F64, "val_type must be SKM_F32 or SKM_F64");
    const int G = m->ndev;
    std::vector<skm_dataset *> shards(G, nullptr);
    // every worker rebases its slice of the column pointers and uploads its block concurrently
    const int rc = run_all(m, [&](int g) -> int {
        const int64_t lo = n * g / G, hi = n * (g + 1) / G, nc = hi - lo;
        const int64_t t0 = hidx(jc, jc_type, lo);
        std::vector<int64_t> cp((size_t)nc + 1);
        for (int64_t j = 0; j <= nc; ++j) cp[j] = hidx(jc, jc_type, lo + j) - t0;
        return skm_dataset_create_csc(m->ctx[g], p, nc, cp.data(), SKM_I64,
                                      ir ? (const char *)ir + (size_t)t0 * tsz(ir_type) : nullptr, ir_type,
                                      val ? (const char *)val + (size_t)t0 * tsz(val_type) : nullptr, val_type,
                                      store_dtype, 0, &shards[g]);
    });
    if (rc != SKM_OK) { for (skm_dataset *d : shards) skm_dataset_destroy(d); return rc; }
    const int rc2 = skm_multi_dataset_from_shards(m, shards.data(), out);
    if (rc2 != SKM_OK) for (skm_dataset *d : shards) skm_dataset_destroy(d);
    return rc2;
}

extern "C" skm_dataset *skm_multi_dataset_shard(skm_multi_dataset *md, int i, int64_t *col0)
{
    if (!md || i < 0 || i >= md->m->ndev) return nullptr;
    if (col0) *col0 = md->col0[i];
    return md->shard[i];
}

extern "C" int skm_multi_dataset_get_info(const skm_multi_dataset *md, skm_dataset_info *info)
{
    SKM_REQUIRE(md && info, "NULL argument");
    memset(info, 0, sizeof *info);
    info->p = md->p; info->n = md->n;
    for (skm_dataset *d : md->shard) {
        skm_dataset_info s;
        SKM_TRY(skm_dataset_get_info(d, &s));
        info->nnz += s.nnz; info->device_bytes += s.device_bytes; info->stream_bytes += s.stream_bytes;
        if (s.max_col_nnz > info->max_col_nnz) info->max_col_nnz = s.max_col_nnz;
        info->store_dtype = s.store_dtype;
    }
    return SKM_OK;
}

static int owner_of(const skm_multi_dataset *md, int64_t j)
{
    for (int g = 0; g < md->m->ndev; ++g) if (j >= md->col0[g] && j < md->col0[g + 1]) return g;
    return -1;
}

extern "C" int skm_multi_dataset_get_column(skm_multi_dataset *md, int64_t j, double *out)
{
    SKM_REQUIRE(md && out, "NULL argument");
    const int g = owner_of(md, j);
    SKM_REQUIRE(g >= 0, "column %lld out of range", (long long)j);
    return skm_dataset_get_column(md->shard[g], j - md->col0[g], out);
}

// k-means++ over the shards (private/Arthur_initialization.m:39-53): local running minima, sums combined in device order
extern "C" int skm_multi_kpp_update(skm_multi_dataset *md, const double *center, int has_gamma, double gamma, int first,
                                    int sparse_center, double *sum_d2)
{
    SKM_REQUIRE(md && center, "NULL argument");
    md->kpp_sums.assign(md->m->ndev, 0.0);
    SKM_TRY(run_all(md->m, [&](int g) -> int {
        if (md->shard[g]->n == 0) return SKM_OK;
        return sparse_center ? skm_kpp_update_sparse(md->shard[g], center, first, &md->kpp_sums[g])
                             : skm_kpp_update(md->shard[g], center, has_gamma, gamma, first, &md->kpp_sums[g]);
    }));
    double tot = 0.0;
    for (double v : md->kpp_sums) tot += v;
    if (sum_d2) *sum_d2 = tot;
    return SKM_OK;
}

extern "C" int skm_multi_kpp_pick(skm_multi_dataset *md, double target, int64_t *j)
{
    SKM_REQUIRE(md && j, "NULL argument");
    if ((int)md->kpp_sums.size() != md->m->ndev) { skm_set_error("skm_multi_kpp_pick before skm_multi_kpp_update"); return SKM_ERR_STATE; }
    double prefix = 0.0;
    int q = md->m->ndev - 1;
    for (int g = 0; g < md->m->ndev; ++g) {
        if (target < prefix + md->kpp_sums[g] && md->kpp_sums[g] > 0) { q = g; break; }
        if (g + 1 < md->m->ndev) prefix += md->kpp_sums[g];
    }
    if (q == md->m->ndev - 1) { prefix = 0.0; for (int g = 0; g < q; ++g) prefix += md->kpp_sums[g]; }
    while (q > 0 && md->shard[q]->n == 0) --q;
    int64_t jl = 0;
    SKM_TRY(skm_kpp_pick(md->shard[q], target - prefix, &jl));
    *j = md->col0[q] + jl;
    return SKM_OK;
}

// ---------------------------------------------------------------------------------------------- Lloyd
extern "C" void skm_multi_lloyd_destroy(skm_multi_lloyd *ML)
{
    if (!ML) return;
    skm_multi *m = ML->md->m;
    for (int g = 0; g < m->ndev; ++g) {
        cudaSetDevice(m->ctx[g]->device);
        cudaStreamSynchronize(m->ctx[g]->stream);
    }
    for (int g = 0; g < m->ndev; ++g) {
        cudaSetDevice(m->ctx[g]->device);
        if (g < (int)ML->ev_acc.size() && ML->ev_acc[g]) cudaEventDestroy(ML->ev_acc[g]);
        if (g < (int)ML->ev_fin.size() && ML->ev_fin[g]) cudaEventDestroy(ML->ev_fin[g]);
        if (g < (int)ML->d_parts.size()) cudaFree(ML->d_parts[g]);
        if (g < (int)ML->tail.size()) cudaFree(ML->tail[g]);
        if (g < (int)ML->stage.size()) cudaFree(ML->stage[g]);
        if (g < (int)ML->L.size()) skm_lloyd_destroy(ML->L[g]);
    }
    delete ML;
}

extern "C" int skm_multi_lloyd_create(skm_multi_dataset *md, int64_t K, skm_multi_lloyd **out)
{
    SKM_REQUIRE(md && out, "NULL argument");
    *out = nullptr;
    skm_multi *m = md->m;
    const int G = m->ndev;
    skm_multi_lloyd *ML = new (std::nothrow) skm_multi_lloyd();
    if (!ML) { skm_set_error("out of host memory"); return SKM_ERR_NOMEM; }
    ML->md = md; ML->K = K;
    ML->L.assign(G, nullptr); ML->ev_acc.assign(G, nullptr); ML->ev_fin.assign(G, nullptr);
    ML->d_parts.assign(G, nullptr); ML->tail.assign(G, nullptr); ML->stage.assign(G, nullptr);
    ML->counts.assign(K, 0); ML->rechecked.assign(G, 0);
    const int64_t npart = 2 * md->p * K + K + 1;
    int rc = run_all(m, [&](int g) -> int {
        SKM_TRY(skm_lloyd_create(md->shard[g], K, &ML->L[g]));
        SKM_CUDA(cudaEventCreateWithFlags(&ML->ev_acc[g], cudaEventDisableTiming));
        SKM_CUDA(cudaEventCreateWithFlags(&ML->ev_fin[g], cudaEventDisableTiming));
        SKM_CUDA(cudaMalloc((void **)&ML->d_parts[g], sizeof(double *) * G));
        SKM_CUDA(cudaMalloc((void **)&ML->tail[g], sizeof(double) * (K + 1)));
        if (!m->peer && G > 1) SKM_CUDA(cudaMalloc((void **)&ML->stage[g], sizeof(double) * npart * G));
        return SKM_OK;
    });
    if (rc == SKM_OK)
        rc = run_all(m, [&](int g) -> int {
            std::vector<double *> ptrs(G);
            for (int h = 0; h < G; ++h)
                ptrs[h] = (m->peer || h == g) ? ML->L[h]->partials : ML->stage[g] + (size_t)h * npart;
            SKM_CUDA(cudaMemcpy(ML->d_parts[g], ptrs.data(), sizeof(double *) * G, cudaMemcpyHostToDevice));
            return SKM_OK;
        });
    if (rc != SKM_OK) { skm_multi_lloyd_destroy(ML); return rc; }
    *out = ML;
    return SKM_OK;
}

extern "C" int skm_multi_lloyd_set_modes(skm_multi_lloyd *ML, int update_mode, int assign_mode)
{
    SKM_REQUIRE(ML, "NULL argument");
    return run_all(ML->md->m, [&](int g) -> int {
        SKM_TRY(skm_lloyd_set_update_mode(ML->L[g], update_mode));
        if (ML->md->shard[g]->store_dtype == SKM_F32 || assign_mode == 0) SKM_TRY(skm_lloyd_set_assign_mode(ML->L[g], assign_mode));
        return SKM_OK;
    });
}

extern "C" int skm_multi_lloyd_set_centers(skm_multi_lloyd *ML, const double *centers)
{
    SKM_REQUIRE(ML && centers, "NULL argument");
    return run_all(ML->md->m, [&](int g) -> int { return skm_lloyd_set_centers(ML->L[g], centers); });
}

extern "C" int skm_multi_lloyd_set_center_column(skm_multi_lloyd *ML, int64_t k, const double *col)
{
    SKM_REQUIRE(ML && col, "NULL argument");
    return run_all(ML->md->m, [&](int g) -> int { return skm_lloyd_set_center_column(ML->L[g], k, col); });
}

extern "C" int skm_multi_lloyd_get_centers(skm_multi_lloyd *ML, double *centers)
{
    SKM_REQUIRE(ML && centers, "NULL argument");
    return skm_lloyd_get_centers(ML->L[0], centers);          // identical on every device by construction
}

extern "C" int skm_multi_lloyd_get_centers_old(skm_multi_lloyd *ML, double *centers)
{
    SKM_REQUIRE(ML && centers, "NULL argument");
    return skm_lloyd_get_centers_old(ML->L[0], centers);
}

// centres of device g (tests: every device must hold the same bits)
extern "C" int skm_multi_lloyd_get_centers_of(skm_multi_lloyd *ML, int g, double *centers)
{
    SKM_REQUIRE(ML && centers && g >= 0 && g < ML->md->m->ndev, "bad argument");
    return skm_lloyd_get_centers(ML->L[g], centers);
}

static int multi_read_stats(skm_multi_lloyd *ML, int g, std::vector<double> &tail, double *hs)
{
    skm_lloyd *L = ML->L[g];
    skm_ctx *ctx = L->ctx;
    int nf = 0;
    SKM_CUDA(cudaMemcpyAsync(hs, L->stats, sizeof(double) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SKM_CUDA(cudaMemcpyAsync(&nf, L->nflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SKM_CUDA(cudaMemcpyAsync(tail.data(), ML->tail[g], sizeof(double) * (ML->K + 1), cudaMemcpyDeviceToHost, ctx->stream));
    SKM_CUDA(cudaStreamSynchronize(ctx->stream));
    ML->rechecked[g] = nf;
    return SKM_OK;
}

static int multi_finish_stats(skm_multi_lloyd *ML, const std::vector<std::vector<double>> &tails,
                              const std::vector<double> &hs, skm_iter_stats *stats)
{
    const int64_t K = ML->K;
    int64_t n_empty = 0, npts = 0, rech = 0;
    for (int64_t k = 0; k < K; ++k) {
        ML->counts[k] = (int64_t)llround(tails[0][k]);
        npts += ML->counts[k];
        n_empty += ML->counts[k] == 0;
    }
    for (int g = 0; g < ML->md->m->ndev; ++g) {
        rech += ML->rechecked[g];
        ML->L[g]->last_rechecked = ML->rechecked[g];
        for (int64_t k = 0; k < K; ++k) ML->L[g]->h_counts[k] = ML->counts[k];
    }
    if (stats) {
        stats->dff = sqrt(hs[0]);
        stats->sumsq = tails[0][K];
        stats->n_empty = n_empty;
        stats->n_rechecked = rech;
        stats->n_points = npts;
        stats->has_nan = hs[1] != 0.0;
        stats->reserved = 0;
    }
    return SKM_OK;
}

extern "C" int skm_multi_lloyd_step(skm_multi_lloyd *ML, int has_gamma, double gamma_dist, double gamma_update,
                                    int ml_correction, int sparse_centers, skm_iter_stats *stats)
{
    SKM_REQUIRE(ML, "NULL argument");
    skm_multi *m = ML->md->m;
    const int G = m->ndev;
    const int64_t p = ML->md->p, K = ML->K, npart = 2 * p * K + K + 1;
    // phase 1: K1 + K2 of every shard; a shard's partials may only be rewritten once every peer has read the
    // previous iteration's (ev_fin), and are ready for the peers when ev_acc fires
    SKM_TRY(run_all(m, [&](int g) -> int {
        skm_lloyd *L = ML->L[g];
        cudaStream_t s = L->ctx->stream;
        if (ML->fin_recorded)
            for (int h = 0; h < G; ++h) if (h != g) SKM_CUDA(cudaStreamWaitEvent(s, ML->ev_fin[h], 0));
        if (sparse_centers) SKM_TRY(skm_lloyd_assign_sparse(L, has_gamma, gamma_dist));
        else SKM_TRY(skm_lloyd_assign(L, has_gamma, gamma_dist));
        SKM_TRY(skm_lloyd_accumulate(L));
        SKM_CUDA(cudaEventRecord(ML->ev_acc[g], s));
        return SKM_OK;
    }));
    // phase 2: fused peer-memory reduction + K3 on every device, statistics back
    std::vector<std::vector<double>> tails(G, std::vector<double>(K + 1));
    std::vector<double> hs(8 * G);
    SKM_TRY(run_all(m, [&](int g) -> int {
        skm_lloyd *L = ML->L[g];
        skm_ctx *ctx = L->ctx;
        cudaStream_t s = ctx->stream;
        for (int h = 0; h < G; ++h) if (h != g) SKM_CUDA(cudaStreamWaitEvent(s, ML->ev_acc[h], 0));
        if (!m->peer)
            for (int h = 0; h < G; ++h) if (h != g)
                SKM_CUDA(cudaMemcpyPeerAsync(ML->stage[g] + (size_t)h * npart, ctx->device, ML->L[h]->partials,
                                             m->ctx[h]->device, sizeof(double) * npart, s));
        {
            SkmTimed t(ctx, SKM_T_FINAL);
            SKM_TRY(skm_launch_finalize_peers(ctx, p, K, G, ML->d_parts[g], gamma_update, ml_correction, L->centers,
                                              L->centers_old, L->stats, ML->tail[g]));
        }
        SKM_CUDA(cudaEventRecord(ML->ev_fin[g], s));
        return multi_read_stats(ML, g, tails[g], &hs[8 * g]);
    }));
    ML->fin_recorded = true;
    return multi_finish_stats(ML, tails, hs, stats);
}

extern "C" int skm_multi_lloyd_refresh_diff(skm_multi_lloyd *ML, skm_iter_stats *stats)
{
    SKM_REQUIRE(ML, "NULL argument");
    std::vector<skm_iter_stats> st(ML->md->m->ndev);
    SKM_TRY(run_all(ML->md->m, [&](int g) -> int { return skm_lloyd_refresh_diff(ML->L[g], &st[g]); }));
    if (stats) { stats->dff = st[0].dff; stats->has_nan = st[0].has_nan; }
    return SKM_OK;
}

extern "C" int skm_multi_lloyd_get_counts(skm_multi_lloyd *ML, int64_t *counts)
{
    SKM_REQUIRE(ML && counts, "NULL argument");
    for (int64_t k = 0; k < ML->K; ++k) counts[k] = ML->counts[k];
    return SKM_OK;
}

extern "C" int skm_multi_lloyd_get_assignments(skm_multi_lloyd *ML, int32_t *assign_out, double *dist_out)
{
    SKM_REQUIRE(ML, "NULL argument");
    skm_multi_dataset *md = ML->md;
    return run_all(md->m, [&](int g) -> int {
        if (md->shard[g]->n == 0) return SKM_OK;
        return skm_lloyd_get_assignments(ML->L[g], assign_out ? assign_out + md->col0[g] : nullptr,
                                         dist_out ? dist_out + md->col0[g] : nullptr);
    });
}

// first column (global index) attaining the largest distance (kmeans_sparsified.m:435): shards are ascending
// column blocks, so the lowest device wins ties
extern "C" int skm_multi_lloyd_argmax_distance(skm_multi_lloyd *ML, double *maxdist, int64_t *j)
{
    SKM_REQUIRE(ML && maxdist && j, "NULL argument");
    skm_multi_dataset *md = ML->md;
    const int G = md->m->ndev;
    std::vector<double> v(G, -1.0);
    std::vector<int64_t> jj(G, -1);
    SKM_TRY(run_all(md->m, [&](int g) -> int {
        if (md->shard[g]->n == 0) return SKM_OK;
        return skm_lloyd_argmax_distance(ML->L[g], &v[g], &jj[g]);
    }));
    int best = -1;
    for (int g = 0; g < G; ++g) {
        if (jj[g] < 0 || v[g] != v[g]) continue;
        if (best < 0 || v[g] > v[best]) best = g;
    }
    if (best < 0) { *maxdist = nan(""); *j = 0; return SKM_OK; }
    *maxdist = v[best];
    *j = md->col0[best] + jj[best];
    return SKM_OK;
}

extern "C" int skm_multi_lloyd_launch_count(skm_multi_lloyd *ML, int64_t *launches)
{
    SKM_REQUIRE(ML && launches, "NULL argument");
    int64_t t = 0;
    for (skm_ctx *c : ML->md->m->ctx) t += c->launches;
    *launches = t;
    return SKM_OK;
}
