/*
 * skm_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's arithmetic on the sparsified
 * K-means hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (libskm_b200.so and sparsifiedkmeans_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here
 * bit-for-bit against the reference's own C files compiled unmodified into
 * oracle/_ref (see oracle/Makefile) and against the golden vectors in
 * tests/golden/ that were generated from those binaries.
 *
 * All matrices are column-major (MATLAB convention).  CSC index arrays are
 * int64 (the reference uses mwIndex = 64-bit under -largeArrayDims,
 * setup_kmeans.m:19).  Build: gcc -O2 -ffp-contract=off (no FMA contraction,
 * matching the stock `mex -O` x86-64 build of the reference).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t idx_t;

/* Masked Euclidean distance of every sparse column to every dense centre.
 * out[k + K*j] = sqrt( sum_{t in col j} (x[t] - c[ir[t] + p*k])^2 ), terms added
 * in stored order into a zero-initialised double.
 * Follows /root/reference/private/SparseMatrixMinusCluster.c:169-182 (general K;
 * the K=1,2,3 special cases at :133-168 perform the same operations per k). */
void skmo_masked_dist(idx_t p, idx_t n, idx_t K, const idx_t *jc, const idx_t *ir,
                      const double *x, const double *c, double *out)
{
    for (idx_t j = 0; j < n; ++j) {
        double *o = out + (size_t)j * K;
        for (idx_t k = 0; k < K; ++k) {
            const double *ck = c + (size_t)k * p;
            double s = 0.0;
            for (idx_t t = jc[j]; t < jc[j + 1]; ++t) {
                double d = x[t] - ck[ir[t]];
                s += d * d;
            }
            o[k] = sqrt(s);
        }
    }
}

/* beta variant, single centre only:
 * out[j] = sqrt( sum_t x^2 + (-2*beta)*x*c + c*c ), evaluated left to right.
 * Follows SparseMatrixMinusCluster.c:118-129. */
void skmo_masked_dist_beta(idx_t p, idx_t n, const idx_t *jc, const idx_t *ir,
                           const double *x, const double *c, double beta, double *out)
{
    (void)p;
    double b = beta * -2.;
    for (idx_t j = 0; j < n; ++j) {
        double s = 0.0;
        for (idx_t t = jc[j]; t < jc[j + 1]; ++t) {
            double xv = x[t], cv = c[ir[t]];
            s += xv * xv + b * xv * cv + cv * cv;
        }
        out[j] = sqrt(s);
    }
}

/* Column-wise minimum with MATLAB `min(D,[],1)` semantics: first occurrence
 * wins ties, NaNs are skipped unless the whole column is NaN (then NaN, index 1).
 * Follows /root/reference/private/findClusterAssignments.m:168-171.
 * assign is 1-based, as MATLAB returns it. */
void skmo_colmin(idx_t K, idx_t n, const double *D, double *dmin, idx_t *assign)
{
    for (idx_t j = 0; j < n; ++j) {
        const double *d = D + (size_t)j * K;
        idx_t best = -1;
        double bv = 0.0;
        for (idx_t k = 0; k < K; ++k) {
            double v = d[k];
            if (v != v) continue;
            if (best < 0 || v < bv) { best = k; bv = v; }
        }
        if (best < 0) { dmin[j] = NAN; assign[j] = 1; }
        else { dmin[j] = bv; assign[j] = best + 1; }
    }
}

/* Fused distance + argmin without the K x n temporary; bit-identical to
 * skmo_masked_dist followed by skmo_colmin (sqrt is applied before comparing).
 * threads<=1: serial (what the reference does); threads>1: contiguous column
 * slices on pthreads ("reference kernel x all cores", which the reference itself
 * never does -- it is single-threaded). */
typedef struct {
    idx_t p, K, j0, j1;
    const idx_t *jc, *ir;
    const double *x, *c;
    double *dmin;
    idx_t *assign;
} assign_job_t;

static void assign_range(const assign_job_t *a)
{
    for (idx_t j = a->j0; j < a->j1; ++j) {
        idx_t best = -1;
        double bv = 0.0;
        for (idx_t k = 0; k < a->K; ++k) {
            const double *ck = a->c + (size_t)k * a->p;
            double s = 0.0;
            for (idx_t t = a->jc[j]; t < a->jc[j + 1]; ++t) {
                double d = a->x[t] - ck[a->ir[t]];
                s += d * d;
            }
            double v = sqrt(s);
            if (v != v) continue;
            if (best < 0 || v < bv) { best = k; bv = v; }
        }
        if (best < 0) { a->dmin[j] = NAN; a->assign[j] = 1; }
        else { a->dmin[j] = bv; a->assign[j] = best + 1; }
    }
}

static void *assign_worker(void *arg) { assign_range((const assign_job_t *)arg); return NULL; }

void skmo_assign(idx_t p, idx_t n, idx_t K, const idx_t *jc, const idx_t *ir,
                 const double *x, const double *c, double *dmin, idx_t *assign,
                 int threads)
{
    if (threads < 1) threads = 1;
    if ((idx_t)threads > n) threads = (int)(n > 0 ? n : 1);
    assign_job_t *jobs = (assign_job_t *)malloc(sizeof(assign_job_t) * threads);
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    for (int w = 0; w < threads; ++w) {
        assign_job_t jb = { p, K, n * w / threads, n * (w + 1) / threads, jc, ir, x, c, dmin, assign };
        jobs[w] = jb;
    }
    if (threads == 1) {
        assign_range(&jobs[0]);
    } else {
        for (int w = 0; w < threads; ++w) pthread_create(&tid[w], NULL, assign_worker, &jobs[w]);
        for (int w = 0; w < threads; ++w) pthread_join(tid[w], NULL);
    }
    free(jobs);
    free(tid);
}

/* innerProd[j] = sum_t x[t]*c[ir[t]], normX2[j] = sum_t x[t]^2.
 * Follows /root/reference/private/SparseMatrixInnerProduct.c:87-100. */
void skmo_inner_product(idx_t n, const idx_t *jc, const idx_t *ir, const double *x,
                        const double *c, double *ip, double *nrm2)
{
    for (idx_t j = 0; j < n; ++j) {
        double a = 0.0, b = 0.0;
        for (idx_t t = jc[j]; t < jc[j + 1]; ++t) {
            a += x[t] * c[ir[t]];
            b += x[t] * x[t];
        }
        ip[j] = a;
        nrm2[j] = b;
    }
}

/* normX2[j] = sum_t x[t]^2.
 * Follows /root/reference/private/SparseMatrixColumnNormSq.c:71-77. */
void skmo_colnormsq(idx_t n, const idx_t *jc, const double *x, double *nrm2)
{
    for (idx_t j = 0; j < n; ++j) {
        double b = 0.0;
        for (idx_t t = jc[j]; t < jc[j + 1]; ++t) b += x[t] * x[t];
        nrm2[j] = b;
    }
}

/* Unnormalised natural-order (Sylvester) Walsh-Hadamard transform of each
 * column: radix-2 butterflies, strides 1,2,4,...  (a,b) -> (a+b, a-b).
 * Follows /root/reference/private/hadamard.c:57-92 (and the identical butterfly
 * of hadamard_pthreads.c:69-90).  Only additions/subtractions occur, so the
 * rounding of every output is fixed by the butterfly network, not the loop order. */
void skmo_hadamard(idx_t m, idx_t n, const double *x, double *y)
{
    for (idx_t j = 0; j < n; ++j) {
        const double *xi = x + (size_t)j * m;
        double *yo = y + (size_t)j * m;
        memcpy(yo, xi, (size_t)m * sizeof(double));
        for (idx_t h = 1; h < m; h <<= 1) {
            for (idx_t base = 0; base < m; base += 2 * h) {
                for (idx_t i = base; i < base + h; ++i) {
                    double a = yo[i], b = yo[i + h];
                    yo[i] = a + b;
                    yo[i + h] = a - b;
                }
            }
        }
    }
}

/* Per-cluster row sums and support counts, then the ML-corrected centre
 *   C(:,k) = gamma * S(:,k) ./ (N(:,k) + 1e-16)
 * with S(:,k) = sum over members j of X(:,j) and N(:,k) = sum of spones(X)(:,j),
 * members visited in ascending column order.
 * Follows /root/reference/kmeans_sparsified.m:430-453 (formula at :448).
 * assign is 1-based; entries outside 1..K are ignored.  counts[k] receives the
 * number of members; columns of `centers` for empty clusters are left untouched
 * (the reference's EmptyAction branch handles them, :432-445).
 * If ml_correction == 0: plain mean over members (kmeans_sparsified.m:450). */
void skmo_centroid_update(idx_t p, idx_t n, idx_t K, const idx_t *jc, const idx_t *ir,
                          const double *x, const idx_t *assign, double gamma,
                          int ml_correction, double *S, double *N, idx_t *counts,
                          double *centers)
{
    memset(S, 0, (size_t)p * K * sizeof(double));
    memset(N, 0, (size_t)p * K * sizeof(double));
    memset(counts, 0, (size_t)K * sizeof(idx_t));
    for (idx_t j = 0; j < n; ++j) {
        idx_t k = assign[j] - 1;
        if (k < 0 || k >= K) continue;
        counts[k] += 1;
        double *Sk = S + (size_t)k * p, *Nk = N + (size_t)k * p;
        for (idx_t t = jc[j]; t < jc[j + 1]; ++t) {
            Sk[ir[t]] += x[t];
            Nk[ir[t]] += 1.0;
        }
    }
    for (idx_t k = 0; k < K; ++k) {
        if (counts[k] == 0) continue;
        double *ck = centers + (size_t)k * p;
        const double *Sk = S + (size_t)k * p, *Nk = N + (size_t)k * p;
        if (ml_correction) {
            for (idx_t i = 0; i < p; ++i) ck[i] = gamma * Sk[i] / (Nk[i] + 1e-16);
        } else {
            double cnt = (double)counts[k];
            for (idx_t i = 0; i < p; ++i) ck[i] = Sk[i] / cnt;
        }
    }
}
