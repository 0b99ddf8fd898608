/*
 * CPU oracle — TEST / BASELINE INFRASTRUCTURE ONLY (never linked into libtermgpu.so).
 *
 * Plain-C restatement of the aggregates the reference's DataFusion plans compute for the numeric
 * business-rules suite, used (a) by tests as a second, independent checker at sizes where the numpy
 * oracle is slow and (b) by bench.py's cpu_baseline / --impl reference legs as "CPU restatement of the
 * DataFusion path, N cores" (BASELINE.md §2.2). It follows the reference's schedule: ONE full scan per
 * constraint (ValidationSuite::run_sequential, term-guard/src/core/suite.rs:67-100), each scan
 * parallelised over all host cores the way DataFusion splits a query over target_partitions
 * (core/context.rs:32-34). SQL being restated, by function:
 *   to_count_valid   COUNT(c)                         constraints/completeness.rs:158-163
 *   to_min_max_f64   MIN(c) / MAX(c)                  constraints/statistics.rs:263
 *   to_sum_f64       SUM(c), AVG = SUM/COUNT          constraints/statistics.rs:263
 *   to_var_f64       VARIANCE/STDDEV (sample, Welford per partition + Chan merge, as DataFusion's
 *                    VarianceAccumulator does)        constraints/statistics.rs:263, tests/property_tests.rs:784
 *   to_corr_f64      CORR(x,y) over pairwise-complete rows, ONE pass (online co-moments per partition + Chan merge,
 *                    the shape of DataFusion's CorrelationAccumulator)   constraints/correlation.rs:313-316
 *   to_fused_c2      the whole C2 literal suite in one scan (the fused variant BASELINE.md §2.2 reports beside the
 *                    per-constraint schedule)
 *   to_pred_gt_lt    COUNT(CASE WHEN f > a AND i < b THEN 1 END)   constraints/custom_sql.rs:203-209
 * Parity of this file is pinned by tests/test_oracle_c.py against the numpy oracle, which is pinned
 * against the reference's golden vectors.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline int bit(const uint8_t* v, int64_t i) { return v ? (v[i >> 3] >> (i & 7)) & 1 : 1; }

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the baseline legs ask for the host's cores explicitly */
void to_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int to_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int64_t to_count_valid(const uint8_t* validity, int64_t n) {
    if (!validity) return n;
    int64_t c = 0;
#pragma omp parallel for reduction(+ : c) schedule(static)
    for (int64_t w = 0; w < n / 64; ++w) {
        uint64_t x;
        memcpy(&x, validity + w * 8, 8);
        c += __builtin_popcountll(x);
    }
    for (int64_t i = (n / 64) * 64; i < n; ++i) c += bit(validity, i);
    return c;
}

void to_min_max_f64(const double* v, const uint8_t* validity, int64_t n, double* mn, double* mx, int64_t* cnt) {
    double lo = INFINITY, hi = -INFINITY;
    int64_t c = 0;
#pragma omp parallel for reduction(min : lo) reduction(max : hi) reduction(+ : c) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        if (bit(validity, i)) {
            lo = v[i] < lo ? v[i] : lo;
            hi = v[i] > hi ? v[i] : hi;
            ++c;
        }
    }
    *mn = lo;
    *mx = hi;
    *cnt = c;
}

void to_min_max_sum_i64(const int64_t* v, const uint8_t* validity, int64_t n, int64_t* mn, int64_t* mx,
                        int64_t* wrapping_sum, double* fsum, int64_t* cnt) {
    int64_t lo = INT64_MAX, hi = INT64_MIN, c = 0;
    uint64_t s = 0;
    double fs = 0;
#pragma omp parallel for reduction(min : lo) reduction(max : hi) reduction(+ : c, s, fs) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        if (bit(validity, i)) {
            lo = v[i] < lo ? v[i] : lo;
            hi = v[i] > hi ? v[i] : hi;
            s += (uint64_t)v[i];
            fs += (double)v[i];
            ++c;
        }
    }
    *mn = lo;
    *mx = hi;
    *wrapping_sum = (int64_t)s;
    *fsum = fs;
    *cnt = c;
}

void to_sum_f64(const double* v, const uint8_t* validity, int64_t n, double* sum, int64_t* cnt) {
    double s = 0;
    int64_t c = 0;
#pragma omp parallel for reduction(+ : s, c) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        if (bit(validity, i)) {
            s += v[i];
            ++c;
        }
    }
    *sum = s;
    *cnt = c;
}

/* sample variance: Welford within a partition, Chan merge across partitions */
void to_var_f64(const double* v, const uint8_t* validity, int64_t n, double* var_samp, int64_t* cnt) {
    int nt = to_num_threads();
    double* mean = calloc(nt, sizeof(double));
    double* m2 = calloc(nt, sizeof(double));
    int64_t* cn = calloc(nt, sizeof(int64_t));
#pragma omp parallel num_threads(nt)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
        double mu = 0, q = 0;
        int64_t c = 0;
        for (int64_t i = lo; i < hi; ++i) {
            if (bit(validity, i)) {
                ++c;
                double d = v[i] - mu;
                mu += d / (double)c;
                q += d * (v[i] - mu);
            }
        }
        mean[t] = mu;
        m2[t] = q;
        cn[t] = c;
    }
    double mu = 0, q = 0;
    int64_t c = 0;
    for (int t = 0; t < nt; ++t) {
        if (!cn[t]) continue;
        int64_t nc = c + cn[t];
        double d = mean[t] - mu;
        q += m2[t] + d * d * (double)c * (double)cn[t] / (double)nc;
        mu += d * (double)cn[t] / (double)nc;
        c = nc;
    }
    *cnt = c;
    *var_samp = c > 1 ? q / (double)(c - 1) : NAN;
    free(mean);
    free(m2);
    free(cn);
}

/* online co-moment state of one partition: what DataFusion's CorrelationAccumulator keeps (a covariance accumulator
 * and two variance accumulators, each updated row by row) */
typedef struct {
    int64_t n;
    double mx, my, cxx, cyy, cxy;
} comoment;

static inline void comoment_add(comoment* s, double x, double y) {
    s->n += 1;
    double dx = x - s->mx, dy = y - s->my;
    s->mx += dx / (double)s->n;
    s->my += dy / (double)s->n;
    s->cxx += dx * (x - s->mx);
    s->cyy += dy * (y - s->my);
    s->cxy += dx * (y - s->my);
}

static inline void comoment_merge(comoment* a, const comoment* b) {
    if (!b->n) return;
    if (!a->n) {
        *a = *b;
        return;
    }
    int64_t n = a->n + b->n;
    double dx = b->mx - a->mx, dy = b->my - a->my, w = (double)a->n * (double)b->n / (double)n;
    a->cxx += b->cxx + dx * dx * w;
    a->cyy += b->cyy + dy * dy * w;
    a->cxy += b->cxy + dx * dy * w;
    a->mx += dx * (double)b->n / (double)n;
    a->my += dy * (double)b->n / (double)n;
    a->n = n;
}

static void comoment_finish(const comoment* s, double* corr, double* covar_samp, int64_t* cnt) {
    *cnt = s->n;
    if (s->n < 2) {
        *corr = NAN;
        *covar_samp = NAN;
        return;
    }
    double den = sqrt(s->cxx * s->cyy);
    *corr = den > 0 ? s->cxy / den : NAN;
    *covar_samp = s->cxy / (double)(s->n - 1);
}

/* CORR(x, y) over rows where both are valid: ONE pass, online co-moments per partition merged pairwise (Chan), the
 * shape of DataFusion's accumulator (the columns are read once: 16.25 B/row) */
void to_corr_f64(const double* x, const uint8_t* vx, const double* y, const uint8_t* vy, int64_t n, double* corr,
                 double* covar_samp, int64_t* cnt) {
    int nt = to_num_threads();
    comoment* part = calloc((size_t)nt, sizeof(comoment));
#pragma omp parallel num_threads(nt)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
        comoment s = {0, 0, 0, 0, 0, 0};
        for (int64_t i = lo; i < hi; ++i)
            if (bit(vx, i) && bit(vy, i)) comoment_add(&s, x[i], y[i]);
        part[t] = s;
    }
    comoment all = {0, 0, 0, 0, 0, 0};
    for (int t = 0; t < nt; ++t) comoment_merge(&all, &part[t]);
    free(part);
    comoment_finish(&all, corr, covar_samp, cnt);
}

/* The C2 literal suite as ONE fused scan (BASELINE.md §2.2's "fused single-scan CPU variant"): what the reference's
 * optimizer/ would do if it fused the suite — every referenced column is read once (32.5 B/row). out[0] = MIN(f0),
 * out[1] = AVG(f1), out[2] = CORR(f0, f1), out[3] = satisfied / n of (f2 > a AND i0 < b) */
void to_fused_c2(const double* f0, const uint8_t* v0, const double* f1, const uint8_t* v1, const double* f2, const uint8_t* v2,
                 const int64_t* i0, const uint8_t* vi, double a, int64_t b, int64_t n, double* out) {
    int nt = to_num_threads();
    comoment* part = calloc((size_t)nt, sizeof(comoment));
    double* pmin = calloc((size_t)nt, sizeof(double));
    double* psum = calloc((size_t)nt, sizeof(double));
    int64_t* pcnt = calloc((size_t)nt, sizeof(int64_t));
    int64_t* ppred = calloc((size_t)nt, sizeof(int64_t));
#pragma omp parallel num_threads(nt)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
        comoment s = {0, 0, 0, 0, 0, 0};
        double mn = INFINITY, sum1 = 0;
        int64_t c1 = 0, pred = 0;
        for (int64_t i = lo; i < hi; ++i) {
            const int b0 = bit(v0, i), b1 = bit(v1, i);
            if (b0) mn = f0[i] < mn ? f0[i] : mn;
            if (b1) {
                sum1 += f1[i];
                ++c1;
            }
            if (b0 && b1) comoment_add(&s, f0[i], f1[i]);
            pred += (bit(v2, i) && bit(vi, i) && f2[i] > a && i0[i] < b) ? 1 : 0;
        }
        part[t] = s;
        pmin[t] = mn;
        psum[t] = sum1;
        pcnt[t] = c1;
        ppred[t] = pred;
    }
    comoment all = {0, 0, 0, 0, 0, 0};
    double mn = INFINITY, sum1 = 0;
    int64_t c1 = 0, pred = 0, cn;
    for (int t = 0; t < nt; ++t) {
        comoment_merge(&all, &part[t]);
        mn = pmin[t] < mn ? pmin[t] : mn;
        sum1 += psum[t];
        c1 += pcnt[t];
        pred += ppred[t];
    }
    double cov;
    out[0] = mn;
    out[1] = c1 ? sum1 / (double)c1 : NAN;
    comoment_finish(&all, &out[2], &cov, &cn);
    out[3] = n ? (double)pred / (double)n : NAN;
    free(part);
    free(pmin);
    free(psum);
    free(pcnt);
    free(ppred);
}

/* COUNT(CASE WHEN f > a AND i < b THEN 1 END): NULL operands make the row unsatisfied unless the other
 * conjunct is FALSE (either way not counted) */
int64_t to_pred_gt_lt(const double* f, const uint8_t* vf, double a, const int64_t* iv, const uint8_t* vi, int64_t b,
                      int64_t n) {
    int64_t c = 0;
#pragma omp parallel for reduction(+ : c) schedule(static)
    for (int64_t i = 0; i < n; ++i) c += (bit(vf, i) && bit(vi, i) && f[i] > a && iv[i] < b) ? 1 : 0;
    return c;
}
