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
 *   to_corr_f64      CORR(x,y) over pairwise-complete rows   constraints/correlation.rs:313-316
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

/* CORR(x, y): two-pass centred sums over rows where both are valid */
void to_corr_f64(const double* x, const uint8_t* vx, const double* y, const uint8_t* vy, int64_t n, double* corr,
                 double* covar_samp, int64_t* cnt) {
    double sx = 0, sy = 0;
    int64_t c = 0;
#pragma omp parallel for reduction(+ : sx, sy, c) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        if (bit(vx, i) && bit(vy, i)) {
            sx += x[i];
            sy += y[i];
            ++c;
        }
    }
    *cnt = c;
    if (c < 2) {
        *corr = NAN;
        *covar_samp = NAN;
        return;
    }
    double mx = sx / (double)c, my = sy / (double)c, sxx = 0, syy = 0, sxy = 0;
#pragma omp parallel for reduction(+ : sxx, syy, sxy) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        if (bit(vx, i) && bit(vy, i)) {
            double dx = x[i] - mx, dy = y[i] - my;
            sxx += dx * dx;
            syy += dy * dy;
            sxy += dx * dy;
        }
    }
    double den = sqrt(sxx * syy);
    *corr = den > 0 ? sxy / den : NAN;
    *covar_samp = sxy / (double)(c - 1);
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
