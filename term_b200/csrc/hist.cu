// Equal-width histogram of a Float64 column over [min, max] (analyzers/advanced/histogram.rs:256-345).
//
// The reference issues a second query whose CASE expression lists the bucket bounds computed from the first query's
// MIN / MAX:  lower_i = min + i * w,  upper_i = min + (i + 1) * w  (last bucket: max + 0.001 w),  w = (max - min) / nb
// (1.0 when the range is empty or nb = 1);  `WHEN c >= lower_i AND c < upper_i THEN i + 1 ... ELSE nb`. Here the
// bounds are recomputed with the same f64 expressions, a value's bucket is guessed as floor((v - min) / w) and then
// corrected against those exact bounds, and the counts go through a shared-memory histogram. The min / max come from
// the column's NUM aggregate of the same plan (evaluated by the fused scan just before).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "engine.hpp"

namespace tg {

constexpr int HIST_THREADS = 256, HIST_ILP = 4, HIST_MAX_BUCKETS = 1000;

__global__ void __launch_bounds__(HIST_THREADS) hist_kernel(const double* __restrict__ values, const uint32_t* __restrict__ validity, int64_t n,
                                                            double mn, double width, int nb, unsigned long long* out) {
    __shared__ uint32_t s_hist[HIST_MAX_BUCKETS];
    for (int i = threadIdx.x; i < nb; i += HIST_THREADS) s_hist[i] = 0;
    __syncthreads();
    const double inv = 1.0 / width;
    for (int64_t base = (int64_t)blockIdx.x * HIST_THREADS * HIST_ILP; base < n; base += (int64_t)gridDim.x * HIST_THREADS * HIST_ILP) {
        double v[HIST_ILP];
        uint32_t vw[HIST_ILP];
#pragma unroll
        for (int k = 0; k < HIST_ILP; ++k) {
            const int64_t row = base + (int64_t)k * HIST_THREADS + threadIdx.x;
            const int64_t rc = row < n ? row : n - 1;
            v[k] = __ldg(values + rc);
            vw[k] = validity ? __ldg(validity + (rc >> 5)) : 0xffffffffu;
        }
#pragma unroll
        for (int k = 0; k < HIST_ILP; ++k) {
            const int64_t row = base + (int64_t)k * HIST_THREADS + threadIdx.x;
            if (row >= n || !((vw[k] >> (row & 31)) & 1u)) continue;
            const double x = v[k];
            int i = nb - 1;  // ELSE branch (NaN compares false everywhere)
            if (x == x) {
                const double g = floor((x - mn) * inv);
                i = g < 0.0 ? 0 : (g > (double)(nb - 1) ? nb - 1 : (int)g);
                // exact bounds, as the reference's SQL text carries them: product and sum rounded separately (an FMA
                // would round once and move values that sit exactly on a bound)
                while (i > 0 && x < __dadd_rn(mn, __dmul_rn((double)i, width))) --i;
                while (i < nb - 1 && x >= __dadd_rn(mn, __dmul_rn((double)(i + 1), width))) ++i;
            }
            atomicAdd(&s_hist[i], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += HIST_THREADS)
        if (s_hist[i]) atomicAdd(&out[i], (unsigned long long)s_hist[i]);
}

// counts of the Float64 column `c` over nb equal-width buckets of [mn, mx] -> counts[nb] (host); returns launches
static int hist_count(Engine& e, Plan& p, const Column* c, int64_t n_rows, double mn, double mx, int nb, uint64_t* counts) {
    const double range = mx - mn, width = (range > 0.0 && nb > 1) ? range / (double)nb : 1.0;
    // algorithmic bytes: the column is read a second time (the reference scans it twice as well)
    p.stats.bytes_scanned += (uint64_t)n_rows * 8 + (c->validity.p ? (uint64_t)(n_rows + 7) / 8 : 0);
    uint8_t* scr = e.scratch((size_t)HIST_MAX_BUCKETS * 8 + 256);
    unsigned long long* d_out = (unsigned long long*)scr;
    TG_CUDA(cudaEventRecord(e.ev[6], e.stream));
    TG_CUDA(cudaMemsetAsync(d_out, 0, (size_t)nb * 8, e.stream));
    const int64_t per_cta = (int64_t)HIST_THREADS * HIST_ILP;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_rows + per_cta - 1) / per_cta, (int64_t)e.sm_count * 8));
    hist_kernel<<<grid, HIST_THREADS, 0, e.stream>>>((const double*)c->values.p, (const uint32_t*)c->validity.p, n_rows, mn, width, nb, d_out);
    TG_CUDA(cudaGetLastError());
    TG_CUDA(cudaMemcpyAsync(counts, d_out, (size_t)nb * 8, cudaMemcpyDeviceToHost, e.stream));
    TG_CUDA(cudaEventRecord(e.ev[7], e.stream));
    TG_CUDA(cudaStreamSynchronize(e.stream));
    float ms = 0;
    TG_CUDA(cudaEventElapsedTime(&ms, e.ev[6], e.ev[7]));
    p.stats.sketch_ms += ms;
    p.stats.gpu_ms += ms;
    p.stats.launches += 1;
    e.launches += 1;
    return 1;
}

static const Agg* hist_num_agg(const Plan& p, const Agg& a) {
    const Agg* num = nullptr;
    for (auto& o : p.aggs)
        if (o.kind == A_NUM && o.cols.size() == 1 && o.cols[0] == a.cols[0]) num = &o;
    return num;
}

void exec_hist_job(Engine& e, Table& t, Plan& p, int agg_id) {
    Agg& a = p.aggs[agg_id];
    Column* c = t.find(a.cols[0]);
    if (!c) throw Error(TG_ERR_COLUMN_NOT_FOUND, "Schema error: No field named " + a.cols[0] + ". Valid fields are " + t.valid_fields() + ".");
    const int nb = std::min(std::max(a.iparam, 1), HIST_MAX_BUCKETS);
    a.blob.assign((size_t)nb * 8, 0);
    if (c->dtype != TG_FLOAT64) return;  // the slot reports the reference's Float64 downcast error from the NUM aggregate
    const Agg* num = hist_num_agg(p, a);
    if (!num || num->err != TG_OK || num->u[0] == 0 || t.n_rows == 0) {
        a.blob.clear();  // a shard without values has no range of its own: it merges into any range
        return;
    }
    a.f[0] = num->f[3];
    a.f[1] = num->f[4];
    hist_count(e, p, c, t.n_rows, a.f[0], a.f[1], nb, (uint64_t*)a.blob.data());
}

// Second phase of a row-sharded histogram (analyzers/advanced/histogram.rs:184-290 takes the bucket bounds from the
// table-wide MIN / MAX): after the shards' partials were merged, the plan's NUM aggregate holds the GLOBAL min / max;
// this re-counts the LOCAL shard `t` against those bounds. The host layer sums the counts of all shards (a u64
// all-reduce) and installs them (Plan::histogram_install).
void hist_rebucket(Engine& e, Table* t, Plan& p, int agg_id, uint64_t* counts, int nb_out) {
    std::lock_guard<std::mutex> g(e.mu);
    TG_CUDA(cudaSetDevice(e.device));
    e.sync_copies();
    if (agg_id < 0 || agg_id >= (int)p.aggs.size() || p.aggs[agg_id].kind != A_HIST) throw Error(TG_ERR_INVALID_ARG, "not a histogram aggregate");
    Agg& a = p.aggs[agg_id];
    const int nb = std::min(std::max(a.iparam, 1), HIST_MAX_BUCKETS);
    if (nb_out != nb || !counts) throw Error(TG_ERR_INVALID_ARG, "histogram has " + std::to_string(nb) + " buckets");
    for (int i = 0; i < nb; ++i) counts[i] = 0;
    const Agg* num = hist_num_agg(p, a);
    if (!num || num->err != TG_OK || num->u[0] == 0 || num->u[4]) return;
    if (!t || t->n_rows == 0) return;
    Column* c = t->find(a.cols[0]);
    if (!c) throw Error(TG_ERR_COLUMN_NOT_FOUND, "Schema error: No field named " + a.cols[0] + ". Valid fields are " + t->valid_fields() + ".");
    if (c->dtype != TG_FLOAT64) return;
    hist_count(e, p, c, t->n_rows, num->f[3], num->f[4], nb, counts);
}

}  // namespace tg
