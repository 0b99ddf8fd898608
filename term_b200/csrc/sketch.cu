// K4 — quantile sketch (KLL-style) on the device.
//
// Stands where KllSketch::{update, cascade_compact, get_quantile, merge} stand in the reference
// (analyzers/advanced/kll_sketch.rs:166-366). The reference streams values one by one through per-level
// compactors that sort a buffer and keep every other item with doubled weight. The device version keeps the
// two ideas that matter for the rank-error contract (1.65/sqrt(k), :397-399) and maps them to HBM-speed passes:
//   level 0 = a sampler: each lane keeps ONE uniformly chosen value out of every `s` valid values it
//             streams (weight s) — what a KLL compactor ladder of log2(s) levels does to a buffer in
//             expectation, without sorting anything at scan rate;
//   level 1 = one exact sort of the <= ~0.8M weighted samples, a prefix sum of their weights and a systematic
//             resample (sort + keep every j-th = a compactor applied to a sorted buffer) down to 8k items,
//             all on the device; only the 8k items cross PCIe.
// min / max / count are exact (kll_sketch.rs:201-203); NaN is skipped (:197-199).
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "engine.hpp"

namespace tg {

constexpr int KLL_THREADS = 256;
constexpr int KLL_CTAS_PER_SM = 4;
constexpr int64_t KLL_TARGET_SAMPLES = 1 << 17;
constexpr int KLL_UNROLL = 4;

struct KllCounters {
    unsigned long long n;         // valid, non-NaN values == total weight of the samples
    unsigned long long pad;
    unsigned long long min_bits;  // order-preserving keys (see f64_key)
    unsigned long long max_bits;
};

__device__ __forceinline__ uint64_t f64_key(double v) {
    uint64_t b = (uint64_t)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double key_f64(uint64_t k) {
    uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    double d;
    memcpy(&d, &b, 8);
    return d;
}

// Level 0. Thread t streams the row pairs t, t+T, t+2T, .. (one 128-bit load each, KLL_UNROLL loads in flight).
//   EXACT   (n <= KLL_TARGET_SAMPLES): every valid value becomes a sample of weight 1.
//   sampled: the thread's pairs form groups of g pairs; ONE position inside each group is drawn up front and the
//            value at that position, if valid, becomes a sample of weight 2g (the rows the group stands for). Every
//            valid value is thus kept with probability 1/2g and weight 2g — an unbiased (Horvitz–Thompson) weighted
//            sample whatever the placement of the NULLs — and the per-pair cost is one compare, not a per-value
//            reservoir update.
// The e-th sample of thread t goes to slot e*T + t: coalesced, no atomics, and the sample set is a pure function of
// (data, seed, grid) — run-to-run reproducible. Slots never written keep weight 0. min / max / count are exact.
template <bool IS_I64, bool EXACT>
__global__ void __launch_bounds__(KLL_THREADS) kll_sample_kernel(const uint8_t* __restrict__ values,
                                                                 const uint32_t* __restrict__ validity, int64_t n,
                                                                 uint32_t g, uint32_t seed, double* __restrict__ out_vals,
                                                                 uint32_t* __restrict__ out_w, uint32_t max_emit,
                                                                 KllCounters* ctr) {
    const int64_t T = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_pairs = (n + 1) >> 1;
    uint32_t rng = (uint32_t)(tid * 2654435761u) ^ seed ^ 0x9e3779b9u;
    rng = rng ? rng : 1u;
    auto draw = [&](uint32_t range) {  // uniform in 0..range-1
        rng ^= rng << 13;
        rng ^= rng >> 17;
        rng ^= rng << 5;
        return __umulhi(rng, range);
    };
    uint32_t emitted = 0, cnt = 0;
    double mn = INFINITY, mx = -INFINITY;
    auto emit = [&](double v, uint32_t w) {
        if (emitted < max_emit) {
            out_vals[(int64_t)emitted * T + tid] = v;
            out_w[(int64_t)emitted * T + tid] = w;
        }
        ++emitted;
    };
    // group state (sampled mode)
    int64_t remaining = tid < n_pairs ? (n_pairs - tid + T - 1) / T : 0;  // pairs this thread will see
    uint32_t glen = (uint32_t)(remaining < (int64_t)g ? remaining : (int64_t)g), q = 0, jt = glen ? draw(2 * glen) : 0;
    double cand = 0.0;
    bool cok = false;
    const ulonglong2* v2 = reinterpret_cast<const ulonglong2*>(values);
    for (int64_t p0 = tid; p0 < n_pairs; p0 += T * KLL_UNROLL) {
        // wave 1: all loads of the iteration, branch-free (clamped indices), nothing consumed yet
        ulonglong2 w[KLL_UNROLL];
        uint32_t vraw[KLL_UNROLL], bits[KLL_UNROLL];
#pragma unroll
        for (int u = 0; u < KLL_UNROLL; ++u) {
            const int64_t p = p0 + (int64_t)u * T;
            const int64_t pc = p < n_pairs ? p : n_pairs - 1;
            w[u] = __ldg(v2 + pc);
            vraw[u] = validity ? __ldg(validity + (pc >> 4)) : 0xffffffffu;  // row = 2 pc, word = row >> 5
        }
        // wave 2: validity bits of the pair; bit 2 = the pair exists
#pragma unroll
        for (int u = 0; u < KLL_UNROLL; ++u) {
            const int64_t p = p0 + (int64_t)u * T;
            const bool in = p < n_pairs;
            uint32_t b = (vraw[u] >> ((uint32_t)(p << 1) & 31u)) & 3u;
            if (p * 2 + 1 >= n) b &= 1u;
            bits[u] = in ? (b | 4u) : 0u;
        }
#pragma unroll
        for (int u = 0; u < KLL_UNROLL; ++u) {
            const double x0 = IS_I64 ? (double)(long long)w[u].x : __longlong_as_double((long long)w[u].x);
            const double x1 = IS_I64 ? (double)(long long)w[u].y : __longlong_as_double((long long)w[u].y);
            const bool ok0 = (bits[u] & 1u) && x0 == x0, ok1 = (bits[u] & 2u) && x1 == x1;
            if (ok0) {
                ++cnt;
                mn = fmin(mn, x0);
                mx = fmax(mx, x0);
            }
            if (ok1) {
                ++cnt;
                mn = fmin(mn, x1);
                mx = fmax(mx, x1);
            }
            if (EXACT) {
                if (ok0) emit(x0, 1u);
                if (ok1) emit(x1, 1u);
            } else if (bits[u] & 4u) {
                if (q == (jt >> 1)) {
                    cand = (jt & 1u) ? x1 : x0;
                    cok = (jt & 1u) ? ok1 : ok0;
                }
                if (++q == glen) {
                    if (cok) emit(cand, 2u * glen);
                    remaining -= glen;
                    glen = (uint32_t)(remaining < (int64_t)g ? remaining : (int64_t)g);
                    q = 0;
                    cok = false;
                    jt = glen ? draw(2 * glen) : 0;
                }
            }
        }
    }
    uint64_t kmin = f64_key(mn), kmax = f64_key(mx);
    unsigned long long cnt64 = cnt;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        cnt64 += __shfl_xor_sync(0xffffffffu, cnt64, m);
        const uint64_t omin = __shfl_xor_sync(0xffffffffu, (unsigned long long)kmin, m), omax = __shfl_xor_sync(0xffffffffu, (unsigned long long)kmax, m);
        kmin = omin < kmin ? omin : kmin;
        kmax = omax > kmax ? omax : kmax;
    }
    if ((threadIdx.x & 31) == 0 && cnt64) {
        atomicAdd(&ctr->n, cnt64);
        atomicMin(&ctr->min_bits, (unsigned long long)kmin);
        atomicMax(&ctr->max_bits, (unsigned long long)kmax);
    }
}

// Level 1, after the sort: systematic resample of the value-sorted weighted samples down to `cap` items of
// (almost) equal weight — output i stands for ranks (i*W/cap, (i+1)*W/cap] and takes the sample holding the
// midpoint rank (same rule as kll_compact in kll_host.cpp). cum[] = inclusive prefix sums of the weights.
__global__ void kll_pick_kernel(const double* __restrict__ sorted_vals, const unsigned long long* __restrict__ cum, int64_t m,
                                const KllCounters* __restrict__ ctr, uint32_t cap, double* __restrict__ out_v,
                                unsigned long long* __restrict__ out_w) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    const unsigned long long W = cum[m - 1];  // total weight of the samples (== the valid count in EXACT mode)
    const unsigned long long lo = (unsigned long long)i * W / cap, hi = (unsigned long long)(i + 1) * W / cap;
    out_w[i] = hi - lo;
    if (hi == lo) return;
    const unsigned long long target = lo + (hi - lo + 1) / 2;  // 1-based rank
    int64_t a = 0, b = m - 1;                                     // first j with cum[j] >= target
    while (a < b) {
        const int64_t mid = (a + b) >> 1;
        if (cum[mid] >= target) b = mid;
        else a = mid + 1;
    }
    out_v[i] = sorted_vals[a];
}

struct U32ToU64 {
    __host__ __device__ __forceinline__ unsigned long long operator()(const uint32_t& x) const { return x; }
};

static size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static void write_kll_blob(Agg& a, uint64_t cap, uint64_t cnt, double mn, double mx, const std::vector<double>& v,
                           const std::vector<uint64_t>& w) {
    uint64_t m = v.size();
    a.blob.resize(40 + m * 16);
    memcpy(a.blob.data(), &cnt, 8);
    memcpy(a.blob.data() + 8, &mn, 8);
    memcpy(a.blob.data() + 16, &mx, 8);
    memcpy(a.blob.data() + 24, &cap, 8);
    memcpy(a.blob.data() + 32, &m, 8);
    for (uint64_t i = 0; i < m; ++i) {
        memcpy(a.blob.data() + 40 + i * 16, &v[i], 8);
        memcpy(a.blob.data() + 48 + i * 16, &w[i], 8);
    }
}

// All quantile sketches of a plan in one go: every column's sampler -> sort -> prefix sum -> resample pipeline is
// queued back to back on the stream (own slice of the scratch block each), ONE synchronisation at the end, then the
// <= 8k picked items per column are turned into blobs on the host.
void exec_kll_jobs(Engine& e, Table& t, Plan& p, const std::vector<int>& agg_ids) {
    struct Job {
        int agg;
        Column* c;
        uint64_t cap;
        size_t host_off;
    };
    std::vector<Job> jobs;
    const int64_t n = t.n_rows;
    for (int id : agg_ids) {
        Agg& a = p.aggs[id];
        Column* c = t.find(a.cols[0]);
        if (!c) {
            a.err = TG_ERR_COLUMN_NOT_FOUND;
            a.err_msg = "Schema error: No field named " + a.cols[0] + ". Valid fields are " + t.valid_fields() + ".";
            continue;
        }
        if (c->dtype != TG_INT64 && c->dtype != TG_FLOAT64) {
            a.err = TG_ERR_TYPE_MISMATCH;
            a.err_msg = "quantile sketch requires a numeric (Int64 / Float64) column";
            continue;
        }
        // sketch capacity 8k items; above 2^20 the resample arithmetic (i * W) would need 128 bits on the device
        const uint64_t cap = (uint64_t)std::min<int64_t>(std::max<int64_t>(8 * (int64_t)a.iparam, 64), (int64_t)1 << 20);
        p.stats.bytes_scanned += (uint64_t)n * 8 + (c->validity.p ? (uint64_t)(n + 7) / 8 : 0);
        if (n == 0) {
            write_kll_blob(a, cap, 0, INFINITY, -INFINITY, {}, {});
            continue;
        }
        jobs.push_back(Job{id, c, cap, 0});
    }
    if (jobs.empty()) return;
    const int64_t n_pairs = (n + 1) / 2;
    // EXACT below KLL_TARGET_SAMPLES rows, otherwise groups of g pairs (2g rows) -> about KLL_TARGET_SAMPLES samples
    const bool exact = n <= KLL_TARGET_SAMPLES;
    const uint32_t g = exact ? 1u : (uint32_t)std::max<int64_t>(1, (n_pairs + KLL_TARGET_SAMPLES - 1) / KLL_TARGET_SAMPLES);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_pairs + KLL_THREADS - 1) / KLL_THREADS, (int64_t)e.sm_count * KLL_CTAS_PER_SM));
    const int64_t T = (int64_t)grid * KLL_THREADS;
    const int64_t pairs_per_thread = (n_pairs + T - 1) / T;
    const uint32_t max_emit = (uint32_t)(exact ? 2 * pairs_per_thread : pairs_per_thread / g + 1);
    const int64_t m = T * (int64_t)max_emit;  // sample slots (unused ones have weight 0)
    const size_t v_b = round_up((size_t)m * 8, 256), w_b = round_up((size_t)m * 4, 256), cum_b = round_up((size_t)m * 8, 256);
    size_t sort_b = 0, scan_b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_b, (const double*)nullptr, (double*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    m, 0, 64, e.stream);
    cub::TransformInputIterator<unsigned long long, U32ToU64, const uint32_t*> w_it((const uint32_t*)nullptr, U32ToU64());
    cub::DeviceScan::InclusiveSum(nullptr, scan_b, w_it, (unsigned long long*)nullptr, m, e.stream);
    const size_t tmp_b = round_up(std::max(sort_b, scan_b), 256);
    size_t per_job = 2 * v_b + 2 * w_b + cum_b + tmp_b + 256, host_total = 0;
    for (auto& j : jobs) {
        per_job = std::max(per_job, 2 * v_b + 2 * w_b + cum_b + tmp_b + 256 + 2 * round_up((size_t)j.cap * 8, 256));
        j.host_off = host_total;
        host_total += 2 * round_up((size_t)j.cap * 8, 256) + 64;
    }
    uint8_t* scr = e.scratch(per_job * jobs.size());
    uint8_t* hs = e.host_scratch(host_total);
    TG_CUDA(cudaEventRecord(e.ev[6], e.stream));
    for (size_t ji = 0; ji < jobs.size(); ++ji) {
        Job& j = jobs[ji];
        const size_t pick_b = round_up((size_t)j.cap * 8, 256);
        uint8_t* q = scr + ji * per_job;
        double* v_in = (double*)q; q += v_b;
        double* v_out = (double*)q; q += v_b;
        uint32_t* w_in = (uint32_t*)q; q += w_b;
        uint32_t* w_out = (uint32_t*)q; q += w_b;
        unsigned long long* d_cum = (unsigned long long*)q; q += cum_b;
        double* d_pick_v = (double*)q; q += pick_b;
        unsigned long long* d_pick_w = (unsigned long long*)q; q += pick_b;
        uint8_t* d_tmp = q; q += tmp_b;
        KllCounters* d_ctr = (KllCounters*)q;
        KllCounters* h_init = (KllCounters*)(hs + j.host_off + 2 * pick_b);  // pinned: also the landing place of the result
        *h_init = KllCounters{0, 0, ~0ull, 0ull};
        TG_CUDA(cudaMemcpyAsync(d_ctr, h_init, sizeof(KllCounters), cudaMemcpyHostToDevice, e.stream));
        TG_CUDA(cudaMemsetAsync(v_in, 0x7f, v_b, e.stream));
        TG_CUDA(cudaMemsetAsync(w_in, 0, w_b, e.stream));
        const uint32_t seed = 0x5eed0000u + (uint32_t)j.agg;
        typedef void (*Sampler)(const uint8_t*, const uint32_t*, int64_t, uint32_t, uint32_t, double*, uint32_t*, uint32_t, KllCounters*);
        const bool i64 = j.c->dtype == TG_INT64;
        const Sampler sampler = i64 ? (exact ? (Sampler)kll_sample_kernel<true, true> : (Sampler)kll_sample_kernel<true, false>)
                                    : (exact ? (Sampler)kll_sample_kernel<false, true> : (Sampler)kll_sample_kernel<false, false>);
        sampler<<<grid, KLL_THREADS, 0, e.stream>>>(j.c->values.p, (const uint32_t*)j.c->validity.p, n, g, seed, v_in, w_in, max_emit, d_ctr);
        TG_CUDA(cudaGetLastError());
        // exact mode sorts on all 64 key bits; a sampled sketch only needs the order down to the top 32 bits (sign,
        // exponent, 20 mantissa bits: values closer than 1e-6 relative may swap, far inside the rank-error bound),
        // which halves the radix passes — each is launch-latency bound at this size
        TG_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, sort_b, v_in, v_out, w_in, w_out, m, exact ? 0 : 32, 64, e.stream));
        cub::TransformInputIterator<unsigned long long, U32ToU64, const uint32_t*> w_sorted(w_out, U32ToU64());
        TG_CUDA(cub::DeviceScan::InclusiveSum(d_tmp, scan_b, w_sorted, d_cum, m, e.stream));
        kll_pick_kernel<<<(unsigned)((j.cap + 255) / 256), 256, 0, e.stream>>>(v_out, d_cum, m, d_ctr, (uint32_t)j.cap, d_pick_v, d_pick_w);
        TG_CUDA(cudaGetLastError());
        // the D2H of the counters must not race with the H2D that initialised them from the same pinned slot: both
        // are stream-ordered
        TG_CUDA(cudaMemcpyAsync(hs + j.host_off, d_pick_v, (size_t)j.cap * 8, cudaMemcpyDeviceToHost, e.stream));
        TG_CUDA(cudaMemcpyAsync(hs + j.host_off + pick_b, d_pick_w, (size_t)j.cap * 8, cudaMemcpyDeviceToHost, e.stream));
        TG_CUDA(cudaMemcpyAsync(h_init, d_ctr, sizeof(KllCounters), cudaMemcpyDeviceToHost, e.stream));
        p.stats.launches += 8;
        e.launches += 8;
    }
    TG_CUDA(cudaEventRecord(e.ev[7], e.stream));
    TG_CUDA(cudaStreamSynchronize(e.stream));
    float ms = 0;
    TG_CUDA(cudaEventElapsedTime(&ms, e.ev[6], e.ev[7]));
    p.stats.sketch_ms += ms;
    p.stats.gpu_ms += ms;
    for (auto& j : jobs) {
        Agg& a = p.aggs[j.agg];
        const size_t pick_b = round_up((size_t)j.cap * 8, 256);
        const double* h_v = (const double*)(hs + j.host_off);
        const unsigned long long* h_w = (const unsigned long long*)(hs + j.host_off + pick_b);
        const KllCounters h = *(const KllCounters*)(hs + j.host_off + 2 * pick_b);
        a.u[0] = h.n;
        if (h.n == 0) {
            write_kll_blob(a, j.cap, 0, INFINITY, -INFINITY, {}, {});
            continue;
        }
        std::vector<double> ov;
        std::vector<uint64_t> ow;
        for (uint64_t i = 0; i < j.cap; ++i)
            if (h_w[i]) {
                ov.push_back(h_v[i]);
                ow.push_back(h_w[i]);
            }
        write_kll_blob(a, j.cap, h.n, key_f64(h.min_bits), key_f64(h.max_bits), ov, ow);
    }
}

}  // namespace tg
