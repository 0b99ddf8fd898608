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
constexpr int64_t KLL_TARGET_SAMPLES = 1 << 19;
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

// Level 0. Thread t streams the row pairs t, t+T, t+2T, .. (one 128-bit load each, KLL_UNROLL loads in flight),
// keeps one uniformly chosen value out of every `s` valid values it sees (the position inside the group is drawn
// once per group) and writes its e-th sample to slot e*T + t: coalesced, no atomics, and the sample set is a
// pure function of (data, seed, grid) — run-to-run reproducible. Slots that are never written keep weight 0.
__global__ void __launch_bounds__(KLL_THREADS) kll_sample_kernel(const uint8_t* __restrict__ values,
                                                                 const uint32_t* __restrict__ validity, int64_t n, int is_i64,
                                                                 uint32_t s, uint32_t seed, double* __restrict__ out_vals,
                                                                 uint32_t* __restrict__ out_w, uint32_t max_emit,
                                                                 KllCounters* ctr) {
    const int64_t T = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_pairs = (n + 1) >> 1;
    uint32_t rng = (uint32_t)(tid * 2654435761u) ^ seed ^ 0x9e3779b9u;
    rng = rng ? rng : 1u;
    auto draw = [&]() {
        rng ^= rng << 13;
        rng ^= rng >> 17;
        rng ^= rng << 5;
        return 1u + __umulhi(rng, s);  // uniform in 1..s
    };
    uint32_t c = 0, target = draw(), emitted = 0;
    double cand = 0.0, mn = INFINITY, mx = -INFINITY;
    unsigned long long cnt = 0;
    auto take = [&](double x, bool valid) {
        const bool ok = valid && x == x;
        cnt += ok;
        mn = fmin(mn, ok ? x : mn);
        mx = fmax(mx, ok ? x : mx);
        c += ok;
        if (ok && c == target) cand = x;
        if (c == s) {  // only reachable on an ok value
            if (emitted < max_emit) {
                out_vals[(int64_t)emitted * T + tid] = cand;
                out_w[(int64_t)emitted * T + tid] = s;
            }
            ++emitted;
            c = 0;
            target = draw();
        }
    };
    const double2* v2 = reinterpret_cast<const double2*>(values);
    const longlong2* i2 = reinterpret_cast<const longlong2*>(values);
    for (int64_t p0 = tid; p0 < n_pairs; p0 += T * KLL_UNROLL) {
        double x[KLL_UNROLL][2];
        uint32_t bits[KLL_UNROLL];
#pragma unroll
        for (int u = 0; u < KLL_UNROLL; ++u) {
            const int64_t p = p0 + (int64_t)u * T;
            bits[u] = 0;
            x[u][0] = x[u][1] = 0.0;
            if (p < n_pairs) {
                const int64_t row = p * 2;
                bits[u] = validity ? (__ldg(validity + (row >> 5)) >> (row & 31)) & 3u : 3u;
                if (row + 1 >= n) bits[u] &= 1u;
                if (is_i64) {
                    const longlong2 w = __ldg(i2 + p);
                    x[u][0] = (double)w.x;
                    x[u][1] = (double)w.y;
                } else {
                    const double2 w = __ldg(v2 + p);
                    x[u][0] = w.x;
                    x[u][1] = w.y;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < KLL_UNROLL; ++u) {
            take(x[u][0], bits[u] & 1u);
            take(x[u][1], bits[u] & 2u);
        }
    }
    if (c && emitted < max_emit) {  // partial last group, weight = its size
        out_vals[(int64_t)emitted * T + tid] = cand;
        out_w[(int64_t)emitted * T + tid] = c;
    }
    uint64_t kmin = f64_key(mn), kmax = f64_key(mx);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, m);
        const uint64_t omin = __shfl_xor_sync(0xffffffffu, (unsigned long long)kmin, m), omax = __shfl_xor_sync(0xffffffffu, (unsigned long long)kmax, m);
        kmin = omin < kmin ? omin : kmin;
        kmax = omax > kmax ? omax : kmax;
    }
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicAdd(&ctr->n, cnt);
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
    const unsigned long long W = ctr->n;
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

void exec_kll_job(Engine& e, Table& t, Plan& p, int agg_id) {
    Agg& a = p.aggs[agg_id];
    Column* c = t.find(a.cols[0]);
    if (!c) throw Error(TG_ERR_COLUMN_NOT_FOUND, "Schema error: No field named " + a.cols[0] + ". Valid fields are " + t.valid_fields() + ".");
    if (c->dtype != TG_INT64 && c->dtype != TG_FLOAT64)
        throw Error(TG_ERR_TYPE_MISMATCH, "quantile sketch requires a numeric (Int64 / Float64) column");
    const int64_t n = t.n_rows;
    // sketch capacity 8k items; above 2^20 the resample arithmetic (i * W) would need 128 bits on the device
    const uint64_t cap = (uint64_t)std::min<int64_t>(std::max<int64_t>(8 * (int64_t)a.iparam, 64), (int64_t)1 << 20);
    p.stats.bytes_scanned += (uint64_t)n * 8 + (c->validity.p ? (uint64_t)(n + 7) / 8 : 0);
    auto write_blob = [&](uint64_t cnt, double mn, double mx, const std::vector<double>& v, const std::vector<uint64_t>& w) {
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
    };
    if (n == 0) {
        write_blob(0, INFINITY, -INFINITY, {}, {});
        return;
    }
    const int64_t n_pairs = (n + 1) / 2;
    const uint32_t s = (uint32_t)std::max<int64_t>(1, (n + KLL_TARGET_SAMPLES - 1) / KLL_TARGET_SAMPLES);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_pairs + KLL_THREADS - 1) / KLL_THREADS, (int64_t)e.sm_count * KLL_CTAS_PER_SM));
    const int64_t T = (int64_t)grid * KLL_THREADS;
    const int64_t rows_per_thread = 2 * ((n_pairs + T - 1) / T);
    const uint32_t max_emit = (uint32_t)(rows_per_thread / s + 1);
    const int64_t m = T * (int64_t)max_emit;  // sample slots (unused ones have weight 0)
    const size_t v_b = round_up((size_t)m * 8, 256), w_b = round_up((size_t)m * 4, 256), cum_b = round_up((size_t)m * 8, 256);
    const size_t pick_b = round_up((size_t)cap * 8, 256);
    size_t sort_b = 0, scan_b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_b, (const double*)nullptr, (double*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    m, 0, 64, e.stream);
    cub::TransformInputIterator<unsigned long long, U32ToU64, const uint32_t*> w_it((const uint32_t*)nullptr, U32ToU64());
    cub::DeviceScan::InclusiveSum(nullptr, scan_b, w_it, (unsigned long long*)nullptr, m, e.stream);
    const size_t tmp_b = round_up(std::max(sort_b, scan_b), 256);
    uint8_t* scr = e.scratch(2 * v_b + 2 * w_b + cum_b + 2 * pick_b + tmp_b + 256);
    uint8_t* q = scr;
    double* v_in = (double*)q; q += v_b;
    double* v_out = (double*)q; q += v_b;
    uint32_t* w_in = (uint32_t*)q; q += w_b;
    uint32_t* w_out = (uint32_t*)q; q += w_b;
    unsigned long long* d_cum = (unsigned long long*)q; q += cum_b;
    double* d_pick_v = (double*)q; q += pick_b;
    unsigned long long* d_pick_w = (unsigned long long*)q; q += pick_b;
    uint8_t* d_tmp = q; q += tmp_b;
    KllCounters* d_ctr = (KllCounters*)q;
    KllCounters h{0, 0, ~0ull, 0ull};
    TG_CUDA(cudaEventRecord(e.ev[6], e.stream));
    TG_CUDA(cudaMemcpyAsync(d_ctr, &h, sizeof(h), cudaMemcpyHostToDevice, e.stream));
    TG_CUDA(cudaMemsetAsync(v_in, 0x7f, v_b, e.stream));
    TG_CUDA(cudaMemsetAsync(w_in, 0, w_b, e.stream));
    kll_sample_kernel<<<grid, KLL_THREADS, 0, e.stream>>>(c->values.p, (const uint32_t*)c->validity.p, n, c->dtype == TG_INT64, s,
                                                          0x5eed0000u + (uint32_t)agg_id, v_in, w_in, max_emit, d_ctr);
    TG_CUDA(cudaGetLastError());
    TG_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, sort_b, v_in, v_out, w_in, w_out, m, 0, 64, e.stream));
    cub::TransformInputIterator<unsigned long long, U32ToU64, const uint32_t*> w_sorted(w_out, U32ToU64());
    TG_CUDA(cub::DeviceScan::InclusiveSum(d_tmp, scan_b, w_sorted, d_cum, m, e.stream));
    kll_pick_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, e.stream>>>(v_out, d_cum, m, d_ctr, (uint32_t)cap, d_pick_v, d_pick_w);
    TG_CUDA(cudaGetLastError());
    const int launches = 2 + 4 + 2;
    uint8_t* hs = e.host_scratch(2 * pick_b + 64);
    double* h_v = (double*)hs;
    unsigned long long* h_w = (unsigned long long*)(hs + pick_b);
    KllCounters* h_ctr = (KllCounters*)(hs + 2 * pick_b);
    TG_CUDA(cudaMemcpyAsync(h_v, d_pick_v, (size_t)cap * 8, cudaMemcpyDeviceToHost, e.stream));
    TG_CUDA(cudaMemcpyAsync(h_w, d_pick_w, (size_t)cap * 8, cudaMemcpyDeviceToHost, e.stream));
    TG_CUDA(cudaMemcpyAsync(h_ctr, d_ctr, sizeof(KllCounters), cudaMemcpyDeviceToHost, e.stream));
    TG_CUDA(cudaEventRecord(e.ev[7], e.stream));
    TG_CUDA(cudaStreamSynchronize(e.stream));
    float ms = 0;
    TG_CUDA(cudaEventElapsedTime(&ms, e.ev[6], e.ev[7]));
    p.stats.sketch_ms += ms;
    p.stats.gpu_ms += ms;
    p.stats.launches += launches;
    e.launches += launches;
    h = *h_ctr;
    a.u[0] = h.n;
    if (h.n == 0) {
        write_blob(0, INFINITY, -INFINITY, {}, {});
        return;
    }
    std::vector<double> ov;
    std::vector<uint64_t> ow;
    for (uint64_t i = 0; i < cap; ++i)
        if (h_w[i]) {
            ov.push_back(h_v[i]);
            ow.push_back(h_w[i]);
        }
    write_blob(h.n, key_f64(h.min_bits), key_f64(h.max_bits), ov, ow);
}

}  // namespace tg
