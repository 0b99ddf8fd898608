// K4 — quantile sketch (KLL-style) on the device.
//
// Stands where KllSketch::{update, cascade_compact, get_quantile, merge} stand in the reference
// (analyzers/advanced/kll_sketch.rs:166-366). The reference streams values one by one through per-level
// compactors that sort a buffer and keep every other item with doubled weight. The device version keeps the
// two ideas that matter for the rank-error contract (1.65/sqrt(k), :397-399) and maps them to HBM-speed passes:
//   level 0 = a sampler: each lane keeps ONE uniformly chosen value out of every `s` valid values it
//             streams (reservoir of size 1, weight s) — what a KLL compactor ladder of log2(s) levels does
//             to a buffer in expectation, without sorting anything at scan rate;
//   level 1 = one exact sort of the <= ~2M weighted samples, then a systematic resample (sort + keep every
//             j-th = a compactor applied to a sorted buffer) down to 8k items on the host.
// min / max / count are exact (kll_sketch.rs:201-203); NaN is skipped (:197-199).
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "engine.hpp"

namespace tg {

constexpr int KLL_THREADS = 256;
constexpr int64_t KLL_TARGET_SAMPLES = 1 << 21;

struct KllCounters {
    unsigned long long n;         // valid, non-NaN values
    unsigned long long n_samples;
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

__global__ void __launch_bounds__(KLL_THREADS) kll_sample_kernel(const uint8_t* values, const uint32_t* validity, int64_t n,
                                                                 int is_i64, uint32_t s, uint32_t seed, double* out_vals,
                                                                 uint32_t* out_w, uint64_t out_cap, KllCounters* ctr) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    uint32_t rng = (uint32_t)(tid * 2654435761u) ^ seed ^ 0x9e3779b9u;
    rng = rng ? rng : 1u;
    uint32_t c = 0;
    double cand = 0.0;
    unsigned long long cnt = 0;
    uint64_t kmin = ~0ull, kmax = 0ull;
    auto emit = [&](double v, uint32_t w) {
        const unsigned long long i = atomicAdd(&ctr->n_samples, 1ull);
        if (i < out_cap) {
            out_vals[i] = v;
            out_w[i] = w;
        }
    };
    for (int64_t row = tid; row < n; row += stride) {
        if (validity && !((validity[row >> 5] >> (row & 31)) & 1u)) continue;
        const double x = is_i64 ? (double)reinterpret_cast<const int64_t*>(values)[row] : reinterpret_cast<const double*>(values)[row];
        if (x != x) continue;
        ++cnt;
        const uint64_t k = f64_key(x);
        kmin = k < kmin ? k : kmin;
        kmax = k > kmax ? k : kmax;
        ++c;
        rng ^= rng << 13;
        rng ^= rng >> 17;
        rng ^= rng << 5;
        if (__umulhi(rng, c) == 0) cand = x;  // replace with probability 1/c
        if (c == s) {
            emit(cand, s);
            c = 0;
        }
    }
    if (c) emit(cand, c);
    // block reduce of count / min / max
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, m);
        const uint64_t omin = __shfl_xor_sync(0xffffffffu, (unsigned long long)kmin, m), omax = __shfl_xor_sync(0xffffffffu, (unsigned long long)kmax, m);
        kmin = omin < kmin ? omin : kmin;
        kmax = omax > kmax ? omax : kmax;
    }
    if ((threadIdx.x & 31) == 0) {
        if (cnt) atomicAdd(&ctr->n, cnt);
        atomicMin(&ctr->min_bits, (unsigned long long)kmin);
        atomicMax(&ctr->max_bits, (unsigned long long)kmax);
    }
}

static size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

void exec_kll_job(Engine& e, Table& t, Plan& p, int agg_id) {
    Agg& a = p.aggs[agg_id];
    Column* c = t.find(a.cols[0]);
    if (!c) throw Error(TG_ERR_COLUMN_NOT_FOUND, "Schema error: No field named " + a.cols[0] + ". Valid fields are " + t.valid_fields() + ".");
    if (c->dtype != TG_INT64 && c->dtype != TG_FLOAT64)
        throw Error(TG_ERR_TYPE_MISMATCH, "quantile sketch requires a numeric (Int64 / Float64) column");
    const int64_t n = t.n_rows;
    const uint64_t cap = (uint64_t)std::max(8 * a.iparam, 64);
    p.stats.bytes_scanned += (uint64_t)n * 8 + (c->validity.p ? (uint64_t)(n + 7) / 8 : 0);
    // empty sketch blob
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
    const uint32_t s = (uint32_t)std::max<int64_t>(1, (n + KLL_TARGET_SAMPLES - 1) / KLL_TARGET_SAMPLES);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + KLL_THREADS - 1) / KLL_THREADS, (int64_t)e.sm_count * 8));
    const uint64_t out_cap = (uint64_t)(n / s) + (uint64_t)grid * KLL_THREADS + 64;
    const size_t v_b = round_up(out_cap * 8, 256), w_b = round_up(out_cap * 4, 256);
    size_t tmp_b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_b, (const double*)nullptr, (double*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (int64_t)out_cap, 0, 64, e.stream);
    tmp_b = round_up(tmp_b, 256);
    uint8_t* scr = e.scratch(2 * v_b + 2 * w_b + tmp_b + 256);
    double* v_in = (double*)scr;
    double* v_out = (double*)(scr + v_b);
    uint32_t* w_in = (uint32_t*)(scr + 2 * v_b);
    uint32_t* w_out = (uint32_t*)(scr + 2 * v_b + w_b);
    uint8_t* d_tmp = scr + 2 * v_b + 2 * w_b;
    KllCounters* d_ctr = (KllCounters*)(scr + 2 * v_b + 2 * w_b + tmp_b);
    KllCounters h{0, 0, ~0ull, 0ull};
    TG_CUDA(cudaEventRecord(e.ev[6], e.stream));
    TG_CUDA(cudaMemcpyAsync(d_ctr, &h, sizeof(h), cudaMemcpyHostToDevice, e.stream));
    kll_sample_kernel<<<grid, KLL_THREADS, 0, e.stream>>>(c->values.p, (const uint32_t*)c->validity.p, n, c->dtype == TG_INT64, s,
                                                          0x5eed0000u + (uint32_t)agg_id, v_in, w_in, out_cap, d_ctr);
    TG_CUDA(cudaGetLastError());
    TG_CUDA(cudaMemcpyAsync(&h, d_ctr, sizeof(h), cudaMemcpyDeviceToHost, e.stream));
    TG_CUDA(cudaStreamSynchronize(e.stream));
    const int64_t m = (int64_t)std::min<uint64_t>(h.n_samples, out_cap);
    int launches = 1;
    std::vector<double> hv((size_t)m);
    std::vector<uint32_t> hw((size_t)m);
    if (m > 0) {
        TG_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_b, v_in, v_out, w_in, w_out, m, 0, 64, e.stream));
        launches += 4;
        TG_CUDA(cudaMemcpyAsync(hv.data(), v_out, (size_t)m * 8, cudaMemcpyDeviceToHost, e.stream));
        TG_CUDA(cudaMemcpyAsync(hw.data(), w_out, (size_t)m * 4, cudaMemcpyDeviceToHost, e.stream));
    }
    TG_CUDA(cudaEventRecord(e.ev[7], e.stream));
    TG_CUDA(cudaStreamSynchronize(e.stream));
    float ms = 0;
    TG_CUDA(cudaEventElapsedTime(&ms, e.ev[6], e.ev[7]));
    p.stats.sketch_ms += ms;
    p.stats.gpu_ms += ms;
    p.stats.launches += launches;
    e.launches += launches;
    a.u[0] = h.n;
    if (h.n == 0) {
        write_blob(0, INFINITY, -INFINITY, {}, {});
        return;
    }
    // systematic resample of the sorted weighted samples down to `cap` items (same rule as kll_host.cpp)
    std::vector<double> ov;
    std::vector<uint64_t> ow;
    if ((uint64_t)m <= cap) {
        ov.assign(hv.begin(), hv.end());
        ow.assign(hw.begin(), hw.end());
    } else {
        uint64_t W = 0;
        for (auto w : hw) W += w;
        size_t j = 0;
        uint64_t cum = hw[0];
        for (uint64_t i = 0; i < cap; ++i) {
            const uint64_t lo = (uint64_t)((__uint128_t)i * W / cap), hi = (uint64_t)((__uint128_t)(i + 1) * W / cap);
            if (hi == lo) continue;
            const uint64_t target = lo + (hi - lo + 1) / 2;
            while (cum < target && j + 1 < (size_t)m) cum += hw[++j];
            ov.push_back(hv[j]);
            ow.push_back(hi - lo);
        }
    }
    write_blob(h.n, key_f64(h.min_bits), key_f64(h.max_bits), ov, ow);
}

}  // namespace tg
