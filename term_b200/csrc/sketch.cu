// K4 — KLL quantile sketch on the device.
//
// Stands where KllSketch::{update, cascade_compact, get_quantile, merge} stand in the reference
// (analyzers/advanced/kll_sketch.rs:166-366). The reference streams values one by one through a ladder of compactors:
// a full level is sorted, every other item (odd or even positions, by a coin) moves up one level and doubles its
// weight. The device version builds the SAME state — levels of items, an item of level h standing for 2^h values —
// with passes that run at HBM / launch speed:
//   kll_sample_kernel   the bottom h0 levels as the KLL paper's sampler: ONE position out of every 2^h0 rows, so every
//                       valid value becomes a level-h0 item with probability 2^-h0 (exact, h0 = 0, below 2^19 rows);
//   rs_sort_pairs       hand-written radix sort of the <= ~0.6M item keys (radix_sort.cu);
//   kll_compact_kernel  the compactor ladder applied to the sorted level: while more than `capacity` items remain, pair
//                       neighbours, promote the odd or the even one (counter-based coin per level) one level up and leave
//                       an unpaired last item behind at its level. On a sorted buffer t cascaded compactions are a
//                       strided gather (stride 2^t, offset = the coins), so the whole ladder is ONE small kernel.
// The sketch that leaves the device is {n, min, max, levels[h] = sorted items of weight 2^h}: KllSketch's own fields.
// It merges level by level and answers get_quantile with the reference's rule (kll_host.cpp). min / max / count are
// exact (kll_sketch.rs:201-203); NaN is skipped (:197-199).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "engine.hpp"
#include "radix_sort.cuh"

namespace tg {

constexpr int KLL_THREADS = 256;
constexpr int KLL_CTAS_PER_SM = 4;
constexpr int64_t KLL_TARGET_SAMPLES = 1 << 19;  // level-h0 items kept per column (exact below that many rows)
constexpr int KLL_UNROLL = 4;

struct KllCounters {
    unsigned long long n;         // valid, non-NaN values (exact count)
    unsigned long long n_items;   // items the sampler emitted (each stands for 2^h0 values)
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

// The sampler (the bottom h0 levels of the KLL ladder). Thread t streams the row pairs t, t+T, t+2T, .. (one 128-bit load
// each, KLL_UNROLL loads in flight).
//   EXACT   (n <= KLL_TARGET_SAMPLES): every valid value becomes a level-0 item (weight 1).
//   sampled: the thread's pairs form groups of g pairs, 2g = 2^h0 rows; ONE of the 2g positions of each group is drawn
//            up front and the value there, if it exists and is valid, becomes a level-h0 item (weight 2^h0). Every valid
//            value is thus kept with probability 2^-h0 wherever the NULLs sit and however short the thread's last group
//            is — the Horvitz–Thompson argument of the KLL paper's sampler — and the per-pair cost is one compare.
// Items are written as order-preserving 64-bit keys; the e-th item of thread t goes to slot e*T + t: coalesced, no
// atomics, and the item set is a pure function of (data, seed, grid) — run-to-run reproducible. Slots never written keep
// the all-ones key (greater than every real key) and sort behind the items. min / max / count are exact.
template <bool IS_I64, bool EXACT>
__global__ void __launch_bounds__(KLL_THREADS) kll_sample_kernel(const uint8_t* __restrict__ values,
                                                                 const uint32_t* __restrict__ validity, int64_t n,
                                                                 uint32_t g, uint32_t seed, uint64_t* __restrict__ out_keys,
                                                                 uint32_t max_emit, KllCounters* ctr) {
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
    auto emit = [&](double v) {
        if (emitted < max_emit) out_keys[(int64_t)emitted * T + tid] = f64_key(v == 0.0 ? 0.0 : v);  // -0.0 -> +0.0
        ++emitted;
    };
    // group state (sampled mode)
    int64_t remaining = tid < n_pairs ? (n_pairs - tid + T - 1) / T : 0;  // pairs this thread will see
    uint32_t glen = (uint32_t)(remaining < (int64_t)g ? remaining : (int64_t)g), q = 0, jt = glen ? draw(2 * g) : 0;
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
                if (ok0) emit(x0);
                if (ok1) emit(x1);
            } else if (bits[u] & 4u) {
                if (q == (jt >> 1)) {
                    cand = (jt & 1u) ? x1 : x0;
                    cok = (jt & 1u) ? ok1 : ok0;
                }
                if (++q == glen) {
                    if (cok) emit(cand);  // a short last group draws from all 2g positions too: same inclusion probability
                    remaining -= glen;
                    glen = (uint32_t)(remaining < (int64_t)g ? remaining : (int64_t)g);
                    q = 0;
                    cok = false;
                    jt = glen ? draw(2 * g) : 0;
                }
            }
        }
    }
    uint64_t kmin = f64_key(mn), kmax = f64_key(mx);
    unsigned long long cnt64 = cnt, em64 = emitted < max_emit ? emitted : max_emit;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        cnt64 += __shfl_xor_sync(0xffffffffu, cnt64, m);
        em64 += __shfl_xor_sync(0xffffffffu, em64, m);
        const uint64_t omin = __shfl_xor_sync(0xffffffffu, (unsigned long long)kmin, m), omax = __shfl_xor_sync(0xffffffffu, (unsigned long long)kmax, m);
        kmin = omin < kmin ? omin : kmin;
        kmax = omax > kmax ? omax : kmax;
    }
    if ((threadIdx.x & 31) == 0 && cnt64) {
        atomicAdd(&ctr->n, cnt64);
        atomicAdd(&ctr->n_items, em64);
        atomicMin(&ctr->min_bits, (unsigned long long)kmin);
        atomicMax(&ctr->max_bits, (unsigned long long)kmax);
    }
}

// The compactor ladder over the sorted level-h0 items (see the header). Output block:
//   KllLadder header | leftover items (one per level that had an odd count) | `cap` slots of top-level items
struct KllLadder {
    uint32_t top_level, top_count, n_left, h0;
    uint32_t left_level[64];
    double left_val[64];
};

__host__ __device__ __forceinline__ uint32_t kll_coin(uint32_t seed, uint32_t level) {
    uint32_t x = seed ^ (level * 0x9e3779b9u);
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x & 1u;
}

__global__ void kll_compact_kernel(const RsControl* ctl, const uint64_t* k0, const uint64_t* k1, const KllCounters* __restrict__ ctr,
                                   uint32_t h0, uint32_t cap, uint32_t seed, KllLadder* out, double* __restrict__ out_top) {
    const uint64_t* __restrict__ sorted = ctl->result ? k1 : k0;
    unsigned long long cnt = ctr->n_items, off = 0, stride = 1;
    uint32_t level = h0, n_left = 0;
    const bool writer = blockIdx.x == 0 && threadIdx.x == 0;
    while (cnt > cap) {
        if (cnt & 1ull) {  // the unpaired last item stays at this level
            if (writer && n_left < 64) {
                out->left_level[n_left] = level;
                out->left_val[n_left] = key_f64(sorted[off + (cnt - 1) * stride]);
            }
            ++n_left;
        }
        off += kll_coin(seed, level) ? stride : 0ull;
        stride <<= 1;
        cnt >>= 1;
        ++level;
    }
    if (writer) {
        out->top_level = level;
        out->top_count = (uint32_t)cnt;
        out->n_left = n_left < 64 ? n_left : 64;
        out->h0 = h0;
    }
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt) out_top[i] = key_f64(sorted[off + (unsigned long long)i * stride]);
}

static size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// blob (kll_host.cpp): u64 n | f64 min | f64 max | u64 k | u64 cap | u64 n_levels | per level: u64 count, count x f64 ascending
static void write_kll_blob(Agg& a, uint64_t cap, uint64_t cnt, double mn, double mx, const std::vector<std::vector<double>>& levels) {
    size_t bytes = 48;
    for (auto& l : levels) bytes += 8 + l.size() * 8;
    a.blob.resize(bytes);
    uint8_t* p = a.blob.data();
    const uint64_t k = (uint64_t)a.iparam, nl = levels.size();
    memcpy(p, &cnt, 8);
    memcpy(p + 8, &mn, 8);
    memcpy(p + 16, &mx, 8);
    memcpy(p + 24, &k, 8);
    memcpy(p + 32, &cap, 8);
    memcpy(p + 40, &nl, 8);
    p += 48;
    for (auto& l : levels) {
        const uint64_t c = l.size();
        memcpy(p, &c, 8);
        if (c) memcpy(p + 8, l.data(), c * 8);
        p += 8 + c * 8;
    }
}

// All quantile sketches of a plan in one go: every column's sampler -> sort -> compactor-ladder pipeline is queued back
// to back on the stream (own slice of the scratch block each), ONE synchronisation at the end, then the <= 8k items per
// column are turned into blobs on the host.
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
        c = c->temporal ? nullptr : numeric_view(e, c);  // (Int32 / Float32: the widened shadow)
        if (!c) {
            a.err = TG_ERR_TYPE_MISMATCH;
            a.err_msg = "quantile sketch requires a numeric column";
            continue;
        }
        // sketch capacity: 8k items in total (k = the reference's accuracy parameter)
        const uint64_t cap = (uint64_t)std::min<int64_t>(std::max<int64_t>(8 * (int64_t)a.iparam, 64), (int64_t)1 << 20);
        p.stats.bytes_scanned += (uint64_t)n * 8 + (c->validity.p ? (uint64_t)(n + 7) / 8 : 0);
        if (n == 0) {
            write_kll_blob(a, cap, 0, INFINITY, -INFINITY, {});
            continue;
        }
        jobs.push_back(Job{id, c, cap, 0});
    }
    if (jobs.empty()) return;
    const int64_t n_pairs = (n + 1) / 2;
    // EXACT below KLL_TARGET_SAMPLES rows, otherwise groups of g pairs, 2g = 2^h0 rows -> KLL_TARGET_SAMPLES / 2 .. KLL_TARGET_SAMPLES items
    const bool exact = n <= KLL_TARGET_SAMPLES;
    uint32_t g = 1, h0 = 0;
    if (!exact) {
        h0 = 1;
        while (((int64_t)g << 1) * KLL_TARGET_SAMPLES < n) {  // 2g * TARGET >= n
            g <<= 1;
            ++h0;
        }
    }
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_pairs + KLL_THREADS - 1) / KLL_THREADS, (int64_t)e.sm_count * KLL_CTAS_PER_SM));
    const int64_t T = (int64_t)grid * KLL_THREADS;
    const int64_t pairs_per_thread = (n_pairs + T - 1) / T;
    const uint32_t max_emit = (uint32_t)(exact ? 2 * pairs_per_thread : pairs_per_thread / g + 1);
    const int64_t m = T * (int64_t)max_emit;  // item slots (unused ones keep the all-ones key)
    const int n_passes = exact ? 8 : 4;        // see below
    const size_t k_b = round_up((size_t)m * 8, 256), tmp_b = round_up(rs_temp_bytes(m, n_passes), 256);
    const size_t lad_b = round_up(sizeof(KllLadder), 256);
    size_t per_job = 0, host_total = 0;
    for (auto& j : jobs) {
        const size_t out_b = lad_b + round_up((size_t)j.cap * 8, 256) + 256;
        per_job = std::max(per_job, 2 * k_b + tmp_b + out_b + 256);
        j.host_off = host_total;
        host_total += out_b;
    }
    uint8_t* scr = e.scratch(per_job * jobs.size());
    uint8_t* hs = e.host_scratch(host_total);
    // The samplers run back to back on the engine's stream at HBM speed; a column's sort / ladder / copy-out is a chain
    // of small launch-latency-bound kernels, so it goes to a side stream and overlaps the next column's sampler.
    e.ensure_side_streams();
    TG_CUDA(cudaEventRecord(e.ev[6], e.stream));
    for (size_t ji = 0; ji < jobs.size(); ++ji) {
        Job& j = jobs[ji];
        cudaStream_t side = e.side[ji % 4];
        const size_t top_b = round_up((size_t)j.cap * 8, 256);
        uint8_t* q = scr + ji * per_job;
        uint64_t* kb[2];
        kb[0] = (uint64_t*)q; q += k_b;
        kb[1] = (uint64_t*)q; q += k_b;
        uint8_t* d_tmp = q; q += tmp_b;
        KllLadder* d_lad = (KllLadder*)q; q += lad_b;
        double* d_top = (double*)q; q += top_b;
        KllCounters* d_ctr = (KllCounters*)q;
        KllCounters* h_ctr = (KllCounters*)(hs + j.host_off + lad_b + top_b);  // pinned: also the landing place of the result
        *h_ctr = KllCounters{0, 0, ~0ull, 0ull};
        TG_CUDA(cudaMemcpyAsync(d_ctr, h_ctr, sizeof(KllCounters), cudaMemcpyHostToDevice, e.stream));
        TG_CUDA(cudaMemsetAsync(kb[0], 0xff, k_b, e.stream));
        TG_CUDA(cudaMemsetAsync(d_lad, 0, lad_b, e.stream));
        const uint32_t seed = 0x5eed0000u + (uint32_t)j.agg;
        typedef void (*Sampler)(const uint8_t*, const uint32_t*, int64_t, uint32_t, uint32_t, uint64_t*, uint32_t, KllCounters*);
        const bool i64 = j.c->dtype == TG_INT64;
        const Sampler sampler = i64 ? (exact ? (Sampler)kll_sample_kernel<true, true> : (Sampler)kll_sample_kernel<true, false>)
                                    : (exact ? (Sampler)kll_sample_kernel<false, true> : (Sampler)kll_sample_kernel<false, false>);
        sampler<<<grid, KLL_THREADS, 0, e.stream>>>(j.c->values.p, (const uint32_t*)j.c->validity.p, n, g, seed, kb[0], max_emit, d_ctr);
        TG_CUDA(cudaGetLastError());
        // exact mode sorts on all 64 key bits; a sampled level only needs the order down to the top 32 bits (sign,
        // exponent, 20 mantissa bits: items closer than 1e-6 relative may swap, far inside the rank-error bound),
        // which halves the radix passes — each is launch-latency bound at this size
        TG_CUDA(cudaEventRecord(e.side_ev[ji % 4], e.stream));
        TG_CUDA(cudaStreamWaitEvent(side, e.side_ev[ji % 4], 0));
        const RsTemp RT = rs_temp_carve(d_tmp, m, n_passes);
        uint64_t* no_vals[2] = {nullptr, nullptr};
        int launches = 1 + rs_sort_pairs<uint64_t>(side, kb, no_vals, m, exact ? 0 : 32, n_passes, false, RT, e.sm_count);
        TG_CUDA(cudaGetLastError());
        kll_compact_kernel<<<(unsigned)((j.cap + 255) / 256), 256, 0, side>>>(RT.ctl, kb[0], kb[1], d_ctr, h0, (uint32_t)j.cap, seed, d_lad, d_top);
        TG_CUDA(cudaGetLastError());
        ++launches;
        // the D2H of the counters must not race with the H2D that initialised them from the same pinned slot: the side
        // stream waited for the engine's stream above
        TG_CUDA(cudaMemcpyAsync(hs + j.host_off, d_lad, lad_b + top_b, cudaMemcpyDeviceToHost, side));
        TG_CUDA(cudaMemcpyAsync(h_ctr, d_ctr, sizeof(KllCounters), cudaMemcpyDeviceToHost, side));
        p.stats.launches += launches;
        e.launches += launches;
    }
    // join: the engine's stream continues only after every side chain (a side stream is reused by every 4th column in order)
    for (int k = 0; k < 4 && k < (int)jobs.size(); ++k) {
        TG_CUDA(cudaEventRecord(e.side_ev[k], e.side[k]));
        TG_CUDA(cudaStreamWaitEvent(e.stream, e.side_ev[k], 0));
    }
    TG_CUDA(cudaEventRecord(e.ev[7], e.stream));
    TG_CUDA(cudaStreamSynchronize(e.stream));
    float ms = 0;
    TG_CUDA(cudaEventElapsedTime(&ms, e.ev[6], e.ev[7]));
    p.stats.sketch_ms += ms;
    p.stats.gpu_ms += ms;
    for (auto& j : jobs) {
        Agg& a = p.aggs[j.agg];
        const size_t top_b = round_up((size_t)j.cap * 8, 256);
        const KllLadder& lad = *(const KllLadder*)(hs + j.host_off);
        const double* h_top = (const double*)(hs + j.host_off + lad_b);
        const KllCounters h = *(const KllCounters*)(hs + j.host_off + lad_b + top_b);
        a.u[0] = h.n;
        if (h.n == 0) {
            write_kll_blob(a, j.cap, 0, INFINITY, -INFINITY, {});
            continue;
        }
        std::vector<std::vector<double>> levels(lad.top_count ? lad.top_level + 1 : 0);
        for (uint32_t i = 0; i < lad.n_left; ++i) {
            if (lad.left_level[i] >= levels.size()) levels.resize(lad.left_level[i] + 1);
            levels[lad.left_level[i]].push_back(lad.left_val[i]);
        }
        if (lad.top_count) levels[lad.top_level].assign(h_top, h_top + lad.top_count);
        write_kll_blob(a, j.cap, h.n, key_f64(h.min_bits), key_f64(h.max_bits), levels);
    }
}

}  // namespace tg
