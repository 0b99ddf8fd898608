// Host side of the quantile sketch (K4): blob layout, merge and query.
//
// The reference's KllSketch (analyzers/advanced/kll_sketch.rs:142-400) keeps per-level buffers whose
// items weigh 2^level and answers get_quantile(phi) with the first item, in value order, whose
// cumulative weight reaches ceil(phi * W) (:246-322); phi = 0 / 1 return the exact min / max (:260-265);
// merge concatenates and re-compacts (:327-366). The device sketch is a weighted, value-sorted item list
// with the same query rule; compaction is "sort, then keep every j-th item (systematic resample)",
// which is what a KLL compactor does to a sorted buffer, so the same rank-error argument applies and
// the bound 1.65/sqrt(k) (:397-399) holds with a wide margin (capacity = 8k items).
//
// blob: u64 n | f64 min | f64 max | u64 capacity | u64 m | m x { f64 value, u64 weight } (sorted by value)
#include <algorithm>
#include <cmath>
#include <cstring>

#include "plan.hpp"

namespace tg {

struct KllItem {
    double v;
    uint64_t w;
};
struct KllHost {
    uint64_t n = 0;
    double mn = INFINITY, mx = -INFINITY;
    uint64_t cap = 0;
    std::vector<KllItem> items;
};

static bool kll_parse(const std::vector<uint8_t>& b, KllHost& k) {
    if (b.size() < 40) return false;
    uint64_t m;
    memcpy(&k.n, b.data(), 8);
    memcpy(&k.mn, b.data() + 8, 8);
    memcpy(&k.mx, b.data() + 16, 8);
    memcpy(&k.cap, b.data() + 24, 8);
    memcpy(&m, b.data() + 32, 8);
    if (b.size() < 40 + m * 16) return false;
    k.items.resize(m);
    if (m) memcpy(k.items.data(), b.data() + 40, m * 16);
    return true;
}
static void kll_write(const KllHost& k, std::vector<uint8_t>& b) {
    uint64_t m = k.items.size();
    b.resize(40 + m * 16);
    memcpy(b.data(), &k.n, 8);
    memcpy(b.data() + 8, &k.mn, 8);
    memcpy(b.data() + 16, &k.mx, 8);
    memcpy(b.data() + 24, &k.cap, 8);
    memcpy(b.data() + 32, &m, 8);
    if (m) memcpy(b.data() + 40, k.items.data(), m * 16);
}

// systematic resample of a value-sorted weighted list down to `cap` items of (almost) equal weight
static void kll_compact(KllHost& k) {
    if (k.cap == 0 || k.items.size() <= k.cap) return;
    uint64_t W = 0;
    for (auto& it : k.items) W += it.w;
    const uint64_t M = k.cap;
    std::vector<KllItem> out;
    out.reserve(M);
    // output i represents ranks (i*W/M, (i+1)*W/M]; pick the item holding the midpoint rank
    size_t j = 0;
    uint64_t cum = k.items.empty() ? 0 : k.items[0].w;
    for (uint64_t i = 0; i < M; ++i) {
        const uint64_t lo = (uint64_t)((__uint128_t)i * W / M), hi = (uint64_t)((__uint128_t)(i + 1) * W / M);
        if (hi == lo) continue;
        const uint64_t target = lo + (hi - lo + 1) / 2;  // 1-based rank
        while (cum < target && j + 1 < k.items.size()) cum += k.items[++j].w;
        out.push_back(KllItem{k.items[j].v, hi - lo});
    }
    k.items.swap(out);
}

void kll_blob_merge(std::vector<uint8_t>& into, const std::vector<uint8_t>& other) {
    KllHost a, b;
    const bool ha = kll_parse(into, a), hb = kll_parse(other, b);
    if (!hb) return;
    if (!ha) {
        into = other;
        return;
    }
    a.n += b.n;
    a.mn = std::fmin(a.mn, b.mn);
    a.mx = std::fmax(a.mx, b.mx);
    a.cap = std::max(a.cap, b.cap);
    std::vector<KllItem> merged(a.items.size() + b.items.size());
    std::merge(a.items.begin(), a.items.end(), b.items.begin(), b.items.end(), merged.begin(),
               [](const KllItem& x, const KllItem& y) { return x.v < y.v; });
    a.items.swap(merged);
    kll_compact(a);
    kll_write(a, into);
}

// kll_sketch.rs:246-322
bool kll_blob_query(const std::vector<uint8_t>& blob, double phi, double* out) {
    KllHost k;
    if (!kll_parse(blob, k) || k.n == 0) return false;
    if (phi == 0.0) {
        *out = k.mn;
        return true;
    }
    if (phi == 1.0) {
        *out = k.mx;
        return true;
    }
    if (k.items.empty()) return false;
    uint64_t W = 0;
    for (auto& it : k.items) W += it.w;
    const double target = std::ceil(phi * (double)W);
    uint64_t cum = 0;
    for (auto& it : k.items) {
        cum += it.w;
        if ((double)cum >= target) {
            *out = it.v;
            return true;
        }
    }
    *out = k.mx;
    return true;
}

void kll_blob_summary(const std::vector<uint8_t>& blob, uint64_t* n, double* mn, double* mx) {
    KllHost k;
    if (!kll_parse(blob, k)) {
        *n = 0;
        *mn = 0;
        *mx = 0;
        return;
    }
    *n = k.n;
    *mn = k.mn;
    *mx = k.mx;
}

}  // namespace tg
