// Host side of the KLL quantile sketch (K4): state layout, merge and query.
//
// The state is the reference's KllSketch (analyzers/advanced/kll_sketch.rs:142-160): n, min_value, max_value and a
// stack of compactors, the items of level h standing for 2^h values each. The device fills it (sketch.cu: sampler for
// the bottom levels, sort, compactor ladder); this file merges states and answers queries:
//   merge  (:327-366)  per level: concatenate the items; then, while the sketch holds more than its capacity, the
//                      lowest level with at least two items is compacted: sorted, neighbours paired, the odd or the even
//                      one of every pair (a counter-based coin) promoted one level up, an unpaired last item stays. This
//                      is Compactor::compact (:57-76) with the weight kept exact — the reference keeps BOTH halves (one
//                      at the old level, one promoted), which inflates its total weight; the rank-error contract
//                      (1.65 / sqrt(k), :397-399) is what the tests pin, and a weight-preserving ladder meets it with a
//                      wide margin (capacity = 8k items: every compaction moves a rank by at most one item weight).
//   query  (:246-322)  all items with weight 2^level, in value order; the first whose cumulative weight reaches
//                      ceil(phi * W); phi = 0 / 1 answer the exact min / max (:260-265).
//
// blob: u64 n | f64 min | f64 max | u64 k | u64 capacity | u64 n_levels | per level: u64 count, count x f64 (ascending)
#include <algorithm>
#include <cmath>
#include <cstring>

#include "plan.hpp"

namespace tg {

struct KllHost {
    uint64_t n = 0;
    double mn = INFINITY, mx = -INFINITY;
    uint64_t k = 0, cap = 0;
    std::vector<std::vector<double>> levels;
    size_t items() const {
        size_t c = 0;
        for (auto& l : levels) c += l.size();
        return c;
    }
};

static bool kll_parse(const std::vector<uint8_t>& b, KllHost& s) {
    if (b.size() < 48) return false;
    uint64_t nl;
    memcpy(&s.n, b.data(), 8);
    memcpy(&s.mn, b.data() + 8, 8);
    memcpy(&s.mx, b.data() + 16, 8);
    memcpy(&s.k, b.data() + 24, 8);
    memcpy(&s.cap, b.data() + 32, 8);
    memcpy(&nl, b.data() + 40, 8);
    if (nl > 64) return false;
    size_t off = 48;
    s.levels.assign(nl, {});
    for (uint64_t l = 0; l < nl; ++l) {
        if (off + 8 > b.size()) return false;
        uint64_t c;
        memcpy(&c, b.data() + off, 8);
        off += 8;
        if (c > (b.size() - off) / 8) return false;
        s.levels[l].resize(c);
        if (c) memcpy(s.levels[l].data(), b.data() + off, c * 8);
        off += c * 8;
    }
    return true;
}

static void kll_write(const KllHost& s, std::vector<uint8_t>& b) {
    size_t bytes = 48;
    for (auto& l : s.levels) bytes += 8 + l.size() * 8;
    b.resize(bytes);
    uint8_t* p = b.data();
    const uint64_t nl = s.levels.size();
    memcpy(p, &s.n, 8);
    memcpy(p + 8, &s.mn, 8);
    memcpy(p + 16, &s.mx, 8);
    memcpy(p + 24, &s.k, 8);
    memcpy(p + 32, &s.cap, 8);
    memcpy(p + 40, &nl, 8);
    p += 48;
    for (auto& l : s.levels) {
        const uint64_t c = l.size();
        memcpy(p, &c, 8);
        if (c) memcpy(p + 8, l.data(), c * 8);
        p += 8 + c * 8;
    }
}

static uint32_t kll_host_coin(uint64_t n, uint32_t level, uint64_t count) {
    uint64_t x = n * 0x9e3779b97f4a7c15ull ^ ((uint64_t)level << 32) ^ count;
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return (uint32_t)(x & 1u);
}

// cascade_compact (kll_sketch.rs:219-236) for a merged state: levels are kept sorted
static void kll_compact(KllHost& s) {
    if (s.cap == 0) return;
    while (s.items() > s.cap) {
        size_t lv = 0;
        while (lv < s.levels.size() && s.levels[lv].size() < 2) ++lv;
        if (lv == s.levels.size()) return;  // nothing left to pair (cannot happen while items > cap >= 64)
        if (lv + 1 == s.levels.size()) s.levels.emplace_back();
        std::vector<double>& cur = s.levels[lv];
        const size_t pairs = cur.size() / 2;
        const uint32_t coin = kll_host_coin(s.n, (uint32_t)lv, cur.size());
        std::vector<double> promoted(pairs);
        for (size_t i = 0; i < pairs; ++i) promoted[i] = cur[2 * i + coin];
        std::vector<double> left;
        if (cur.size() & 1) left.push_back(cur.back());
        cur.swap(left);
        std::vector<double>& up = s.levels[lv + 1];
        std::vector<double> merged(up.size() + promoted.size());
        std::merge(up.begin(), up.end(), promoted.begin(), promoted.end(), merged.begin());
        up.swap(merged);
    }
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
    a.k = std::max(a.k, b.k);
    a.cap = std::max(a.cap, b.cap);
    if (b.levels.size() > a.levels.size()) a.levels.resize(b.levels.size());
    for (size_t l = 0; l < b.levels.size(); ++l) {
        std::vector<double> merged(a.levels[l].size() + b.levels[l].size());
        std::merge(a.levels[l].begin(), a.levels[l].end(), b.levels[l].begin(), b.levels[l].end(), merged.begin());
        a.levels[l].swap(merged);
    }
    kll_compact(a);
    kll_write(a, into);
}

// kll_sketch.rs:246-322
bool kll_blob_query(const std::vector<uint8_t>& blob, double phi, double* out) {
    KllHost s;
    if (!kll_parse(blob, s) || s.n == 0) return false;
    if (phi == 0.0) {
        *out = s.mn;
        return true;
    }
    if (phi == 1.0) {
        *out = s.mx;
        return true;
    }
    struct Item {
        double v;
        uint64_t w;
    };
    std::vector<Item> items;
    items.reserve(s.items());
    for (size_t l = 0; l < s.levels.size(); ++l) {
        const uint64_t w = l >= 63 ? (UINT64_MAX / 2) : ((uint64_t)1 << l);
        for (double v : s.levels[l]) items.push_back(Item{v, w});
    }
    if (items.empty()) return false;
    std::stable_sort(items.begin(), items.end(), [](const Item& x, const Item& y) { return x.v < y.v; });
    uint64_t W = 0;
    for (auto& it : items) W = W + it.w < W ? UINT64_MAX : W + it.w;  // saturating, like the reference
    const double target = std::ceil(phi * (double)W);
    uint64_t cum = 0;
    for (auto& it : items) {
        cum = cum + it.w < cum ? UINT64_MAX : cum + it.w;
        if ((double)cum >= target) {
            *out = it.v;
            return true;
        }
    }
    *out = s.mx;
    return true;
}

void kll_blob_summary(const std::vector<uint8_t>& blob, uint64_t* n, double* mn, double* mx) {
    KllHost s;
    if (!kll_parse(blob, s)) {
        *n = 0;
        *mn = 0;
        *mx = 0;
        return;
    }
    *n = s.n;
    *mn = s.mn;
    *mx = s.mx;
}

// KllSketch's fields for a host that wants to hand the state on (tg_plan_kll_levels): number of levels, and per level
// the item count / the items
int kll_blob_levels(const std::vector<uint8_t>& blob, std::vector<std::vector<double>>& levels, uint64_t* k) {
    KllHost s;
    if (!kll_parse(blob, s)) return -1;
    levels = s.levels;
    if (k) *k = s.k;
    return (int)s.levels.size();
}

}  // namespace tg
