"""CPU oracle — TEST INFRASTRUCTURE ONLY. Line-by-line restatement of the reference's in-tree KllSketch
(term-guard/src/analyzers/advanced/kll_sketch.rs:15-400), including its deterministic compaction coin
(`select_compaction_strategy`, :78-102, non-`test-utils` build): SipHash-1-3 with zero keys (Rust's
`DefaultHasher`) over `(items.len() as usize, items[0] as u64)`, keep-odd iff hash is odd.

Note the reference sketch is NOT weight preserving: `compact` (:57-76) keeps one half at the same level
and promotes the other half with doubled weight, so its total weight drifts above n; its own tests
accept large value errors (:459-468). The product's contract is therefore the rank-error bound
1.65/sqrt(k) (:397-399) against exact quantiles, not bit parity with this sketch.
"""
import math

MASK = (1 << 64) - 1


def _rotl(x, b):
    return ((x << b) | (x >> (64 - b))) & MASK


def siphash13(words):
    """SipHash-1-3, k0 = k1 = 0, over a message made of 8-byte little-endian words."""
    v0, v1, v2, v3 = 0x736F6D6570736575, 0x646F72616E646F6D, 0x6C7967656E657261, 0x7465646279746573

    def rnd(v0, v1, v2, v3):
        v0 = (v0 + v1) & MASK
        v1 = _rotl(v1, 13) ^ v0
        v0 = _rotl(v0, 32)
        v2 = (v2 + v3) & MASK
        v3 = _rotl(v3, 16) ^ v2
        v0 = (v0 + v3) & MASK
        v3 = _rotl(v3, 21) ^ v0
        v2 = (v2 + v1) & MASK
        v1 = _rotl(v1, 17) ^ v2
        v2 = _rotl(v2, 32)
        return v0, v1, v2, v3

    for m in words:
        v3 ^= m
        v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
        v0 ^= m
    b = ((len(words) * 8) & 0xFF) << 56
    v3 ^= b
    v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
    v0 ^= b
    v2 ^= 0xFF
    for _ in range(3):
        v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
    return v0 ^ v1 ^ v2 ^ v3


def _f64_as_u64(x: float) -> int:
    """Rust `as u64`: saturating, NaN -> 0"""
    if math.isnan(x) or x <= 0:
        return 0
    if x >= 18446744073709551615.0:
        return MASK
    return int(x)


class Compactor:  # kll_sketch.rs:15-118
    def __init__(self, capacity):
        self.capacity, self.items = capacity, []

    def is_full(self):
        return len(self.items) >= self.capacity

    def compact(self):
        self.items.sort()
        words = [len(self.items)]
        if self.items:
            words.append(_f64_as_u64(self.items[0]))
        keep_odd = (siphash13(words) % 2) == 1
        kept = [x for i, x in enumerate(self.items) if (i % 2 == 1) == keep_odd]
        compacted = [x for i, x in enumerate(self.items) if (i % 2 == 1) != keep_odd]
        self.items = kept
        return compacted


class KllSketch:  # kll_sketch.rs:142-400
    def __init__(self, k):
        if k < 2:
            raise ValueError("k must be at least 2")
        self.k, self.compactors, self.n = k, [Compactor(k)], 0
        self.min_value, self.max_value = math.inf, -math.inf

    def level_capacity(self, level):  # :183-192
        k = self.k
        return [k, max(8, (k * 2) // 3), max(4, k // 2), max(4, k // 4), max(4, k // 8)][level] if level < 5 else 4

    def update(self, value):  # :195-210
        if math.isnan(value):
            return
        self.n += 1
        self.min_value = min(self.min_value, value)
        self.max_value = max(self.max_value, value)
        self.compactors[0].items.append(value)
        level = 0
        while level < len(self.compactors) and self.compactors[level].is_full():  # cascade_compact :212-229
            if level + 1 >= len(self.compactors):
                self.compactors.append(Compactor(self.level_capacity(level + 1)))
            self.compactors[level + 1].items.extend(self.compactors[level].compact())
            level += 1

    def get_quantile(self, phi):  # :246-322
        if self.n == 0:
            raise ValueError("Cannot compute quantile on empty sketch")
        if not (0.0 <= phi <= 1.0):
            raise ValueError("phi")
        if phi == 0.0:
            return self.min_value
        if phi == 1.0:
            return self.max_value
        weighted = []
        for level, c in enumerate(self.compactors):
            w = (MASK // 2) if level >= 63 else (1 << level)
            weighted.extend((x, w) for x in sorted(c.items))
        weighted.sort(key=lambda t: t[0])
        total = min(sum(w for _, w in weighted), MASK)
        target = math.ceil(phi * float(total))
        cum = 0
        for v, w in weighted:
            cum = min(cum + w, MASK)
            if float(cum) >= target:
                return v
        return self.max_value

    def merge(self, other):  # :327-366
        if self.k != other.k:
            raise ValueError("Cannot merge sketches with different k values")
        self.n += other.n
        self.min_value = min(self.min_value, other.min_value)
        self.max_value = max(self.max_value, other.max_value)
        for level, oc in enumerate(other.compactors):
            while level >= len(self.compactors):
                self.compactors.append(Compactor(self.level_capacity(level)))
            self.compactors[level].items.extend(oc.items)
        for level in range(len(self.compactors)):  # Rust evaluates `0..len` once (:351)
            while self.compactors[level].is_full():
                if level + 1 >= len(self.compactors):
                    self.compactors.append(Compactor(self.level_capacity(level + 1)))
                self.compactors[level + 1].items.extend(self.compactors[level].compact())

    def count(self):
        return self.n

    def relative_error_bound(self):  # :397-399
        return 1.65 / math.sqrt(self.k)
