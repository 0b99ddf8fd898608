// K6 — Spearman rank correlation: RANK() OVER (ORDER BY x) (competition / minimum ranks) for both
// columns over the pairwise-complete rows, then the Pearson co-moments of the ranks
// (analyzers/advanced/correlation.rs:334-350, metric :407-427).
//
// Pipeline (no library code, no random access to HBM):
//   1. rk_count / rk_offsets / rk_compact_keys   the pairwise-complete rows, compacted in row order (deterministic),
//                                                as order-preserving 64-bit keys (kx, ky)
//   2. sort by kx carrying ky as the payload     hand-written onesweep LSD radix sort (radix_sort.cu); digit positions on
//                                                which every key agrees (sign / exponent bytes) are skipped
//   3. rk_tile_heads / rk_tile_scan / rk_rank_x  the minimum rank of a sorted position is 1 + the position of the first
//                                                key of its run: per-tile "last run head", a one-block max-scan of the
//                                                tile results, an in-tile max-scan with the carry. Emits (ky, rank_x).
//   4. sort by ky carrying rank_x
//   5. rk_tile_heads / rk_tile_scan / rk_rank_y_moments   rank_y the same way, and the shifted co-moments of
//                                                (rank_x, rank_y) accumulated on the fly in a fixed order
// Across GPUs the same stages run as a sample sort: each sort is preceded by an all-to-all that brings
// every key range to one rank (splitters from evenly spaced samples of the locally sorted shards), and the run heads
// become GLOBAL ranks by adding the number of keys on the lower ranks (tg_rank_* entry points, distributed.py).
// Ranks travel with the rows through the two sorts instead of being scattered back by row id (a random 4-byte write
// per row costs more than a whole sort pass).
// The reference accumulates rank products in UInt64 and overflows above ~3.8M rows; here
// ranks are exact integers carried as f64 and the sums are centred at (n+1)/2.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "engine.hpp"
#include "ptx.cuh"
#include "radix_sort.cuh"

namespace tg {

constexpr int RK_THREADS = 256;

__device__ __forceinline__ uint64_t order_key(uint64_t bits, int is_i64) {
    if (is_i64) return bits ^ 0x8000000000000000ull;
    // canonical -0.0 -> +0.0 so they tie like SQL equality
    if ((bits << 1) == 0) bits = 0;
    return (bits & 0x8000000000000000ull) ? ~bits : (bits | 0x8000000000000000ull);
}

// ---- 1. compaction of the pairwise-complete rows. A block owns RK_CTILE = 8192 consecutive rows: thread t the 32 rows
// of validity word t of the tile.
constexpr int RK_CTILE = RK_THREADS * 32;

__device__ __forceinline__ uint32_t rk_pair_word(const uint32_t* __restrict__ vx, const uint32_t* __restrict__ vy, int64_t word, int64_t n) {
    const int64_t row0 = word * 32;
    if (row0 >= n) return 0u;
    uint32_t w = 0xffffffffu;
    if (vx) w &= __ldg(vx + word);
    if (vy) w &= __ldg(vy + word);
    if (n - row0 < 32) w &= (1u << (uint32_t)(n - row0)) - 1u;
    return w;
}

__global__ void __launch_bounds__(RK_THREADS) rk_count_kernel(const uint32_t* __restrict__ vx, const uint32_t* __restrict__ vy, int64_t n,
                                                              uint32_t* __restrict__ tile_count) {
    const int64_t word = (int64_t)blockIdx.x * RK_THREADS + threadIdx.x;
    uint32_t c = (uint32_t)__popc(rk_pair_word(vx, vy, word, n));
    __shared__ uint32_t red[RK_THREADS / 32];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) c += __shfl_xor_sync(0xffffffffu, c, m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int w = 0; w < RK_THREADS / 32; ++w) s += red[w];
        tile_count[blockIdx.x] = s;
    }
}

// one block: exclusive sum of the tile counts; total -> *n_pairs
__global__ void __launch_bounds__(1024) rk_offsets_kernel(const uint32_t* __restrict__ tile_count, int64_t n_tiles, uint32_t* __restrict__ tile_off,
                                                          unsigned long long* n_pairs) {
    __shared__ unsigned long long s_part[1024];
    const int64_t per = (n_tiles + 1023) / 1024;
    const int64_t lo = (int64_t)threadIdx.x * per, hi = lo + per < n_tiles ? lo + per : n_tiles;
    unsigned long long m = 0;
    for (int64_t t = lo; t < hi; ++t) m += tile_count[t];
    s_part[threadIdx.x] = m;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned long long y = (int)threadIdx.x >= o ? s_part[threadIdx.x - o] : 0ull;
        __syncthreads();
        s_part[threadIdx.x] += y;
        __syncthreads();
    }
    unsigned long long run = threadIdx.x ? s_part[threadIdx.x - 1] : 0ull;
    for (int64_t t = lo; t < hi; ++t) {
        tile_off[t] = (uint32_t)run;
        run += tile_count[t];
    }
    if (threadIdx.x == 1023) *n_pairs = s_part[1023];
}

// warp w of the block owns validity words w*32 .. w*32+31 of the tile (1024 rows): for each word the 32 lanes read its 32
// rows coalesced and the complete ones are written behind each other — row order is kept, so the result is deterministic
__global__ void __launch_bounds__(RK_THREADS) rk_compact_keys_kernel(const uint64_t* __restrict__ x, const uint32_t* __restrict__ vx, int x_i64,
                                                                     const uint64_t* __restrict__ y, const uint32_t* __restrict__ vy, int y_i64,
                                                                     int64_t n, const uint32_t* __restrict__ tile_off,
                                                                     uint64_t* __restrict__ kx, uint64_t* __restrict__ ky) {
    __shared__ uint32_t s_warp[RK_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t word = (int64_t)blockIdx.x * RK_THREADS + threadIdx.x;
    const uint32_t w = rk_pair_word(vx, vy, word, n);
    // exclusive scan of popc(w) over the block
    const uint32_t c = (uint32_t)__popc(w);
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t base = tile_off[blockIdx.x];
    for (int i = 0; i < warp; ++i) base += s_warp[i];
    const uint32_t my_off = base + incl - c;  // first output slot of this lane's word
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
        const uint32_t wj = __shfl_sync(0xffffffffu, w, j);
        const uint32_t oj = __shfl_sync(0xffffffffu, my_off, j);
        if (wj == 0u) continue;  // warp-uniform
        if ((wj >> lane) & 1u) {
            const int64_t r = (word - lane + j) * 32 + lane;
            const uint32_t dst = oj + (uint32_t)__popc(wj & lt);
            kx[dst] = order_key(__ldg(x + r), x_i64);
            ky[dst] = order_key(__ldg(y + r), y_i64);
        }
    }
}

// ---- 3 / 5. minimum ranks from sorted keys. A position's run head = the largest p' <= p with key[p'] != key[p' - 1]
// (p' = 0 counts); stored 1-based so that 0 means "no head in this range" and the combine is a plain max.
constexpr int RK_ITEMS = 8, RK_TILE = RK_THREADS * RK_ITEMS;

// tile_last[t] = 1-based position of the last run head inside tile t (0: the tile starts inside a run and never leaves it)
// what the runs are runs of: the whole key after a full sort, rs_quant(key) after a quantised one
struct RkRunKey {
    bool quantised;
    RsQuant Q;
    __device__ __forceinline__ RkRunKey(const RsControl* ctl, const RsQuant* quant) : quantised(ctl->quantised != 0), Q{} {
        if (quantised) Q = *quant;
    }
    __device__ __forceinline__ uint64_t operator()(uint64_t k) const { return quantised ? (uint64_t)rs_quant(Q, k) : k; }
};

__global__ void __launch_bounds__(RK_THREADS) rk_tile_heads_kernel(const RsControl* ctl, const RsQuant* quant, const uint64_t* k0, const uint64_t* k1,
                                                                   int64_t n, uint32_t* tile_last) {
    const uint64_t* __restrict__ ks = ctl->result ? k1 : k0;
    const RkRunKey rk(ctl, quant);
    const int64_t base = (int64_t)blockIdx.x * RK_TILE;
    uint32_t best = 0;
#pragma unroll 4
    for (int i = 0; i < RK_ITEMS; ++i) {
        const int64_t p = base + (int64_t)i * RK_THREADS + threadIdx.x;
        if (p < n && (p == 0 || rk(ks[p]) != rk(ks[p - 1]))) best = (uint32_t)p + 1u;  // positions grow with i
    }
    __shared__ uint32_t red[RK_THREADS / 32];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, m));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t b = 0;
        for (int w = 0; w < RK_THREADS / 32; ++w) b = max(b, red[w]);
        tile_last[blockIdx.x] = b;
    }
}

// exclusive max-scan over the tiles (one block): carry[t] = last head before tile t
__global__ void __launch_bounds__(1024) rk_tile_scan_kernel(const uint32_t* __restrict__ tile_last, int64_t n_tiles, uint32_t* __restrict__ carry) {
    __shared__ uint32_t s_part[1024];
    const int64_t per = (n_tiles + 1023) / 1024;
    const int64_t lo = (int64_t)threadIdx.x * per, hi = lo + per < n_tiles ? lo + per : n_tiles;
    uint32_t m = 0;
#pragma unroll 8
    for (int64_t t = lo; t < hi; ++t) m = max(m, tile_last[t]);
    s_part[threadIdx.x] = m;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // inclusive max-scan of the 1024 partials (Hillis-Steele)
        const uint32_t y = (int)threadIdx.x >= o ? s_part[threadIdx.x - o] : 0u;
        __syncthreads();
        s_part[threadIdx.x] = max(s_part[threadIdx.x], y);
        __syncthreads();
    }
    uint32_t run = threadIdx.x ? s_part[threadIdx.x - 1] : 0u;
#pragma unroll 8
    for (int64_t t = lo; t < hi; ++t) {
        carry[t] = run;
        run = max(run, tile_last[t]);
    }
}

// the ranks of one tile of sorted keys: warp-striped positions (coalesced), run heads by comparing with the left
// neighbour, inclusive max-scan across the warp for every item row, carried from row to row and from the earlier warps
// (shared memory) / earlier tiles (carry). rank[i] of position warp_base + i * 32 + lane.
struct RkTileRanks {
    uint32_t rank[RK_ITEMS];
};
// After a QUANTISED sort (RsControl::quantised) the runs are runs of equal rs_quant(key), inside which the keys are in
// input order: the minimum rank of a key is then (run head) + (keys of its run that are smaller). Runs are short when
// the 32-bit quantisation nearly identifies the key (continuous data: one or two keys), so every position simply walks its run from
// the head: up to RK_RUN_LIMIT keys, neighbours in sorted order (cache hits). A longer run is fine when all its keys are
// equal (duplicates); a position of a longer run that sees a key different from its own within the first RK_RUN_LIMIT keys
// raises *fallback and the host repeats the phase with a full sort (every impure run contains such a position: one whose
// key differs from the run's first key).
constexpr int RK_RUN_LIMIT = 32;
// The tile's keys and their quantised values are staged ONCE in shared memory, with RK_RUN_LIMIT positions of halo on
// both sides: the head test, the singleton test (the common case of continuous data: a run of one) and the run walk read
// the staged copy; only a run that reaches further than the halo (long runs of duplicates) goes back to global memory.
constexpr int RK_HALO = RK_RUN_LIMIT, RK_WIN = RK_TILE + 2 * RK_HALO;
struct RkStage {
    uint64_t k[RK_WIN];
    uint32_t q[RK_WIN];
    uint32_t warp_last[RK_THREADS / 32];
};
__device__ __forceinline__ RkTileRanks rk_tile_ranks(const uint64_t* __restrict__ ks, int64_t n, const uint32_t* __restrict__ carry,
                                                     RkStage& S, const RkRunKey& rk, uint32_t* fallback) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t win0 = (int64_t)blockIdx.x * RK_TILE - RK_HALO;  // global position of S.k[0]
    const bool quantised = rk.quantised;
    for (int o = threadIdx.x; o < RK_WIN; o += RK_THREADS) {
        const int64_t j = win0 + o;
        if (j >= 0 && j < n) {
            const uint64_t k = ks[j];
            S.k[o] = k;
            if (quantised) S.q[o] = rs_quant(rk.Q, k);
        }
    }
    __syncthreads();
    const int obase = RK_HALO + warp * 32 * RK_ITEMS + lane;  // window index of item 0
    RkTileRanks R;
    uint32_t run = 0;  // last head seen so far inside this warp's range (warp-uniform after each row)
#pragma unroll
    for (int i = 0; i < RK_ITEMS; ++i) {
        const int o = obase + i * 32;
        const int64_t p = win0 + o;
        uint32_t h = 0;
        if (p < n) {
            const bool head = p == 0 || (quantised ? S.q[o] != S.q[o - 1] : S.k[o] != S.k[o - 1]);
            if (head) h = (uint32_t)p + 1u;
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, h, d);
            if (lane >= d) h = max(h, v);
        }
        h = max(h, run);
        R.rank[i] = h;
        run = __shfl_sync(0xffffffffu, h, 31);
    }
    if (lane == 0) S.warp_last[warp] = run;
    __syncthreads();
    uint32_t pre = carry[blockIdx.x];
    for (int w = 0; w < warp; ++w) pre = max(pre, S.warp_last[w]);
#pragma unroll
    for (int i = 0; i < RK_ITEMS; ++i) R.rank[i] = max(R.rank[i], pre);
    if (quantised) {
        bool bad = false;
        auto key_at = [&](int64_t j) {
            const int64_t o = j - win0;
            return o >= 0 && o < RK_WIN ? S.k[o] : ks[j];
        };
        auto quant_at = [&](int64_t j) {
            const int64_t o = j - win0;
            return o >= 0 && o < RK_WIN ? S.q[o] : rs_quant(rk.Q, ks[j]);
        };
#pragma unroll
        for (int i = 0; i < RK_ITEMS; ++i) {
            const int o = obase + i * 32;
            const int64_t p = win0 + o;
            if (p >= n) continue;
            const int64_t head = (int64_t)R.rank[i] - 1;  // 0-based first position of the run
            const uint32_t q = S.q[o];
            if (head == p && (p + 1 >= n || S.q[o + 1] != q)) continue;  // a run of one
            const uint64_t k = S.k[o];
            uint32_t less = 0;
            bool differ = false;
            int64_t j = head;
            for (int s = 0; s < RK_RUN_LIMIT && j < n; ++s, ++j) {
                if (quant_at(j) != q) break;
                const uint64_t kj = key_at(j);
                less += kj < k;
                differ |= kj != k;
            }
            const bool ended = j >= n || quant_at(j) != q;
            if (!ended && differ) bad = true;  // a long run with different keys: the prefix order is not enough
            R.rank[i] += less;                   // (long runs of equal keys: less == 0)
        }
        if (bad) atomicExch(fallback, 1u);
    }
    return R;
}

// after the sort by kx (payload ky): emit (ky, rank_base + rank_x) in the sorted-by-x order — the input of the second sort.
// The sorted data sits in buffer ctl->result of the two ping-pong buffers; the output goes to the OTHER buffer pair
// (keys: the 64-bit key buffer, ranks: the payload buffer reused as 32-bit words).
__global__ void __launch_bounds__(RK_THREADS) rk_rank_x_kernel(const RsControl* ctl, const RsQuant* quant, uint64_t* k0, uint64_t* k1, uint64_t* p0,
                                                               uint64_t* p1, int64_t n, const uint32_t* __restrict__ carry, uint32_t rank_base) {
    __shared__ RkStage S;
    const bool r = ctl->result != 0;
    const uint64_t* __restrict__ ks = r ? k1 : k0;
    const uint64_t* __restrict__ ys = r ? p1 : p0;
    uint64_t* __restrict__ out_key = r ? k0 : k1;
    uint32_t* __restrict__ out_rank = reinterpret_cast<uint32_t*>(r ? p0 : p1);
    const RkTileRanks R = rk_tile_ranks(ks, n, carry, S, RkRunKey(ctl, quant), const_cast<uint32_t*>(&ctl->fallback));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t wbase = (int64_t)blockIdx.x * RK_TILE + (int64_t)warp * 32 * RK_ITEMS;
#pragma unroll
    for (int i = 0; i < RK_ITEMS; ++i) {
        const int64_t p = wbase + i * 32 + lane;
        if (p < n) {
            out_key[p] = ys[p];
            out_rank[p] = rank_base + R.rank[i];
        }
    }
}

// after the sort by ky (payload rank_x, 32-bit words in the payload buffers): rank_y on the fly and the block's partial
// sums of the shifted rank co-moments (fixed tile -> block mapping, fixed reduction shape: run-to-run reproducible)
__global__ void __launch_bounds__(RK_THREADS) rk_rank_y_moments_kernel(const RsControl* ctl, const RsQuant* quant, const uint64_t* k0, const uint64_t* k1,
                                                                       const uint32_t* r0, const uint32_t* r1, int64_t n,
                                                                       const uint32_t* __restrict__ carry, uint32_t rank_base, double K,
                                                                       double* __restrict__ partial /* [grid][5] */) {
    __shared__ RkStage S;
    __shared__ double red[5][RK_THREADS / 32];
    const uint64_t* __restrict__ ks = ctl->result ? k1 : k0;
    const uint32_t* __restrict__ rxs = ctl->result ? r1 : r0;
    const RkTileRanks R = rk_tile_ranks(ks, n, carry, S, RkRunKey(ctl, quant), const_cast<uint32_t*>(&ctl->fallback));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t wbase = (int64_t)blockIdx.x * RK_TILE + (int64_t)warp * 32 * RK_ITEMS;
    double s[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < RK_ITEMS; ++i) {
        const int64_t p = wbase + i * 32 + lane;
        if (p < n) {
            const double dx = (double)rxs[p] - K, dy = (double)(rank_base + R.rank[i]) - K;
            s[0] += dx;
            s[1] += dy;
            s[2] = fma(dx, dx, s[2]);
            s[3] = fma(dy, dy, s[3]);
            s[4] = fma(dx, dy, s[4]);
        }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        double v = s[k];
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if (lane == 0) red[k][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double v = 0;
        for (int w = 0; w < RK_THREADS / 32; ++w) v += red[threadIdx.x][w];
        partial[(size_t)blockIdx.x * 5 + threadIdx.x] = v;
    }
}

// smallest / largest key (out = {~0, 0} before): the range the quantised sort spreads its 32 bits over
__global__ void __launch_bounds__(RK_THREADS) rk_minmax_kernel(const uint64_t* __restrict__ keys, int64_t n, unsigned long long* out) {
    uint64_t lo = ~0ull, hi = 0ull;
    for (int64_t i = (int64_t)blockIdx.x * RK_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * RK_THREADS) {
        const uint64_t k = __ldg(keys + i);
        lo = k < lo ? k : lo;
        hi = k > hi ? k : hi;
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        const uint64_t a = shfl_xor_u64(lo, m), b = shfl_xor_u64(hi, m);
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out, (unsigned long long)lo);
        atomicMax(out + 1, (unsigned long long)hi);
    }
}

// ---- pieces of the distributed sort (sample sort): evenly spaced keys of the sorted shard, and the
// first position behind each splitter
__global__ void rk_sample_kernel(const uint64_t* __restrict__ sorted, int64_t n, int32_t m, uint64_t* __restrict__ out) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = sorted[(int64_t)(((__int128)(2 * i + 1) * n) / (2 * (int64_t)m))];
}
__global__ void rk_split_kernel(const uint64_t* __restrict__ sorted, int64_t n, const uint64_t* __restrict__ splitters, int32_t n_split,
                                long long* __restrict__ pos /* [n_split]: number of keys <= splitter */) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_split) return;
    const uint64_t s = splitters[i];
    int64_t lo = 0, hi = n;  // first index with key > s
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted[mid] <= s) lo = mid + 1;
        else hi = mid;
    }
    pos[i] = lo;
}

// fixed-order sum of the block partials: one block per moment, thread t sums the blocks t, t + 1024, .. (four independent
// chains), then a fixed shared-memory tree — the same order on every run
__global__ void __launch_bounds__(1024) rk_final_kernel(const double* __restrict__ partial, int64_t n_blocks, double* out) {
    __shared__ double s_sum[1024];
    const int k = blockIdx.x;
    double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
    int64_t b = threadIdx.x;
    for (; b + 3 * 1024 < n_blocks; b += 4 * 1024) {
        v0 += partial[(size_t)b * 5 + k];
        v1 += partial[(size_t)(b + 1024) * 5 + k];
        v2 += partial[(size_t)(b + 2048) * 5 + k];
        v3 += partial[(size_t)(b + 3072) * 5 + k];
    }
    for (; b < n_blocks; b += 1024) v0 += partial[(size_t)b * 5 + k];
    s_sum[threadIdx.x] = (v0 + v1) + (v2 + v3);
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[k] = s_sum[0];
}

static size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------------------------------
// Rank session: the state of one Spearman evaluation between its stages. The single-GPU job runs the stages back to back;
// the multi-GPU host layer (term_b200/distributed.py, or a Rust shim over the tg_rank_* entry points) puts a sample sort
// between them: local sort -> splitters from evenly spaced samples of every shard -> all-to-all by key range -> the
// receiving rank sorts its range and turns local run heads into GLOBAL minimum ranks with the number of keys on the
// lower ranks as the offset. Equal keys always land on one rank (a key goes to the first part whose splitter is >= it).
// ---------------------------------------------------------------------------------------------------------------------
struct RankArena {
    uint8_t* base = nullptr;
    size_t bytes = 0;
    int64_t cap = 0;
    uint64_t* K[2] = {nullptr, nullptr};  // keys (ping-pong)
    uint64_t* V[2] = {nullptr, nullptr};  // payload (ping-pong): 64-bit words, or 32-bit ranks in the same space
    uint32_t* tile_last = nullptr;
    uint32_t* carry = nullptr;
    double* partial = nullptr;
    uint8_t* tmp = nullptr;               // sort temp + small outputs
    size_t tmp_bytes = 0;
};
struct RankSession {
    RankArena cur, next;
    int64_t n = 0;        // elements in cur
    int which = 0;        // buffer of cur holding the data
    int payload32 = 0;    // payload is 32-bit ranks (second phase)
    int x_i64 = 0, y_i64 = 0;  // the columns' types (how their order keys map back to values)
};

static void arena_free(Engine& e, RankArena& a) {
    if (a.base) e.dev_free(a.base, a.bytes);
    a = RankArena{};
}
static void arena_alloc(Engine& e, RankArena& a, int64_t cap) {
    arena_free(e, a);
    cap = std::max<int64_t>(cap, 1);
    // The ranges a rank receives differ by a few per cent from phase to phase and from step to step; an arena sized
    // exactly would miss the engine's exact-size block cache every time (cudaMalloc + cudaFree of ~4 GB: milliseconds).
    // Arenas therefore take the largest size seen so far, padded by 1/8.
    if (cap > e.rank_cap_hint) e.rank_cap_hint = cap + cap / 8;
    cap = e.rank_cap_hint;
    const size_t k_b = round_up((size_t)cap * 8, 256);
    const int64_t r_tiles = (cap + RK_TILE - 1) / RK_TILE;
    const size_t t_b = round_up((size_t)r_tiles * 4, 256), part_b = round_up((size_t)r_tiles * 40, 256);
    const size_t tmp_b = round_up(rs_temp_bytes(cap, RS_MAX_PASSES), 256) + 4096;
    a.bytes = round_up(4 * k_b + 2 * t_b + part_b + tmp_b, (size_t)1 << 20);  // rounded: the block cache is keyed by size
    a.base = e.dev_alloc(a.bytes);
    a.cap = cap;
    uint8_t* q = a.base;
    a.K[0] = (uint64_t*)q; q += k_b;
    a.K[1] = (uint64_t*)q; q += k_b;
    a.V[0] = (uint64_t*)q; q += k_b;
    a.V[1] = (uint64_t*)q; q += k_b;
    a.tile_last = (uint32_t*)q; q += t_b;
    a.carry = (uint32_t*)q; q += t_b;
    a.partial = (double*)q; q += part_b;
    a.tmp = q;
    a.tmp_bytes = tmp_b;
}

static RankSession& session(Engine& e) {
    if (!e.rank_session) e.rank_session = new RankSession();
    return *e.rank_session;
}
void rank_session_destroy(Engine& e) {
    if (!e.rank_session) return;
    arena_free(e, e.rank_session->cur);
    arena_free(e, e.rank_session->next);
    delete e.rank_session;
    e.rank_session = nullptr;
}

// stage 0: the pairwise-complete rows of (cx, cy) as order-preserving keys: keys = kx, payload = ky. Returns the pair count.
static int64_t rank_begin_locked(Engine& e, Table& t, const std::string& nx, const std::string& ny, int* launches) {
    Column* cx = t.find(nx);
    Column* cy = t.find(ny);
    if (!cx) throw Error(TG_ERR_COLUMN_NOT_FOUND, "Schema error: No field named " + nx + ". Valid fields are " + t.valid_fields() + ".");
    if (!cy) throw Error(TG_ERR_COLUMN_NOT_FOUND, "Schema error: No field named " + ny + ". Valid fields are " + t.valid_fields() + ".");
    cx = cx->temporal ? nullptr : numeric_view(e, cx);  // (Int32 / Float32: the widened shadows)
    cy = cy->temporal ? nullptr : numeric_view(e, cy);
    if (!cx || !cy) throw Error(TG_ERR_TYPE_MISMATCH, "Spearman correlation requires numeric columns");
    const int64_t n = t.n_rows;
    if (n >= (int64_t)1 << 30) throw Error(TG_ERR_UNSUPPORTED, "Spearman: 2^30 or more rows per shard");
    RankSession& S = session(e);
    arena_free(e, S.next);
    arena_alloc(e, S.cur, n);
    S.n = 0;
    S.which = 0;
    S.payload32 = 0;
    S.x_i64 = cx->dtype == TG_INT64;
    S.y_i64 = cy->dtype == TG_INT64;
    if (n == 0) return 0;
    const int64_t c_tiles = (n + RK_CTILE - 1) / RK_CTILE;
    // compaction scratch lives in the arena's temp area
    uint32_t* tile_count = (uint32_t*)e.scratch(2 * round_up((size_t)c_tiles * 4, 256) + 256);
    uint32_t* tile_off = tile_count + round_up((size_t)c_tiles * 4, 256) / 4;
    unsigned long long* d_np = (unsigned long long*)(tile_off + round_up((size_t)c_tiles * 4, 256) / 4);
    const uint32_t* vx = (const uint32_t*)cx->validity.p;
    const uint32_t* vy = (const uint32_t*)cy->validity.p;
    rk_count_kernel<<<(unsigned)c_tiles, RK_THREADS, 0, e.stream>>>(vx, vy, n, tile_count);
    rk_offsets_kernel<<<1, 1024, 0, e.stream>>>(tile_count, c_tiles, tile_off, d_np);
    rk_compact_keys_kernel<<<(unsigned)c_tiles, RK_THREADS, 0, e.stream>>>((const uint64_t*)cx->values.p, vx, cx->dtype == TG_INT64,
                                                                           (const uint64_t*)cy->values.p, vy, cy->dtype == TG_INT64, n, tile_off,
                                                                           S.cur.K[0], S.cur.V[0]);
    TG_CUDA(cudaGetLastError());
    unsigned long long n_pairs = 0;
    TG_CUDA(cudaMemcpyAsync(&n_pairs, d_np, 8, cudaMemcpyDeviceToHost, e.stream));
    TG_CUDA(cudaStreamSynchronize(e.stream));
    *launches += 3;
    S.n = (int64_t)n_pairs;
    return S.n;
}

// queue the sort of the session's current (keys, payload); the result buffer index is only known on the device (ctl)
// where the quantisation of the session's current sort lives (the arena's temp area, behind the sort's own pieces)
static RsQuant* rank_quant(RankArena& A) { return (RsQuant*)(A.tmp + A.tmp_bytes - 1024); }

// quantised: finish_x / finish_y sort 32 value-linear bits (4 passes instead of 8), then fix the short runs of equal
// quantised keys; key_is_i64 says how the keys of this phase came about (the x / y column's type)
static RsTemp rank_sort_queue(Engine& e, RankSession& S, int* launches, bool quantised = false, int key_is_i64 = 0) {
    RankArena& A = S.cur;
    const RsTemp T = rs_temp_carve(A.tmp, std::max<int64_t>(S.n, 1), RS_MAX_PASSES);
    uint64_t* keys[2] = {A.K[S.which], A.K[S.which ^ 1]};
    const RsQuant* quant = nullptr;
    if (quantised) {
        unsigned long long* mm = (unsigned long long*)(A.tmp + A.tmp_bytes - 512);
        TG_CUDA(cudaMemsetAsync(mm, 0xFF, 8, e.stream));
        TG_CUDA(cudaMemsetAsync(mm + 1, 0, 8, e.stream));
        rk_minmax_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>((S.n + RK_THREADS * 8 - 1) / (RK_THREADS * 8), (int64_t)e.sm_count * 8)), RK_THREADS, 0,
                           e.stream>>>(keys[0], S.n, mm);
        rs_quant_setup(e.stream, mm, key_is_i64, rank_quant(A));
        *launches += 2;
        quant = rank_quant(A);
    }
    if (S.payload32) {
        uint32_t* vals[2] = {(uint32_t*)A.V[S.which], (uint32_t*)A.V[S.which ^ 1]};
        *launches += rs_sort_pairs<uint32_t>(e.stream, keys, vals, S.n, 0, RS_MAX_PASSES, false, T, e.sm_count, quant);
    } else {
        uint64_t* vals[2] = {A.V[S.which], A.V[S.which ^ 1]};
        *launches += rs_sort_pairs<uint64_t>(e.stream, keys, vals, S.n, 0, RS_MAX_PASSES, false, T, e.sm_count, quant);
    }
    TG_CUDA(cudaGetLastError());
    return T;
}
// after the queued work: which physical buffer holds the result
// Returns false when the consumer of a prefix-sorted result asked for the full order (S.which then names the buffer that
// holds the prefix-sorted input, untouched by the consumer: the caller repeats the phase with a full sort).
static bool rank_sort_settle(Engine& e, RankSession& S, const RsTemp& T, bool flipped_by_consumer) {
    RsControl ctl;
    TG_CUDA(cudaMemcpyAsync(&ctl, T.ctl, sizeof(ctl), cudaMemcpyDeviceToHost, e.stream));
    TG_CUDA(cudaStreamSynchronize(e.stream));
    int r = ctl.result ? (S.which ^ 1) : S.which;
    if (ctl.fallback) {
        S.which = r;
        return false;
    }
    if (flipped_by_consumer) r ^= 1;  // the consumer kernel wrote the next stage's data into the other buffer pair
    S.which = r;
    return true;
}

static void rank_local_sort_locked(Engine& e, int* launches) {
    RankSession& S = session(e);
    if (S.n <= 0) return;
    const RsTemp T = rank_sort_queue(e, S, launches);
    rank_sort_settle(e, S, T, false);
}

// sort the current (kx, ky), turn run heads into ranks (+ rank_base), leave (ky, rank_x) as the session's data
static void rank_finish_x_locked(Engine& e, uint64_t rank_base, int* launches) {
    RankSession& S = session(e);
    if (S.payload32) throw Error(TG_ERR_INVALID_ARG, "rank session: x phase already finished");
    if (S.n > 0) {
        if (rank_base + (uint64_t)S.n >= ((uint64_t)1 << 32)) throw Error(TG_ERR_UNSUPPORTED, "Spearman: 2^32 or more pairs");
        RankArena& A = S.cur;
        for (int prefix = getenv("TG_RANK_FULL_SORT") ? 0 : 1;; prefix = 0) {
            const RsTemp T = rank_sort_queue(e, S, launches, prefix != 0, S.x_i64);
            const int64_t r_tiles = (S.n + RK_TILE - 1) / RK_TILE;
            uint64_t* k0 = A.K[S.which];
            uint64_t* k1 = A.K[S.which ^ 1];
            uint64_t* p0 = A.V[S.which];
            uint64_t* p1 = A.V[S.which ^ 1];
            rk_tile_heads_kernel<<<(unsigned)r_tiles, RK_THREADS, 0, e.stream>>>(T.ctl, rank_quant(A), k0, k1, S.n, A.tile_last);
            rk_tile_scan_kernel<<<1, 1024, 0, e.stream>>>(A.tile_last, r_tiles, A.carry);
            rk_rank_x_kernel<<<(unsigned)r_tiles, RK_THREADS, 0, e.stream>>>(T.ctl, rank_quant(A), k0, k1, p0, p1, S.n, A.carry, (uint32_t)rank_base);
            TG_CUDA(cudaGetLastError());
            *launches += 3;
            if (rank_sort_settle(e, S, T, true) || prefix == 0) break;  // else: long runs of different keys — full sort
        }
    }
    S.payload32 = 1;
}

// sort the current (ky, rank_x), rank_y = rank_base + run head, shifted co-moments around K -> sums[5]; ends the session
static void rank_finish_y_locked(Engine& e, uint64_t rank_base, double K, uint64_t* n_out, double* sums, int* launches) {
    RankSession& S = session(e);
    if (!S.payload32) throw Error(TG_ERR_INVALID_ARG, "rank session: finish the x phase first");
    for (int k = 0; k < 5; ++k) sums[k] = 0.0;
    *n_out = (uint64_t)S.n;
    if (S.n > 0) {
        if (rank_base + (uint64_t)S.n >= ((uint64_t)1 << 32)) throw Error(TG_ERR_UNSUPPORTED, "Spearman: 2^32 or more pairs");
        RankArena& A = S.cur;
        double* h_sums = (double*)e.host_scratch(256);
        for (int prefix = getenv("TG_RANK_FULL_SORT") ? 0 : 1;; prefix = 0) {
            const RsTemp T = rank_sort_queue(e, S, launches, prefix != 0, S.y_i64);
            const int64_t r_tiles = (S.n + RK_TILE - 1) / RK_TILE;
            const uint64_t* k0 = A.K[S.which];
            const uint64_t* k1 = A.K[S.which ^ 1];
            const uint32_t* r0 = (const uint32_t*)A.V[S.which];
            const uint32_t* r1 = (const uint32_t*)A.V[S.which ^ 1];
            double* d_out = (double*)(A.tmp + A.tmp_bytes - 2048);
            rk_tile_heads_kernel<<<(unsigned)r_tiles, RK_THREADS, 0, e.stream>>>(T.ctl, rank_quant(A), k0, k1, S.n, A.tile_last);
            rk_tile_scan_kernel<<<1, 1024, 0, e.stream>>>(A.tile_last, r_tiles, A.carry);
            rk_rank_y_moments_kernel<<<(unsigned)r_tiles, RK_THREADS, 0, e.stream>>>(T.ctl, rank_quant(A), k0, k1, r0, r1, S.n, A.carry, (uint32_t)rank_base, K, A.partial);
            rk_final_kernel<<<5, 1024, 0, e.stream>>>(A.partial, r_tiles, d_out);
            TG_CUDA(cudaGetLastError());
            *launches += 4;
            TG_CUDA(cudaMemcpyAsync(h_sums, d_out, 40, cudaMemcpyDeviceToHost, e.stream));
            if (rank_sort_settle(e, S, T, false) || prefix == 0) break;  // (the settle synchronises the stream)
        }
        for (int k = 0; k < 5; ++k) sums[k] = h_sums[k];
    }
    arena_free(e, S.cur);
    arena_free(e, S.next);
    S.n = 0;
}

void exec_spearman_job(Engine& e, Table& t, Plan& p, int agg_id) {
    Agg& a = p.aggs[agg_id];
    const int64_t n = t.n_rows;
    Column* cx = t.find(a.cols[0]);
    Column* cy = t.find(a.cols[1]);
    if (cx && cy)
        p.stats.bytes_scanned += 2 * (uint64_t)n * 8 + (cx->validity.p ? (uint64_t)(n + 7) / 8 : 0) + (cy->validity.p ? (uint64_t)(n + 7) / 8 : 0);
    int launches = 0;
    TG_CUDA(cudaEventRecord(e.ev[6], e.stream));
    struct Cleanup {
        Engine& e;
        ~Cleanup() { rank_session_destroy(e); }
    } cleanup{e};
    const int64_t n_pairs = rank_begin_locked(e, t, a.cols[0], a.cols[1], &launches);
    a.u[0] = (uint64_t)n_pairs;
    const double K = ((double)n_pairs + 1.0) / 2.0;
    a.f[0] = K;
    a.f[1] = K;
    if (n_pairs >= 2) {
        rank_finish_x_locked(e, 0, &launches);
        uint64_t n_out = 0;
        double h[5];
        rank_finish_y_locked(e, 0, K, &n_out, h, &launches);
        for (int k = 0; k < 5; ++k) a.f[2 + k] = h[k];
    }
    TG_CUDA(cudaEventRecord(e.ev[7], e.stream));
    TG_CUDA(cudaStreamSynchronize(e.stream));
    float ms = 0;
    TG_CUDA(cudaEventElapsedTime(&ms, e.ev[6], e.ev[7]));
    p.stats.sketch_ms += ms;
    p.stats.gpu_ms += ms;
    p.stats.launches += launches;
    e.launches += launches;
}

// ---- the stages as the host layer of the distributed sort sees them (capi.cpp: tg_rank_*) ----
struct RankLock {
    std::lock_guard<std::mutex> g;
    explicit RankLock(Engine& e) : g(e.mu) {
        cudaError_t ce = cudaSetDevice(e.device);
        if (ce != cudaSuccess) throw Error(TG_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(ce));
    }
};

int64_t rank_begin(Engine& e, const std::string& table, const std::string& cx, const std::string& cy) {
    RankLock l(e);
    e.sync_copies();
    auto it = e.tables.find(table);
    if (it == e.tables.end()) throw Error(TG_ERR_TABLE_NOT_FOUND, "Error during planning: table 'datafusion.public." + table + "' not found");
    int launches = 0;
    const int64_t n = rank_begin_locked(e, *it->second, cx, cy, &launches);
    e.launches += launches;
    return n;
}
void rank_local_sort(Engine& e) {
    RankLock l(e);
    int launches = 0;
    rank_local_sort_locked(e, &launches);
    e.launches += launches;
}
// up to m evenly spaced keys of the sorted shard; returns how many were written
int32_t rank_sample(Engine& e, int32_t m, uint64_t* out) {
    RankLock l(e);
    RankSession& S = session(e);
    if (S.n <= 0 || m <= 0) return 0;
    m = (int32_t)std::min<int64_t>(m, S.n);
    uint64_t* d = (uint64_t*)e.scratch((size_t)m * 8 + 256);
    rk_sample_kernel<<<(m + 255) / 256, 256, 0, e.stream>>>(S.cur.K[S.which], S.n, m, d);
    TG_CUDA(cudaGetLastError());
    TG_CUDA(cudaMemcpyAsync(out, d, (size_t)m * 8, cudaMemcpyDeviceToHost, e.stream));
    TG_CUDA(cudaStreamSynchronize(e.stream));
    e.launches += 1;
    return m;
}
// counts[p] = keys of the sorted shard that go to part p: part p takes the keys in (splitter[p-1], splitter[p]]
void rank_split(Engine& e, const uint64_t* splitters, int32_t n_parts, int64_t* counts) {
    RankLock l(e);
    RankSession& S = session(e);
    if (n_parts < 1) throw Error(TG_ERR_INVALID_ARG, "n_parts must be positive");
    std::vector<long long> pos((size_t)n_parts, S.n);
    if (S.n > 0 && n_parts > 1) {
        uint8_t* scr = e.scratch((size_t)n_parts * 16 + 512);
        uint64_t* d_s = (uint64_t*)scr;
        long long* d_p = (long long*)(scr + round_up((size_t)n_parts * 8, 256));
        TG_CUDA(cudaMemcpyAsync(d_s, splitters, (size_t)(n_parts - 1) * 8, cudaMemcpyHostToDevice, e.stream));
        rk_split_kernel<<<(n_parts + 254) / 255, 256, 0, e.stream>>>(S.cur.K[S.which], S.n, d_s, n_parts - 1, d_p);
        TG_CUDA(cudaGetLastError());
        TG_CUDA(cudaMemcpyAsync(pos.data(), d_p, (size_t)(n_parts - 1) * 8, cudaMemcpyDeviceToHost, e.stream));
        TG_CUDA(cudaStreamSynchronize(e.stream));
        e.launches += 1;
    }
    if (S.n <= 0)
        for (auto& x : pos) x = 0;
    long long prev = 0;
    for (int32_t i = 0; i < n_parts; ++i) {
        const long long hi = i == n_parts - 1 ? (long long)std::max<int64_t>(S.n, 0) : std::max(pos[i], prev);
        counts[i] = hi - prev;
        prev = hi;
    }
}
void rank_send_buffers(Engine& e, const void** keys, const void** payload, int32_t* payload_bytes) {
    RankLock l(e);
    RankSession& S = session(e);
    *keys = S.cur.K[S.which];
    *payload = S.cur.V[S.which];
    *payload_bytes = S.payload32 ? 4 : 8;
}
void rank_recv_buffers(Engine& e, int64_t n_recv, void** keys, void** payload) {
    RankLock l(e);
    RankSession& S = session(e);
    if (n_recv < 0 || n_recv >= ((int64_t)1 << 30)) throw Error(TG_ERR_UNSUPPORTED, "Spearman: 2^30 or more keys in one rank's range");
    arena_alloc(e, S.next, n_recv);
    *keys = S.next.K[0];
    *payload = S.next.V[0];
}
void rank_recv_commit(Engine& e, int64_t n_recv) {
    RankLock l(e);
    RankSession& S = session(e);
    if (!S.next.base || n_recv > S.next.cap) throw Error(TG_ERR_INVALID_ARG, "rank session: no receive buffers of that size");
    TG_CUDA(cudaStreamSynchronize(e.stream));
    arena_free(e, S.cur);
    S.cur = S.next;
    S.next = RankArena{};
    S.n = n_recv;
    S.which = 0;
}
void rank_finish_x(Engine& e, uint64_t rank_base) {
    RankLock l(e);
    int launches = 0;
    rank_finish_x_locked(e, rank_base, &launches);
    e.launches += launches;
}
void rank_finish_y(Engine& e, uint64_t rank_base, double K, uint64_t* n_out, double* sums) {
    RankLock l(e);
    int launches = 0;
    rank_finish_y_locked(e, rank_base, K, n_out, sums, &launches);
    e.launches += launches;
}
// ---- hooks of the in-library exchange (comm.cpp rank_exchange; the caller holds the engine's lock)
void rank_current_unlocked(Engine& e, const uint64_t** keys, const void** payload, int* pay_bytes, int64_t* n) {
    RankSession& S = session(e);
    *keys = S.cur.K[S.which];
    *payload = S.cur.V[S.which];
    *pay_bytes = S.payload32 ? 4 : 8;
    *n = S.n;
}
int32_t rank_sample_unlocked(Engine& e, int32_t m, uint64_t* out) {
    RankSession& S = session(e);
    if (S.n <= 0 || m <= 0) return 0;
    m = (int32_t)std::min<int64_t>(m, S.n);
    uint64_t* d = (uint64_t*)e.scratch((size_t)m * 8 + 256);
    rk_sample_kernel<<<(m + 255) / 256, 256, 0, e.stream>>>(S.cur.K[S.which], S.n, m, d);
    TG_CUDA(cudaGetLastError());
    TG_CUDA(cudaMemcpyAsync(out, d, (size_t)m * 8, cudaMemcpyDeviceToHost, e.stream));
    TG_CUDA(cudaStreamSynchronize(e.stream));
    e.launches += 1;
    return m;
}
// the session's data now lives in external buffers (this rank's receive buffers of the push exchange, capacity >= n_recv
// elements); a fresh arena supplies the ping-pong partners
void rank_adopt_received_unlocked(Engine& e, uint64_t* keys, void* payload, int64_t n_recv) {
    RankSession& S = session(e);
    if (n_recv < 0 || n_recv >= ((int64_t)1 << 30)) throw Error(TG_ERR_UNSUPPORTED, "Spearman: 2^30 or more keys in one rank's range");
    TG_CUDA(cudaStreamSynchronize(e.stream));
    arena_alloc(e, S.next, n_recv);
    arena_free(e, S.cur);
    S.cur = S.next;
    S.next = RankArena{};
    S.cur.K[0] = keys;
    S.cur.V[0] = (uint64_t*)payload;
    S.n = n_recv;
    S.which = 0;
}
int rank_phase_unlocked(Engine& e) { return session(e).payload32 ? 1 : 0; }

void rank_abort(Engine& e) {
    RankLock l(e);
    rank_session_destroy(e);
}

}  // namespace tg
