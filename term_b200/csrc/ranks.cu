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
// Ranks travel with the rows through the two sorts instead of being scattered back by row id (a random 4-byte write
// per row costs more than a whole sort pass).
// The reference accumulates rank products in UInt64 and overflows above ~3.8M rows; here
// ranks are exact integers carried as f64 and the sums are centred at (n+1)/2.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "engine.hpp"
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
constexpr int RK_ITEMS = 16, RK_TILE = RK_THREADS * RK_ITEMS;

// tile_last[t] = 1-based position of the last run head inside tile t (0: the tile starts inside a run and never leaves it)
__global__ void __launch_bounds__(RK_THREADS) rk_tile_heads_kernel(const RsControl* ctl, const uint64_t* k0, const uint64_t* k1, int64_t n,
                                                                   uint32_t* tile_last) {
    const uint64_t* __restrict__ ks = ctl->result ? k1 : k0;
    const int64_t base = (int64_t)blockIdx.x * RK_TILE;
    uint32_t best = 0;
#pragma unroll 4
    for (int i = 0; i < RK_ITEMS; ++i) {
        const int64_t p = base + (int64_t)i * RK_THREADS + threadIdx.x;
        if (p < n && (p == 0 || ks[p] != ks[p - 1])) best = (uint32_t)p + 1u;  // positions grow with i
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
__device__ __forceinline__ RkTileRanks rk_tile_ranks(const uint64_t* __restrict__ ks, int64_t n, const uint32_t* __restrict__ carry,
                                                     uint32_t* s_warp /* [RK_THREADS / 32] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t wbase = (int64_t)blockIdx.x * RK_TILE + (int64_t)warp * 32 * RK_ITEMS;
    RkTileRanks R;
    uint32_t run = 0;  // last head seen so far inside this warp's range (warp-uniform after each row)
#pragma unroll
    for (int i = 0; i < RK_ITEMS; ++i) {
        const int64_t p = wbase + i * 32 + lane;
        uint32_t h = 0;
        if (p < n) {
            const uint64_t k = ks[p];
            if (p == 0 || k != ks[p - 1]) h = (uint32_t)p + 1u;
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, h, o);
            if (lane >= o) h = max(h, v);
        }
        h = max(h, run);
        R.rank[i] = h;
        run = __shfl_sync(0xffffffffu, h, 31);
    }
    if (lane == 0) s_warp[warp] = run;
    __syncthreads();
    uint32_t pre = carry[blockIdx.x];
    for (int w = 0; w < warp; ++w) pre = max(pre, s_warp[w]);
#pragma unroll
    for (int i = 0; i < RK_ITEMS; ++i) R.rank[i] = max(R.rank[i], pre);
    return R;
}

// after the sort by kx (payload ky): emit (ky, rank_x) in the sorted-by-x order — the input of the second sort
__global__ void __launch_bounds__(RK_THREADS) rk_rank_x_kernel(const RsControl* ctl, const uint64_t* k0, const uint64_t* k1, const uint64_t* p0,
                                                               const uint64_t* p1, int64_t n, const uint32_t* __restrict__ carry,
                                                               uint64_t* __restrict__ out_key, uint32_t* __restrict__ out_rank) {
    __shared__ uint32_t s_warp[RK_THREADS / 32];
    const uint64_t* __restrict__ ks = ctl->result ? k1 : k0;
    const uint64_t* __restrict__ ys = ctl->result ? p1 : p0;
    const RkTileRanks R = rk_tile_ranks(ks, n, carry, s_warp);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t wbase = (int64_t)blockIdx.x * RK_TILE + (int64_t)warp * 32 * RK_ITEMS;
#pragma unroll
    for (int i = 0; i < RK_ITEMS; ++i) {
        const int64_t p = wbase + i * 32 + lane;
        if (p < n) {
            out_key[p] = ys[p];
            out_rank[p] = R.rank[i];
        }
    }
}

// after the sort by ky (payload rank_x): rank_y on the fly and the block's partial sums of the shifted rank co-moments
// (fixed tile -> block mapping, fixed reduction shape: run-to-run reproducible)
__global__ void __launch_bounds__(RK_THREADS) rk_rank_y_moments_kernel(const RsControl* ctl, const uint64_t* k0, const uint64_t* k1,
                                                                       const uint32_t* r0, const uint32_t* r1, int64_t n,
                                                                       const uint32_t* __restrict__ carry, double K,
                                                                       double* __restrict__ partial /* [grid][5] */) {
    __shared__ uint32_t s_warp[RK_THREADS / 32];
    __shared__ double red[5][RK_THREADS / 32];
    const uint64_t* __restrict__ ks = ctl->result ? k1 : k0;
    const uint32_t* __restrict__ rxs = ctl->result ? r1 : r0;
    const RkTileRanks R = rk_tile_ranks(ks, n, carry, s_warp);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t wbase = (int64_t)blockIdx.x * RK_TILE + (int64_t)warp * 32 * RK_ITEMS;
    double s[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < RK_ITEMS; ++i) {
        const int64_t p = wbase + i * 32 + lane;
        if (p < n) {
            const double dx = (double)rxs[p] - K, dy = (double)R.rank[i] - K;
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

// fixed-order sum of the block partials: 5 warps, one per moment, lanes stride the blocks, shuffle tree at the end
__global__ void __launch_bounds__(160) rk_final_kernel(const double* __restrict__ partial, int64_t n_blocks, double* out) {
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double v = 0;
    for (int64_t b = lane; b < n_blocks; b += 32) v += partial[(size_t)b * 5 + k];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if (lane == 0) out[k] = v;
}

static size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

void exec_spearman_job(Engine& e, Table& t, Plan& p, int agg_id) {
    Agg& a = p.aggs[agg_id];
    Column* cx = t.find(a.cols[0]);
    Column* cy = t.find(a.cols[1]);
    for (int i = 0; i < 2; ++i)
        if (!(i ? cy : cx))
            throw Error(TG_ERR_COLUMN_NOT_FOUND, "Schema error: No field named " + a.cols[i] + ". Valid fields are " + t.valid_fields() + ".");
    for (Column* c : {cx, cy})
        if (c->dtype != TG_INT64 && c->dtype != TG_FLOAT64)
            throw Error(TG_ERR_TYPE_MISMATCH, "Spearman correlation requires numeric (Int64 / Float64) columns");
    const int64_t n = t.n_rows;
    p.stats.bytes_scanned += 2 * (uint64_t)n * 8 + (cx->validity.p ? (uint64_t)(n + 7) / 8 : 0) + (cy->validity.p ? (uint64_t)(n + 7) / 8 : 0);
    if (n == 0) return;
    if (n >= (int64_t)1 << 30) throw Error(TG_ERR_UNSUPPORTED, "Spearman: 2^30 or more rows per shard");
    const size_t k_b = round_up((size_t)n * 8, 256), i_b = round_up((size_t)n * 4, 256);
    const int64_t c_tiles = (n + RK_CTILE - 1) / RK_CTILE;    // compaction tiles (rows)
    const int64_t r_tiles_max = (n + RK_TILE - 1) / RK_TILE;  // rank tiles (pairs <= rows)
    const size_t ct_b = round_up((size_t)c_tiles * 4, 256), rt_b = round_up((size_t)r_tiles_max * 4, 256);
    const size_t tmp_b = round_up(rs_temp_bytes(n, RS_MAX_PASSES), 256);
    const size_t part_b = round_up((size_t)r_tiles_max * 40, 256);
    // A0/A1: keys of sort 1, B0/B1: its payload (ky), C: keys of sort 2 (its partner is A0), R0/R1: payload of sort 2 (rank_x)
    uint8_t* scr = e.scratch(5 * k_b + 2 * i_b + 2 * ct_b + 2 * rt_b + tmp_b + part_b + 512);
    uint8_t* q = scr;
    uint64_t* A0 = (uint64_t*)q; q += k_b;
    uint64_t* A1 = (uint64_t*)q; q += k_b;
    uint64_t* B0 = (uint64_t*)q; q += k_b;
    uint64_t* B1 = (uint64_t*)q; q += k_b;
    uint64_t* C = (uint64_t*)q; q += k_b;
    uint32_t* R0 = (uint32_t*)q; q += i_b;
    uint32_t* R1 = (uint32_t*)q; q += i_b;
    uint32_t* tile_count = (uint32_t*)q; q += ct_b;
    uint32_t* tile_off = (uint32_t*)q; q += ct_b;
    uint32_t* tile_last = (uint32_t*)q; q += rt_b;
    uint32_t* carry = (uint32_t*)q; q += rt_b;
    uint8_t* d_tmp = q; q += tmp_b;
    double* partial = (double*)q; q += part_b;
    double* d_out = (double*)q;
    unsigned long long* d_np = (unsigned long long*)(q + 64);
    const uint32_t* vx = (const uint32_t*)cx->validity.p;
    const uint32_t* vy = (const uint32_t*)cy->validity.p;
    TG_CUDA(cudaEventRecord(e.ev[6], e.stream));
    rk_count_kernel<<<(unsigned)c_tiles, RK_THREADS, 0, e.stream>>>(vx, vy, n, tile_count);
    rk_offsets_kernel<<<1, 1024, 0, e.stream>>>(tile_count, c_tiles, tile_off, d_np);
    rk_compact_keys_kernel<<<(unsigned)c_tiles, RK_THREADS, 0, e.stream>>>((const uint64_t*)cx->values.p, vx, cx->dtype == TG_INT64,
                                                                           (const uint64_t*)cy->values.p, vy, cy->dtype == TG_INT64, n, tile_off, A0, B0);
    TG_CUDA(cudaGetLastError());
    unsigned long long n_pairs = 0;
    TG_CUDA(cudaMemcpyAsync(&n_pairs, d_np, 8, cudaMemcpyDeviceToHost, e.stream));
    TG_CUDA(cudaStreamSynchronize(e.stream));
    int launches = 3;
    a.u[0] = n_pairs;
    const double K = ((double)n_pairs + 1.0) / 2.0;
    a.f[0] = K;
    a.f[1] = K;
    if (n_pairs >= 2) {
        const int64_t m = (int64_t)n_pairs;
        const int64_t r_tiles = (m + RK_TILE - 1) / RK_TILE;
        const RsTemp T = rs_temp_carve(d_tmp, m, RS_MAX_PASSES);
        {
            uint64_t* keys[2] = {A0, A1};
            uint64_t* vals[2] = {B0, B1};
            launches += rs_sort_pairs<uint64_t>(e.stream, keys, vals, m, 0, RS_MAX_PASSES, false, T, e.sm_count);
            TG_CUDA(cudaGetLastError());
            rk_tile_heads_kernel<<<(unsigned)r_tiles, RK_THREADS, 0, e.stream>>>(T.ctl, A0, A1, m, tile_last);
            rk_tile_scan_kernel<<<1, 1024, 0, e.stream>>>(tile_last, r_tiles, carry);
            rk_rank_x_kernel<<<(unsigned)r_tiles, RK_THREADS, 0, e.stream>>>(T.ctl, A0, A1, B0, B1, m, carry, C, R0);
            TG_CUDA(cudaGetLastError());
            launches += 3;
        }
        {
            uint64_t* keys[2] = {C, A0};
            uint32_t* vals[2] = {R0, R1};
            launches += rs_sort_pairs<uint32_t>(e.stream, keys, vals, m, 0, RS_MAX_PASSES, false, T, e.sm_count);
            TG_CUDA(cudaGetLastError());
            rk_tile_heads_kernel<<<(unsigned)r_tiles, RK_THREADS, 0, e.stream>>>(T.ctl, C, A0, m, tile_last);
            rk_tile_scan_kernel<<<1, 1024, 0, e.stream>>>(tile_last, r_tiles, carry);
            rk_rank_y_moments_kernel<<<(unsigned)r_tiles, RK_THREADS, 0, e.stream>>>(T.ctl, C, A0, R0, R1, m, carry, K, partial);
            rk_final_kernel<<<1, 160, 0, e.stream>>>(partial, r_tiles, d_out);
            TG_CUDA(cudaGetLastError());
            launches += 4;
        }
        double h[5];
        TG_CUDA(cudaMemcpyAsync(h, d_out, 40, cudaMemcpyDeviceToHost, e.stream));
        TG_CUDA(cudaStreamSynchronize(e.stream));
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

}  // namespace tg
