// K6 — Spearman rank correlation: RANK() OVER (ORDER BY x) (competition / minimum ranks) for both
// columns over the pairwise-complete rows, then the Pearson co-moments of the ranks
// (analyzers/advanced/correlation.rs:334-350, metric :407-427).
//
// Pipeline (no library code): order-preserving 64-bit keys (rows with a NULL on either side get the maximum key and
// sort last) -> hand-written onesweep LSD radix sort of (key, row id) (radix_sort.cu; digit positions on which every
// key agrees — sign / exponent bytes — are skipped, the row ids of the first executed pass are synthesised) -> the
// minimum rank of a sorted position is 1 + the position of the first key of its run: per-tile "last run head", a
// one-block max-scan of the tile results, then an in-tile max-scan with the carry -> ranks scattered back by row id
// -> deterministic two-level reduction of the shifted rank moments.
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

__global__ void rk_keys_kernel(const uint64_t* x, const uint32_t* vx, int x_i64, const uint64_t* y, const uint32_t* vy, int y_i64,
                               int64_t n, uint64_t* kx, uint64_t* ky, unsigned long long* n_pairs) {
    unsigned long long c = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const bool ok = (!vx || ((vx[r >> 5] >> (r & 31)) & 1u)) && (!vy || ((vy[r >> 5] >> (r & 31)) & 1u));
        kx[r] = ok ? order_key(x[r], x_i64) : ~0ull;
        ky[r] = ok ? order_key(y[r], y_i64) : ~0ull;
        c += ok;
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) c += __shfl_xor_sync(0xffffffffu, c, m);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(n_pairs, c);
}

// ---- minimum ranks from the sorted keys. A position's run head = the largest p' <= p with key[p'] != key[p' - 1]
// (p' = 0 counts); stored 1-based so that 0 means "no head in this range" and the combine is a plain max.
constexpr int RK_ITEMS = 16, RK_TILE = RK_THREADS * RK_ITEMS;

__device__ __forceinline__ const uint64_t* rk_sorted_keys(const RsControl* ctl, const uint64_t* k0, const uint64_t* k1) {
    return ctl->result ? k1 : k0;
}

// tile_last[t] = 1-based position of the last run head inside tile t (0: the tile starts inside a run and never leaves it)
__global__ void __launch_bounds__(RK_THREADS) rk_tile_heads_kernel(const RsControl* ctl, const uint64_t* k0, const uint64_t* k1, int64_t n,
                                                                   uint32_t* tile_last) {
    const uint64_t* __restrict__ ks = rk_sorted_keys(ctl, k0, k1);
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

// rank of every sorted position = its run head (1-based == the SQL competition rank), scattered back to its row. Keys
// equal to the sentinel may belong to valid rows (Int64 max / an all-ones NaN): the validity bitmaps decide.
__global__ void __launch_bounds__(RK_THREADS) rk_rank_scatter_kernel(const RsControl* ctl, const uint64_t* k0, const uint64_t* k1,
                                                                     const uint32_t* i0, const uint32_t* i1, int64_t n,
                                                                     const uint32_t* __restrict__ carry, const uint32_t* __restrict__ vx,
                                                                     const uint32_t* __restrict__ vy, uint32_t* __restrict__ rank_of_row) {
    const uint64_t* __restrict__ ks = rk_sorted_keys(ctl, k0, k1);
    const uint32_t* __restrict__ idx = ctl->result ? i1 : i0;
    __shared__ uint32_t s_warp[RK_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // blocked arrangement: thread t owns positions base + 0 .. ITEMS - 1 (a serial max-scan in registers)
    const int64_t base = (int64_t)blockIdx.x * RK_TILE + (int64_t)threadIdx.x * RK_ITEMS;
    uint64_t k[RK_ITEMS];
    uint64_t prev = 0;
    if (base > 0 && base < n) prev = ks[base - 1];
#pragma unroll
    for (int i = 0; i < RK_ITEMS; ++i) k[i] = base + i < n ? ks[base + i] : 0ull;
    uint32_t head[RK_ITEMS];
    uint32_t run = 0;
#pragma unroll
    for (int i = 0; i < RK_ITEMS; ++i) {
        const int64_t p = base + i;
        if (p < n && (p == 0 || k[i] != prev)) run = (uint32_t)p + 1u;
        head[i] = run;
        prev = k[i];
    }
    // exclusive max-scan of the threads' totals
    uint32_t x = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x = max(x, y);
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    uint32_t pre = carry[blockIdx.x];
    for (int w = 0; w < warp; ++w) pre = max(pre, s_warp[w]);
    const uint32_t up = __shfl_up_sync(0xffffffffu, x, 1);
    if (lane > 0) pre = max(pre, up);
#pragma unroll
    for (int i = 0; i < RK_ITEMS; ++i) {
        const int64_t p = base + i;
        if (p < n) {
            const uint32_t row = idx[p];
            bool ok = true;
            if (k[i] == ~0ull) ok = (!vx || ((vx[row >> 5] >> (row & 31)) & 1u)) && (!vy || ((vy[row >> 5] >> (row & 31)) & 1u));
            if (ok) rank_of_row[row] = max(head[i], pre);
        }
    }
}

// block partials of the shifted rank co-moments, reduced in a fixed order by rk_final_kernel
__global__ void __launch_bounds__(RK_THREADS) rk_moments_kernel(const uint32_t* rx, const uint32_t* ry, int64_t n,
                                                                double K, double* partial /* [grid][5] */) {
    double s[5] = {0, 0, 0, 0, 0};
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t a = rx[r], b = ry[r];
        if (a == 0u) continue;  // not a pairwise-complete row: its rank slots keep the memset's 0 (ranks start at 1)
        const double dx = (double)a - K, dy = (double)b - K;
        s[0] += dx;
        s[1] += dy;
        s[2] = fma(dx, dx, s[2]);
        s[3] = fma(dy, dy, s[3]);
        s[4] = fma(dx, dy, s[4]);
    }
    __shared__ double red[5][RK_THREADS / 32];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        double v = s[k];
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double v = 0;
        for (int w = 0; w < RK_THREADS / 32; ++w) v += red[threadIdx.x][w];
        partial[(size_t)blockIdx.x * 5 + threadIdx.x] = v;
    }
}
__global__ void rk_final_kernel(const double* partial, int n_blocks, double* out) {
    if (threadIdx.x < 5) {
        double v = 0;
        for (int b = 0; b < n_blocks; ++b) v += partial[(size_t)b * 5 + threadIdx.x];
        out[threadIdx.x] = v;
    }
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
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + RK_THREADS - 1) / RK_THREADS, (int64_t)e.sm_count * 8));
    const size_t k_b = round_up((size_t)n * 8, 256), i_b = round_up((size_t)n * 4, 256);
    const int64_t n_tiles = (n + RK_TILE - 1) / RK_TILE;
    const size_t t_b = round_up((size_t)n_tiles * 4, 256);
    const size_t tmp_b = round_up(rs_temp_bytes(n, RS_MAX_PASSES), 256);
    // kx, ky, k_alt (ping-pong partner of whichever side is being sorted), idx x2, rank_x, rank_y, tile heads / carry, sort temp, partials
    uint8_t* scr = e.scratch(3 * k_b + 4 * i_b + 2 * t_b + tmp_b + round_up((size_t)grid * 40, 256) + 512);
    uint8_t* q = scr;
    uint64_t* kx = (uint64_t*)q; q += k_b;
    uint64_t* ky = (uint64_t*)q; q += k_b;
    uint64_t* kalt = (uint64_t*)q; q += k_b;
    uint32_t* idx0 = (uint32_t*)q; q += i_b;
    uint32_t* idx1 = (uint32_t*)q; q += i_b;
    uint32_t* rx = (uint32_t*)q; q += i_b;
    uint32_t* ry = (uint32_t*)q; q += i_b;
    uint32_t* tile_last = (uint32_t*)q; q += t_b;
    uint32_t* carry = (uint32_t*)q; q += t_b;
    uint8_t* d_tmp = q; q += tmp_b;
    double* partial = (double*)q; q += round_up((size_t)grid * 40, 256);
    double* d_out = (double*)q;
    unsigned long long* d_np = (unsigned long long*)(q + 64);
    const uint32_t* vx = (const uint32_t*)cx->validity.p;
    const uint32_t* vy = (const uint32_t*)cy->validity.p;
    TG_CUDA(cudaEventRecord(e.ev[6], e.stream));
    TG_CUDA(cudaMemsetAsync(d_np, 0, 8, e.stream));
    TG_CUDA(cudaMemsetAsync(rx, 0, 2 * i_b, e.stream));  // rank 0 = "row is not pairwise complete"
    rk_keys_kernel<<<grid, RK_THREADS, 0, e.stream>>>((const uint64_t*)cx->values.p, vx, cx->dtype == TG_INT64,
                                                      (const uint64_t*)cy->values.p, vy, cy->dtype == TG_INT64, n, kx, ky, d_np);
    TG_CUDA(cudaGetLastError());
    unsigned long long n_pairs = 0;
    TG_CUDA(cudaMemcpyAsync(&n_pairs, d_np, 8, cudaMemcpyDeviceToHost, e.stream));
    TG_CUDA(cudaStreamSynchronize(e.stream));
    int launches = 1;
    a.u[0] = n_pairs;
    const double K = ((double)n_pairs + 1.0) / 2.0;
    a.f[0] = K;
    a.f[1] = K;
    if (n_pairs >= 2) {
        const RsTemp T = rs_temp_carve(d_tmp, n, RS_MAX_PASSES);
        for (int side = 0; side < 2; ++side) {
            uint64_t* keys[2] = {side ? ky : kx, kalt};
            uint32_t* vals[2] = {idx0, idx1};
            launches += rs_sort_pairs<uint32_t>(e.stream, keys, vals, n, 0, RS_MAX_PASSES, /*iota_values=*/true, T, e.sm_count);
            TG_CUDA(cudaGetLastError());
            rk_tile_heads_kernel<<<(unsigned)n_tiles, RK_THREADS, 0, e.stream>>>(T.ctl, keys[0], keys[1], n, tile_last);
            rk_tile_scan_kernel<<<1, 1024, 0, e.stream>>>(tile_last, n_tiles, carry);
            rk_rank_scatter_kernel<<<(unsigned)n_tiles, RK_THREADS, 0, e.stream>>>(T.ctl, keys[0], keys[1], idx0, idx1, n, carry, vx, vy, side ? ry : rx);
            TG_CUDA(cudaGetLastError());
            launches += 3;
        }
        rk_moments_kernel<<<grid, RK_THREADS, 0, e.stream>>>(rx, ry, n, K, partial);
        rk_final_kernel<<<1, 32, 0, e.stream>>>(partial, grid, d_out);
        TG_CUDA(cudaGetLastError());
        launches += 2;
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
