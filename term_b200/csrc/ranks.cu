// K6 — Spearman rank correlation: RANK() OVER (ORDER BY x) (competition / minimum ranks) for both
// columns over the pairwise-complete rows, then the Pearson co-moments of the ranks
// (analyzers/advanced/correlation.rs:334-350, metric :407-427).
//
// Pipeline: order-preserving 64-bit keys (rows with a NULL on either side get the maximum key and sort
// last) -> radix sort (key, row id) [CUB DeviceRadixSort: library sort, like cuBLAS for GEMM] -> head
// flags + inclusive max-scan give every sorted position the position of the first equal key (= min rank)
// -> scatter ranks back by row id -> deterministic two-level reduction of the shifted rank moments.
// The reference accumulates rank products in UInt64 and overflows above ~3.8M rows; here
// ranks are exact integers carried as f64 and the sums are centred at (n+1)/2.
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "engine.hpp"

namespace tg {

constexpr int RK_THREADS = 256;

__device__ __forceinline__ uint64_t order_key(uint64_t bits, int is_i64) {
    if (is_i64) return bits ^ 0x8000000000000000ull;
    // canonical -0.0 -> +0.0 so they tie like SQL equality
    if ((bits << 1) == 0) bits = 0;
    return (bits & 0x8000000000000000ull) ? ~bits : (bits | 0x8000000000000000ull);
}

__global__ void rk_keys_kernel(const uint64_t* x, const uint32_t* vx, int x_i64, const uint64_t* y, const uint32_t* vy, int y_i64,
                               int64_t n, uint64_t* kx, uint64_t* ky, uint32_t* idx, unsigned long long* n_pairs) {
    unsigned long long c = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const bool ok = (!vx || ((vx[r >> 5] >> (r & 31)) & 1u)) && (!vy || ((vy[r >> 5] >> (r & 31)) & 1u));
        kx[r] = ok ? order_key(x[r], x_i64) : ~0ull;
        ky[r] = ok ? order_key(y[r], y_i64) : ~0ull;
        idx[r] = (uint32_t)r;
        c += ok;
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) c += __shfl_xor_sync(0xffffffffu, c, m);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(n_pairs, c);
}

__global__ void rk_heads_kernel(const uint64_t* sorted_keys, int64_t n, uint32_t* head_pos) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x)
        head_pos[p] = (p == 0 || sorted_keys[p] != sorted_keys[p - 1]) ? (uint32_t)p : 0u;
}

__global__ void rk_scatter_kernel(const uint32_t* sorted_idx, const uint32_t* min_pos, int64_t n_pairs, uint32_t* rank_of_row) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += (int64_t)gridDim.x * blockDim.x)
        rank_of_row[sorted_idx[p]] = min_pos[p] + 1u;
}

// block partials of the shifted rank co-moments, reduced in a fixed order by rk_final_kernel
__global__ void __launch_bounds__(RK_THREADS) rk_moments_kernel(const uint32_t* rx, const uint32_t* ry, const uint64_t* kx, int64_t n,
                                                                double K, double* partial /* [grid][5] */) {
    double s[5] = {0, 0, 0, 0, 0};
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        if (kx[r] == ~0ull) continue;  // not a pairwise-complete row (kx here is the UNSORTED key array)
        const double dx = (double)rx[r] - K, dy = (double)ry[r] - K;
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
    if (n >= (int64_t)1 << 32) throw Error(TG_ERR_UNSUPPORTED, "Spearman: more than 2^32 rows per shard");
    p.stats.bytes_scanned += 2 * (uint64_t)n * 8 + (cx->validity.p ? (uint64_t)(n + 7) / 8 : 0) + (cy->validity.p ? (uint64_t)(n + 7) / 8 : 0);
    if (n == 0) return;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + RK_THREADS - 1) / RK_THREADS, (int64_t)e.sm_count * 8));
    const size_t k_b = round_up((size_t)n * 8, 256), i_b = round_up((size_t)n * 4, 256);
    size_t sort_b = 0, scan_b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_b, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, n, 0, 64, e.stream);
    cub::DeviceScan::InclusiveScan(nullptr, scan_b, (const uint32_t*)nullptr, (uint32_t*)nullptr, cub::Max(), n, e.stream);
    const size_t tmp_b = round_up(std::max(sort_b, scan_b), 256);
    // kx, ky (unsorted), ks (sorted keys), idx, idx_sorted, head/minpos, rank_x, rank_y, partials
    uint8_t* scr = e.scratch(3 * k_b + 5 * i_b + tmp_b + round_up((size_t)grid * 40, 256) + 512);
    uint8_t* q = scr;
    uint64_t* kx = (uint64_t*)q; q += k_b;
    uint64_t* ky = (uint64_t*)q; q += k_b;
    uint64_t* ks = (uint64_t*)q; q += k_b;
    uint32_t* idx = (uint32_t*)q; q += i_b;
    uint32_t* idx_s = (uint32_t*)q; q += i_b;
    uint32_t* pos = (uint32_t*)q; q += i_b;
    uint32_t* rx = (uint32_t*)q; q += i_b;
    uint32_t* ry = (uint32_t*)q; q += i_b;
    uint8_t* d_tmp = q; q += tmp_b;
    double* partial = (double*)q; q += round_up((size_t)grid * 40, 256);
    double* d_out = (double*)q;
    unsigned long long* d_np = (unsigned long long*)(q + 64);
    TG_CUDA(cudaEventRecord(e.ev[6], e.stream));
    TG_CUDA(cudaMemsetAsync(d_np, 0, 8, e.stream));
    rk_keys_kernel<<<grid, RK_THREADS, 0, e.stream>>>((const uint64_t*)cx->values.p, (const uint32_t*)cx->validity.p, cx->dtype == TG_INT64,
                                                      (const uint64_t*)cy->values.p, (const uint32_t*)cy->validity.p, cy->dtype == TG_INT64, n,
                                                      kx, ky, idx, d_np);
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
        for (int side = 0; side < 2; ++side) {
            const uint64_t* keys = side ? ky : kx;
            uint32_t* rank = side ? ry : rx;
            size_t sb = tmp_b;
            TG_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, sb, keys, ks, idx, idx_s, n, 0, 64, e.stream));
            rk_heads_kernel<<<grid, RK_THREADS, 0, e.stream>>>(ks, (int64_t)n_pairs, pos);
            sb = tmp_b;
            TG_CUDA(cub::DeviceScan::InclusiveScan(d_tmp, sb, pos, pos, cub::Max(), (int64_t)n_pairs, e.stream));
            rk_scatter_kernel<<<grid, RK_THREADS, 0, e.stream>>>(idx_s, pos, (int64_t)n_pairs, rank);
            TG_CUDA(cudaGetLastError());
            launches += 8;
        }
        rk_moments_kernel<<<grid, RK_THREADS, 0, e.stream>>>(rx, ry, kx, n, K, partial);
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
