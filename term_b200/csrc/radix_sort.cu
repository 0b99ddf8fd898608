// Onesweep LSD radix sort (see radix_sort.cuh). 64-bit keys, optional 32- or 64-bit values.
#include "radix_sort.cuh"

#include <algorithm>
#include <type_traits>

namespace tg {

static size_t rs_round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static int64_t rs_tiles(int64_t n) { return (n + RS_TILE - 1) / RS_TILE; }

size_t rs_temp_bytes(int64_t n, int n_passes) {
    const size_t fixed = rs_round_up(sizeof(RsControl), 256) + 2 * rs_round_up((size_t)RS_MAX_PASSES * RS_BINS * 8, 256);
    return fixed + rs_round_up((size_t)std::max<int64_t>(rs_tiles(n), 1) * RS_BINS * 4, 256) * (size_t)n_passes + 256;
}

RsTemp rs_temp_carve(uint8_t* temp, int64_t n, int n_passes) {
    RsTemp T;
    uint8_t* q = temp;
    T.ctl = (RsControl*)q;
    q += rs_round_up(sizeof(RsControl), 256);
    T.hist = (unsigned long long*)q;
    q += rs_round_up((size_t)RS_MAX_PASSES * RS_BINS * 8, 256);
    T.base = (unsigned long long*)q;
    q += rs_round_up((size_t)RS_MAX_PASSES * RS_BINS * 8, 256);
    T.status = (uint32_t*)q;
    T.status_words_per_pass = rs_round_up((size_t)std::max<int64_t>(rs_tiles(n), 1) * RS_BINS * 4, 256) / 4;
    (void)n_passes;
    return T;
}

// ---------------------------------------------------------------------------------------------- histogram ----
constexpr int RSH_THREADS = 512;

// One read of the keys, all digit positions at once. A digit that is the same across the warp (the sign / exponent
// bytes of a numeric column, the high bytes of small integers) costs ONE shared atomic instead of a 32-way conflict.
template <bool QUANT>
__global__ void __launch_bounds__(RSH_THREADS) rs_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int begin_bit, int n_passes,
                                                              unsigned long long* __restrict__ hist, const RsQuant* __restrict__ quant) {
    __shared__ uint32_t s_hist[RS_MAX_PASSES][RS_BINS];
    RsQuant Q{};
    if constexpr (QUANT) Q = *quant;
    for (int i = threadIdx.x; i < RS_MAX_PASSES * RS_BINS; i += RSH_THREADS) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // chunks of 2^17 keys per flush: 32-bit shared counters cannot overflow, enough chunks to fill the machine
    const int64_t chunk = (int64_t)1 << 17;
    for (int64_t c0 = (int64_t)blockIdx.x * chunk; c0 < n; c0 += (int64_t)gridDim.x * chunk) {
        const int64_t c1 = c0 + chunk < n ? c0 + chunk : n;
        for (int64_t r0 = c0 + (int64_t)(threadIdx.x & ~31); r0 < c1; r0 += RSH_THREADS) {
            const int64_t r = r0 + lane;
            const bool in = r < c1;
            uint64_t k = in ? __ldg(keys + r) : 0ull;
            const unsigned live = __ballot_sync(0xffffffffu, in);
            if (!in) continue;
            if constexpr (QUANT) k = rs_quant(Q, k);
            const int leader = __ffs(live) - 1;
            for (int p = 0; p < n_passes; ++p) {
                const uint32_t d = (uint32_t)(k >> (begin_bit + p * RS_RADIX_BITS)) & (RS_BINS - 1);
                const uint32_t d0 = __shfl_sync(live, d, leader);
                if (__all_sync(live, d == d0)) {
                    if (lane == leader) atomicAdd(&s_hist[p][d], (uint32_t)__popc(live));
                } else {
                    atomicAdd(&s_hist[p][d], 1u);
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n_passes * RS_BINS; i += RSH_THREADS) {
            const uint32_t v = (&s_hist[0][0])[i];
            if (v) {
                atomicAdd(&hist[i], (unsigned long long)v);
                (&s_hist[0][0])[i] = 0;
            }
        }
        __syncthreads();
    }
}

// hist -> exclusive bin bases per pass; trivial passes; ping-pong parity
__global__ void __launch_bounds__(RS_BINS) rs_scan_kernel(const unsigned long long* __restrict__ hist, unsigned long long* __restrict__ base,
                                                          int64_t n, int n_passes, RsControl* ctl, uint32_t quantised) {
    __shared__ unsigned long long s_warp[RS_BINS / 32];
    __shared__ uint32_t s_trivial[RS_MAX_PASSES];
    const int d = threadIdx.x, lane = d & 31, w = d >> 5;
    if (d < RS_MAX_PASSES) s_trivial[d] = 0;
    __syncthreads();
    for (int p = 0; p < n_passes; ++p) {
        const unsigned long long c = hist[p * RS_BINS + d];
        if (c == (unsigned long long)n) s_trivial[p] = 1;
        unsigned long long x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[w] = x;
        __syncthreads();
        unsigned long long pre = 0;
        for (int i = 0; i < w; ++i) pre += s_warp[i];
        base[p * RS_BINS + d] = pre + x - c;
        __syncthreads();
    }
    if (d == 0) {
        ctl->quantised = quantised;
        uint32_t cur = 0, first = (uint32_t)n_passes;
        for (int p = 0; p < n_passes; ++p) {
            ctl->skip[p] = s_trivial[p];
            ctl->src[p] = cur;
            if (!s_trivial[p]) {
                cur ^= 1u;
                if (first == (uint32_t)n_passes) first = (uint32_t)p;
            }
        }
        ctl->first_exec = first;
        ctl->src[n_passes] = cur;
        ctl->result = cur;
        ctl->n_passes = (uint32_t)n_passes;
    }
}

// ------------------------------------------------------------------------------------------------- one pass ----
// Shared memory of a tile. The staging buffer holds the tile's keys in tile-sorted order, then (after they have left)
// its values: what the second write-out needs from the keys — their digit — is kept as one byte per position.
struct RsSmem {
    uint64_t stage[RS_TILE];               // keys, then values (V is at most 8 bytes)
    uint8_t dig[RS_TILE];                  // digit of the key at tile-sorted position j
    uint32_t hist[RS_WARPS][RS_BINS];      // per-warp digit counts -> exclusive offsets of the warp inside the tile's bin
    uint32_t block_hist[RS_BINS];          // the tile's digit counts (published early as the look-back aggregate)
    uint32_t digit_start[RS_BINS];         // first tile-sorted position of the bin
    long long global_base[RS_BINS];        // destination index of tile-sorted position j of bin d = global_base[d] + j
    uint32_t warp_tot[RS_BINS / 32];
    uint32_t tile;
};

enum : uint32_t { RS_FLAG_AGG = 1u, RS_FLAG_PREFIX = 2u };
constexpr int RS_LOOKBACK_BATCH = 4;  // predecessor status words fetched speculatively per look-back round trip

template <typename V, bool HAS_V, bool IOTA, bool QUANT>
__global__ void __launch_bounds__(RS_THREADS, RS_MIN_BLOCKS) rs_pass_kernel(uint64_t* k0, uint64_t* k1, V* v0, V* v1, int64_t n, int shift, int pass,
                                                                            RsControl* ctl, const unsigned long long* __restrict__ bin_base,
                                                                            uint32_t* status, const RsQuant* __restrict__ quant) {
    extern __shared__ __align__(16) uint8_t rs_smem_raw[];
    RsSmem& S = *reinterpret_cast<RsSmem*>(rs_smem_raw);
    if (ctl->skip[pass]) return;
    const uint32_t which = ctl->src[pass];
    const uint64_t* __restrict__ ksrc = which ? k1 : k0;
    uint64_t* __restrict__ kdst = which ? k0 : k1;
    const V* __restrict__ vsrc = which ? v1 : v0;
    V* __restrict__ vdst = which ? v0 : v1;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) S.tile = atomicAdd(&ctl->tile_counter[pass], 1u);
    for (int i = tid; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&S.hist[0][0])[i] = 0;
    if (tid < RS_BINS) S.block_hist[tid] = 0;
    __syncthreads();
    const uint32_t tile = S.tile;
    const int64_t tile_base = (int64_t)tile * RS_TILE;
    const int n_tile = (int)((n - tile_base) < (int64_t)RS_TILE ? (n - tile_base) : (int64_t)RS_TILE);

    // ---- warp-striped load: warp w owns positions [w * 32 * ITEMS, (w + 1) * 32 * ITEMS), item i of lane l = + i * 32 + l.
    // Positions past the end are padded with the maximum key: they rank behind every real key of bin 255.
    uint64_t key[RS_ITEMS];
    uint32_t rank[RS_ITEMS];
    uint32_t dg[(RS_ITEMS + 3) / 4];  // the items' digits, computed once, four to a register
    const int wbase = warp * 32 * RS_ITEMS + lane;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int pos = wbase + i * 32;
        key[i] = pos < n_tile ? ksrc[tile_base + pos] : ~0ull;
    }
    {
        RsQuant Q{};
        if constexpr (QUANT) Q = *quant;
#pragma unroll
        for (int i = 0; i < (RS_ITEMS + 3) / 4; ++i) dg[i] = 0;
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            uint32_t d;
            if constexpr (QUANT) d = wbase + i * 32 < n_tile ? (rs_quant(Q, key[i]) >> shift) & (RS_BINS - 1) : (uint32_t)(RS_BINS - 1);
            else d = (uint32_t)(key[i] >> shift) & (RS_BINS - 1);
            dg[i >> 2] |= d << (8 * (i & 3));
        }
    }
#define RS_DIGIT(i) ((dg[(i) >> 2] >> (8 * ((i)&3))) & 255u)
    // ---- the tile's digit counts first (plain shared atomics, no ordering needed): the aggregate the successors' look-back
    // waits for is published before the expensive stable ranking starts
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) atomicAdd(&S.block_hist[RS_DIGIT(i)], 1u);
    __syncthreads();
    uint32_t tile_count = 0, real_count = 0;
    volatile uint32_t* st = status + (size_t)tile * RS_BINS + tid;
    if (tid < RS_BINS) {
        tile_count = S.block_hist[tid];
        // the padding keys are not data: they never leave the tile and are not counted
        real_count = tile_count - (tid == RS_BINS - 1 ? (uint32_t)(RS_TILE - n_tile) : 0u);
        *st = (real_count << 2) | (tile == 0 ? RS_FLAG_PREFIX : RS_FLAG_AGG);
        // exclusive scan of tile_count over the bins -> first tile-sorted position of each bin
        uint32_t x = tile_count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) S.warp_tot[warp] = x;
        // (bins live in warps 0..7 only: a named barrier over those 256 threads orders warp_tot)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        uint32_t pre = 0;
        for (int i = 0; i < warp; ++i) pre += S.warp_tot[i];
        S.digit_start[tid] = pre + x - tile_count;
    }
    // ---- stable ranking inside the warp: items with equal digits are ordered by (i, lane) = by position
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const uint32_t d = RS_DIGIT(i);
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (lane == leader) {
            pre = S.hist[warp][d];
            S.hist[warp][d] = pre + (uint32_t)__popc(peers);
        }
        pre = __shfl_sync(0xffffffffu, pre, leader);
        rank[i] = pre + (uint32_t)__popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();

    // ---- per bin (thread d): exclusive scan of the warps' counts; decoupled look-back over the predecessor tiles
    if (tid < RS_BINS) {
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const uint32_t t = S.hist[w][tid];
            S.hist[w][tid] = sum;
            sum += t;
        }
        unsigned long long excl = 0;
        if (tile > 0) {
            int64_t t = (int64_t)tile - 1;
            bool done = false;
            while (!done) {
                uint32_t v[RS_LOOKBACK_BATCH];
#pragma unroll
                for (int k = 0; k < RS_LOOKBACK_BATCH; ++k)
                    v[k] = t - k >= 0 ? *(volatile uint32_t*)(status + (size_t)(t - k) * RS_BINS + tid) : RS_FLAG_PREFIX;
#pragma unroll
                for (int k = 0; k < RS_LOOKBACK_BATCH; ++k) {
                    const uint32_t f = v[k] & 3u;
                    if (f == 0) break;  // not published yet: look again from here
                    excl += v[k] >> 2;
                    --t;
                    if (f & RS_FLAG_PREFIX) {
                        done = true;
                        break;
                    }
                }
            }
            *st = ((uint32_t)(excl + real_count) << 2) | RS_FLAG_PREFIX;
        }
        S.global_base[tid] = (long long)(bin_base[tid] + excl) - (long long)S.digit_start[tid];
    }
    __syncthreads();

    // ---- reorder the tile in shared memory (tile-sorted order), then every bin's run is written contiguously
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const uint32_t d = RS_DIGIT(i);
        const uint32_t p = S.digit_start[d] + S.hist[warp][d] + rank[i];
        rank[i] = p;
        S.stage[p] = key[i];
        if constexpr (QUANT) S.dig[p] = (uint8_t)d;  // (plain sorts read the digit back from the staged key)
    }
#undef RS_DIGIT
    __syncthreads();
    for (int j = tid; j < n_tile; j += RS_THREADS) {
        const uint64_t k = S.stage[j];
        uint32_t d;
        if constexpr (QUANT) d = S.dig[j];
        else {
            d = (uint32_t)(k >> shift) & (RS_BINS - 1);
            if constexpr (HAS_V) S.dig[j] = (uint8_t)d;
        }
        kdst[S.global_base[d] + j] = k;
    }
    if constexpr (HAS_V) {
        __syncthreads();
        V* sv = reinterpret_cast<V*>(S.stage);
        // the first executed pass of an "iota" sort synthesises the values: the positions
        const bool iota = IOTA && (uint32_t)pass == ctl->first_exec;
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const int pos = wbase + i * 32;
            V v = V();
            if (pos < n_tile) v = iota ? (V)(tile_base + pos) : vsrc[tile_base + pos];
            sv[rank[i]] = v;
        }
        __syncthreads();
        for (int j = tid; j < n_tile; j += RS_THREADS) vdst[S.global_base[S.dig[j]] + j] = sv[j];
    }
}

template <typename V>
__global__ void rs_iota_if_unsorted_kernel(V* vals, int64_t n, const RsControl* ctl) {
    if (ctl->first_exec != ctl->n_passes) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) vals[i] = (V)i;
}

__global__ void rs_quant_setup_kernel(const unsigned long long* __restrict__ minmax, int is_i64, RsQuant* q) {
    const uint64_t kmin = minmax[0], kmax = minmax[1];
    RsQuant r{};
    r.mode = 0;
    r.kmin = kmin;
    if (kmax >= kmin) {
        if (is_i64) {
            const uint64_t range = kmax - kmin;
            const int bits = 64 - __clzll((long long)range);  // 0 when range == 0
            r.mode = 1;
            r.shift = bits > 32 ? (uint32_t)(bits - 32) : 0u;
        } else {
            auto value = [](uint64_t k) {
                const uint64_t bits = (k & 0x8000000000000000ull) ? (k ^ 0x8000000000000000ull) : ~k;
                return __longlong_as_double((long long)bits);
            };
            const double lo = value(kmin), hi = value(kmax), range = hi - lo;
            // One sign and at most 64 binades: the order keys themselves are value-linear inside a binade and every binade
            // gets >= 2^26 of the 2^32 levels — integer arithmetic only (the fp64 chain below costs the pass kernel ~25 %).
            const bool same_sign = ((kmin ^ kmax) >> 63) == 0;
            if (same_sign && (kmax >> 52) - (kmin >> 52) < 64 && isfinite(lo) && isfinite(hi)) {
                const uint64_t r2 = kmax - kmin;
                const int bits = 64 - __clzll((long long)r2);
                r.mode = 1;
                r.shift = bits > 32 ? (uint32_t)(bits - 32) : 0u;
            } else if (isfinite(lo) && isfinite(hi) && isfinite(range)) {
                // NaN / infinite ends or an overflowing range: the raw prefix (mode 0) and, most likely, the full-sort fallback
                r.mode = 2;
                r.xmin = lo;
                r.scale = range > 0.0 ? 4294967295.0 / range : 0.0;
                if (!isfinite(r.scale)) r.scale = 0.0;
            }
        }
    }
    *q = r;
}
void rs_quant_setup(cudaStream_t stream, const unsigned long long* minmax, int is_i64, RsQuant* q) {
    rs_quant_setup_kernel<<<1, 1, 0, stream>>>(minmax, is_i64, q);
}

template <typename V>
int rs_sort_pairs(cudaStream_t stream, uint64_t* const keys[2], V* const vals[2], int64_t n, int begin_bit, int n_passes,
                  bool iota_values, const RsTemp& T, int sm_count, const RsQuant* quant, bool hist_ready) {
    if (quant) {
        begin_bit = 0;
        n_passes = 4;
    }
    if (n_passes > RS_MAX_PASSES) n_passes = RS_MAX_PASSES;
    const int64_t tiles = rs_tiles(n);
    cudaMemsetAsync(T.ctl, 0, sizeof(RsControl), stream);
    if (!hist_ready) cudaMemsetAsync(T.hist, 0, (size_t)RS_MAX_PASSES * RS_BINS * 8, stream);
    if (n <= 0 || n_passes <= 0) return 0;
    cudaMemsetAsync(T.status, 0, T.status_words_per_pass * 4 * (size_t)n_passes, stream);
    const int hgrid = (int)std::max<int64_t>(1, std::min<int64_t>((n + ((int64_t)1 << 16) - 1) >> 16, (int64_t)sm_count * 4));
    if (hist_ready) {
    } else if (quant) rs_hist_kernel<true><<<hgrid, RSH_THREADS, 0, stream>>>(keys[0], n, begin_bit, n_passes, T.hist, quant);
    else rs_hist_kernel<false><<<hgrid, RSH_THREADS, 0, stream>>>(keys[0], n, begin_bit, n_passes, T.hist, nullptr);
    rs_scan_kernel<<<1, RS_BINS, 0, stream>>>(T.hist, T.base, n, n_passes, T.ctl, quant ? 1u : 0u);
    const bool has_v = vals[0] != nullptr || vals[1] != nullptr;
    using K = void (*)(uint64_t*, uint64_t*, V*, V*, int64_t, int, int, RsControl*, const unsigned long long*, uint32_t*, const RsQuant*);
    K kern;
    const size_t smem = sizeof(RsSmem);
    if (quant) kern = !has_v ? (K)rs_pass_kernel<V, false, false, true> : iota_values ? (K)rs_pass_kernel<V, true, true, true> : (K)rs_pass_kernel<V, true, false, true>;
    else kern = !has_v ? (K)rs_pass_kernel<V, false, false, false> : iota_values ? (K)rs_pass_kernel<V, true, true, false> : (K)rs_pass_kernel<V, true, false, false>;
    int launches = hist_ready ? 1 : 2;
    if (has_v && iota_values) {  // every pass trivial (all keys equal): nobody synthesises the positions
        rs_iota_if_unsorted_kernel<V><<<(unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)sm_count * 8)), 256, 0, stream>>>(vals[0], n, T.ctl);
        ++launches;
    }
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int p = 0; p < n_passes; ++p) {
        kern<<<(unsigned)tiles, RS_THREADS, smem, stream>>>(keys[0], keys[1], vals[0], vals[1], n, begin_bit + p * RS_RADIX_BITS, p, T.ctl,
                                                            T.base + (size_t)p * RS_BINS, T.status + (size_t)p * T.status_words_per_pass, quant);
        ++launches;
    }
    return launches;
}

template int rs_sort_pairs<uint32_t>(cudaStream_t, uint64_t* const[2], uint32_t* const[2], int64_t, int, int, bool, const RsTemp&, int, const RsQuant*, bool);
template int rs_sort_pairs<uint64_t>(cudaStream_t, uint64_t* const[2], uint64_t* const[2], int64_t, int, int, bool, const RsTemp&, int, const RsQuant*, bool);

}  // namespace tg
