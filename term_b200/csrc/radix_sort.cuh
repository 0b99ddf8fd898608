// Hand-written LSD radix sort for sm_100a ("onesweep": ONE histogram pass over the keys for all digit positions, then
// ONE read + ONE write of the data per 8-bit digit, the tiles of a pass chained by decoupled look-back on per-tile,
// per-bin status words). Used by K6 (Spearman min-ranks, ranks.cu) and K4 (quantile sketch, sketch.cu); no library
// sort runs on any product path.
//
//   rs_hist_kernel      keys -> hist[pass][256]           (warp-uniform digits cost one shared atomic per warp)
//   rs_scan_kernel      hist -> bin bases, which passes are trivial (every key in one bin: skipped), buffer parity
//   rs_pass_kernel      one digit: warp-striped load, stable in-warp ranking with MATCH.ANY, cross-warp scan, look-back,
//                       tile reordered in shared memory so every bin's run leaves the SM coalesced
//
// Everything is stream-ordered; which of the two ping-pong buffers ends up holding the result is only known on the
// device (skipped passes do not flip it), so consumers read RsControl::result (or the host reads it after its sync).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace tg {

constexpr int RS_RADIX_BITS = 8, RS_BINS = 256, RS_MAX_PASSES = 8;
#ifndef TG_RS_THREADS
#define TG_RS_THREADS 256
#define TG_RS_ITEMS 14
#define TG_RS_MIN_BLOCKS 4
#endif
constexpr int RS_THREADS = TG_RS_THREADS, RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = TG_RS_ITEMS;              // keys per thread
constexpr int RS_MIN_BLOCKS = TG_RS_MIN_BLOCKS;    // resident tiles per SM the pass kernel is compiled for
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;     // 5376 keys per tile

struct RsControl {
    uint32_t skip[RS_MAX_PASSES];      // pass p is trivial (all keys share the digit)
    uint32_t src[RS_MAX_PASSES + 1];   // buffer (0 / 1) holding the data BEFORE pass p; [n_passes] = after the last
    uint32_t result;                   // == src[n_passes]
    uint32_t n_passes;
    uint32_t tile_counter[RS_MAX_PASSES];
    uint32_t first_exec;               // first non-trivial pass (== n_passes when every pass is trivial)
    uint32_t quantised;                // the data ends up sorted by rs_quant(key) only (0: by the key bits)
    uint32_t fallback;                 // raised by a consumer of a quantised result that needs the full order after all
    uint32_t pad[2];
};

// "Quantised" sorts order the data by a 32-bit monotone function Q of the key instead of the key's own bits: for numeric
// columns Q is LINEAR IN THE VALUE between the column's minimum and maximum, so its 32 bits are spent evenly over the
// value range (the raw bits of a double spend 11 of them on the exponent: a column that spans zero collapses whole binades
// into one prefix). Q is monotone (every step is a monotone floating-point / integer operation), so the result is ordered
// by Q with equal Q in input order; the consumer orders the short runs of equal Q by the full keys (ranks.cu).
struct RsQuant {       // device-resident, written by rs_quant_setup
    uint32_t mode;     // 0: Q = key >> 32; 1: Q = (key - kmin) >> shift (integer keys; f64 keys of one sign within 64 binades);
                       // 2: f64 order keys, Q = (x - xmin) * scale
    uint32_t shift;
    uint64_t kmin;
    double xmin, scale;
};
__device__ __forceinline__ uint32_t rs_quant(const RsQuant& q, uint64_t k) {
    if (q.mode == 2) {
        const uint64_t bits = (k & 0x8000000000000000ull) ? (k ^ 0x8000000000000000ull) : ~k;
        const double x = __longlong_as_double((long long)bits);
        return __double2uint_rz((x - q.xmin) * q.scale);  // saturating; the keys lie inside [xmin, xmax]
    }
    if (q.mode == 1) return (uint32_t)((k - q.kmin) >> q.shift);
    return (uint32_t)(k >> 32);
}
// minmax[0] / [1]: smallest / largest key (order-preserving u64 keys of an Int64 (is_i64) or Float64 column)
void rs_quant_setup(cudaStream_t stream, const unsigned long long* minmax, int is_i64, RsQuant* q);

struct RsTemp {
    RsControl* ctl;
    unsigned long long* hist;  // [RS_MAX_PASSES][256]
    unsigned long long* base;  // [RS_MAX_PASSES][256] exclusive bin bases
    uint32_t* status;          // [n_passes][n_tiles][256]
    size_t status_words_per_pass;
};

size_t rs_temp_bytes(int64_t n, int n_passes);
// carve `temp` (256-byte aligned, rs_temp_bytes(n, n_passes) bytes) into the pieces above
RsTemp rs_temp_carve(uint8_t* temp, int64_t n, int n_passes);

// Sort n (key, value) pairs by key bits [begin_bit, begin_bit + 8 * n_passes). keys[0] / vals[0] hold the input, [1] is
// the other ping-pong buffer. vals may be {nullptr, nullptr} (keys only). iota_values: the values of the input are the
// positions 0..n-1 and vals[0] is never read (it is filled with them when every pass turns out trivial). n < 2^30. Returns the number of kernels launched; the buffer index of the
// result is T.ctl->result (device memory).
// quant != nullptr: a quantised sort (4 passes over the 32 bits of rs_quant(*quant, key); begin_bit / n_passes ignored).
// hist_ready: the producer of keys[0] has already counted the digits into T.hist ([pass][256], zeroed by the caller before):
// no histogram pass over the keys.
template <typename V>
int rs_sort_pairs(cudaStream_t stream, uint64_t* const keys[2], V* const vals[2], int64_t n, int begin_bit, int n_passes,
                  bool iota_values, const RsTemp& T, int sm_count, const RsQuant* quant = nullptr, bool hist_ready = false);

}  // namespace tg
