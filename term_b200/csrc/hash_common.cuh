// Shared device helpers of the hash jobs (hashing.cu: single-table path for small inputs, Utf8 / composite keys and
// grouped counts; hashpart.cu: radix-partitioned path for large Int64 / Float64 key columns and the multi-GPU shuffle).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tg {

constexpr unsigned long long EMPTY64 = 0xFFFFFFFFFFFFFFFFull;
constexpr int HASH_THREADS = 256;

__host__ __device__ __forceinline__ uint64_t fmix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return k;
}
__host__ __device__ __forceinline__ uint64_t canon_f64(uint64_t bits) {
    // -0.0 == +0.0 and all NaNs compare as one value when grouping
    if ((bits << 1) == 0) return 0;
    if ((bits & 0x7ff0000000000000ull) == 0x7ff0000000000000ull && (bits & 0x000fffffffffffffull)) return 0x7ff8000000000000ull;
    return bits;
}
__device__ __forceinline__ bool row_valid(const uint32_t* validity, int64_t row) {
    return !validity || ((validity[row >> 5] >> (row & 31)) & 1u);
}

// How the 64 hash bits of a key are spent (disjoint fields, so the levels are independent):
//   bits  0..31  slot inside a (bucket-local) hash table
//   bits 32..41  local radix bucket (<= 1024 buckets per GPU)
//   bits 42..63  destination rank of the multi-GPU shuffle: (top22 * world) >> 22
__host__ __device__ __forceinline__ uint32_t hash_bucket(uint64_t h, uint32_t pmask) { return (uint32_t)(h >> 32) & pmask; }
__host__ __device__ __forceinline__ uint32_t hash_rank(uint64_t h, uint32_t world) { return (uint32_t)(((h >> 42) * world) >> 22); }

struct HashCounters {
    unsigned long long distinct_all;      // groups, NULL as a value
    unsigned long long distinct_nonnull;  // groups whose key has no NULL component
    unsigned long long singles_plus;      // +1 on first insert
    unsigned long long singles_minus;     // +1 on second insert
    unsigned long long any_null_rows;
    unsigned long long special;           // rows whose exact key equals the EMPTY sentinel (path A)
    unsigned long long violations;        // FK
    unsigned long long null_children;     // FK
    unsigned long long n_examples;
    unsigned long long overflow;          // a bucket table filled up (skewed hash): the caller falls back
    unsigned long long pad[6];
};

// block-level reduction of per-thread counters, one atomic per counter per warp
__device__ __forceinline__ void flush_counter(unsigned long long v, unsigned long long* dst) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(dst, v);
}

// the same with ONE global atomic per block (every thread of the block must call it): the bucket kernels are launched per
// bucket with ~1000 blocks each, and a warp-granular flush made thousands of same-address atomics the tail of every launch
__device__ __forceinline__ void flush_counter_block(unsigned long long v, unsigned long long* dst) {
    __shared__ unsigned long long s_part[32];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    const int warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31) >> 5;
    __syncthreads();  // (s_part may still be read by the previous counter's flush)
    if ((threadIdx.x & 31) == 0) s_part[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < n_warps; ++w) t += s_part[w];
        if (t) atomicAdd(dst, t);
    }
}

// ---- 128-bit fingerprints of Utf8 / composite keys ----
struct Fp {
    uint64_t h1, h2;
};
constexpr uint64_t NULL_TAG1 = 0x9ae16a3b2f90404full, NULL_TAG2 = 0xc3a5c85c97cb3127ull;

__device__ __forceinline__ void fp_combine(Fp& acc, uint64_t a, uint64_t b, bool first) {
    if (first) {
        acc.h1 = fmix64(a ^ 0x2545f4914f6cdd1dull);
        acc.h2 = fmix64(b + 0x9e3779b97f4a7c15ull);
    } else {
        acc.h1 = fmix64(acc.h1 * 0x9e3779b97f4a7c15ull + a);
        acc.h2 = fmix64((acc.h2 ^ b) * 0xd6e8feb86659fd93ull + 0x632be59bd9b4e019ull);
    }
}


struct Table128 {
    unsigned long long* h1;
    unsigned long long* h2;
    uint32_t* counts;
    uint64_t mask;
};

// find-or-insert a fingerprint; returns the slot and whether this call created it
__device__ __forceinline__ uint64_t upsert128(const Table128& t, Fp f, bool& created) {
    uint64_t a = f.h1 == EMPTY64 ? 0 : f.h1, b = f.h2 == EMPTY64 ? 0 : f.h2;
    uint64_t slot = (a ^ (b >> 32)) & t.mask;
    created = false;
    while (true) {
        const unsigned long long prev = atomicCAS(&t.h1[slot], EMPTY64, (unsigned long long)a);
        if (prev == EMPTY64) {
            // claimed: publish the second word
            atomicExch(&t.h2[slot], (unsigned long long)b);
            created = true;
            return slot;
        }
        if (prev == a) {
            unsigned long long v;
            do {
                v = *reinterpret_cast<volatile unsigned long long*>(&t.h2[slot]);
            } while (v == EMPTY64);
            if (v == b) return slot;
        }
        slot = (slot + 1) & t.mask;
    }
}


}  // namespace tg
