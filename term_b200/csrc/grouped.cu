// K5 — grouped completeness: GROUP BY g.. -> COUNT(*), COUNT(c) in ONE pass over the group columns.
//
// Replaces  SELECT g.., COUNT(*), COUNT(c) FROM t GROUP BY g.. [ORDER BY .. LIMIT ..]
//           (analyzers/basic/grouped_completeness.rs:131-176)
//
// One fused kernel reads, per row, the group columns (Utf8: offsets + bytes, fixed width: the value; validity
// bits) and the target column's validity bit — exactly the algorithmic bytes of SURVEY §8d, nothing is
// materialised per row. The group tuple is reduced to a 128-bit fingerprint in registers and counted in a
// per-CTA shared-memory table (the common case is a handful to a few thousand groups): a warp first agrees on
// which lanes share a slot (__match_any_sync) so each distinct group costs one shared atomic per warp. A CTA whose
// table fills up spills the row to the global table directly; at the end every CTA folds its table into the global
// one. The first row of every group is kept so the host can fetch the group's key values for the report.
#include <algorithm>
#include <cstring>

#include "engine.hpp"
#include "hash_common.cuh"
#include "ptx.cuh"

namespace tg {

constexpr int GRP_THREADS = 1024;
constexpr int GRP_SLOTS = 8192;     // per-CTA shared table (power of two): 24 B per slot = 192 KB of dynamic shared memory
constexpr int GRP_MAX_PROBE = 24;   // after this many probes the row goes to the global table
constexpr int GRP_MAX_COLS = 8;

struct GrpCol {
    const uint8_t* values;
    const int32_t* offsets;
    const uint32_t* validity;
    int32_t dtype;
    int32_t pad;
};
struct GrpParams {
    GrpCol cols[GRP_MAX_COLS];
    int32_t n_cols;
    int32_t pad;
    int64_t n_rows;
    const uint32_t* target_validity;
    Table128 gt;  // global table (h1, h2) + parallel arrays below
    unsigned long long* totals;
    unsigned long long* nonnull;
    long long* first_row;
    unsigned long long* n_groups;
};

constexpr int GRP_ILP = 4;  // rows per thread per iteration: their loads are issued together

// Fingerprints of GRP_ILP rows at once. Per group column the loads go out in waves (validity + offsets or values,
// then the first 8 key bytes) for all rows before anything is hashed, so a thread has GRP_ILP independent memory
// round trips in flight instead of a dependent chain per row.
__device__ __forceinline__ void row_fingerprints(const GrpParams& P, const int64_t (&row)[GRP_ILP], const bool (&act)[GRP_ILP],
                                                 Fp (&acc)[GRP_ILP]) {
#pragma unroll
    for (int k = 0; k < GRP_ILP; ++k) acc[k] = Fp{0, 0};
    for (int i = 0; i < P.n_cols; ++i) {
        const GrpCol& c = P.cols[i];
        const bool first = i == 0;
        bool valid[GRP_ILP];
#pragma unroll
        for (int k = 0; k < GRP_ILP; ++k) valid[k] = act[k] && row_valid(c.validity, row[k]);
        if (c.dtype == TG_UTF8) {
            int32_t b[GRP_ILP], e[GRP_ILP];
            uint64_t w[GRP_ILP];
#pragma unroll
            for (int k = 0; k < GRP_ILP; ++k) {
                b[k] = valid[k] ? __ldg(c.offsets + row[k]) : 0;
                e[k] = valid[k] ? __ldg(c.offsets + row[k] + 1) : 0;
            }
#pragma unroll
            for (int k = 0; k < GRP_ILP; ++k) w[k] = e[k] > b[k] ? load_upto8(c.values, b[k], min(8, e[k] - b[k])) : 0ull;
#pragma unroll
            for (int k = 0; k < GRP_ILP; ++k) {
                if (!valid[k]) {
                    fp_combine(acc[k], NULL_TAG1, NULL_TAG2, first);
                    continue;
                }
                uint64_t x = 0x736f6d6570736575ull ^ (uint64_t)(e[k] - b[k]), y = 0x646f72616e646f6dull + (uint64_t)(e[k] - b[k]) * 0x100000001b3ull;
                for (int32_t p = b[k]; p < e[k]; p += 8) {
                    const uint64_t ww = p == b[k] ? w[k] : load_upto8(c.values, p, min(8, e[k] - p));
                    x = (x ^ ww) * 0x9fb21c651e98df25ull;
                    x ^= x >> 29;
                    y = (y + ww) * 0xc2b2ae3d27d4eb4full;
                    y ^= y >> 31;
                }
                fp_combine(acc[k], x, y, first);  // fp_combine applies the finalising mix
            }
        } else {
            uint64_t v[GRP_ILP];
#pragma unroll
            for (int k = 0; k < GRP_ILP; ++k) {
                v[k] = 0;
                if (!valid[k]) continue;
                const int64_t r = row[k];
                switch (c.dtype) {
                    case TG_INT64: v[k] = __ldg(reinterpret_cast<const unsigned long long*>(c.values) + r); break;
                    case TG_FLOAT64: v[k] = canon_f64(__ldg(reinterpret_cast<const unsigned long long*>(c.values) + r)); break;
                    case TG_INT32: v[k] = (uint64_t)(int64_t)__ldg(reinterpret_cast<const int32_t*>(c.values) + r); break;
                    case TG_FLOAT32: v[k] = canon_f64((uint64_t)__double_as_longlong((double)__ldg(reinterpret_cast<const float*>(c.values) + r))); break;
                    default: v[k] = (__ldg(reinterpret_cast<const uint32_t*>(c.values) + (r >> 5)) >> (r & 31)) & 1u; break;  // TG_BOOL
                }
            }
#pragma unroll
            for (int k = 0; k < GRP_ILP; ++k) {
                if (valid[k]) fp_combine(acc[k], v[k], v[k] ^ 0x5851f42d4c957f2dull, first);
                else fp_combine(acc[k], NULL_TAG1, NULL_TAG2, first);
            }
        }
    }
}

__device__ __noinline__ void global_count(const GrpParams& P, Fp f, unsigned long long tot, unsigned long long nn, long long first) {
    bool created;
    const uint64_t slot = upsert128(P.gt, f, created);
    if (created) atomicAdd(P.n_groups, 1ull);
    atomicAdd(&P.totals[slot], tot);
    if (nn) atomicAdd(&P.nonnull[slot], nn);
    atomicMin(&P.first_row[slot], first);
}

// find-or-insert (a, b) in the CTA's shared table; returns the slot or -1 when the probe limit is hit. Kept out of
// line: it is called GRP_ILP times per iteration and inlining every copy overflows the instruction cache.
__device__ __noinline__ int smem_upsert(unsigned long long* s_h1, unsigned long long* s_h2, unsigned long long a, unsigned long long b,
                                        bool& created) {
    created = false;
    uint32_t s = (uint32_t)(a ^ (b >> 32)) & (GRP_SLOTS - 1);
    for (int probe = 0; probe < GRP_MAX_PROBE; ++probe, s = (s + 1) & (GRP_SLOTS - 1)) {
        unsigned long long v = s_h1[s];
        bool claimed = false;
        if (v == EMPTY64) {
            v = atomicCAS(&s_h1[s], EMPTY64, a);
            claimed = v == EMPTY64;
        }
        if (claimed || v == a) {
            // the first arrival publishes the second word; a different second word = another group
            unsigned long long w = s_h2[s];
            if (w == EMPTY64) w = atomicCAS(&s_h2[s], EMPTY64, b);
            if (w == EMPTY64 || w == b) {
                created = claimed;
                return (int)s;
            }
        }
    }
    return -1;
}

__global__ void __launch_bounds__(GRP_THREADS) group_count_fused_kernel(const __grid_constant__ GrpParams P) {
    extern __shared__ __align__(16) unsigned long long grp_smem[];
    unsigned long long* s_h1 = grp_smem;
    unsigned long long* s_h2 = grp_smem + GRP_SLOTS;
    uint32_t* s_tot = reinterpret_cast<uint32_t*>(grp_smem + 2 * GRP_SLOTS);
    uint32_t* s_nn = s_tot + GRP_SLOTS;
    for (int i = threadIdx.x; i < GRP_SLOTS; i += GRP_THREADS) {
        s_h1[i] = EMPTY64;
        s_h2[i] = EMPTY64;
        s_tot[i] = 0;
        s_nn[i] = 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int64_t base0 = (int64_t)blockIdx.x * GRP_THREADS * GRP_ILP; base0 < P.n_rows; base0 += (int64_t)gridDim.x * GRP_THREADS * GRP_ILP) {
        int64_t row[GRP_ILP];
        bool act[GRP_ILP], okv[GRP_ILP];
        Fp fps[GRP_ILP];
#pragma unroll
        for (int k = 0; k < GRP_ILP; ++k) {
            row[k] = base0 + (int64_t)k * GRP_THREADS + threadIdx.x;
            act[k] = row[k] < P.n_rows;
        }
#pragma unroll
        for (int k = 0; k < GRP_ILP; ++k) okv[k] = act[k] && row_valid(P.target_validity, row[k]);
        row_fingerprints(P, row, act, fps);
#pragma unroll
        for (int k = 0; k < GRP_ILP; ++k) {
            const bool ok = okv[k];
            int slot = -1;  // shared slot, or -1: not active / spilled
            if (act[k]) {
                // EMPTY64 is the shared table's vacancy marker in both words
                const unsigned long long a = fps[k].h1 == EMPTY64 ? 0ull : fps[k].h1, b = fps[k].h2 == EMPTY64 ? 0ull : fps[k].h2;
                bool created;
                slot = smem_upsert(s_h1, s_h2, a, b, created);
                // a new group of this CTA registers itself (and a representative row) in the global table right away;
                // a row that finds the shared table full is counted there directly
                if (slot < 0) global_count(P, fps[k], 1ull, ok ? 1ull : 0ull, (long long)row[k]);
                else if (created) global_count(P, fps[k], 0ull, 0ull, (long long)row[k]);
            }
            // one shared atomic per distinct slot per warp
            const unsigned peers = __match_any_sync(0xffffffffu, slot);
            const unsigned ok_peers = __ballot_sync(0xffffffffu, ok) & peers;
            if (slot >= 0 && lane == __ffs(peers) - 1) {
                atomicAdd(&s_tot[slot], (uint32_t)__popc(peers));
                const int c = __popc(ok_peers);
                if (c) atomicAdd(&s_nn[slot], (uint32_t)c);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < GRP_SLOTS; i += GRP_THREADS) {
        if (s_tot[i] == 0) continue;
        // EMPTY64 was mapped to 0 above exactly like upsert128 maps it, so the pair can be handed over as is
        global_count(P, Fp{s_h1[i], s_h2[i]}, s_tot[i], s_nn[i], INT64_MAX);
    }
}

__global__ void group_collect_kernel(Table128 t, const unsigned long long* totals, const unsigned long long* nonnull,
                                     const long long* first_row, uint64_t cap, unsigned long long* out_n, uint64_t max_out,
                                     unsigned long long* out /* [max_out][3] = first_row, total, nonnull */) {
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += (uint64_t)gridDim.x * blockDim.x) {
        if (t.h1[s] == EMPTY64) continue;
        const unsigned long long i = atomicAdd(out_n, 1ull);
        if (i < max_out) {
            out[i * 3 + 0] = (unsigned long long)first_row[s];
            out[i * 3 + 1] = totals[s];
            out[i * 3 + 2] = nonnull[s];
        }
    }
}

// key values of the groups' first rows: per (group, column) -> {valid, start, length} (Utf8) or {valid, value bits, 0}
__global__ void group_keys_kernel(GrpParams P, const unsigned long long* groups /* [n][3] */, uint64_t n_groups,
                                  unsigned long long* meta /* [n][n_cols][3] */) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_groups * (uint64_t)P.n_cols) return;
    const uint64_t g = i / P.n_cols;
    const int c = (int)(i % P.n_cols);
    const int64_t row = (int64_t)groups[g * 3];
    const GrpCol& col = P.cols[c];
    unsigned long long valid = row_valid(col.validity, row) ? 1 : 0, x = 0, y = 0;
    if (valid) {
        switch (col.dtype) {
            case TG_UTF8: x = (unsigned long long)col.offsets[row]; y = (unsigned long long)(col.offsets[row + 1] - col.offsets[row]); break;
            case TG_INT64: case TG_FLOAT64: x = reinterpret_cast<const unsigned long long*>(col.values)[row]; break;
            case TG_INT32: x = (unsigned long long)(long long)reinterpret_cast<const int32_t*>(col.values)[row]; break;
            case TG_FLOAT32: x = (unsigned long long)__double_as_longlong((double)reinterpret_cast<const float*>(col.values)[row]); break;
            default: x = (reinterpret_cast<const uint32_t*>(col.values)[row >> 5] >> (row & 31)) & 1u; break;
        }
    }
    meta[i * 3 + 0] = valid;
    meta[i * 3 + 1] = x;
    meta[i * 3 + 2] = y;
}
// packs the Utf8 key bytes: one warp per (group, column) entry
__global__ void group_key_bytes_kernel(GrpParams P, const unsigned long long* meta, const unsigned long long* dst_off, uint64_t n_entries,
                                       uint8_t* dst) {
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n_entries) return;
    const GrpCol& col = P.cols[(int)(i % P.n_cols)];
    if (col.dtype != TG_UTF8 || !meta[i * 3]) return;
    const uint64_t src = meta[i * 3 + 1], len = meta[i * 3 + 2], d = dst_off[i];
    for (uint64_t k = lane; k < len; k += 32) dst[d + k] = col.values[src + k];
}

static size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static uint64_t pow2_at_least(uint64_t x) {
    uint64_t p = 1024;
    while (p < x) p <<= 1;
    return p;
}

void exec_grouped_job(Engine& e, Table& t, Plan& p, int agg_id) {
    Agg& a = p.aggs[agg_id];
    auto need_col = [&](const std::string& name) -> Column* {
        Column* c = t.find(name);
        if (!c) throw Error(TG_ERR_COLUMN_NOT_FOUND, "Schema error: No field named " + name + ". Valid fields are " + t.valid_fields() + ".");
        return c;
    };
    Column* target = need_col(a.cols[0]);
    std::vector<Column*> gcols;
    for (size_t i = 1; i < a.cols.size(); ++i) gcols.push_back(need_col(a.cols[i]));
    if (gcols.size() > (size_t)GRP_MAX_COLS) throw Error(TG_ERR_UNSUPPORTED, "grouped completeness: more than 8 grouping columns");
    const int64_t n = t.n_rows;
    for (auto* c : gcols) {
        uint64_t b = c->validity.p ? (uint64_t)(n + 7) / 8 : 0;
        if (c->dtype == TG_UTF8) b += (uint64_t)(n + 1) * 4 + (uint64_t)c->value_bytes;
        else if (c->dtype == TG_BOOL) b += (uint64_t)(n + 7) / 8;
        else b += (uint64_t)n * c->elem_bytes();
        p.stats.bytes_scanned += b;
    }
    if (target->validity.p) p.stats.bytes_scanned += (uint64_t)(n + 7) / 8;
    uint64_t zero = 0;
    a.blob.assign((uint8_t*)&zero, (uint8_t*)&zero + 8);
    if (n == 0) return;
    const uint64_t max_groups_dev = 1u << 20;
    const uint64_t cap = pow2_at_least(std::min<uint64_t>((uint64_t)n * 2, max_groups_dev * 4));
    const size_t h_b = cap * 8, out_b = round_up((size_t)max_groups_dev * 24, 256);
    uint8_t* scr = e.scratch(5 * h_b + out_b + 256);
    uint8_t* q = scr;
    GrpParams P{};
    P.gt = Table128{(unsigned long long*)q, (unsigned long long*)(q + h_b), nullptr, cap - 1}; q += 2 * h_b;
    P.totals = (unsigned long long*)q; q += h_b;
    P.nonnull = (unsigned long long*)q; q += h_b;
    P.first_row = (long long*)q; q += h_b;
    unsigned long long* d_out = (unsigned long long*)q; q += out_b;
    unsigned long long* d_n = (unsigned long long*)q;
    P.n_groups = d_n;
    P.n_cols = (int)gcols.size();
    P.n_rows = n;
    P.target_validity = (const uint32_t*)target->validity.p;
    for (size_t i = 0; i < gcols.size(); ++i)
        P.cols[i] = GrpCol{gcols[i]->values.p, (const int32_t*)gcols[i]->offsets.p, (const uint32_t*)gcols[i]->validity.p, gcols[i]->dtype, 0};
    cudaEventRecord(e.ev[4], e.stream);
    TG_CUDA(cudaMemsetAsync(P.gt.h1, 0xFF, 2 * h_b, e.stream));
    TG_CUDA(cudaMemsetAsync(P.totals, 0, 2 * h_b, e.stream));
    TG_CUDA(cudaMemsetAsync(P.first_row, 0x7F, h_b, e.stream));
    TG_CUDA(cudaMemsetAsync(d_n, 0, 256, e.stream));
    const size_t smem = (size_t)GRP_SLOTS * 24;
    TG_CUDA(cudaFuncSetAttribute(group_count_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  // per device
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + GRP_THREADS * GRP_ILP - 1) / (GRP_THREADS * GRP_ILP), (int64_t)e.sm_count));
    group_count_fused_kernel<<<grid, GRP_THREADS, smem, e.stream>>>(P);
    TG_CUDA(cudaGetLastError());
    const int cgrid = (int)std::max<uint64_t>(1, std::min<uint64_t>((cap + 255) / 256, (uint64_t)e.sm_count * 8));
    group_collect_kernel<<<cgrid, 256, 0, e.stream>>>(P.gt, P.totals, P.nonnull, P.first_row, cap, d_n + 1, max_groups_dev, d_out);
    TG_CUDA(cudaGetLastError());
    unsigned long long h_n[2] = {0, 0};
    TG_CUDA(cudaMemcpyAsync(h_n, d_n, 16, cudaMemcpyDeviceToHost, e.stream));
    cudaEventRecord(e.ev[5], e.stream);
    TG_CUDA(cudaStreamSynchronize(e.stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, e.ev[4], e.ev[5]);
    p.stats.hash_ms += ms;
    p.stats.gpu_ms += ms;
    p.stats.launches += 2;
    e.launches += 2;
    const unsigned long long n_groups = h_n[0];
    if (n_groups > max_groups_dev) throw Error(TG_ERR_UNSUPPORTED, "grouped completeness: more than 1048576 groups");
    // ---- key values of the groups (batched: two small kernels + two copies, however many groups) ----
    const size_t nc = gcols.size(), n_entries = (size_t)n_groups * nc;
    std::vector<unsigned long long> out((size_t)n_groups * 3), meta(n_entries * 3), dst_off(n_entries, 0);
    std::vector<uint8_t> key_bytes;
    if (n_groups) {
        TG_CUDA(cudaMemcpy(out.data(), d_out, out.size() * 8, cudaMemcpyDeviceToHost));
        // meta / offsets / packed bytes live in the engine's grow-only auxiliary block (no cudaMalloc per execute)
        unsigned long long* d_meta = (unsigned long long*)e.aux(n_entries * 3 * 8 + n_entries * 8 + 256);
        unsigned long long* d_dst = d_meta + n_entries * 3;
        group_keys_kernel<<<(unsigned)((n_entries + 255) / 256), 256, 0, e.stream>>>(P, d_out, n_groups, d_meta);
        TG_CUDA(cudaGetLastError());
        TG_CUDA(cudaMemcpyAsync(meta.data(), d_meta, meta.size() * 8, cudaMemcpyDeviceToHost, e.stream));
        TG_CUDA(cudaStreamSynchronize(e.stream));
        uint64_t total = 0;
        for (size_t i = 0; i < n_entries; ++i) {
            dst_off[i] = total;
            if (gcols[i % nc]->dtype == TG_UTF8 && meta[i * 3]) total += meta[i * 3 + 2];
        }
        key_bytes.resize((size_t)total);
        if (total) {
            // the packed bytes go where d_out's copy-out already happened: the scratch block (>= max_groups * 24 bytes)
            uint8_t* d_bytes = total <= out_b ? (uint8_t*)d_out : nullptr;
            const bool own = d_bytes == nullptr;
            if (own) TG_CUDA(cudaMalloc(&d_bytes, (size_t)total));
            TG_CUDA(cudaMemcpyAsync(d_dst, dst_off.data(), n_entries * 8, cudaMemcpyHostToDevice, e.stream));
            group_key_bytes_kernel<<<(unsigned)((n_entries * 32 + 255) / 256), 256, 0, e.stream>>>(P, d_meta, d_dst, n_entries, d_bytes);
            cudaError_t ce = cudaGetLastError();
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(key_bytes.data(), d_bytes, (size_t)total, cudaMemcpyDeviceToHost, e.stream);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(e.stream);
            if (own) cudaFree(d_bytes);
            TG_CUDA(ce);
        }
        p.stats.launches += 2;
        e.launches += 2;
    }
    // blob: [u64 n_groups] then per group: u32 key_len, key (group column values joined by \x1f), u64 total, u64 non_null
    a.blob.resize(8);
    memcpy(a.blob.data(), &n_groups, 8);
    for (unsigned long long g = 0; g < n_groups; ++g) {
        std::string key;
        for (size_t i = 0; i < nc; ++i) {
            if (i) key += '\x1f';
            const size_t en = (size_t)g * nc + i;
            if (!meta[en * 3]) {
                key += "NULL";
                continue;
            }
            const unsigned long long x = meta[en * 3 + 1];
            switch (gcols[i]->dtype) {
                case TG_UTF8: key.append((const char*)key_bytes.data() + dst_off[en], (size_t)meta[en * 3 + 2]); break;
                case TG_INT64: case TG_INT32: key += fmt_i64((int64_t)x); break;
                case TG_FLOAT64: case TG_FLOAT32: {
                    double d;
                    memcpy(&d, &x, 8);
                    key += fmt_f64(d);
                } break;
                default: break;  // Boolean group values print as the empty string, as before
            }
        }
        uint32_t L = (uint32_t)key.size();
        size_t o = a.blob.size();
        a.blob.resize(o + 4 + L + 16);
        memcpy(a.blob.data() + o, &L, 4);
        memcpy(a.blob.data() + o + 4, key.data(), L);
        memcpy(a.blob.data() + o + 4 + L, &out[g * 3 + 1], 8);
        memcpy(a.blob.data() + o + 4 + L + 8, &out[g * 3 + 2], 8);
    }
}

}  // namespace tg
