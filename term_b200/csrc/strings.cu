// K2 — string pattern kernel: host-compiled byte DFAs run over Arrow Utf8 (offsets + value bytes).
//
// Replaces `COUNT(CASE WHEN [TRIM(]c[)] ~|~* 'pat' ... THEN 1 END), COUNT(*)`
// (constraints/format.rs:762-776 -> arrow-string regexp_is_match -> regex crate) for every pattern that a
// suite applies to the same column, in ONE pass over that column's bytes.
//
// Layout / mapping: thread-per-string, 256-thread CTAs walking blocks of consecutive rows (adjacent
// lanes read adjacent strings, so a warp's loads fall in a handful of contiguous 128-byte lines that L1
// serves to all lanes); transition tables of up to 4 DFAs live in shared memory and share ONE joint
// byte-class map so each input byte costs one class lookup plus one table lookup per DFA still alive.
// A DFA leaves the alive set as soon as it reaches DEAD (no match possible) or MATCH (match found).
#include <algorithm>
#include <cstring>
#include <map>

#include "engine.hpp"
#include "regex_dfa.hpp"

namespace tg {

constexpr int STR_MAX_DFA = 4;
constexpr int STR_THREADS = 256;

struct DfaDev {
    uint32_t n_classes;   // joint classes
    uint32_t start;
    uint32_t table_off;   // offset (in uint16) of next[] inside the shared table blob
    uint32_t accept_off;  // offset (in uint16) of accept_end[] (one uint16 per state)
    uint32_t trim;
};
struct StrParams {
    const int32_t* offsets;
    const uint8_t* bytes;
    const uint32_t* validity;
    int64_t n_rows;
    int32_t n_dfa;
    uint32_t blob_u16;        // total uint16 entries in the table blob (after the 256-byte class map)
    const uint16_t* g_blob;   // [128 uint16 = 256-byte joint class map][tables...]
    unsigned long long* out;  // [n_dfa] match counts, [STR_MAX_DFA] = valid (non-null) rows
    DfaDev dfa[STR_MAX_DFA];
};

extern __shared__ __align__(16) uint16_t str_smem[];

__global__ void __launch_bounds__(STR_THREADS) dfa_kernel(const __grid_constant__ StrParams P) {
    // stage class map + tables
    {
        const uint32_t total = 128 + P.blob_u16;
        for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) str_smem[i] = P.g_blob[i];
    }
    __syncthreads();
    const uint8_t* jclass = reinterpret_cast<const uint8_t*>(str_smem);
    const uint16_t* tab = str_smem + 128;

    unsigned long long cnt[STR_MAX_DFA] = {0, 0, 0, 0};
    unsigned long long nvalid = 0;
    const uint32_t* words = reinterpret_cast<const uint32_t*>(P.bytes);

    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < P.n_rows; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = base + threadIdx.x;
        if (row >= P.n_rows) continue;
        if (P.validity && !((P.validity[row >> 5] >> (row & 31)) & 1u)) continue;
        ++nvalid;
        const int32_t b = P.offsets[row], e = P.offsets[row + 1];
        // TRIM(c): ASCII space only, both ends
        int32_t tb = b, te = e;
        bool any_trim = false;
#pragma unroll
        for (int i = 0; i < STR_MAX_DFA; ++i) any_trim |= (i < P.n_dfa) && P.dfa[i].trim;
        if (any_trim) {
            while (tb < te && P.bytes[tb] == ' ') ++tb;
            while (te > tb && P.bytes[te - 1] == ' ') --te;
        }
        uint32_t st[STR_MAX_DFA];
        uint32_t alive = 0;
#pragma unroll
        for (int i = 0; i < STR_MAX_DFA; ++i) {
            st[i] = i < P.n_dfa ? P.dfa[i].start : DFA_DEAD;
            if (i < P.n_dfa && st[i] > DFA_MATCH) alive |= 1u << i;
        }
        uint32_t w = 0;
        int32_t wi = -1;
        for (int32_t p = b; p < e && alive; ++p) {
            if ((p >> 2) != wi) {
                wi = p >> 2;
                w = __ldg(words + wi);
            }
            const uint32_t byte = (w >> ((p & 3) * 8)) & 0xffu;
            const uint32_t c = jclass[byte];
#pragma unroll
            for (int i = 0; i < STR_MAX_DFA; ++i) {
                if (alive & (1u << i)) {
                    const bool inside = !P.dfa[i].trim || (p >= tb && p < te);
                    if (inside) {
                        st[i] = tab[P.dfa[i].table_off + st[i] * P.dfa[i].n_classes + c];
                        if (st[i] <= DFA_MATCH) alive &= ~(1u << i);
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < STR_MAX_DFA; ++i) {
            if (i < P.n_dfa) {
                const bool m = st[i] == DFA_MATCH || (st[i] != DFA_DEAD && tab[P.dfa[i].accept_off + st[i]] != 0);
                cnt[i] += m;
            }
        }
    }
    // block reduction -> one atomic per counter per CTA
    __shared__ unsigned long long red[STR_MAX_DFA + 1][STR_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i <= STR_MAX_DFA; ++i) {
        unsigned long long v = i < STR_MAX_DFA ? cnt[i] : nvalid;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if (lane == 0) red[i][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x <= STR_MAX_DFA) {
        unsigned long long v = 0;
        for (int k = 0; k < STR_THREADS / 32; ++k) v += red[threadIdx.x][k];
        if (v) atomicAdd(P.out + threadIdx.x, v);
    }
}

static size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static void run_string_pass(Engine& e, Table& t, Plan& p, Column& c, const std::vector<int>& ids) {
    const int nd = (int)ids.size();
    std::vector<Dfa> dfas;
    for (int id : ids) {
        Agg& a = p.aggs[id];
        dfas.push_back(compile_regex(a.text, (a.flags & 1) != 0));
    }
    // joint byte classes
    std::map<std::vector<uint8_t>, uint8_t> sig_ids;
    uint8_t jclass[256];
    std::vector<uint8_t> rep;
    for (int b = 0; b < 256; ++b) {
        std::vector<uint8_t> sig;
        for (auto& d : dfas) sig.push_back(d.class_of[b]);
        auto it = sig_ids.find(sig);
        if (it == sig_ids.end()) {
            it = sig_ids.emplace(sig, (uint8_t)rep.size()).first;
            rep.push_back((uint8_t)b);
        }
        jclass[b] = it->second;
    }
    const uint32_t njc = (uint32_t)rep.size();
    std::vector<uint16_t> blob(128);
    memcpy(blob.data(), jclass, 256);
    StrParams P{};
    P.n_dfa = nd;
    for (int i = 0; i < nd; ++i) {
        const Dfa& d = dfas[i];
        P.dfa[i].n_classes = njc;
        P.dfa[i].start = d.start;
        P.dfa[i].trim = (p.aggs[ids[i]].flags & 2) ? 1 : 0;
        P.dfa[i].table_off = (uint32_t)blob.size() - 128;
        for (uint32_t s = 0; s < d.n_states; ++s)
            for (uint32_t jc = 0; jc < njc; ++jc)
                blob.push_back(d.next[(size_t)s * d.n_classes + d.class_of[rep[jc]]]);
        P.dfa[i].accept_off = (uint32_t)blob.size() - 128;
        for (uint32_t s = 0; s < d.n_states; ++s) blob.push_back(d.accept_end[s]);
    }
    P.blob_u16 = (uint32_t)blob.size() - 128;
    const size_t smem = blob.size() * 2;
    if (smem > 200 * 1024) throw Error(TG_ERR_UNSUPPORTED, "DFA tables exceed shared memory");
    const size_t blob_bytes = round_up(smem, 256);
    uint8_t* scr = e.scratch(blob_bytes + 256);
    TG_CUDA(cudaMemcpyAsync(scr, blob.data(), smem, cudaMemcpyHostToDevice, e.stream));
    unsigned long long* d_out = reinterpret_cast<unsigned long long*>(scr + blob_bytes);
    TG_CUDA(cudaMemsetAsync(d_out, 0, 64, e.stream));
    P.g_blob = reinterpret_cast<const uint16_t*>(scr);
    P.out = d_out;
    P.offsets = reinterpret_cast<const int32_t*>(c.offsets.p);
    P.bytes = c.values.p;
    P.validity = reinterpret_cast<const uint32_t*>(c.validity.p);
    P.n_rows = t.n_rows;
    static bool attr = false;
    if (!attr) {
        TG_CUDA(cudaFuncSetAttribute(dfa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    int per_sm = 1;
    TG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dfa_kernel, STR_THREADS, smem));
    per_sm = std::max(per_sm, 1);
    const int64_t blocks_needed = (t.n_rows + STR_THREADS - 1) / STR_THREADS;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(blocks_needed, (int64_t)e.sm_count * per_sm));
    TG_CUDA(cudaEventRecord(e.ev[2], e.stream));
    dfa_kernel<<<grid, STR_THREADS, smem, e.stream>>>(P);
    TG_CUDA(cudaGetLastError());
    TG_CUDA(cudaEventRecord(e.ev[3], e.stream));
    e.launches += 1;
    p.stats.launches += 1;
    unsigned long long h_out[8];
    TG_CUDA(cudaMemcpyAsync(h_out, d_out, 64, cudaMemcpyDeviceToHost, e.stream));
    TG_CUDA(cudaStreamSynchronize(e.stream));
    float ms = 0;
    TG_CUDA(cudaEventElapsedTime(&ms, e.ev[2], e.ev[3]));
    p.stats.string_ms += ms;
    p.stats.gpu_ms += ms;
    for (int i = 0; i < nd; ++i) {
        Agg& a = p.aggs[ids[i]];
        a.u[0] = h_out[i];
        a.u[1] = (uint64_t)t.n_rows - h_out[STR_MAX_DFA];
        a.u[2] = (uint64_t)t.n_rows;
    }
}

void exec_string_jobs(Engine& e, Table& t, Plan& p, const std::vector<int>& agg_ids) {
    // group by column, then passes of up to STR_MAX_DFA patterns whose tables fit in shared memory
    std::map<Column*, std::vector<int>> by_col;
    for (int id : agg_ids) {
        Agg& a = p.aggs[id];
        Column* c = t.find(a.cols[0]);
        if (!c) {
            a.err = TG_ERR_COLUMN_NOT_FOUND;
            a.err_msg = "Schema error: No field named " + a.cols[0] + ". Valid fields are " + t.valid_fields() + ".";
            continue;
        }
        if (c->dtype != TG_UTF8) {
            a.err = TG_ERR_TYPE_MISMATCH;
            a.err_msg = "Error during planning: regular expression match requires a Utf8 column, '" + c->name + "' is not";
            continue;
        }
        by_col[c].push_back(id);
    }
    for (auto& kv : by_col) {
        Column& c = *kv.first;
        // algorithmic bytes: offsets + value bytes + validity, once per column
        p.stats.bytes_scanned += (uint64_t)(t.n_rows + 1) * 4 + (uint64_t)c.value_bytes +
                                 (c.validity.p ? (uint64_t)(t.n_rows + 7) / 8 : 0);
        if (t.n_rows == 0) {
            for (int id : kv.second) p.aggs[id].u[2] = 0;
            continue;
        }
        std::vector<int> pass;
        size_t est = 0;
        auto flush = [&]() {
            if (!pass.empty()) run_string_pass(e, t, p, c, pass);
            pass.clear();
            est = 0;
        };
        for (int id : kv.second) {
            Dfa d = compile_regex(p.aggs[id].text, (p.aggs[id].flags & 1) != 0);
            // joint classes can exceed each DFA's own count; bound by 2x as a planning estimate
            size_t bytes = (size_t)d.n_states * std::min<size_t>(256, (size_t)d.n_classes * 2) * 2 + d.n_states * 2;
            if (!pass.empty() && (pass.size() >= (size_t)STR_MAX_DFA || est + bytes > 150 * 1024)) flush();
            pass.push_back(id);
            est += bytes;
        }
        flush();
    }
}

}  // namespace tg
