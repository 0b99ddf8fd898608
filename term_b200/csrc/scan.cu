// K1 — fused single-pass numeric scan (sm_100a).
//
// Replaces, for a whole suite at once, the per-constraint DataFusion aggregates
//   COUNT(*) / COUNT(c)            constraints/completeness.rs:158, analyzers/basic/completeness.rs:106
//   MIN/MAX/AVG/SUM/STDDEV/VARIANCE constraints/statistics.rs:263, analyzers/basic/{mean,min_max,sum}.rs
//   CORR / COVAR_SAMP / raw co-sums constraints/correlation.rs:313, analyzers/advanced/correlation.rs:239
//   COUNT(CASE WHEN pred THEN 1 END) constraints/custom_sql.rs:203, analyzers/advanced/compliance.rs:153
//
// Design (B200-first, HBM-bound):
//   * persistent CTAs, one per SM; each CTA walks row tiles round-robin.
//   * warp 0 is a TMA producer: for every tile it issues one cp.async.bulk (UBLKCP) per referenced
//     buffer (values and validity words of each column) into a multi-stage shared-memory ring,
//     completion tracked by mbarrier transaction bytes. Every input byte crosses HBM->SM exactly once
//     no matter how many constraints reference it.
//   * 16 consumer warps each own a fixed list of "units" (one aggregate over a row slice of the
//     tile); they read the tile from shared memory conflict-free (lane-contiguous 8-byte elements),
//     keep per-lane accumulators in shared memory between tiles, and release the stage through an
//     "empty" mbarrier, so fast warps run up to n_stages-1 tiles ahead of slow ones.
//   * descriptor tables (columns, units, predicate code) are copied to shared memory once; every
//     kind / flag decision is hoisted out of the row loops (template specialisations).
//   * moments are accumulated as shifted sums Σ(x-K), Σ(x-K)² with K an element of the column; NULL rows
//     are replaced by K (contributing exactly 0), so the inner loops have no predicated accumulates.
//   * per-CTA partials go to global memory; scan_finalize_kernel reduces them in a fixed order, so a
//     given (grid, plan) is bit-reproducible run to run.
#include <cfloat>
#include <cstddef>
#include <cstdio>
#include <math_constants.h>

#include "ptx.cuh"
#include "scan_defs.h"

namespace tg {

extern __shared__ __align__(128) uint8_t scan_smem[];

__device__ __forceinline__ uint32_t tail_mask(int base_row, int rows_in_tile) {
    const int rem = rows_in_tile - base_row;
    return rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
}
__device__ __forceinline__ double u2d(uint64_t u) { return __longlong_as_double((long long)u); }
__device__ __forceinline__ uint64_t d2u(double d) { return (uint64_t)__double_as_longlong(d); }

// per-lane state in shared memory: st[slot * 32 + lane]
__device__ __forceinline__ void unit_init(const ScanUnitDesc& u, uint64_t* st, int lane) {
#pragma unroll
    for (int k = 0; k < SCAN_STATE_SLOTS; ++k) st[k * 32 + lane] = 0;
    if (u.kind == UNIT_NUM_F64) {
        st[S_MIN * 32 + lane] = d2u(CUDART_INF);
        st[S_MAX * 32 + lane] = d2u(-CUDART_INF);
    } else if (u.kind == UNIT_NUM_I64) {
        st[S_MIN * 32 + lane] = (uint64_t)INT64_MAX;
        st[S_MAX * 32 + lane] = (uint64_t)INT64_MIN;
    }
}

// validity word (32 rows) of a column inside the stage; all-ones when the column has no bitmap
__device__ __forceinline__ uint32_t vword(const ScanColDesc& c, const uint8_t* stage, int word) {
    return c.validity ? reinterpret_cast<const uint32_t*>(stage + c.smem_bits_off)[word] : 0xffffffffu;
}

// ======================================================================================================
// Units. Every unit kind is a small object with begin(state) / tile(stage) / end(state): a warp that owns ONE
// unit keeps the object (descriptor fields + accumulators) in registers across its whole tile loop; a warp that
// owns several runs begin/tile/end per tile against the per-lane state in shared memory.
// ======================================================================================================

// ---- COUNT: popcount of validity words ----
struct CountUnit {
    uint32_t bits_off;
    int w0, nw, row0, lane;
    uint64_t n;
    __device__ __forceinline__ void begin(const ScanTables& T, const ScanUnitDesc& u, const uint64_t* st, int lane_) {
        const ScanColDesc& c = T.cols[u.c0];
        bits_off = c.validity ? c.smem_bits_off : 0xffffffffu;
        w0 = u.row0 >> 5;
        nw = u.nrows >> 5;
        row0 = u.row0;
        lane = lane_;
        n = st[S_N * 32 + lane];
    }
    __device__ __forceinline__ void tile(const uint8_t* stage, int rows, bool) {
        for (int j = lane; j < nw; j += 32) {
            uint32_t w = bits_off != 0xffffffffu ? reinterpret_cast<const uint32_t*>(stage + bits_off)[w0 + j] : 0xffffffffu;
            n += __popc(w & tail_mask(row0 + 32 * j, rows));
        }
    }
    __device__ __forceinline__ void end(uint64_t* st) { st[S_N * 32 + lane] = n; }
};

// ---- NUM ----
// NULL rows (and rows past the end of the table) are replaced by the pivot element K: they add 0 to Σd and
// Σd², cannot change min/max (K is a value of the column), and the wrapping integer sum is corrected on the
// host by (rows_processed - n)·K. Valid rows are counted by popcounting validity words (one word per lane).
template <bool IS_I64, int FLAGS>
struct NumUnit {
    static constexpr bool MOM = (FLAGS & UF_MOMENTS) != 0, MM = (FLAGS & UF_MINMAX) != 0, ISUM = IS_I64 && (FLAGS & UF_ISUM) != 0;
    uint32_t val_off, bits_off;  // bits_off == ~0u: no bitmap
    int row0, nw, lane;
    double K;
    uint64_t Kbits, n, isum;
    double sd, sdd, fmn, fmx;
    int64_t imn, imx;

    __device__ __forceinline__ void begin(const ScanTables& T, const ScanUnitDesc& u, const uint64_t* st, int lane_) {
        const ScanColDesc& c = T.cols[u.c0];
        val_off = c.smem_val_off;
        bits_off = c.validity ? c.smem_bits_off : 0xffffffffu;
        row0 = u.row0;
        nw = u.nrows >> 5;
        lane = lane_;
        K = c.pivot;
        Kbits = IS_I64 ? (uint64_t)c.ipivot : d2u(c.pivot);
        n = st[S_N * 32 + lane];
        sd = u2d(st[S_SD * 32 + lane]);
        sdd = u2d(st[S_SDD * 32 + lane]);
        fmn = u2d(st[S_MIN * 32 + lane]);
        fmx = u2d(st[S_MAX * 32 + lane]);
        imn = (int64_t)st[S_MIN * 32 + lane];
        imx = (int64_t)st[S_MAX * 32 + lane];
        isum = st[S_ISUM * 32 + lane];
    }
    template <int MASK>  // 0 dense full tile, 1 bitmap, 2 bitmap and/or tail
    __device__ __forceinline__ void loop(const uint8_t* stage, int rows) {
        const uint64_t* vals = reinterpret_cast<const uint64_t*>(stage + val_off) + row0 + lane;
        const uint32_t* bits = reinterpret_cast<const uint32_t*>(stage + (bits_off != 0xffffffffu ? bits_off : 0u)) + (row0 >> 5);
        const uint32_t lanebit = 1u << lane;
        const bool has_bits = bits_off != 0xffffffffu;
#pragma unroll 4
        for (int j = 0; j < nw; ++j) {
            uint64_t raw = vals[32 * j];
            if (MASK == 1) {
                raw = (bits[j] & lanebit) ? raw : Kbits;
            } else if (MASK == 2) {
                uint32_t w = has_bits ? bits[j] : 0xffffffffu;
                w &= tail_mask(row0 + 32 * j, rows);
                raw = (w & lanebit) ? raw : Kbits;
            }
            if (IS_I64) {
                const int64_t xi = (int64_t)raw;
                if (MOM) {
                    const double d = (double)xi - K;
                    sd += d;
                    sdd = fma(d, d, sdd);
                }
                if (ISUM) isum += (uint64_t)xi;
                if (MM) {
                    imn = xi < imn ? xi : imn;
                    imx = xi > imx ? xi : imx;
                }
            } else {
                const double x = u2d(raw);
                if (MOM) {
                    const double d = x - K;
                    sd += d;
                    sdd = fma(d, d, sdd);
                }
                if (MM) {
                    fmn = x < fmn ? x : fmn;
                    fmx = x > fmx ? x : fmx;
                }
            }
        }
    }
    __device__ __forceinline__ void tile(const uint8_t* stage, int rows, bool partial) {
        // valid rows: each lane popcounts different words
        if (bits_off != 0xffffffffu || partial) {
            for (int j = lane; j < nw; j += 32) {
                uint32_t w = bits_off != 0xffffffffu ? reinterpret_cast<const uint32_t*>(stage + bits_off)[(row0 >> 5) + j] : 0xffffffffu;
                n += __popc(w & tail_mask(row0 + 32 * j, rows));
            }
        } else if (lane == 0) {
            n += (uint64_t)nw * 32;
        }
        if (FLAGS == 0) return;
        if (partial) loop<2>(stage, rows);
        else if (bits_off != 0xffffffffu) loop<1>(stage, rows);
        else loop<0>(stage, rows);
    }
    __device__ __forceinline__ void end(uint64_t* st) {
        st[S_N * 32 + lane] = n;
        if (MOM) {
            st[S_SD * 32 + lane] = d2u(sd);
            st[S_SDD * 32 + lane] = d2u(sdd);
        }
        if (MM) {
            st[S_MIN * 32 + lane] = IS_I64 ? (uint64_t)imn : d2u(fmn);
            st[S_MAX * 32 + lane] = IS_I64 ? (uint64_t)imx : d2u(fmx);
        }
        if (ISUM) st[S_ISUM * 32 + lane] = isum;
    }
};

// slow path: the pivot is not an element of the column (no finite valid value was found when the column was
// registered, e.g. all NULL): rows are masked explicitly
template <bool IS_I64>
struct NumSlowUnit {
    uint32_t val_off, bits_off;
    int row0, nw, lane;
    double K;
    uint64_t n, isum;
    double sd, sdd, fmn, fmx;
    int64_t imn, imx;
    __device__ __forceinline__ void begin(const ScanTables& T, const ScanUnitDesc& u, const uint64_t* st, int lane_) {
        const ScanColDesc& c = T.cols[u.c0];
        val_off = c.smem_val_off;
        bits_off = c.validity ? c.smem_bits_off : 0xffffffffu;
        row0 = u.row0;
        nw = u.nrows >> 5;
        lane = lane_;
        K = c.pivot;
        n = st[S_N * 32 + lane];
        sd = u2d(st[S_SD * 32 + lane]);
        sdd = u2d(st[S_SDD * 32 + lane]);
        fmn = u2d(st[S_MIN * 32 + lane]);
        fmx = u2d(st[S_MAX * 32 + lane]);
        imn = (int64_t)st[S_MIN * 32 + lane];
        imx = (int64_t)st[S_MAX * 32 + lane];
        isum = st[S_ISUM * 32 + lane];
    }
    __device__ __noinline__ void tile(const uint8_t* stage, int rows, bool) {
        const uint64_t* vals = reinterpret_cast<const uint64_t*>(stage + val_off) + row0 + lane;
        for (int j = 0; j < nw; ++j) {
            uint32_t w = bits_off != 0xffffffffu ? reinterpret_cast<const uint32_t*>(stage + bits_off)[(row0 >> 5) + j] : 0xffffffffu;
            w &= tail_mask(row0 + 32 * j, rows);
            const bool ok = (w >> lane) & 1u;
            n += ok;
            const uint64_t raw = vals[32 * j];
            if (IS_I64) {
                const int64_t xi = (int64_t)raw;
                const double d = ok ? (double)xi - K : 0.0;
                sd += d;
                sdd = fma(d, d, sdd);
                isum += ok ? (uint64_t)xi : 0ull;
                imn = min(imn, ok ? xi : INT64_MAX);
                imx = max(imx, ok ? xi : INT64_MIN);
            } else {
                const double x = u2d(raw);
                const double d = ok ? x - K : 0.0;
                sd += d;
                sdd = fma(d, d, sdd);
                fmn = fmin(fmn, ok ? x : CUDART_INF);
                fmx = fmax(fmx, ok ? x : -CUDART_INF);
            }
        }
    }
    __device__ __forceinline__ void end(uint64_t* st) {
        st[S_N * 32 + lane] = n;
        st[S_SD * 32 + lane] = d2u(sd);
        st[S_SDD * 32 + lane] = d2u(sdd);
        st[S_MIN * 32 + lane] = IS_I64 ? (uint64_t)imn : d2u(fmn);
        st[S_MAX * 32 + lane] = IS_I64 ? (uint64_t)imx : d2u(fmx);
        st[S_ISUM * 32 + lane] = isum;
    }
};

// ---- PAIR: rows where either side is NULL are replaced by (Kx, Ky) => dx = dy = 0 ----
template <bool XI, bool YI>
struct PairUnit {
    uint32_t vx_off, vy_off, bx_off, by_off;
    int row0, nw, lane;
    double Kx, Ky, sx, sy, sxx, syy, sxy;
    uint64_t n;
    __device__ __forceinline__ void begin(const ScanTables& T, const ScanUnitDesc& u, const uint64_t* st, int lane_) {
        const ScanColDesc& cx = T.cols[u.c0];
        const ScanColDesc& cy = T.cols[u.c1];
        vx_off = cx.smem_val_off;
        vy_off = cy.smem_val_off;
        bx_off = cx.validity ? cx.smem_bits_off : 0xffffffffu;
        by_off = cy.validity ? cy.smem_bits_off : 0xffffffffu;
        row0 = u.row0;
        nw = u.nrows >> 5;
        lane = lane_;
        Kx = cx.pivot;
        Ky = cy.pivot;
        n = st[P_N * 32 + lane];
        sx = u2d(st[P_SX * 32 + lane]);
        sy = u2d(st[P_SY * 32 + lane]);
        sxx = u2d(st[P_SXX * 32 + lane]);
        syy = u2d(st[P_SYY * 32 + lane]);
        sxy = u2d(st[P_SXY * 32 + lane]);
    }
    __device__ __forceinline__ uint32_t both(const uint8_t* stage, int word) const {
        uint32_t w = 0xffffffffu;
        if (bx_off != 0xffffffffu) w &= reinterpret_cast<const uint32_t*>(stage + bx_off)[word];
        if (by_off != 0xffffffffu) w &= reinterpret_cast<const uint32_t*>(stage + by_off)[word];
        return w;
    }
    template <int MASK>
    __device__ __forceinline__ void loop(const uint8_t* stage, int rows) {
        const uint64_t* vx = reinterpret_cast<const uint64_t*>(stage + vx_off) + row0 + lane;
        const uint64_t* vy = reinterpret_cast<const uint64_t*>(stage + vy_off) + row0 + lane;
        const uint32_t lanebit = 1u << lane;
        const int w0 = row0 >> 5;
#pragma unroll 4
        for (int j = 0; j < nw; ++j) {
            const uint64_t rx = vx[32 * j], ry = vy[32 * j];
            double x = XI ? (double)(int64_t)rx : u2d(rx);
            double y = YI ? (double)(int64_t)ry : u2d(ry);
            if (MASK) {
                uint32_t w = both(stage, w0 + j);
                if (MASK == 2) w &= tail_mask(row0 + 32 * j, rows);
                const bool ok = w & lanebit;
                x = ok ? x : Kx;
                y = ok ? y : Ky;
            }
            const double dx = x - Kx, dy = y - Ky;
            sx += dx;
            sy += dy;
            sxx = fma(dx, dx, sxx);
            syy = fma(dy, dy, syy);
            sxy = fma(dx, dy, sxy);
        }
    }
    __device__ __forceinline__ void tile(const uint8_t* stage, int rows, bool partial) {
        const bool has_bits = bx_off != 0xffffffffu || by_off != 0xffffffffu;
        if (has_bits || partial) {
            for (int j = lane; j < nw; j += 32) n += __popc(both(stage, (row0 >> 5) + j) & tail_mask(row0 + 32 * j, rows));
        } else if (lane == 0) {
            n += (uint64_t)nw * 32;
        }
        if (partial) loop<2>(stage, rows);
        else if (has_bits) loop<1>(stage, rows);
        else loop<0>(stage, rows);
    }
    __device__ __forceinline__ void end(uint64_t* st) {
        st[P_N * 32 + lane] = n;
        st[P_SX * 32 + lane] = d2u(sx);
        st[P_SY * 32 + lane] = d2u(sy);
        st[P_SXX * 32 + lane] = d2u(sxx);
        st[P_SYY * 32 + lane] = d2u(syy);
        st[P_SXY * 32 + lane] = d2u(sxy);
    }
};

// ---- TERMS: AND / OR of <= 4 comparison terms ----
// One pass per term over a chunk of TG groups (32 rows each); the running TRUE masks of the chunk stay in
// registers; the comparison operator is selected OUTSIDE the row loop, so a term costs one load, one compare and
// one ballot per 32 rows.
constexpr int TG = 8;  // 256 rows per chunk

template <int KIND, int OP>  // OP = cmp_mask: 1 <, 2 ==, 3 <=, 4 >, 5 <> (13 with unordered), 6 >=
__device__ __forceinline__ void term_cmp_pass(uint64_t imm, uint32_t val_off, uint32_t bits_off, const uint8_t* stage,
                                              int row0, int lane, bool is_or, uint32_t (&m)[TG]) {
    const uint64_t* vals = reinterpret_cast<const uint64_t*>(stage + val_off) + row0 + lane;
    const uint32_t* bits = reinterpret_cast<const uint32_t*>(stage + (bits_off != 0xffffffffu ? bits_off : 0u)) + (row0 >> 5);
    const bool has_bits = bits_off != 0xffffffffu;
#pragma unroll
    for (int g = 0; g < TG; ++g) {
        const uint64_t raw = vals[32 * g];
        bool r;
        if (KIND == TK_I64) {
            const int64_t x = (int64_t)raw, c = (int64_t)imm;
            r = OP == 1 ? x < c : OP == 2 ? x == c : OP == 3 ? x <= c : OP == 4 ? x > c : OP == 6 ? x >= c : x != c;
        } else {
            // floats: NaN sorts above every number (Arrow total order): > and >= and <> are true for NaN
            const double x = KIND == TK_F64 ? u2d(raw) : (double)(int64_t)raw, c = u2d(imm);
            r = OP == 1 ? x < c : OP == 2 ? x == c : OP == 3 ? x <= c : OP == 4 ? !(x <= c) : OP == 6 ? !(x < c) : x != c;
        }
        uint32_t tm = __ballot_sync(0xffffffffu, r);
        if (has_bits) tm &= bits[g];
        m[g] = is_or ? (m[g] | tm) : (m[g] & tm);
    }
}
template <int KIND>
__device__ __forceinline__ void term_pass(int cmp_mask, uint64_t imm, uint32_t val_off, uint32_t bits_off, const uint8_t* stage,
                                          int row0, int lane, bool is_or, uint32_t (&m)[TG]) {
    switch (cmp_mask & 7) {
        case 1: term_cmp_pass<KIND, 1>(imm, val_off, bits_off, stage, row0, lane, is_or, m); break;
        case 2: term_cmp_pass<KIND, 2>(imm, val_off, bits_off, stage, row0, lane, is_or, m); break;
        case 3: term_cmp_pass<KIND, 3>(imm, val_off, bits_off, stage, row0, lane, is_or, m); break;
        case 4: term_cmp_pass<KIND, 4>(imm, val_off, bits_off, stage, row0, lane, is_or, m); break;
        case 6: term_cmp_pass<KIND, 6>(imm, val_off, bits_off, stage, row0, lane, is_or, m); break;
        default: term_cmp_pass<KIND, 5>(imm, val_off, bits_off, stage, row0, lane, is_or, m); break;
    }
}

struct TermsUnit {
    uint64_t imm[SCAN_UNIT_TERMS];
    uint32_t val_off[SCAN_UNIT_TERMS], bits_off[SCAN_UNIT_TERMS];
    int kind[SCAN_UNIT_TERMS], cmp[SCAN_UNIT_TERMS];
    int nt, row0, nrows, lane;
    bool is_or;
    uint64_t cnt;
    __device__ __forceinline__ void begin(const ScanTables& T, const ScanUnitDesc& u, const uint64_t* st, int lane_) {
        nt = u.code_len;
#pragma unroll
        for (int k = 0; k < SCAN_UNIT_TERMS; ++k) {
            const ScanTerm& t = T.terms[u.code_off + (k < nt ? k : 0)];
            const ScanColDesc& c = T.cols[t.col];
            imm[k] = t.imm;
            kind[k] = t.kind;
            cmp[k] = t.cmp_mask;
            val_off[k] = c.smem_val_off;
            bits_off[k] = c.validity ? c.smem_bits_off : 0xffffffffu;
        }
        row0 = u.row0;
        nrows = u.nrows;
        lane = lane_;
        is_or = u.flags & 1;
        cnt = st[0];
    }
    __device__ __forceinline__ void tile(const uint8_t* stage, int rows, bool) {
        for (int base = row0; base < row0 + nrows; base += 32 * TG) {
            uint32_t m[TG];
#pragma unroll
            for (int g = 0; g < TG; ++g) m[g] = is_or ? 0u : 0xffffffffu;
#pragma unroll
            for (int k = 0; k < SCAN_UNIT_TERMS; ++k) {
                if (k >= nt) break;
                if (kind[k] == TK_ISNULL || kind[k] == TK_NOTNULL) {
                    const uint32_t* bits = reinterpret_cast<const uint32_t*>(stage + (bits_off[k] != 0xffffffffu ? bits_off[k] : 0u)) + (base >> 5);
#pragma unroll
                    for (int g = 0; g < TG; ++g) {
                        const uint32_t valid = bits_off[k] != 0xffffffffu ? bits[g] : 0xffffffffu;
                        const uint32_t tm = kind[k] == TK_ISNULL ? ~valid : valid;
                        m[g] = is_or ? (m[g] | tm) : (m[g] & tm);
                    }
                } else if (kind[k] == TK_F64) {
                    term_pass<TK_F64>(cmp[k], imm[k], val_off[k], bits_off[k], stage, base, lane, is_or, m);
                } else if (kind[k] == TK_I64) {
                    term_pass<TK_I64>(cmp[k], imm[k], val_off[k], bits_off[k], stage, base, lane, is_or, m);
                } else {
                    term_pass<TK_I64_AS_F64>(cmp[k], imm[k], val_off[k], bits_off[k], stage, base, lane, is_or, m);
                }
            }
#pragma unroll
            for (int g = 0; g < TG; ++g) cnt += __popc(m[g] & tail_mask(base + 32 * g, rows));
        }
    }
    __device__ __forceinline__ void end(uint64_t* st) {
        if (lane == 0) st[0] = cnt;
    }
};

// ---- general predicate evaluator: per-lane numeric temporaries, warp-uniform NULL/TRUE/FALSE masks ----
constexpr int PG = PRED_GROUPS;

struct NumVal {
    uint64_t v[PG];    // per-lane payload, group g = tile row (base + 32 g + lane)
    uint32_t nul[PG];  // warp-uniform NULL masks
};
struct BoolVal {
    uint32_t t[PG], f[PG];  // warp-uniform TRUE / FALSE masks (neither bit set = NULL)
};

__device__ __forceinline__ void fetch_num(const ScanTables& P, const uint8_t* stage, uint8_t kind, uint16_t idx,
                                          uint64_t imm, const NumVal (&nt)[4], int row, NumVal& out) {
    if (kind == PK_NTEMP) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k == idx) out = nt[k];
    } else if (kind == PK_IMM) {
#pragma unroll
        for (int g = 0; g < PG; ++g) {
            out.v[g] = imm;
            out.nul[g] = 0u;
        }
    } else if (kind == PK_NNULL) {
#pragma unroll
        for (int g = 0; g < PG; ++g) {
            out.v[g] = 0;
            out.nul[g] = 0xffffffffu;
        }
    } else {
        const ScanColDesc& c = P.cols[idx];
        const uint64_t* vals = reinterpret_cast<const uint64_t*>(stage + c.smem_val_off) + row;
#pragma unroll
        for (int g = 0; g < PG; ++g) {
            const uint64_t raw = vals[32 * g];
            out.v[g] = (kind == PK_COL_I64_AS_F64) ? d2u((double)(int64_t)raw) : raw;
            out.nul[g] = ~vword(c, stage, (row >> 5) + g);
        }
    }
}

__device__ __forceinline__ void fetch_bool(const ScanTables& P, const uint8_t* stage, uint8_t kind, uint16_t idx,
                                           uint64_t imm, const BoolVal (&bt)[4], int row, BoolVal& out) {
    if (kind == PK_BTEMP) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k == idx) out = bt[k];
    } else if (kind == PK_BIMM) {
#pragma unroll
        for (int g = 0; g < PG; ++g) {
            out.t[g] = imm ? 0xffffffffu : 0u;
            out.f[g] = imm ? 0u : 0xffffffffu;
        }
    } else if (kind == PK_COL_BOOL) {
        const ScanColDesc& c = P.cols[idx];
        const uint32_t* vals = reinterpret_cast<const uint32_t*>(stage + c.smem_val_off);
#pragma unroll
        for (int g = 0; g < PG; ++g) {
            const uint32_t ok = vword(c, stage, (row >> 5) + g), b = vals[(row >> 5) + g];
            out.t[g] = b & ok;
            out.f[g] = ~b & ok;
        }
    } else {  // PK_BNULL / PK_NONE
#pragma unroll
        for (int g = 0; g < PG; ++g) out.t[g] = out.f[g] = 0u;
    }
}

__device__ __noinline__ void unit_pred(const ScanTables& P, const ScanUnitDesc& u, const uint8_t* stage, uint64_t* st,
                                       int lane, int rows_in_tile) {
    uint64_t cnt = 0;
    uint32_t div0 = 0;
    const int nw = u.nrows >> 5;
    for (int j = 0; j < nw; j += PG) {
        NumVal nt[4];
        BoolVal bt[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int g = 0; g < PG; ++g) {
                nt[k].v[g] = 0;
                nt[k].nul[g] = 0xffffffffu;
                bt[k].t[g] = bt[k].f[g] = 0u;
            }
        const int row = u.row0 + 32 * j + lane;  // this lane's row in group 0
        for (int pc = 0; pc < u.code_len; ++pc) {
            const PredInstr ins = P.code[u.code_off + pc];
            if (ins.op < PO_ISNULL_N) {
                NumVal a, b;
                fetch_num(P, stage, ins.a_kind, ins.a_idx, ins.imm, nt, row, a);
                if (ins.b_kind != PK_NONE) fetch_num(P, stage, ins.b_kind, ins.b_idx, ins.imm, nt, row, b);
                else b = a;
                if (ins.op < PO_EQ_F) {
                    NumVal r;
#pragma unroll
                    for (int g = 0; g < PG; ++g) {
                        const double af = u2d(a.v[g]), bf = u2d(b.v[g]);
                        const int64_t ai = (int64_t)a.v[g], bi = (int64_t)b.v[g];
                        uint32_t nm = a.nul[g] | b.nul[g];
                        uint64_t o = 0;
                        switch (ins.op) {
                            case PO_MOVN: o = a.v[g]; break;
                            case PO_ADD_F: o = d2u(af + bf); break;
                            case PO_SUB_F: o = d2u(af - bf); break;
                            case PO_MUL_F: o = d2u(af * bf); break;
                            case PO_DIV_F: o = d2u(af / bf); break;
                            case PO_NEG_F: o = d2u(-af); break;
                            case PO_ABS_F: o = d2u(fabs(af)); break;
                            case PO_ADD_I: o = (uint64_t)ai + (uint64_t)bi; break;
                            case PO_SUB_I: o = (uint64_t)ai - (uint64_t)bi; break;
                            case PO_MUL_I: o = (uint64_t)ai * (uint64_t)bi; break;
                            case PO_DIV_I:
                            case PO_MOD_I: {
                                const uint32_t z = __ballot_sync(0xffffffffu, bi == 0) & ~nm;
                                div0 |= z;
                                nm |= z;
                                const int64_t sb = bi == 0 ? 1 : bi;
                                if (ins.op == PO_DIV_I) o = sb == -1 ? (uint64_t)0 - (uint64_t)ai : (uint64_t)(ai / sb);
                                else o = sb == -1 ? 0 : (uint64_t)(ai % sb);
                            } break;
                            case PO_NEG_I: o = (uint64_t)0 - (uint64_t)ai; break;
                            case PO_ABS_I: o = ai < 0 ? (uint64_t)0 - (uint64_t)ai : (uint64_t)ai; break;
                            case PO_I2F: o = d2u((double)ai); break;
                            default: break;
                        }
                        r.v[g] = o;
                        r.nul[g] = nm;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k == ins.dst) nt[k] = r;
                } else {
                    BoolVal r;
#pragma unroll
                    for (int g = 0; g < PG; ++g) {
                        const double af = u2d(a.v[g]), bf = u2d(b.v[g]);
                        const int64_t ai = (int64_t)a.v[g], bi = (int64_t)b.v[g];
                        bool c = false;
                        switch (ins.op) {
                            case PO_EQ_F: c = af == bf; break;
                            case PO_NE_F: c = af != bf; break;
                            case PO_LT_F: c = af < bf; break;
                            case PO_LE_F: c = af <= bf; break;
                            case PO_GT_F: c = af > bf; break;
                            case PO_GE_F: c = af >= bf; break;
                            case PO_EQ_I: c = ai == bi; break;
                            case PO_NE_I: c = ai != bi; break;
                            case PO_LT_I: c = ai < bi; break;
                            case PO_LE_I: c = ai <= bi; break;
                            case PO_GT_I: c = ai > bi; break;
                            case PO_GE_I: c = ai >= bi; break;
                            default: break;
                        }
                        const uint32_t m = __ballot_sync(0xffffffffu, c), known = ~(a.nul[g] | b.nul[g]);
                        r.t[g] = m & known;
                        r.f[g] = ~m & known;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k == ins.dst) bt[k] = r;
                }
            } else if (ins.op <= PO_ISNOTNULL_N) {
                NumVal a;
                fetch_num(P, stage, ins.a_kind, ins.a_idx, ins.imm, nt, row, a);
                BoolVal r;
#pragma unroll
                for (int g = 0; g < PG; ++g) {
                    r.t[g] = ins.op == PO_ISNULL_N ? a.nul[g] : ~a.nul[g];
                    r.f[g] = ~r.t[g];
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k == ins.dst) bt[k] = r;
            } else {
                BoolVal a, b, r;
                fetch_bool(P, stage, ins.a_kind, ins.a_idx, ins.imm, bt, row, a);
                fetch_bool(P, stage, ins.b_kind, ins.b_idx, ins.imm, bt, row, b);
#pragma unroll
                for (int g = 0; g < PG; ++g) {
                    const uint32_t at = a.t[g], af = a.f[g], btt = b.t[g], bf = b.f[g];
                    uint32_t t = 0, f = 0;
                    switch (ins.op) {
                        case PO_MOVB: t = at; f = af; break;
                        case PO_AND: t = at & btt; f = af | bf; break;
                        case PO_OR: t = at | btt; f = af & bf; break;
                        case PO_NOT: t = af; f = at; break;
                        case PO_ISNULL_B: t = ~(at | af); f = at | af; break;
                        case PO_ISNOTNULL_B: t = at | af; f = ~(at | af); break;
                        case PO_ISTRUE: t = at; f = ~at; break;
                        case PO_ISFALSE: t = af; f = ~af; break;
                        case PO_EQ_B: t = (at & btt) | (af & bf); f = (at & bf) | (af & btt); break;
                        case PO_NE_B: t = (at & bf) | (af & btt); f = (at & btt) | (af & bf); break;
                        default: break;
                    }
                    r.t[g] = t;
                    r.f[g] = f;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k == ins.dst) bt[k] = r;
            }
        }
#pragma unroll
        for (int g = 0; g < PG; ++g) cnt += __popc(bt[0].t[g] & tail_mask(u.row0 + 32 * (j + g), rows_in_tile));
    }
    // masks are warp-uniform: lane 0 carries the counts
    if (lane == 0) {
        st[0 * 32] += cnt;
        st[1 * 32] |= (uint64_t)(div0 != 0);
    }
}

// reduce op of a state slot: 0 u64 add, 1 f64 add, 2 f64 min, 3 f64 max, 4 i64 min, 5 i64 max, 6 or
__host__ __device__ __forceinline__ int slot_op(int kind, int slot) {
    switch (kind) {
        case UNIT_COUNT: return 0;
        case UNIT_TERMS: return 0;
        case UNIT_PRED: return slot == 1 ? 6 : 0;
        case UNIT_PAIR: return slot == P_N ? 0 : 1;
        case UNIT_NUM_F64:
            if (slot == S_N || slot == S_ISUM) return 0;
            if (slot == S_MIN) return 2;
            if (slot == S_MAX) return 3;
            return 1;
        case UNIT_NUM_I64:
            if (slot == S_N || slot == S_ISUM) return 0;
            if (slot == S_MIN) return 4;
            if (slot == S_MAX) return 5;
            return 1;
    }
    return 0;
}
__device__ __forceinline__ uint64_t slot_combine(int op, uint64_t a, uint64_t b) {
    switch (op) {
        case 0: return a + b;
        case 1: return d2u(u2d(a) + u2d(b));
        case 2: return d2u(fmin(u2d(a), u2d(b)));
        case 3: return d2u(fmax(u2d(a), u2d(b)));
        case 4: return (uint64_t)min((int64_t)a, (int64_t)b);
        case 5: return (uint64_t)max((int64_t)a, (int64_t)b);
        default: return a | b;
    }
}
__host__ __device__ __forceinline__ uint64_t slot_identity(int kind, int slot) {
    int op = slot_op(kind, slot);
    switch (op) {
        case 2: return 0x7ff0000000000000ull;  // +inf
        case 3: return 0xfff0000000000000ull;  // -inf
        case 4: return (uint64_t)INT64_MAX;
        case 5: return (uint64_t)INT64_MIN;
        default: return 0;
    }
}

struct PredUnit {
    const ScanTables* T;
    const ScanUnitDesc* u;
    uint64_t* st;
    int lane;
    __device__ __forceinline__ void begin(const ScanTables& T_, const ScanUnitDesc& u_, const uint64_t* st_, int lane_) {
        T = &T_;
        u = &u_;
        st = const_cast<uint64_t*>(st_);
        lane = lane_;
    }
    __device__ __forceinline__ void tile(const uint8_t* stage, int rows, bool) { unit_pred(*T, *u, stage, st, lane, rows); }
    __device__ __forceinline__ void end(uint64_t*) {}
};

// consumer-side view of the stage ring
struct Pipe {
    uint8_t* stages;
    uint64_t* full;
    uint64_t* empty;
    int n_stages, tile_rows;
    uint32_t stage_bytes;
    int64_t n_tiles, n_rows;
};

// a warp that owns exactly one unit: descriptors and accumulators live in registers for the whole scan
template <class U>
__device__ __forceinline__ void run_single(const Pipe& p, const ScanTables& T, const ScanUnitDesc& ud, uint64_t* st, int lane) {
    U unit;
    unit.begin(T, ud, st, lane);
    int s = 0;
    uint32_t ph = 0;
    for (int64_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        mbar_wait(&p.full[s], ph);
        const int64_t row_base = t * p.tile_rows;
        const int rows = (int)min((int64_t)p.tile_rows, p.n_rows - row_base);
        unit.tile(p.stages + (size_t)s * p.stage_bytes, rows, rows < p.tile_rows);
        __syncwarp();
        if (lane == 0) mbar_arrive(&p.empty[s]);
        if (++s == p.n_stages) {
            s = 0;
            ph ^= 1u;
        }
    }
    unit.end(st);
}
template <class U>
__device__ __forceinline__ void run_once(const ScanTables& T, const ScanUnitDesc& ud, uint64_t* st, int lane, const uint8_t* stage,
                                         int rows, bool partial) {
    U unit;
    unit.begin(T, ud, st, lane);
    unit.tile(stage, rows, partial);
    unit.end(st);
}

// kind / flag dispatch, shared by both execution modes (SINGLE: whole tile loop inside; else one tile)
template <bool SINGLE, class U>
__device__ __forceinline__ void run_unit(const Pipe& p, const ScanTables& T, const ScanUnitDesc& ud, uint64_t* st, int lane,
                                         const uint8_t* stage, int rows, bool partial) {
    if (SINGLE) run_single<U>(p, T, ud, st, lane);
    else run_once<U>(T, ud, st, lane, stage, rows, partial);
}
template <bool SINGLE, bool IS_I64>
__device__ __forceinline__ void dispatch_num(const Pipe& p, const ScanTables& T, const ScanUnitDesc& ud, uint64_t* st, int lane,
                                             const uint8_t* stage, int rows, bool partial) {
    if (!T.cols[ud.c0].pivot_is_element) {
        run_unit<SINGLE, NumSlowUnit<IS_I64>>(p, T, ud, st, lane, stage, rows, partial);
        return;
    }
    switch (ud.flags & 7) {
        case 0: run_unit<SINGLE, NumUnit<IS_I64, 0>>(p, T, ud, st, lane, stage, rows, partial); break;
        case 1: run_unit<SINGLE, NumUnit<IS_I64, 1>>(p, T, ud, st, lane, stage, rows, partial); break;
        case 2: run_unit<SINGLE, NumUnit<IS_I64, 2>>(p, T, ud, st, lane, stage, rows, partial); break;
        case 3: run_unit<SINGLE, NumUnit<IS_I64, 3>>(p, T, ud, st, lane, stage, rows, partial); break;
        case 4: run_unit<SINGLE, NumUnit<IS_I64, 4>>(p, T, ud, st, lane, stage, rows, partial); break;
        case 5: run_unit<SINGLE, NumUnit<IS_I64, 5>>(p, T, ud, st, lane, stage, rows, partial); break;
        case 6: run_unit<SINGLE, NumUnit<IS_I64, 6>>(p, T, ud, st, lane, stage, rows, partial); break;
        default: run_unit<SINGLE, NumUnit<IS_I64, 7>>(p, T, ud, st, lane, stage, rows, partial); break;
    }
}
template <bool SINGLE>
__device__ __forceinline__ void dispatch_unit(const Pipe& p, const ScanTables& T, const ScanUnitDesc& ud, uint64_t* st, int lane,
                                              const uint8_t* stage, int rows, bool partial) {
    switch (ud.kind) {
        case UNIT_COUNT: run_unit<SINGLE, CountUnit>(p, T, ud, st, lane, stage, rows, partial); break;
        case UNIT_NUM_F64: dispatch_num<SINGLE, false>(p, T, ud, st, lane, stage, rows, partial); break;
        case UNIT_NUM_I64: dispatch_num<SINGLE, true>(p, T, ud, st, lane, stage, rows, partial); break;
        case UNIT_PAIR:
            switch ((ud.c0_is_i64 ? 1 : 0) | (ud.c1_is_i64 ? 2 : 0)) {
                case 0: run_unit<SINGLE, PairUnit<false, false>>(p, T, ud, st, lane, stage, rows, partial); break;
                case 1: run_unit<SINGLE, PairUnit<true, false>>(p, T, ud, st, lane, stage, rows, partial); break;
                case 2: run_unit<SINGLE, PairUnit<false, true>>(p, T, ud, st, lane, stage, rows, partial); break;
                default: run_unit<SINGLE, PairUnit<true, true>>(p, T, ud, st, lane, stage, rows, partial); break;
            }
            break;
        case UNIT_TERMS: run_unit<SINGLE, TermsUnit>(p, T, ud, st, lane, stage, rows, partial); break;
        default: run_unit<SINGLE, PredUnit>(p, T, ud, st, lane, stage, rows, partial); break;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS, 1) scan_kernel(const __grid_constant__ ScanParams P) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* stages = scan_smem;
    uint64_t* state = reinterpret_cast<uint64_t*>(scan_smem + (size_t)P.n_stages * P.stage_bytes);
    uint64_t* full = state + (size_t)P.n_units * SCAN_STATE_SLOTS * 32;
    uint64_t* empty = full + P.n_stages;
    ScanTables* T = reinterpret_cast<ScanTables*>(empty + P.n_stages);

    // descriptor tables -> shared memory (only the used prefix of each array)
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&P.tab);
        uint32_t* dst = reinterpret_cast<uint32_t*>(T);
        auto copy = [&](size_t off, size_t bytes) {
            for (uint32_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) dst[off / 4 + i] = src[off / 4 + i];
        };
        copy(offsetof(ScanTables, cols), sizeof(ScanColDesc) * P.n_cols);
        copy(offsetof(ScanTables, units), sizeof(ScanUnitDesc) * P.n_units);
        copy(offsetof(ScanTables, code), sizeof(PredInstr) * P.n_code);
        copy(offsetof(ScanTables, terms), sizeof(ScanTerm) * P.n_terms);
        copy(offsetof(ScanTables, warp_units), sizeof(T->warp_units));
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < P.n_stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], SCAN_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();
    Pipe pipe{stages, full, empty, P.n_stages, P.tile_rows, P.stage_bytes, P.n_tiles, P.n_rows};

    if (warp == 0) {
        // ---------------- TMA producer: lane c moves column c ----------------
        const bool active = lane < P.n_cols;
        const ScanColDesc c = T->cols[active ? lane : 0];
        const bool has_v = active && c.values != nullptr, has_b = active && c.validity != nullptr;
        int s = 0;
        uint32_t ph = 0;
        for (int64_t t = blockIdx.x; t < pipe.n_tiles; t += gridDim.x) {
            mbar_wait(&empty[s], ph ^ 1u);
            const int64_t row_base = t * pipe.tile_rows;
            const int rows = (int)min((int64_t)pipe.tile_rows, pipe.n_rows - row_base);
            uint8_t* stage = stages + (size_t)s * pipe.stage_bytes;
            // byte counts padded to 16 (buffers are allocated with that slack)
            const uint32_t bbytes = has_b ? (uint32_t)(((rows + 7) / 8 + 15) & ~15) : 0u;
            const uint32_t vbytes = !has_v ? 0u : (c.kind == SC_BOOL ? (uint32_t)(((rows + 7) / 8 + 15) & ~15)
                                                                      : (uint32_t)((rows * 8 + 15) & ~15));
            uint32_t total = vbytes + bbytes;
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) total += __shfl_xor_sync(0xffffffffu, total, m);
            if (lane == 0) mbar_arrive_expect_tx(&full[s], total);
            __syncwarp();
            if (vbytes) {
                const uint8_t* src = c.kind == SC_BOOL ? c.values + row_base / 8 : c.values + row_base * 8;
                bulk_g2s(stage + c.smem_val_off, src, vbytes, &full[s]);
            }
            if (bbytes) bulk_g2s(stage + c.smem_bits_off, c.validity + row_base / 8, bbytes, &full[s]);
            if (++s == pipe.n_stages) {
                s = 0;
                ph ^= 1u;
            }
        }
    } else {
        // ---------------- consumers ----------------
        const int cw = warp - 1;
        const int n_mine = T->warp_units[cw][0];
        for (int k = 0; k < n_mine; ++k) {
            const int u = T->warp_units[cw][1 + k];
            unit_init(T->units[u], state + (size_t)u * SCAN_STATE_SLOTS * 32, lane);
        }
        __syncwarp();
        if (n_mine == 1) {
            const int u = T->warp_units[cw][1];
            dispatch_unit<true>(pipe, *T, T->units[u], state + (size_t)u * SCAN_STATE_SLOTS * 32, lane, nullptr, 0, false);
        } else {
            int s = 0;
            uint32_t ph = 0;
            for (int64_t t = blockIdx.x; t < pipe.n_tiles; t += gridDim.x) {
                mbar_wait(&full[s], ph);
                const int64_t row_base = t * pipe.tile_rows;
                const int rows = (int)min((int64_t)pipe.tile_rows, pipe.n_rows - row_base);
                const uint8_t* stage = stages + (size_t)s * pipe.stage_bytes;
                for (int k = 0; k < n_mine; ++k) {
                    const int u = T->warp_units[cw][1 + k];
                    dispatch_unit<false>(pipe, *T, T->units[u], state + (size_t)u * SCAN_STATE_SLOTS * 32, lane, stage, rows,
                                         rows < pipe.tile_rows);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                if (++s == pipe.n_stages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
        __syncwarp();
        // cross-lane reduction in a fixed butterfly order, then one record per (CTA, unit)
        for (int k = 0; k < n_mine; ++k) {
            const int u = T->warp_units[cw][1 + k];
            const int kind = T->units[u].kind;
            const uint64_t* st = state + (size_t)u * SCAN_STATE_SLOTS * 32;
            uint64_t* out = P.partials + ((size_t)blockIdx.x * P.n_units + u) * SCAN_STATE_SLOTS;
#pragma unroll
            for (int sl = 0; sl < SCAN_STATE_SLOTS; ++sl) {
                const int op = slot_op(kind, sl);
                uint64_t v = st[sl * 32 + lane];
#pragma unroll
                for (int m = 16; m > 0; m >>= 1) v = slot_combine(op, v, shfl_xor_u64(v, m));
                if (lane == 0) out[sl] = v;
            }
        }
    }
}

// one warp per aggregate: reduce partials[cta][unit] over CTAs and over the units feeding that aggregate
__global__ void scan_finalize_kernel(const __grid_constant__ ScanParams P, int n_ctas, ScanAggOut* out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= P.n_aggs) return;
    int kind = -1;
    for (int u = 0; u < P.n_units; ++u)
        if (P.tab.units[u].agg == warp) kind = P.tab.units[u].kind;
    if (kind < 0) return;
    for (int k = 0; k < SCAN_STATE_SLOTS; ++k) {
        const int op = slot_op(kind, k);
        uint64_t acc = slot_identity(kind, k);
        for (int b = lane; b < n_ctas; b += 32) {
            for (int u = 0; u < P.n_units; ++u) {
                if (P.tab.units[u].agg != warp) continue;
                acc = slot_combine(op, acc, P.partials[((size_t)b * P.n_units + u) * SCAN_STATE_SLOTS + k]);
            }
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) acc = slot_combine(op, acc, shfl_xor_u64(acc, m));
        if (lane == 0) out[warp].s[k] = acc;
    }
}

// ---- host launchers (called from engine.cu) ----
size_t scan_smem_bytes(const ScanParams& P) {
    return (size_t)P.n_stages * P.stage_bytes + (size_t)P.n_units * SCAN_STATE_SLOTS * 32 * 8 +
           (size_t)2 * P.n_stages * 8 + sizeof(ScanTables);
}

cudaError_t scan_launch(const ScanParams& P, int grid, ScanAggOut* d_out, cudaStream_t stream) {
    static bool attr_set = false;
    size_t smem = scan_smem_bytes(P);
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    scan_kernel<<<grid, SCAN_THREADS, smem, stream>>>(P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    int warps = P.n_aggs;
    int threads = 128;
    int blocks = (warps * 32 + threads - 1) / threads;
    scan_finalize_kernel<<<blocks, threads, 0, stream>>>(P, grid, d_out);
    return cudaGetLastError();
}

}  // namespace tg
