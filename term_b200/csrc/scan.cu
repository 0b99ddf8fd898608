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
//   * moments are accumulated as shifted sums Σ(x-K), Σ(x-K)² (K = a value of the column), which
//     merge by plain addition across lanes/CTAs/GPUs and lose at most ~n·eps relative accuracy.
//   * per-CTA partials go to global memory; scan_finalize_kernel reduces them in a fixed order, so a
//     given (grid, plan) is bit-reproducible run to run.
#include <cfloat>
#include <math_constants.h>
#include <cstdio>

#include "ptx.cuh"
#include "scan_defs.h"

namespace tg {

extern __shared__ __align__(128) uint8_t scan_smem[];

__device__ __forceinline__ uint32_t tail_mask(int base_row, int rows_in_tile) {
    int rem = rows_in_tile - base_row;
    return rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
}

// ---- per-lane state in shared memory: state[unit][slot][lane] ----
struct LaneState {
    uint64_t s[SCAN_STATE_SLOTS];
};
__device__ __forceinline__ void state_load(const uint64_t* st, int lane, LaneState& v, int nslots) {
#pragma unroll
    for (int k = 0; k < SCAN_STATE_SLOTS; ++k)
        if (k < nslots) v.s[k] = st[k * 32 + lane];
}
__device__ __forceinline__ void state_store(uint64_t* st, int lane, const LaneState& v, int nslots) {
#pragma unroll
    for (int k = 0; k < SCAN_STATE_SLOTS; ++k)
        if (k < nslots) st[k * 32 + lane] = v.s[k];
}

__device__ __forceinline__ double u2d(uint64_t u) { return __longlong_as_double((long long)u); }
__device__ __forceinline__ uint64_t d2u(double d) { return (uint64_t)__double_as_longlong(d); }

__device__ void unit_init(const ScanUnitDesc& u, uint64_t* st, int lane) {
    LaneState v;
#pragma unroll
    for (int k = 0; k < SCAN_STATE_SLOTS; ++k) v.s[k] = 0;
    if (u.kind == UNIT_NUM_F64) {
        v.s[S_MIN] = d2u(CUDART_INF);
        v.s[S_MAX] = d2u(-CUDART_INF);
    } else if (u.kind == UNIT_NUM_I64) {
        v.s[S_MIN] = (uint64_t)INT64_MAX;
        v.s[S_MAX] = (uint64_t)INT64_MIN;
    }
    state_store(st, lane, v, SCAN_STATE_SLOTS);
}

// validity word j (32 rows) of a column inside the stage; all-ones when the column has no bitmap
__device__ __forceinline__ uint32_t vword(const ScanColDesc& c, const uint8_t* stage, int word) {
    return c.validity ? reinterpret_cast<const uint32_t*>(stage + c.smem_bits_off)[word] : 0xffffffffu;
}

__device__ void unit_count(const ScanParams& P, const ScanUnitDesc& u, const uint8_t* stage, uint64_t* st,
                           int lane, int rows_in_tile) {
    const ScanColDesc& c = P.cols[u.c0];
    uint64_t n = st[S_N * 32 + lane];
    const int w0 = u.row0 >> 5, nw = u.nrows >> 5;
    for (int j = lane; j < nw; j += 32) {
        uint32_t w = vword(c, stage, w0 + j) & tail_mask(u.row0 + 32 * j, rows_in_tile);
        n += __popc(w);
    }
    st[S_N * 32 + lane] = n;
}

template <bool IS_I64>
__device__ void unit_num(const ScanParams& P, const ScanUnitDesc& u, const uint8_t* stage, uint64_t* st,
                         int lane, int rows_in_tile, bool partial) {
    const ScanColDesc& c = P.cols[u.c0];
    const double K = c.pivot;
    LaneState v;
    state_load(st, lane, v, 7);
    uint64_t n = v.s[S_N];
    double sd = u2d(v.s[S_SD]), sdd = u2d(v.s[S_SDD]), sx = u2d(v.s[S_SX]);
    double fmn = u2d(v.s[S_MIN]), fmx = u2d(v.s[S_MAX]);
    int64_t imn = (int64_t)v.s[S_MIN], imx = (int64_t)v.s[S_MAX];
    uint64_t isum = v.s[S_ISUM];
    const int w0 = u.row0 >> 5, nw = u.nrows >> 5;
    const uint64_t* vals = reinterpret_cast<const uint64_t*>(stage + c.smem_val_off) + u.row0 + lane;
#pragma unroll 4
    for (int j = 0; j < nw; ++j) {
        uint32_t w = vword(c, stage, w0 + j);
        if (partial) w &= tail_mask(u.row0 + 32 * j, rows_in_tile);
        const bool ok = (w >> lane) & 1u;
        const uint64_t raw = vals[32 * j];
        n += ok;
        if (IS_I64) {
            const int64_t xi = (int64_t)raw;
            const double x = (double)xi;
            const double d = ok ? x - K : 0.0;
            sd += d;
            sdd = fma(d, d, sdd);
            sx += ok ? x : 0.0;
            isum += ok ? (uint64_t)xi : 0ull;
            imn = min(imn, ok ? xi : INT64_MAX);
            imx = max(imx, ok ? xi : INT64_MIN);
        } else {
            const double x = u2d(raw);
            const double d = ok ? x - K : 0.0;
            sd += d;
            sdd = fma(d, d, sdd);
            sx += ok ? x : 0.0;
            fmn = fmin(fmn, ok ? x : CUDART_INF);
            fmx = fmax(fmx, ok ? x : -CUDART_INF);
        }
    }
    v.s[S_N] = n;
    v.s[S_SD] = d2u(sd);
    v.s[S_SDD] = d2u(sdd);
    v.s[S_SX] = d2u(sx);
    if (IS_I64) {
        v.s[S_MIN] = (uint64_t)imn;
        v.s[S_MAX] = (uint64_t)imx;
        v.s[S_ISUM] = isum;
    } else {
        v.s[S_MIN] = d2u(fmn);
        v.s[S_MAX] = d2u(fmx);
    }
    state_store(st, lane, v, 7);
}

__device__ void unit_pair(const ScanParams& P, const ScanUnitDesc& u, const uint8_t* stage, uint64_t* st,
                          int lane, int rows_in_tile, bool partial) {
    const ScanColDesc& cx = P.cols[u.c0];
    const ScanColDesc& cy = P.cols[u.c1];
    const double Kx = cx.pivot, Ky = cy.pivot;
    LaneState v;
    state_load(st, lane, v, 6);
    uint64_t n = v.s[P_N];
    double sx = u2d(v.s[P_SX]), sy = u2d(v.s[P_SY]), sxx = u2d(v.s[P_SXX]), syy = u2d(v.s[P_SYY]),
           sxy = u2d(v.s[P_SXY]);
    const int w0 = u.row0 >> 5, nw = u.nrows >> 5;
    const uint64_t* vx = reinterpret_cast<const uint64_t*>(stage + cx.smem_val_off) + u.row0 + lane;
    const uint64_t* vy = reinterpret_cast<const uint64_t*>(stage + cy.smem_val_off) + u.row0 + lane;
    const bool xi = u.c0_is_i64, yi = u.c1_is_i64;
#pragma unroll 4
    for (int j = 0; j < nw; ++j) {
        uint32_t w = vword(cx, stage, w0 + j) & vword(cy, stage, w0 + j);
        if (partial) w &= tail_mask(u.row0 + 32 * j, rows_in_tile);
        const bool ok = (w >> lane) & 1u;
        const uint64_t rx = vx[32 * j], ry = vy[32 * j];
        const double x = xi ? (double)(int64_t)rx : u2d(rx);
        const double y = yi ? (double)(int64_t)ry : u2d(ry);
        const double dx = ok ? x - Kx : 0.0, dy = ok ? y - Ky : 0.0;
        n += ok;
        sx += dx;
        sy += dy;
        sxx = fma(dx, dx, sxx);
        syy = fma(dy, dy, syy);
        sxy = fma(dx, dy, sxy);
    }
    v.s[P_N] = n;
    v.s[P_SX] = d2u(sx);
    v.s[P_SY] = d2u(sy);
    v.s[P_SXX] = d2u(sxx);
    v.s[P_SYY] = d2u(syy);
    v.s[P_SXY] = d2u(sxy);
    state_store(st, lane, v, 6);
}

// ---- predicate interpreter: 4 temporaries x PR rows per lane, SQL three-valued logic ----
constexpr int PR = 2;

struct PVal {
    uint64_t v[PR];
    bool nul[PR];
};

__device__ __forceinline__ void pred_fetch(const ScanParams& P, const uint8_t* stage, uint8_t kind, uint16_t idx,
                                           uint64_t imm, const uint64_t (&t)[4][PR], const bool (&tn)[4][PR],
                                           int row, int lane, PVal& out) {
    if (kind == PK_TEMP) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k == idx) {
#pragma unroll
                for (int r = 0; r < PR; ++r) {
                    out.v[r] = t[k][r];
                    out.nul[r] = tn[k][r];
                }
            }
    } else if (kind == PK_IMM) {
#pragma unroll
        for (int r = 0; r < PR; ++r) {
            out.v[r] = imm;
            out.nul[r] = false;
        }
    } else if (kind == PK_NULL) {
#pragma unroll
        for (int r = 0; r < PR; ++r) {
            out.v[r] = 0;
            out.nul[r] = true;
        }
    } else {
        const ScanColDesc& c = P.cols[idx];
#pragma unroll
        for (int r = 0; r < PR; ++r) {
            const int rr = row + 32 * r;  // tile row of this lane
            const uint32_t w = vword(c, stage, rr >> 5);
            out.nul[r] = !((w >> lane) & 1u);
            if (kind == PK_COL_BOOL) {
                const uint32_t bw = reinterpret_cast<const uint32_t*>(stage + c.smem_val_off)[rr >> 5];
                out.v[r] = (bw >> lane) & 1u;
            } else {
                const uint64_t raw = reinterpret_cast<const uint64_t*>(stage + c.smem_val_off)[rr];
                out.v[r] = (kind == PK_COL_I64_AS_F64) ? d2u((double)(int64_t)raw) : raw;
            }
        }
    }
}

__device__ void unit_pred(const ScanParams& P, const ScanUnitDesc& u, const uint8_t* stage, uint64_t* st,
                          int lane, int rows_in_tile) {
    uint64_t cnt = st[0 * 32 + lane];
    uint64_t div0 = st[1 * 32 + lane];
    const int nw = u.nrows >> 5;
    for (int j = 0; j < nw; j += PR) {
        uint64_t t[4][PR];
        bool tn[4][PR];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int r = 0; r < PR; ++r) {
                t[k][r] = 0;
                tn[k][r] = true;
            }
        const int row = u.row0 + 32 * j + lane;
        for (int pc = 0; pc < u.code_len; ++pc) {
            const PredInstr ins = P.code[u.code_off + pc];
            PVal a, b;
            pred_fetch(P, stage, ins.a_kind, ins.a_idx, ins.imm, t, tn, row, lane, a);
            pred_fetch(P, stage, ins.b_kind, ins.b_idx, ins.imm, t, tn, row, lane, b);
            uint64_t res[PR];
            bool rn[PR];
#pragma unroll
            for (int r = 0; r < PR; ++r) {
                const double af = u2d(a.v[r]), bf = u2d(b.v[r]);
                const int64_t ai = (int64_t)a.v[r], bi = (int64_t)b.v[r];
                const bool an = a.nul[r], bn = b.nul[r];
                uint64_t o = 0;
                bool on = an || bn;
                switch (ins.op) {
                    case PO_MOV: o = a.v[r]; on = an; break;
                    case PO_ADD_F: o = d2u(af + bf); break;
                    case PO_SUB_F: o = d2u(af - bf); break;
                    case PO_MUL_F: o = d2u(af * bf); break;
                    case PO_DIV_F: o = d2u(af / bf); break;
                    case PO_NEG_F: o = d2u(-af); on = an; break;
                    case PO_ABS_F: o = d2u(fabs(af)); on = an; break;
                    case PO_ADD_I: o = (uint64_t)ai + (uint64_t)bi; break;
                    case PO_SUB_I: o = (uint64_t)ai - (uint64_t)bi; break;
                    case PO_MUL_I: o = (uint64_t)ai * (uint64_t)bi; break;
                    case PO_DIV_I:
                        if (!on && bi == 0) {
                            div0 = 1;
                            on = true;
                        } else if (!on) {
                            o = (bi == -1) ? (uint64_t)0 - (uint64_t)ai : (uint64_t)(ai / bi);
                        }
                        break;
                    case PO_MOD_I:
                        if (!on && bi == 0) {
                            div0 = 1;
                            on = true;
                        } else if (!on) {
                            o = (bi == -1) ? 0 : (uint64_t)(ai % bi);
                        }
                        break;
                    case PO_NEG_I: o = (uint64_t)0 - (uint64_t)ai; on = an; break;
                    case PO_ABS_I: o = ai < 0 ? (uint64_t)0 - (uint64_t)ai : (uint64_t)ai; on = an; break;
                    case PO_EQ_F: o = af == bf; break;
                    case PO_NE_F: o = af != bf; break;
                    case PO_LT_F: o = af < bf; break;
                    case PO_LE_F: o = af <= bf; break;
                    case PO_GT_F: o = af > bf; break;
                    case PO_GE_F: o = af >= bf; break;
                    case PO_EQ_I: o = ai == bi; break;
                    case PO_NE_I: o = ai != bi; break;
                    case PO_LT_I: o = ai < bi; break;
                    case PO_LE_I: o = ai <= bi; break;
                    case PO_GT_I: o = ai > bi; break;
                    case PO_GE_I: o = ai >= bi; break;
                    case PO_AND: {
                        const bool af0 = !an && a.v[r] == 0, bf0 = !bn && b.v[r] == 0;
                        if (af0 || bf0) { o = 0; on = false; }
                        else if (an || bn) { on = true; }
                        else { o = 1; on = false; }
                    } break;
                    case PO_OR: {
                        const bool at = !an && a.v[r] != 0, bt = !bn && b.v[r] != 0;
                        if (at || bt) { o = 1; on = false; }
                        else if (an || bn) { on = true; }
                        else { o = 0; on = false; }
                    } break;
                    case PO_NOT: o = a.v[r] == 0; on = an; break;
                    case PO_ISNULL: o = an; on = false; break;
                    case PO_ISNOTNULL: o = !an; on = false; break;
                    case PO_ISTRUE: o = !an && a.v[r] != 0; on = false; break;
                    case PO_ISFALSE: o = !an && a.v[r] == 0; on = false; break;
                    case PO_I2F: o = d2u((double)ai); on = an; break;
                    default: break;
                }
                res[r] = o;
                rn[r] = on;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k == ins.dst) {
#pragma unroll
                    for (int r = 0; r < PR; ++r) {
                        t[k][r] = res[r];
                        tn[k][r] = rn[r];
                    }
                }
        }
#pragma unroll
        for (int r = 0; r < PR; ++r) {
            const bool in_range = (row + 32 * r) < rows_in_tile;
            cnt += (in_range && !tn[0][r] && t[0][r] != 0) ? 1 : 0;
        }
    }
    st[0 * 32 + lane] = cnt;
    st[1 * 32 + lane] = div0;
}

// reduce op of a state slot: 0 u64 add, 1 f64 add, 2 f64 min, 3 f64 max, 4 i64 min, 5 i64 max, 6 or
__host__ __device__ __forceinline__ int slot_op(int kind, int slot) {
    switch (kind) {
        case UNIT_COUNT: return 0;
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
        case 2: return 0x7ff0000000000000ull;   // +inf
        case 3: return 0xfff0000000000000ull;   // -inf
        case 4: return (uint64_t)INT64_MAX;
        case 5: return (uint64_t)INT64_MIN;
        default: return 0;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS, 1) scan_kernel(const __grid_constant__ ScanParams P) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* stages = scan_smem;
    uint64_t* state = reinterpret_cast<uint64_t*>(scan_smem + (size_t)P.n_stages * P.stage_bytes);
    uint64_t* full = state + (size_t)P.n_units * SCAN_STATE_SLOTS * 32;
    uint64_t* empty = full + P.n_stages;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.n_stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], SCAN_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    if (warp > 0) {
        for (int u = 0; u < P.n_units; ++u)
            if (P.units[u].warp == warp - 1) unit_init(P.units[u], state + (size_t)u * SCAN_STATE_SLOTS * 32, lane);
    }
    __syncthreads();

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        int it = 0;
        for (int64_t t = blockIdx.x; t < P.n_tiles; t += gridDim.x, ++it) {
            const int s = it % P.n_stages;
            const uint32_t ph = (uint32_t)(it / P.n_stages) & 1u;
            mbar_wait(&empty[s], ph ^ 1u);
            const int64_t row_base = t * P.tile_rows;
            const int rows = (int)min((int64_t)P.tile_rows, P.n_rows - row_base);
            uint8_t* stage = stages + (size_t)s * P.stage_bytes;
            // lane c moves column c; bytes are padded to 16 (buffers are allocated with that slack)
            uint32_t vbytes = 0, bbytes = 0;
            if (lane < P.n_cols) {
                const ScanColDesc& c = P.cols[lane];
                if (c.values) {
                    if (c.kind == SC_BOOL) vbytes = (uint32_t)(((rows + 7) / 8 + 15) & ~15);
                    else vbytes = (uint32_t)((rows * 8 + 15) & ~15);
                }
                if (c.validity) bbytes = (uint32_t)(((rows + 7) / 8 + 15) & ~15);
            }
            uint32_t total = vbytes + bbytes;
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) total += __shfl_xor_sync(0xffffffffu, total, m);
            if (lane == 0) mbar_arrive_expect_tx(&full[s], total);
            __syncwarp();
            if (lane < P.n_cols) {
                const ScanColDesc& c = P.cols[lane];
                if (vbytes) {
                    const uint8_t* src = c.kind == SC_BOOL ? c.values + row_base / 8 : c.values + row_base * 8;
                    bulk_g2s(stage + c.smem_val_off, src, vbytes, &full[s]);
                }
                if (bbytes) bulk_g2s(stage + c.smem_bits_off, c.validity + row_base / 8, bbytes, &full[s]);
            }
        }
    } else {
        // ---------------- consumers ----------------
        const int cw = warp - 1;
        int it = 0;
        for (int64_t t = blockIdx.x; t < P.n_tiles; t += gridDim.x, ++it) {
            const int s = it % P.n_stages;
            const uint32_t ph = (uint32_t)(it / P.n_stages) & 1u;
            mbar_wait(&full[s], ph);
            const int64_t row_base = t * P.tile_rows;
            const int rows = (int)min((int64_t)P.tile_rows, P.n_rows - row_base);
            const bool partial = rows < P.tile_rows;
            const uint8_t* stage = stages + (size_t)s * P.stage_bytes;
            for (int u = 0; u < P.n_units; ++u) {
                const ScanUnitDesc& ud = P.units[u];
                if (ud.warp != cw) continue;
                if (ud.row0 >= rows) continue;
                uint64_t* st = state + (size_t)u * SCAN_STATE_SLOTS * 32;
                switch (ud.kind) {
                    case UNIT_COUNT: unit_count(P, ud, stage, st, lane, rows); break;
                    case UNIT_NUM_F64: unit_num<false>(P, ud, stage, st, lane, rows, partial); break;
                    case UNIT_NUM_I64: unit_num<true>(P, ud, stage, st, lane, rows, partial); break;
                    case UNIT_PAIR: unit_pair(P, ud, stage, st, lane, rows, partial); break;
                    case UNIT_PRED: unit_pred(P, ud, stage, st, lane, rows); break;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        // cross-lane reduction in a fixed butterfly order, then one record per (CTA, unit)
        for (int u = 0; u < P.n_units; ++u) {
            const ScanUnitDesc& ud = P.units[u];
            if (ud.warp != cw) continue;
            const uint64_t* st = state + (size_t)u * SCAN_STATE_SLOTS * 32;
            uint64_t* out = P.partials + ((size_t)blockIdx.x * P.n_units + u) * SCAN_STATE_SLOTS;
#pragma unroll
            for (int k = 0; k < SCAN_STATE_SLOTS; ++k) {
                const int op = slot_op(ud.kind, k);
                uint64_t v = st[k * 32 + lane];
#pragma unroll
                for (int m = 16; m > 0; m >>= 1) v = slot_combine(op, v, shfl_xor_u64(v, m));
                if (lane == 0) out[k] = v;
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
        if (P.units[u].agg == warp) kind = P.units[u].kind;
    if (kind < 0) return;
    for (int k = 0; k < SCAN_STATE_SLOTS; ++k) {
        const int op = slot_op(kind, k);
        uint64_t acc = slot_identity(kind, k);
        for (int b = lane; b < n_ctas; b += 32) {
            for (int u = 0; u < P.n_units; ++u) {
                if (P.units[u].agg != warp) continue;
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
           (size_t)2 * P.n_stages * 8;
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
