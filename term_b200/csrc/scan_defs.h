// Descriptors shared by the host planner and the fused numeric scan kernel (K1).
#pragma once
#include <stdint.h>

namespace tg {

constexpr int SCAN_MAX_COLS = 32;
constexpr int SCAN_MAX_UNITS = 64;
constexpr int SCAN_MAX_CODE = 192;
constexpr int SCAN_CONSUMER_WARPS = 16;
constexpr int SCAN_THREADS = (SCAN_CONSUMER_WARPS + 1) * 32;  // warp 0 = TMA producer
constexpr int SCAN_STATE_SLOTS = 8;                           // 64-bit slots per lane per unit
constexpr int SCAN_MAX_STAGES = 8;

// column kinds inside a tile
enum : int32_t { SC_F64 = 0, SC_I64 = 1, SC_BITS = 2 /* validity only */, SC_BOOL = 3 /* bit-packed values */ };

struct ScanColDesc {
    const uint8_t* values;    // device pointer or null (validity-only column)
    const uint8_t* validity;  // device pointer or null (no nulls)
    int32_t kind;
    uint32_t smem_val_off;   // byte offset of this column's values inside a stage
    uint32_t smem_bits_off;  // byte offset of its validity words inside a stage
    int32_t pad;
    double pivot;            // shift K for the moment sums of this column
};

enum : int32_t {
    UNIT_COUNT = 0,    // popcount of validity                -> slot0 = valid rows
    UNIT_NUM_F64 = 1,  // n, Σd, Σd², min, max, Σx
    UNIT_NUM_I64 = 2,  // n, Σd, Σd², imin, imax, isum(wrapping), Σ(double)x
    UNIT_PAIR = 3,     // n, Σdx, Σdy, Σdx², Σdy², Σdxdy over rows where both are valid
    UNIT_PRED = 4,     // slot0 = rows where predicate is TRUE, slot1 = integer division by zero seen
};

// slot meaning per unit kind (64-bit each)
enum : int32_t {
    S_N = 0,
    S_SD = 1,
    S_SDD = 2,
    S_MIN = 3,
    S_MAX = 4,
    S_SX = 5,
    S_ISUM = 6,
    // pair
    P_N = 0,
    P_SX = 1,
    P_SY = 2,
    P_SXX = 3,
    P_SYY = 4,
    P_SXY = 5,
};

struct ScanUnitDesc {
    int32_t kind;
    int32_t c0, c1;        // tile column indices
    int32_t row0, nrows;   // slice of the tile this unit covers (multiples of 64)
    int32_t agg;           // aggregate this unit accumulates into
    int32_t code_off, code_len;
    int32_t warp;          // owning consumer warp (0-based)
    int32_t c0_is_i64, c1_is_i64;  // pair units: convert on load
    int32_t pad;
};

// predicate 3-address code over 4 temporaries, SQL three-valued logic
enum : uint8_t {
    PK_TEMP = 0,
    PK_COL_F64 = 1,
    PK_COL_I64 = 2,
    PK_COL_I64_AS_F64 = 3,
    PK_IMM = 4,       // payload = imm bits (f64 or i64 or bool)
    PK_NULL = 5,
    PK_COL_BOOL = 6,
};
enum : uint8_t {
    PO_MOV = 0,
    PO_ADD_F, PO_SUB_F, PO_MUL_F, PO_DIV_F, PO_NEG_F,
    PO_ADD_I, PO_SUB_I, PO_MUL_I, PO_DIV_I, PO_MOD_I, PO_NEG_I,
    PO_EQ_F, PO_NE_F, PO_LT_F, PO_LE_F, PO_GT_F, PO_GE_F,
    PO_EQ_I, PO_NE_I, PO_LT_I, PO_LE_I, PO_GT_I, PO_GE_I,
    PO_AND, PO_OR, PO_NOT,
    PO_ISNULL, PO_ISNOTNULL,
    PO_I2F,
    PO_ABS_F, PO_ABS_I,
    PO_ISTRUE, PO_ISFALSE,
};
struct PredInstr {
    uint8_t op;
    uint8_t dst;     // temp index 0..3
    uint8_t a_kind, b_kind;
    uint16_t a_idx, b_idx;  // temp index or tile column index
    uint64_t imm;           // shared immediate (at most one of a/b is PK_IMM)
};
static_assert(sizeof(PredInstr) == 16, "PredInstr layout");

struct ScanParams {
    int64_t n_rows;
    int64_t n_tiles;
    int32_t tile_rows;
    int32_t n_cols;
    int32_t n_units;
    int32_t n_stages;
    uint32_t stage_bytes;
    int32_t n_aggs;
    uint64_t* partials;  // [gridDim.x][n_units][SCAN_STATE_SLOTS]
    ScanColDesc cols[SCAN_MAX_COLS];
    ScanUnitDesc units[SCAN_MAX_UNITS];
    PredInstr code[SCAN_MAX_CODE];
};

// Result of the finalize kernel: one record per aggregate.
struct ScanAggOut {
    uint64_t s[SCAN_STATE_SLOTS];
};

}  // namespace tg
