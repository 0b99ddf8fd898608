// Arrow buffer marshalling: values, validity bitmaps and Utf8 offsets staged into HBM.
// Replaces MemTable registration on a SessionContext (e.g. constraints/completeness.rs:389-391 in the
// reference tests, sources/mod.rs:48-66 in production) for the GPU path.
#include <algorithm>
#include <cstring>

#include "engine.hpp"

namespace tg {

static size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Pivot K of the shifted moment sums: the median of the first (up to) five finite, valid values of the
// column. It must be an ELEMENT of the column (the scan replaces NULL rows by K), and a typical one, so
// that sum(x) = n*K + sum(x-K) keeps full accuracy; a median of five is robust to isolated outliers.
void set_pivot_host(Column& c, int32_t dtype, int64_t n, const void* values, const uint8_t* validity,
                    int64_t bit_offset) {
    if (c.pivot_set || (dtype != TG_INT64 && dtype != TG_FLOAT64)) return;
    double cand[5];
    int64_t icand[5];
    int m = 0;
    for (int64_t i = 0; i < n && i < 65536 && m < 5; ++i) {
        bool ok = !validity || ((validity[(bit_offset + i) >> 3] >> ((bit_offset + i) & 7)) & 1);
        if (!ok) continue;
        if (dtype == TG_INT64) {
            icand[m] = reinterpret_cast<const int64_t*>(values)[i];
            cand[m] = (double)icand[m];
            ++m;
        } else {
            double v = reinterpret_cast<const double*>(values)[i];
            if (v == v && v - v == 0.0) {
                cand[m] = v;
                icand[m] = 0;
                ++m;
            }
        }
    }
    if (m == 0) return;
    for (int i = 1; i < m; ++i)
        for (int j = i; j > 0 && (dtype == TG_INT64 ? icand[j] < icand[j - 1] : cand[j] < cand[j - 1]); --j) {
            std::swap(cand[j], cand[j - 1]);
            std::swap(icand[j], icand[j - 1]);
        }
    c.pivot = cand[m / 2];
    c.ipivot = icand[m / 2];
    c.pivot_set = true;
}

// append `n` bits taken from src starting at bit `src_off` (or all-ones when src == nullptr) to a device
// bitmap that currently holds `have` bits; tail = host mirror of its last partial byte
static void append_bits(Engine& e, DevBuf& buf, uint8_t& tail, int64_t have, const uint8_t* src, int64_t src_off,
                        int64_t n, int64_t* zeros) {
    const int64_t total = have + n;
    e.dev_reserve(buf, (size_t)(total + 7) / 8, (size_t)(have + 7) / 8);
    const int64_t first_byte = have / 8;
    const int shift = (int)(have % 8);
    const size_t out_bytes = (size_t)((total + 7) / 8 - first_byte);
    std::vector<uint8_t> tmp(out_bytes + 8, 0);
    if (shift) tmp[0] = tail & (uint8_t)((1u << shift) - 1);
    int64_t z = 0;
    if (!src) {
        for (int64_t i = 0; i < n; ++i) {
            const int64_t o = shift + i;
            tmp[o >> 3] |= (uint8_t)(1u << (o & 7));
        }
    } else if (shift == 0 && (src_off & 7) == 0) {
        memcpy(tmp.data(), src + src_off / 8, (size_t)(n + 7) / 8);
        if (n & 7) tmp[(n - 1) / 8] &= (uint8_t)((1u << (n & 7)) - 1);
        for (size_t i = 0; i < (size_t)(n + 7) / 8; ++i) z += 8 - __builtin_popcount(tmp[i]);
        z -= (8 - (n & 7)) & 7;
    } else {
        for (int64_t i = 0; i < n; ++i) {
            const int64_t s = src_off + i;
            const int bit = (src[s >> 3] >> (s & 7)) & 1;
            const int64_t o = shift + i;
            tmp[o >> 3] |= (uint8_t)(bit << (o & 7));
            z += !bit;
        }
    }
    if (zeros) *zeros = z;
    tail = (total & 7) ? tmp[out_bytes - 1] : 0;
    // staged copy (tmp is pageable; h2d copies it into the pinned ring before returning)
    e.h2d(buf.p + first_byte, tmp.data(), out_bytes);
}

Column* table_get_or_add(Table& t, const std::string& name, int32_t dtype) {
    Column* c = t.find(name);
    if (c) {
        if (c->dtype != dtype) throw Error(TG_ERR_TYPE_MISMATCH, "column '" + name + "' was registered with a different type");
        return c;
    }
    validate_identifier(name);
    auto nc = std::make_unique<Column>();
    nc->name = name;
    nc->dtype = dtype;
    t.cols.push_back(std::move(nc));
    return t.cols.back().get();
}

// appends the validity of `n` rows (bitmap `validity` read from bit `bit_offset`; nullptr = no NULLs) to a column that
// holds `have` rows; materialises the bitmap lazily the first time a NULL shows up
void append_validity(Engine& e, Column& c, int64_t have, const uint8_t* validity, int64_t bit_offset, int64_t n) {
    bool has_nulls = false;
    bool fast_bits = false;
    int64_t zeros = 0;
    if (validity) {
        if ((bit_offset & 7) == 0 && (have & 7) == 0) {
            // byte-aligned on both sides: count zero bits 64 at a time; the bitmap bytes then go to the device as
            // they are (direct DMA when the source is pinned)
            const uint8_t* src = validity + bit_offset / 8;
            const int64_t full_words = n / 64;
            int64_t ones = 0;
            for (int64_t w = 0; w < full_words; ++w) {
                uint64_t x;
                memcpy(&x, src + 8 * w, 8);
                ones += __builtin_popcountll(x);
            }
            for (int64_t i = full_words * 64; i < n; ++i) ones += (src[i >> 3] >> (i & 7)) & 1;
            zeros = n - ones;
            has_nulls = zeros > 0;
            fast_bits = true;
        } else {
            for (int64_t i = 0; i < n && !has_nulls; ++i) {
                const int64_t s = bit_offset + i;
                has_nulls = !((validity[s >> 3] >> (s & 7)) & 1);
            }
        }
    }
    if (has_nulls && !c.validity.p && have > 0) {
        // earlier batches had no bitmap: materialise all-ones for them
        append_bits(e, c.validity, c.tail_byte, 0, nullptr, 0, have, nullptr);
    }
    if (has_nulls || c.validity.p) {
        if (fast_bits && validity) {
            const size_t nbytes = (size_t)(n + 7) / 8;
            e.dev_reserve(c.validity, (size_t)(have + n + 7) / 8, (size_t)(have + 7) / 8);
            e.h2d(c.validity.p + have / 8, validity + bit_offset / 8, nbytes);
            // host mirror of the last partial byte (bits past the end masked off) for the next unaligned append
            c.tail_byte = (n & 7) ? (uint8_t)(validity[bit_offset / 8 + nbytes - 1] & ((1u << (n & 7)) - 1)) : 0;
        } else {
            append_bits(e, c.validity, c.tail_byte, have, validity, bit_offset, n, &zeros);
        }
        c.null_count += zeros;
    }
}

void table_append_host(Table& t, const std::string& name, int32_t dtype, int64_t n, const void* values,
                       const int32_t* offsets, const uint8_t* validity, int64_t bit_offset) {
    Engine& e = *t.eng;
    std::lock_guard<std::mutex> g(e.mu);
    TG_CUDA(cudaSetDevice(e.device));
    if (n < 0) throw Error(TG_ERR_INVALID_ARG, "negative row count");
    Column& c = *table_get_or_add(t, name, dtype);
    if (c.adopted) throw Error(TG_ERR_INVALID_ARG, "cannot append to an adopted device column");
    const int64_t have = c.n_rows;
    if (n == 0) {
        t.n_rows = std::max(t.n_rows, c.n_rows);
        return;
    }
    // ---- fixed-width values first: their DMA (the bulk of the bytes) is queued before the host walks the validity
    // bitmap below, so counting NULLs overlaps the copy instead of delaying it ----
    const bool fixed_width = dtype == TG_INT64 || dtype == TG_FLOAT64 || dtype == TG_INT32 || dtype == TG_FLOAT32;
    if (fixed_width) {
        const size_t w = (size_t)c.elem_bytes();
        e.dev_reserve(c.values, (size_t)(have + n) * w, (size_t)have * w);
        e.h2d(c.values.p + (size_t)have * w, values, (size_t)n * w);
        c.value_bytes = (have + n) * (int64_t)w;
        set_pivot_host(c, dtype, n, values, validity, bit_offset);
    }
    append_validity(e, c, have, validity, bit_offset, n);
    // ---- values ----
    switch (dtype) {
        case TG_INT64:
        case TG_FLOAT64:
        case TG_INT32:
        case TG_FLOAT32: break;  // queued above
        case TG_BOOL: {
            append_bits(e, c.values, c.tail_vbyte, have, (const uint8_t*)values, bit_offset, n, nullptr);
            c.value_bytes = (have + n + 7) / 8;
        } break;
        case TG_UTF8: {
            if (!offsets) throw Error(TG_ERR_INVALID_ARG, "Utf8 column requires offsets");
            const int32_t o0 = offsets[0], o1 = offsets[n];
            const int64_t nbytes = (int64_t)o1 - o0;
            if (nbytes < 0) throw Error(TG_ERR_INVALID_ARG, "Utf8 offsets are not monotonic");
            if (c.value_bytes + nbytes > INT32_MAX)
                throw Error(TG_ERR_UNSUPPORTED, "Utf8 column exceeds 2 GiB of value bytes (use LargeUtf8 sharding)");
            e.dev_reserve(c.values, (size_t)(c.value_bytes + nbytes), (size_t)c.value_bytes);
            e.h2d(c.values.p + c.value_bytes, (const uint8_t*)values + o0, (size_t)nbytes);
            // rebase offsets onto the concatenated byte buffer
            e.dev_reserve(c.offsets, (size_t)(have + n + 1) * 4, (size_t)(have + 1) * 4);
            std::vector<int32_t> tmp((size_t)n + 1);
            const int32_t delta = (int32_t)c.value_bytes - o0;
            for (int64_t i = 0; i <= n; ++i) tmp[i] = offsets[i] + delta;
            e.h2d(c.offsets.p + (size_t)have * 4, tmp.data(), (size_t)(n + 1) * 4);
            c.value_bytes += nbytes;
        } break;
        default: throw Error(TG_ERR_INVALID_ARG, "unknown dtype");
    }
    c.n_rows = have + n;
    t.n_rows = std::max(t.n_rows, c.n_rows);
}

void table_adopt_device(Table& t, const std::string& name, int32_t dtype, int64_t n, const void* d_values,
                        const int32_t* d_offsets, const uint8_t* d_validity, int64_t n_value_bytes) {
    Engine& e = *t.eng;
    std::lock_guard<std::mutex> g(e.mu);
    TG_CUDA(cudaSetDevice(e.device));
    if (t.find(name)) throw Error(TG_ERR_INVALID_ARG, "column '" + name + "' already exists");
    if (((uintptr_t)d_values & 15) || ((uintptr_t)d_validity & 15) || ((uintptr_t)d_offsets & 15))
        throw Error(TG_ERR_INVALID_ARG, "adopted device buffers must be 16-byte aligned");
    Column& c = *table_get_or_add(t, name, dtype);
    c.adopted = true;
    c.n_rows = n;
    c.values.p = (uint8_t*)d_values;
    c.values.owned = false;
    c.values.cap = 0;
    c.validity.p = (uint8_t*)d_validity;
    c.validity.owned = false;
    c.offsets.p = (uint8_t*)d_offsets;
    c.offsets.owned = false;
    c.value_bytes = dtype == TG_UTF8 ? n_value_bytes : n * c.elem_bytes();
    c.null_count = -1;
    if ((dtype == TG_INT64 || dtype == TG_FLOAT64) && n > 0) {
        const int64_t head = std::min<int64_t>(n, 65536);
        std::vector<uint8_t> hv((size_t)head * 8), hb((size_t)(head + 7) / 8);
        TG_CUDA(cudaMemcpy(hv.data(), d_values, hv.size(), cudaMemcpyDeviceToHost));
        if (d_validity) TG_CUDA(cudaMemcpy(hb.data(), d_validity, hb.size(), cudaMemcpyDeviceToHost));
        set_pivot_host(c, dtype, head, hv.data(), d_validity ? hb.data() : nullptr, 0);
    }
    t.n_rows = std::max(t.n_rows, n);
}

// ---- Arrow C Data Interface (https://arrow.apache.org/docs/format/CDataInterface.html) ----
struct ArrowSchema {
    const char* format;
    const char* name;
    const char* metadata;
    int64_t flags;
    int64_t n_children;
    struct ArrowSchema** children;
    struct ArrowSchema* dictionary;
    void (*release)(struct ArrowSchema*);
    void* private_data;
};
struct ArrowArray {
    int64_t length;
    int64_t null_count;
    int64_t offset;
    int64_t n_buffers;
    int64_t n_children;
    const void** buffers;
    struct ArrowArray** children;
    struct ArrowArray* dictionary;
    void (*release)(struct ArrowArray*);
    void* private_data;
};

// Arrow C Data Interface format string -> the stored type; width / signedness of the delivered values when they are widened
// on the host; DataFusion's name of the delivered type (nullptr: it is the stored type)
struct ArrowFormat {
    int32_t dtype = 0;
    int src_w = 0;
    bool src_unsigned = false, temporal = false, large_offsets = false;
    const char* src_type = nullptr;
    char temporal_unit = 0;
};
static ArrowFormat arrow_format(const std::string& fmt, const char* column) {
    ArrowFormat f;
    auto starts = [&](const char* pre) { return fmt.compare(0, strlen(pre), pre) == 0; };
    if (fmt == "l") f.dtype = TG_INT64;
    else if (fmt == "g") f.dtype = TG_FLOAT64;
    else if (fmt == "u") f.dtype = TG_UTF8;
    else if (fmt == "U") f.dtype = TG_UTF8, f.large_offsets = true;
    else if (fmt == "i") f.dtype = TG_INT32;
    else if (fmt == "f") f.dtype = TG_FLOAT32;
    else if (fmt == "b") f.dtype = TG_BOOL;
    else if (fmt == "c") f.dtype = TG_INT32, f.src_w = 1, f.src_type = "Int8";
    else if (fmt == "C") f.dtype = TG_INT32, f.src_w = 1, f.src_unsigned = true, f.src_type = "UInt8";
    else if (fmt == "s") f.dtype = TG_INT32, f.src_w = 2, f.src_type = "Int16";
    else if (fmt == "S") f.dtype = TG_INT32, f.src_w = 2, f.src_unsigned = true, f.src_type = "UInt16";
    else if (fmt == "I") f.dtype = TG_INT64, f.src_w = 4, f.src_unsigned = true, f.src_type = "UInt32";
    else if (fmt == "L") f.dtype = TG_INT64, f.src_w = 8, f.src_unsigned = true, f.src_type = "UInt64";
    else if (fmt == "tdD") f.dtype = TG_INT32, f.temporal = true, f.src_type = "Date32", f.temporal_unit = 'D';
    else if (fmt == "tdm") f.dtype = TG_INT64, f.temporal = true, f.src_type = "Date64", f.temporal_unit = 'm';
    else if (fmt == "tts" || fmt == "ttm") f.dtype = TG_INT32, f.temporal = true, f.src_type = "Time32";
    else if (fmt == "ttu" || fmt == "ttn") f.dtype = TG_INT64, f.temporal = true, f.src_type = "Time64";
    else if (starts("tss:") || starts("tsm:") || starts("tsu:") || starts("tsn:"))
        f.dtype = TG_INT64, f.temporal = true, f.src_type = "Timestamp", f.temporal_unit = fmt[2];
    else if (fmt == "tDs" || fmt == "tDm" || fmt == "tDu" || fmt == "tDn") f.dtype = TG_INT64, f.temporal = true, f.src_type = "Duration";
    else throw Error(TG_ERR_UNSUPPORTED, std::string("Arrow format '") + fmt + "' of column '" + column + "' is not supported");
    return f;
}

// Declares the Arrow type a stored Int32 / Int64 column stands for when its values arrived already in the stored
// representation (Parquet chunks annotated DATE / TIME / TIMESTAMP / INT(8|16): the physical INT32 / INT64 values ARE the
// Arrow values): the typing rules of table_append_arrow then apply to it.
void table_set_column_arrow_type(Table& t, const std::string& name, const std::string& fmt) {
    Engine& e = *t.eng;
    std::lock_guard<std::mutex> g(e.mu);
    Column* c = t.find(name);
    if (!c) throw Error(TG_ERR_COLUMN_NOT_FOUND, "Schema error: No field named " + name + ". Valid fields are " + t.valid_fields() + ".");
    const ArrowFormat f = arrow_format(fmt, name.c_str());
    if (f.dtype != c->dtype || f.large_offsets) throw Error(TG_ERR_TYPE_MISMATCH, "column '" + name + "' does not store Arrow format '" + fmt + "'");
    if (f.src_w == 4 || f.src_w == 8)  // UInt32 / UInt64 need a conversion of the stored values, not a label
        throw Error(TG_ERR_UNSUPPORTED, "column '" + name + "': Arrow format '" + fmt + "' cannot be declared on stored values");
    c->src_type = f.src_type;
    c->src_unsigned = f.src_unsigned;
    c->temporal = f.temporal;
    c->temporal_unit = f.temporal_unit;
}

void table_append_arrow(Table& t, const void* schema_p, const void* array_p) {
    const ArrowSchema* sc = (const ArrowSchema*)schema_p;
    const ArrowArray* ar = (const ArrowArray*)array_p;
    if (!sc || !ar || !sc->format || strcmp(sc->format, "+s") != 0)
        throw Error(TG_ERR_INVALID_ARG, "expected a struct-typed ArrowArray (a RecordBatch)");
    if (ar->offset != 0) throw Error(TG_ERR_UNSUPPORTED, "sliced struct arrays are not supported");
    for (int64_t i = 0; i < sc->n_children; ++i) {
        const ArrowSchema* cs = sc->children[i];
        const ArrowArray* ca = ar->children[i];
        const std::string fmt = cs->format;
        const ArrowFormat F = arrow_format(fmt, cs->name);
        const int32_t dtype = F.dtype;
        const int src_w = F.src_w;
        const bool src_unsigned = F.src_unsigned, temporal = F.temporal, large_offsets = F.large_offsets;
        const char* src_type = F.src_type;
        const uint8_t* validity = (const uint8_t*)ca->buffers[0];
        if (ca->null_count == 0) validity = nullptr;
        const int64_t off = ca->offset, n = ca->length;
        if (src_type) {  // (checked before anything is appended: a column keeps one delivered type)
            Column* have = t.find(cs->name);
            if (have && (have->src_type == nullptr || strcmp(have->src_type, src_type) != 0))
                throw Error(TG_ERR_TYPE_MISMATCH, std::string("column '") + cs->name + "' was registered with a different type");
        }
        if (dtype == TG_UTF8 && large_offsets) {
            // LargeUtf8: 64-bit offsets narrowed on the host (a batch of 2 GiB or more of string bytes does not fit Utf8)
            const int64_t* o64 = (const int64_t*)ca->buffers[1] + off;
            std::vector<int32_t> o32((size_t)n + 1);
            const int64_t base = n > 0 || ca->buffers[1] ? o64[0] : 0;
            for (int64_t r = 0; r <= n; ++r) {
                const int64_t d = (ca->buffers[1] ? o64[r] : 0) - base;
                if (d > 0x7fffffffll) throw Error(TG_ERR_UNSUPPORTED, std::string("LargeUtf8 column '") + cs->name + "': a batch holds 2 GiB or more of string bytes");
                o32[(size_t)r] = (int32_t)d;
            }
            table_append_host(t, cs->name, dtype, n, (const uint8_t*)ca->buffers[2] + base, o32.data(), validity, off);
        } else if (dtype == TG_UTF8) {
            const int32_t* offs = (const int32_t*)ca->buffers[1] + off;
            table_append_host(t, cs->name, dtype, n, ca->buffers[2], offs, validity, off);
        } else if (dtype == TG_BOOL) {
            // values bitmap shares the array offset; append_bits takes one bit offset for both
            table_append_host(t, cs->name, dtype, n, ca->buffers[1], nullptr, validity, off);
        } else if (src_w && src_w != (dtype == TG_INT64 ? 8 : 4)) {
            // Int8 / Int16 / UInt8 / UInt16 -> Int32, UInt32 -> Int64: widened exactly on the host
            const uint8_t* src = (const uint8_t*)ca->buffers[1] + off * src_w;
            std::vector<uint8_t> wide((size_t)n * (dtype == TG_INT64 ? 8 : 4));
            for (int64_t r = 0; r < n; ++r) {
                if (src_w == 1) {
                    const int32_t v = src_unsigned ? (int32_t)src[r] : (int32_t)(int8_t)src[r];
                    memcpy(wide.data() + r * 4, &v, 4);
                } else if (src_w == 2) {
                    uint16_t u;
                    memcpy(&u, src + r * 2, 2);
                    const int32_t v = src_unsigned ? (int32_t)u : (int32_t)(int16_t)u;
                    memcpy(wide.data() + r * 4, &v, 4);
                } else {
                    uint32_t u;
                    memcpy(&u, src + r * 4, 4);
                    const int64_t v = (int64_t)u;
                    memcpy(wide.data() + r * 8, &v, 8);
                }
            }
            table_append_host(t, cs->name, dtype, n, wide.data(), nullptr, validity, off);
        } else {
            const int w = dtype == TG_INT64 || dtype == TG_FLOAT64 ? 8 : 4;
            const uint8_t* vals = (const uint8_t*)ca->buffers[1] + off * w;
            if (src_unsigned && src_w == 8) {  // UInt64: representable as Int64 unless a valid value has the top bit set
                for (int64_t r = 0; r < n; ++r) {
                    if (validity && !((validity[(off + r) >> 3] >> ((off + r) & 7)) & 1)) continue;
                    if (vals[r * 8 + 7] & 0x80) throw Error(TG_ERR_UNSUPPORTED, std::string("UInt64 column '") + cs->name + "' holds values above the Int64 range");
                }
            }
            table_append_host(t, cs->name, dtype, n, vals, nullptr, validity, off);
        }
        if (src_type) {
            Column* c = t.find(cs->name);
            c->src_type = src_type;
            c->src_unsigned = src_unsigned;
            c->temporal = temporal;
            c->temporal_unit = F.temporal_unit;
        }
    }
}

}  // namespace tg
