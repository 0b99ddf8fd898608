"""Parquet column chunk -> HBM (sources/parquet.rs in the reference). The files are written here with
pyarrow (uncompressed / Snappy, PLAIN / dictionary-encoded, data pages V1 and V2, small pages so that chunks hold many),
pyarrow's own reader is the ground truth. CPU tests: the host page walk (Thrift compact PageHeaders) and the definition-level expansion. GPU
tests: the decoded Arrow buffers in HBM bit for bit, and a suite over the decoded table against the same suite over the
table registered from host Arrow arrays."""
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import term_b200 as T
from term_b200 import _ffi as F

CASES = [(1, 0.0, None), (1000, 0.3, None), (300_000, 0.05, 100_000), (70_001, 1.0, None), (50_000, 0.0, None), (200_000, 0.999, None)]


def _write(tmp_path, n, null_p, row_group, version, seed=1):
    rng = np.random.default_rng(seed + n)
    x, i = rng.normal(0.0, 1.0, n), rng.integers(-50, 50, n)
    f32, i32 = rng.normal(5.0, 2.0, n).astype(np.float32), rng.integers(-9, 9, n).astype(np.int32)
    masks = {c: rng.random(n) < null_p for c in ("x", "i", "f32", "i32")}
    schema = pa.schema([pa.field("x", pa.float64()), pa.field("i", pa.int64()), pa.field("f32", pa.float32()), pa.field("i32", pa.int32()),
                        pa.field("r", pa.int64(), nullable=False)])
    t = pa.table({"x": pa.array(x, mask=masks["x"]), "i": pa.array(i, mask=masks["i"]), "f32": pa.array(f32, mask=masks["f32"]),
                  "i32": pa.array(i32, mask=masks["i32"]), "r": pa.array(i)}, schema=schema)
    path = os.path.join(tmp_path, f"t_{n}_{version}.parquet")
    pq.write_table(t, path, compression="NONE", use_dictionary=False, data_page_version=version, row_group_size=row_group,
                   data_page_size=64 * 1024)
    return path, t


@pytest.mark.parametrize("version", ["1.0", "2.0"])
@pytest.mark.parametrize("n,null_p,row_group", CASES)
def test_page_walk_and_definition_levels_match_pyarrow(built_lib, tmp_path, n, null_p, row_group, version):
    path, t = _write(str(tmp_path), n, null_p, row_group, version)
    md = pq.ParquetFile(path).metadata
    raw = open(path, "rb").read()
    row0 = 0
    for rg in range(md.num_row_groups):
        for ci in range(md.num_columns):
            cm = md.row_group(rg).column(ci)
            name = cm.path_in_schema
            chunk = np.frombuffer(raw[cm.data_page_offset: cm.data_page_offset + cm.total_compressed_size], dtype=np.uint8)
            n_pages = F.lib().tg_parquet_inspect_chunk(chunk.ctypes.data, chunk.size, None, 0)
            assert n_pages > 0, F.last_error()
            pages = (F.tg_parquet_page * n_pages)()
            assert F.lib().tg_parquet_inspect_chunk(chunk.ctypes.data, chunk.size, pages, n_pages) == n_pages
            assert sum(p.num_values for p in pages) == cm.num_values
            assert all(p.encoding == 0 and p.page_type == (0 if version == "1.0" else 3) for p in pages)
            assert pages[0].header_offset == 0 and all(p.body_offset > p.header_offset for p in pages)
            assert pages[-1].body_offset + pages[-1].body_bytes == chunk.size
            if name != "r":
                bits = np.zeros((cm.num_values + 7) // 8, dtype=np.uint8)
                nn = F.lib().tg_parquet_chunk_validity(chunk.ctypes.data, chunk.size, cm.num_values, bits.ctypes.data)
                want = np.asarray(t.column(name).slice(row0, cm.num_values).is_valid())
                got = np.unpackbits(bits, bitorder="little")[: cm.num_values].astype(bool)
                assert nn == int(want.sum()) and (got == want).all(), (name, rg)
        row0 += md.row_group(rg).num_rows


def test_truncated_and_foreign_chunks_are_rejected(built_lib, tmp_path):
    path, _ = _write(str(tmp_path), 5000, 0.2, None, "1.0")
    cm = pq.ParquetFile(path).metadata.row_group(0).column(0)
    raw = open(path, "rb").read()
    chunk = np.frombuffer(raw[cm.data_page_offset: cm.data_page_offset + cm.total_compressed_size], dtype=np.uint8)
    cut = np.ascontiguousarray(chunk[: chunk.size // 2])
    assert F.lib().tg_parquet_inspect_chunk(cut.ctypes.data, cut.size, None, 0) < 0  # a page runs past the chunk
    junk = np.full(64, 0xFF, dtype=np.uint8)
    assert F.lib().tg_parquet_inspect_chunk(junk.ctypes.data, junk.size, None, 0) < 0


# ------------------------------------------------------------------------------------------ GPU ----
def _device_column(ctx, table, column, np_dtype):
    """(values ndarray, validity bool ndarray | None) read back from the column's device buffers"""
    import torch
    from term_b200.distributed import _tensor_from_ptr
    b = ctx.column_buffers(table, column)
    dev = torch.device("cuda", 0)
    n = b["n_rows"]
    w = np.dtype(np_dtype).itemsize
    raw = _tensor_from_ptr(b["values"], n * w, dev, "|u1").cpu().numpy().view(np_dtype)
    valid = None
    if b["validity"]:
        bits = _tensor_from_ptr(b["validity"], (n + 7) // 8, dev, "|u1").cpu().numpy()
        valid = np.unpackbits(bits, bitorder="little")[:n].astype(bool)
    return raw, valid, b


def _check_string_column(ctx, table, column, want_col):
    """offsets / bytes / validity of a decoded Utf8 column against the Arrow array pyarrow reads"""
    import torch
    from term_b200.distributed import _tensor_from_ptr
    b = ctx.column_buffers(table, column)
    dev = torch.device("cuda", 0)
    n = b["n_rows"]
    want = want_col.combine_chunks()
    assert n == len(want)
    offs = _tensor_from_ptr(b["offsets"], (n + 1) * 4, dev, "|u1").cpu().numpy().view(np.int32)
    data = _tensor_from_ptr(b["values"], max(int(offs[-1]), 1), dev, "|u1").cpu().numpy()[: int(offs[-1])]
    want_valid = np.asarray(want.is_valid())
    if b["validity"]:
        bits = _tensor_from_ptr(b["validity"], (n + 7) // 8, dev, "|u1").cpu().numpy()
        valid = np.unpackbits(bits, bitorder="little")[:n].astype(bool)
    else:
        valid = np.ones(n, dtype=bool)
    assert (valid == want_valid).all(), column
    assert b["null_count"] == int((~want_valid).sum())
    assert offs[0] == 0 and (np.diff(offs) >= 0).all()
    filled = want.fill_null("")
    w_offs = np.frombuffer(filled.buffers()[1], dtype=np.int32)[filled.offset: filled.offset + n + 1]
    w_data = np.frombuffer(filled.buffers()[2], dtype=np.uint8) if filled.buffers()[2] is not None else np.zeros(0, dtype=np.uint8)
    assert (np.diff(offs) == np.diff(w_offs)).all(), column          # NULL rows are empty
    assert data.tobytes() == w_data[int(w_offs[0]): int(w_offs[-1])].tobytes(), column
    assert b["n_value_bytes"] == int(offs[-1])


@pytest.mark.gpu
@pytest.mark.parametrize("version", ["1.0", "2.0"])
@pytest.mark.parametrize("n,null_p,row_group", CASES)
def test_parquet_chunks_decode_to_the_arrow_layout(ctx, tmp_path, n, null_p, row_group, version):
    path, t = _write(str(tmp_path), n, null_p, row_group, version)
    name = f"pq_{n}_{version.replace('.', '')}"
    ctx.register_parquet(name, path)
    try:
        assert ctx.num_rows(name) == n
        for col, dt in (("x", np.float64), ("i", np.int64), ("f32", np.float32), ("i32", np.int32), ("r", np.int64)):
            vals, valid, b = _device_column(ctx, name, col, dt)
            want_valid = np.asarray(t.column(col).is_valid())
            want = np.asarray(t.column(col).fill_null(0)).astype(dt)
            if valid is None:
                assert want_valid.all(), col
                valid = np.ones(n, dtype=bool)
            assert (valid == want_valid).all(), col
            assert b["null_count"] == int((~want_valid).sum())
            assert (vals.view(np.uint8).reshape(n, -1)[valid] == want.view(np.uint8).reshape(n, -1)[valid]).all(), col  # bit for bit
            assert (vals[~valid] == 0).all(), col  # NULL rows are written as zero
    finally:
        ctx.deregister_table(name)


@pytest.mark.gpu
def test_suite_over_parquet_equals_suite_over_arrow(ctx, tmp_path):
    path, t = _write(str(tmp_path), 300_000, 0.05, 100_000, "1.0", seed=9)
    path2, t2 = _write(str(tmp_path), 70_001, 0.3, None, "2.0", seed=11)
    ctx.register_parquet("pq_suite", [path2, path])  # two files appended, the second starts at an unaligned row (70 001)
    ctx.register_table("pq_arrow", pa.concat_tables([t2, t]))
    try:
        A = T.Assertion

        def suite(name):
            cb = (T.Check.builder("c").has_size(A.GreaterThan(0.0)).completeness("x", 0.9).completeness(["i", "i32"], 0.5)
                  .has_min("x", A.LessThan(0.0)).has_max("i", A.GreaterThan(0.0)).has_sum("i", A.LessThan(1e18)).has_mean("x", A.Between(-1.0, 1.0))
                  .has_standard_deviation("x", A.Between(0.5, 2.0)).has_correlation("x", "i", A.Between(-1.0, 1.0))
                  .satisfies("x > 0 AND i < 10").validates_uniqueness(["r"], 0.0).has_approx_quantile("x", 0.5, A.Between(-1.0, 1.0))
                  .has_mean("f32", A.Between(4.0, 6.0)).has_max("i32", A.Equals(8.0)))
            return T.ValidationSuite.builder("s").table_name(name).check(cb.build()).build().run(ctx).report.results
        got, want = suite("pq_suite"), suite("pq_arrow")
        assert [(r.name, r.status, r.metric, r.message) for r in got] == [(r.name, r.status, r.metric, r.message) for r in want]
    finally:
        ctx.deregister_table("pq_suite")
        ctx.deregister_table("pq_arrow")


def _write_encoded(tmp_path, n, null_p, version, compression, seed=3, dict_limit=None):
    """low-cardinality columns (dictionary pages + RLE / bit-packed index streams), a constant column (bit width 0), a
    sorted one (long RLE runs), a high-cardinality one (the writer falls back to PLAIN when the dictionary outgrows
    `dict_limit`), nullable and required, 4- and 8-byte types"""
    rng = np.random.default_rng(seed + n)
    cols = {
        "cat": rng.integers(0, 37, n).astype(np.int64) * 1_000_003,
        "const": np.full(n, 42, dtype=np.int64),
        "runs": np.sort(rng.integers(0, 9, n)).astype(np.float64) * 0.5,
        "wide": rng.integers(0, 1 << 40, n).astype(np.int64),
        "f32": rng.integers(0, 5, n).astype(np.float32) * 1.25,
        "i32": rng.integers(-3, 3, n).astype(np.int32),
    }
    masks = {c: rng.random(n) < null_p for c in cols}
    fields = [pa.field(c, pa.from_numpy_dtype(v.dtype)) for c, v in cols.items()] + [pa.field("req", pa.int64(), nullable=False)]
    arrays = {c: pa.array(v, mask=masks[c]) for c, v in cols.items()}
    arrays["req"] = pa.array(cols["cat"])
    # strings: a low-cardinality column (dictionary), a high-cardinality one with multi-byte characters and empty strings
    # (dictionary overflow -> PLAIN when dict_limit is small), a required one
    words = np.array(["", "a", "bb", "région", "naïve café", "x" * 40, "日本語のテキスト", "tail"], dtype=object)
    smask, umask = rng.random(n) < null_p, rng.random(n) < null_p
    arrays["scat"] = pa.array(words[rng.integers(0, len(words), n)].tolist(), type=pa.string(), mask=smask)
    arrays["suniq"] = pa.array([f"user{v}@exämple.com" if v % 7 else "" for v in rng.integers(0, 1 << 30, n)], type=pa.string(), mask=umask)
    arrays["sreq"] = pa.array(words[rng.integers(0, len(words), n)].tolist(), type=pa.string())
    fields += [pa.field("scat", pa.string()), pa.field("suniq", pa.string()), pa.field("sreq", pa.string(), nullable=False)]
    t = pa.table(arrays, schema=pa.schema(fields))
    path = os.path.join(tmp_path, f"enc_{n}_{version}_{compression}.parquet")
    kw = dict(dictionary_pagesize_limit=dict_limit) if dict_limit else {}
    pq.write_table(t, path, compression=compression, use_dictionary=True, data_page_version=version, row_group_size=max(1, n // 2 + 1),
                   data_page_size=32 * 1024, **kw)
    return path, t


def test_snappy_decoder_matches_pyarrow(built_lib):
    rng = np.random.default_rng(5)
    codec = pa.Codec("snappy")
    samples = [b"", b"a", b"abc" * 1000, bytes(rng.integers(0, 256, 70_000, dtype=np.uint8)), bytes(rng.integers(0, 4, 300_000, dtype=np.uint8)),
               np.sort(rng.integers(0, 1 << 20, 50_000)).astype(np.int64).tobytes(), b"\x00" * 200_000, bytes(range(256)) * 300]
    for raw in samples:
        comp = np.frombuffer(codec.compress(raw, asbytes=True), dtype=np.uint8).copy()
        out = np.zeros(len(raw) + 8, dtype=np.uint8)
        got = F.lib().tg_parquet_snappy_decompress(comp.ctypes.data, comp.size, out.ctypes.data, len(raw))
        assert got == len(raw), F.last_error()
        assert out[: len(raw)].tobytes() == raw
        if len(raw) > 64:
            # a stream that claims more than the room it is given, and a truncated one
            assert F.lib().tg_parquet_snappy_decompress(comp.ctypes.data, comp.size, out.ctypes.data, len(raw) - 1) < 0
            assert F.lib().tg_parquet_snappy_decompress(comp.ctypes.data, comp.size // 2, out.ctypes.data, len(raw)) < 0
    # copies that reach before the start of the output
    bad = np.frombuffer(b"\x08" + b"\x01\xff" + b"\x00" * 8, dtype=np.uint8).copy()
    out = np.zeros(64, dtype=np.uint8)
    assert F.lib().tg_parquet_snappy_decompress(bad.ctypes.data, bad.size, out.ctypes.data, 64) < 0


@pytest.mark.parametrize("name,number", [("gzip", 2), ("brotli", 4), ("zstd", 6), ("lz4_raw", 7), ("snappy", 1)])
def test_page_codecs_match_pyarrow(built_lib, name, number):
    """GZIP / BROTLI / ZSTD / LZ4_RAW page bodies through the host's codec libraries (bound at run time), Snappy through
    the library's own decoder: the same dispatch the chunk path applies to compressed pages"""
    if not pa.Codec.is_available(name):
        pytest.skip(f"pyarrow has no {name} codec here")
    rng = np.random.default_rng(11)
    codec = pa.Codec(name)
    samples = [b"a", b"abc" * 1000, bytes(rng.integers(0, 256, 70_000, dtype=np.uint8)), bytes(rng.integers(0, 4, 300_000, dtype=np.uint8)),
               np.sort(rng.integers(0, 1 << 20, 50_000)).astype(np.int64).tobytes(), b"\x00" * 200_000]
    for raw in samples:
        comp = np.frombuffer(codec.compress(raw, asbytes=True), dtype=np.uint8).copy()
        out = np.zeros(len(raw) + 8, dtype=np.uint8)
        got = F.lib().tg_parquet_page_decompress(number, comp.ctypes.data, comp.size, out.ctypes.data, len(raw))
        assert got == len(raw), F.last_error()
        assert out[: len(raw)].tobytes() == raw
        if len(raw) > 64:  # a stream larger than the room its page header announced, and a truncated one
            assert F.lib().tg_parquet_page_decompress(number, comp.ctypes.data, comp.size, out.ctypes.data, len(raw) - 1) < 0
            assert F.lib().tg_parquet_page_decompress(number, comp.ctypes.data, comp.size // 2, out.ctypes.data, len(raw)) < 0
    for bad in (0, 3, 5, 99):  # UNCOMPRESSED is not a page codec; LZO / hadoop LZ4 / unknown numbers have no decoder
        assert F.lib().tg_parquet_page_decompress(bad, comp.ctypes.data, comp.size, out.ctypes.data, len(raw)) == -F.TG_ERR_UNSUPPORTED


def _delta_table(n, seed):
    rng = np.random.default_rng(seed)
    words = ["", "a", "ab", "abc", "prefix-shared-", "prefix-shared-x", "prefix-shared-yy", "é", "你好", "zzz"]
    return pa.table({
        "i64": pa.array(np.cumsum(rng.integers(-3, 1000, n)).astype(np.int64)),
        "i64r": pa.array(rng.integers(-2**62, 2**62, n)),                      # wide deltas (bit widths up to 64)
        "i32": pa.array(rng.integers(-2**31, 2**31, n).astype(np.int32)),     # 32-bit wrap-around deltas
        "const": pa.array(np.full(n, 7, dtype=np.int64)),                      # bit width 0 miniblocks
        "f64": pa.array(rng.normal(0, 1e6, n)), "f32": pa.array(rng.normal(0, 10, n).astype(np.float32)),
        "s": pa.array([words[v] + str(v % 3) * (v % 4) for v in rng.integers(0, len(words), n)], type=pa.string()),
        "sorted": pa.array(sorted(f"key-{v:07d}" for v in rng.integers(0, 10**6, n)), type=pa.string()),
    })


_DELTA_ENCODINGS = {"i64": ("DELTA_BINARY_PACKED", 5, 8), "i64r": ("DELTA_BINARY_PACKED", 5, 8), "i32": ("DELTA_BINARY_PACKED", 5, 4),
                    "const": ("DELTA_BINARY_PACKED", 5, 8), "f64": ("BYTE_STREAM_SPLIT", 9, 8), "f32": ("BYTE_STREAM_SPLIT", 9, 4),
                    "s": ("DELTA_LENGTH_BYTE_ARRAY", 6, 0), "sorted": ("DELTA_BYTE_ARRAY", 7, 0)}


def _plain_bytes(arr):
    """the PLAIN encoding of an Arrow array without NULLs"""
    if pa.types.is_string(arr.type):
        out = bytearray()
        for v in arr.to_pylist():
            b = v.encode()
            out += len(b).to_bytes(4, "little") + b
        return bytes(out)
    return np.asarray(arr).tobytes()


@pytest.mark.parametrize("n", [1, 129, 20_000])
def test_delta_and_byte_stream_split_pages_rewrite_to_plain(built_lib, tmp_path, n):
    """DELTA_BINARY_PACKED / DELTA_LENGTH_BYTE_ARRAY / DELTA_BYTE_ARRAY / BYTE_STREAM_SPLIT value sections written by pyarrow,
    page by page through the host rewrite of the chunk path: the result is the PLAIN encoding of the same values"""
    t = _delta_table(n, seed=n)
    t = t.cast(pa.schema([pa.field(f.name, f.type, nullable=False) for f in t.schema]))  # required columns: no level section
    path = os.path.join(str(tmp_path), f"delta_{n}.parquet")
    pq.write_table(t, path, compression="NONE", use_dictionary=False, data_page_size=4096,
                   column_encoding={c: e[0] for c, e in _DELTA_ENCODINGS.items()})
    md = pq.ParquetFile(path).metadata
    raw = open(path, "rb").read()
    for ci in range(md.num_columns):
        cm = md.row_group(0).column(ci)
        col = cm.path_in_schema
        enc_name, enc, width = _DELTA_ENCODINGS[col]
        assert enc_name in cm.encodings, (col, cm.encodings)
        chunk = np.frombuffer(raw[cm.data_page_offset: cm.data_page_offset + cm.total_compressed_size], dtype=np.uint8)
        n_pages = F.lib().tg_parquet_inspect_chunk(chunk.ctypes.data, chunk.size, None, 0)
        pages = (F.tg_parquet_page * n_pages)()
        assert F.lib().tg_parquet_inspect_chunk(chunk.ctypes.data, chunk.size, pages, n_pages) == n_pages
        row = 0
        for p in pages:
            assert p.encoding == enc, (col, p.encoding)
            want = _plain_bytes(t.column(col).combine_chunks().slice(row, p.num_values))
            body = chunk[p.body_offset: p.body_offset + p.body_bytes]     # required column, V1 page: the body is the value section
            out = np.zeros(len(want) + 16, dtype=np.uint8)
            got = F.lib().tg_parquet_decode_to_plain(enc, width, body.ctypes.data, body.size, p.num_values, out.ctypes.data, out.size)
            assert got == len(want), (col, row, F.last_error())
            assert out[:got].tobytes() == want, (col, row)
            if p.num_values > 40:  # truncated streams and wrong value counts are refused
                assert F.lib().tg_parquet_decode_to_plain(enc, width, body.ctypes.data, body.size // 2, p.num_values, out.ctypes.data, out.size) < 0
                assert F.lib().tg_parquet_decode_to_plain(enc, width, body.ctypes.data, body.size, p.num_values + 1, out.ctypes.data, out.size) < 0
            row += p.num_values
        assert row == n
    junk = np.frombuffer(bytes([0x80, 0x01, 0x04, 0xC8, 0x01, 0x00]) + b"\xff" * 40, dtype=np.uint8).copy()   # bit widths of 255
    out = np.zeros(4096, dtype=np.uint8)
    assert F.lib().tg_parquet_decode_to_plain(5, 8, junk.ctypes.data, junk.size, 100, out.ctypes.data, out.size) < 0
    assert F.lib().tg_parquet_decode_to_plain(4, 8, junk.ctypes.data, junk.size, 100, out.ctypes.data, out.size) == -F.TG_ERR_UNSUPPORTED


@pytest.mark.parametrize("compression", ["NONE", "SNAPPY"])
def test_page_walk_sees_dictionary_pages(built_lib, tmp_path, compression):
    path, t = _write_encoded(str(tmp_path), 20_000, 0.1, "1.0", compression)
    md = pq.ParquetFile(path).metadata
    raw = open(path, "rb").read()
    cm = md.row_group(0).column(0)
    assert cm.has_dictionary_page
    start = cm.dictionary_page_offset
    chunk = np.frombuffer(raw[start: start + cm.total_compressed_size], dtype=np.uint8)
    n_pages = F.lib().tg_parquet_inspect_chunk(chunk.ctypes.data, chunk.size, None, 0)
    pages = (F.tg_parquet_page * n_pages)()
    assert F.lib().tg_parquet_inspect_chunk(chunk.ctypes.data, chunk.size, pages, n_pages) == n_pages
    assert pages[0].page_type == 2 and pages[0].num_values == 37 and pages[0].encoding in (0, 2)
    assert all(p.page_type == 0 and p.encoding in (2, 8) for p in pages[1:])
    assert sum(p.num_values for p in pages[1:]) == cm.num_values
    if compression == "SNAPPY":
        assert pages[0].uncompressed_bytes == 37 * 8


@pytest.mark.gpu
@pytest.mark.parametrize("compression", ["NONE", "SNAPPY"])
@pytest.mark.parametrize("version", ["1.0", "2.0"])
@pytest.mark.parametrize("n,null_p,dict_limit", [(1, 0.0, None), (1000, 0.3, None), (300_000, 0.05, None), (70_001, 1.0, None), (250_000, 0.02, 4096),
                                                  (120_000, 0.999, None)])
def test_dictionary_and_snappy_chunks_decode_to_the_arrow_layout(ctx, tmp_path, n, null_p, dict_limit, version, compression):
    path, t = _write_encoded(str(tmp_path), n, null_p, version, compression, dict_limit=dict_limit)
    name = f"pqe_{n}_{version.replace('.', '')}_{compression}"
    ctx.register_parquet(name, path)
    try:
        assert ctx.num_rows(name) == n
        for col in t.column_names:
            if pa.types.is_string(t.schema.field(col).type):
                _check_string_column(ctx, name, col, t.column(col))
                continue
            dt = t.schema.field(col).type.to_pandas_dtype()
            vals, valid, b = _device_column(ctx, name, col, dt)
            want_valid = np.asarray(t.column(col).is_valid())
            want = np.asarray(t.column(col).fill_null(0)).astype(dt)
            if valid is None:
                assert want_valid.all(), col
                valid = np.ones(n, dtype=bool)
            assert (valid == want_valid).all(), col
            assert b["null_count"] == int((~want_valid).sum())
            assert (vals.view(np.uint8).reshape(n, -1)[valid] == want.view(np.uint8).reshape(n, -1)[valid]).all(), col  # bit for bit
            assert (vals[~valid] == 0).all(), col
    finally:
        ctx.deregister_table(name)


@pytest.mark.gpu
@pytest.mark.parametrize("compression", ["ZSTD", "GZIP", "LZ4", "BROTLI"])
@pytest.mark.parametrize("version", ["1.0", "2.0"])
@pytest.mark.parametrize("n,null_p,dict_limit", [(1000, 0.3, None), (250_000, 0.02, 4096)])
def test_system_codec_chunks_decode_to_the_arrow_layout(ctx, tmp_path, n, null_p, dict_limit, version, compression):
    """GZIP / ZSTD / LZ4_RAW / BROTLI pages (inflated by the host's own codec libraries, bound at run time)"""
    test_dictionary_and_snappy_chunks_decode_to_the_arrow_layout(ctx, tmp_path, n, null_p, dict_limit, version, compression)


@pytest.mark.gpu
@pytest.mark.parametrize("version", ["1.0", "2.0"])
@pytest.mark.parametrize("compression", ["NONE", "ZSTD"])
@pytest.mark.parametrize("n,null_p", [(1, 0.0), (5000, 0.3), (60_000, 0.0), (3000, 1.0)])
def test_delta_encoded_chunks_decode_to_the_arrow_layout(ctx, tmp_path, n, null_p, version, compression):
    """the same encodings through the whole chunk path (levels, NULLs, V1 / V2 pages, compressed pages), decoded buffers bit
    for bit against the Arrow arrays"""
    t0 = _delta_table(n, seed=n + 1)
    rng = np.random.default_rng(n)
    t = pa.table({c: pa.array(t0.column(c).to_pylist(), type=t0.schema.field(c).type, mask=rng.random(n) < null_p) for c in t0.column_names})
    path = os.path.join(str(tmp_path), "delta_gpu.parquet")
    pq.write_table(t, path, compression=compression, use_dictionary=False, data_page_size=4096, data_page_version=version,
                   column_encoding={c: e[0] for c, e in _DELTA_ENCODINGS.items()})
    name = "pq_delta"
    ctx.register_parquet(name, path)
    try:
        assert ctx.num_rows(name) == n
        for col in t.column_names:
            if pa.types.is_string(t.schema.field(col).type):
                _check_string_column(ctx, name, col, t.column(col))
                continue
            dt = t.schema.field(col).type.to_pandas_dtype()
            vals, valid, b = _device_column(ctx, name, col, dt)
            want_valid = np.asarray(t.column(col).is_valid())
            want = np.asarray(t.column(col).fill_null(0)).astype(dt)
            if valid is None:
                assert want_valid.all(), col
                valid = np.ones(n, dtype=bool)
            assert (valid == want_valid).all(), col
            assert (vals.view(np.uint8).reshape(n, -1)[valid] == want.view(np.uint8).reshape(n, -1)[valid]).all(), col
    finally:
        ctx.deregister_table(name)


@pytest.mark.gpu
def test_suite_over_encoded_parquet_equals_suite_over_arrow(ctx, tmp_path):
    path, t = _write_encoded(str(tmp_path), 200_000, 0.05, "1.0", "SNAPPY", seed=21, dict_limit=8192)
    ctx.register_parquet("pqe_suite", path)
    ctx.register_table("pqe_arrow", t)
    try:
        A = T.Assertion

        def suite(name):
            cb = (T.Check.builder("c").has_size(A.GreaterThan(0.0)).completeness("cat", 0.9).has_min("runs", A.LessThan(1.0))
                  .has_max("wide", A.GreaterThan(0.0)).has_sum("cat", A.LessThan(1e18)).has_mean("runs", A.Between(0.0, 5.0))
                  .has_standard_deviation("wide", A.GreaterThan(0.0)).has_correlation("cat", "wide", A.Between(-1.0, 1.0))
                  .satisfies("const = 42").validates_uniqueness(["wide"], 0.0).has_approx_quantile("runs", 0.5, A.Between(0.0, 5.0))
                  .completeness("suniq", 0.5).validates_email("suniq", 0.1).validates_regex("scat", "é", 0.01).validates_uniqueness(["suniq"], 0.1)
                  .has_min_length("sreq", 0).has_max_length("scat", 100))
            return T.ValidationSuite.builder("s").table_name(name).check(cb.build()).build().run(ctx).report.results
        got, want = suite("pqe_suite"), suite("pqe_arrow")
        g2, w2 = [(r.name, r.status, r.metric, r.message) for r in got], [(r.name, r.status, r.metric, r.message) for r in want]
        assert g2 == w2, [(a, b) for a, b in zip(g2, w2) if a != b]
    finally:
        ctx.deregister_table("pqe_suite")
        ctx.deregister_table("pqe_arrow")


@pytest.mark.gpu
def test_unsupported_parquet_features_fail_loudly(ctx, tmp_path):
    t = pa.table({"k": pa.array([1, 2, 3, 1, 2, 3] * 100), "s": pa.array(["a", "b", "c"] * 200)})
    # LZO / hadoop-framed LZ4 chunks: no decoder (pyarrow cannot write them: the C ABI is called with the codec number)
    from term_b200 import _ffi as F
    tab = ctx._create("pq_codec")
    try:
        chunk = np.zeros(64, dtype=np.uint8)
        for codec in (3, 5, 99):
            with pytest.raises(T.TermGpuError, match="codec"):
                F.check(F.lib().tg_table_append_parquet_chunk(tab, b"k", F.TG_INT64, 0, codec, chunk.ctypes.data, chunk.size, 1))
    finally:
        ctx.deregister_table("pq_codec")
    p3 = os.path.join(str(tmp_path), "bin.parquet")
    pq.write_table(pa.table({"b": pa.array([b"\xff\x00", b"x"] * 10, type=pa.binary())}), p3, compression="NONE", use_dictionary=False)
    with pytest.raises(T.TermGpuError, match="STRING"):
        ctx.register_parquet("pq_bad", p3, columns=["b"])
    # annotations that would need CONVERTED values are refused (the reference yields the logical Arrow types) ...
    import decimal
    t2 = pa.table({"u": pa.array([1, 2, 3], pa.uint32()), "u64": pa.array([1, 2, 3], pa.uint64()),
                   "dec": pa.array([decimal.Decimal("1.50"), decimal.Decimal("2.25"), None], pa.decimal128(9, 2)), "ok": pa.array([1, 2, 3], pa.int64())})
    p4 = os.path.join(str(tmp_path), "logical.parquet")
    pq.write_table(t2, p4, compression="NONE", use_dictionary=False)
    for col in ("u", "u64", "dec"):
        with pytest.raises(T.TermGpuError, match="logical type|physical type"):
            ctx.register_parquet("pq_bad", p4, columns=[col])
    ctx.register_parquet("pq_ok", p4, columns=["ok"])
    assert ctx.num_rows("pq_ok") == 3
    ctx.deregister_table("pq_ok")


@pytest.mark.gpu
def test_parquet_logical_types_follow_the_arrow_typing(ctx, tmp_path):
    """... while DATE / TIME / TIMESTAMP / INT(8|16) annotated columns, whose physical values ARE the Arrow values, are declared
    on the column (tg_table_set_column_arrow_type) and behave like the same column registered from Arrow"""
    rng = np.random.default_rng(8)
    n = 5000
    t = pa.table({"ts0": pa.array(rng.integers(0, 10**6, n), type=pa.timestamp("us"), mask=rng.random(n) < 0.1),
                  "ts1": pa.array(rng.integers(0, 10**6, n), type=pa.timestamp("us")),
                  "d": pa.array(rng.integers(19000, 19100, n).astype(np.int32), type=pa.date32(), mask=rng.random(n) < 0.1),
                  "i8": pa.array(rng.integers(-128, 128, n).astype(np.int8)), "u16": pa.array(rng.integers(0, 2**16, n).astype(np.uint16)),
                  "x": pa.array(rng.normal(0, 1, n))})
    path = os.path.join(str(tmp_path), "lt.parquet")
    pq.write_table(t, path, compression="ZSTD", data_page_size=8 * 1024)
    ctx.register_parquet("lt_pq", path)
    ctx.register_table("lt_ar", t)
    try:
        A = T.Assertion

        def suite(name):
            cb = (T.Check.builder("c").completeness("ts0", 0.5).completeness("d", 0.5).satisfies("ts0 <= ts1").satisfies("i8 > 0 AND u16 < 40000")
                  .validates_uniqueness(["d"], 0.0).has_min("i8", A.GreaterThan(-1e300)).has_sum("u16", A.GreaterThan(-1e300)).has_mean("u16", A.GreaterThan(0.0))
                  .has_sum("i8", A.LessThan(1e18)).has_mean("ts0", A.GreaterThan(0.0)).has_standard_deviation("x", A.GreaterThan(0.0)))
            return [(r.name, r.status, r.metric, r.message) for r in T.ValidationSuite.builder("s").table_name(name).check(cb.build()).build().run(ctx).report.results]
        got, want = suite("lt_pq"), suite("lt_ar")
        assert got == want, [(a, b) for a, b in zip(got, want) if a != b]
        by = {r[0]: r for r in got}
        assert by["min"][3] == "Error evaluating constraint: Internal error: Failed to extract statistic value"      # MIN(Int8) is Int8
    finally:
        ctx.deregister_table("lt_pq")
        ctx.deregister_table("lt_ar")


# ---- hand-built pages: run structures pyarrow's writer never emits (long / tiny / unaligned runs, padded groups) ----
def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _zz(v):
    return _varint((v << 1) ^ (v >> 63))


def _page_v1(levels: bytes, values: bytes, num_values: int) -> bytes:
    body = len(levels).to_bytes(4, "little") + levels + values
    dph = b"\x15" + _zz(num_values) + b"\x15" + _zz(0) + b"\x15" + _zz(3) + b"\x15" + _zz(3) + b"\x00"
    return b"\x15" + _zz(0) + b"\x15" + _zz(len(body)) + b"\x15" + _zz(len(body)) + b"\x2c" + dph + b"\x00" + body


def _hybrid(runs):
    """runs: ("rle", count, bit) | ("packed", [bits...]) with len a multiple of 8 except for the stream's last run"""
    out, bits = bytearray(), []
    for r in runs:
        if r[0] == "rle":
            out += _varint(r[1] << 1) + bytes([r[2]])
            bits += [r[2]] * r[1]
        else:
            b = list(r[1])
            groups = (len(b) + 7) // 8
            padded = b + [0] * (groups * 8 - len(b))
            out += _varint((groups << 1) | 1) + np.packbits(np.array(padded, dtype=np.uint8), bitorder="little").tobytes()
            bits += b
    return bytes(out), np.array(bits, dtype=bool)


@pytest.mark.parametrize("seed", range(12))
def test_hybrid_level_streams_of_any_shape_expand_exactly(built_lib, seed):
    rng = np.random.default_rng(seed)
    pages, want_all = [], []
    for _ in range(int(rng.integers(1, 5))):
        runs = []
        for k in range(int(rng.integers(1, 40))):
            if rng.random() < 0.5:
                runs.append(("rle", int(rng.choice([1, 2, 7, 8, 9, 55, 56, 57, 63, 64, 65, 511, 512, 513, 5000])), int(rng.integers(0, 2))))
            else:
                runs.append(("packed", rng.integers(0, 2, 8 * int(rng.integers(1, 70))).tolist()))
        runs.append(("packed", rng.integers(0, 2, int(rng.integers(1, 8))).tolist()))  # padded last group
        levels, want = _hybrid(runs)
        pages.append(_page_v1(levels, b"\x00" * (8 * int(want.sum())), len(want)))
        want_all.append(want)
    want = np.concatenate(want_all)
    chunk = np.frombuffer(b"".join(pages), dtype=np.uint8)
    bits = np.zeros((len(want) + 7) // 8, dtype=np.uint8)
    nn = F.lib().tg_parquet_chunk_validity(chunk.ctypes.data, chunk.size, len(want), bits.ctypes.data)
    assert nn == int(want.sum()), F.last_error()
    assert (np.unpackbits(bits, bitorder="little")[: len(want)].astype(bool) == want).all()
    n_pages = F.lib().tg_parquet_inspect_chunk(chunk.ctypes.data, chunk.size, None, 0)
    assert n_pages == len(pages)


# ---- malformed page headers: every count / length in a page header is file-controlled ----
def _raw_page_v1(num_values: int, levels_len_field: int, payload: bytes) -> bytes:
    body = (levels_len_field & 0xFFFFFFFF).to_bytes(4, "little") + payload
    dph = b"\x15" + _zz(num_values) + b"\x15" + _zz(0) + b"\x15" + _zz(3) + b"\x15" + _zz(3) + b"\x00"
    return b"\x15" + _zz(0) + b"\x15" + _zz(len(body)) + b"\x15" + _zz(len(body)) + b"\x2c" + dph + b"\x00" + body


def _raw_page_v2(num_values: int, num_nulls: int, def_bytes: int, rep_bytes: int, body: bytes) -> bytes:
    # DataPageHeaderV2 {1 num_values, 2 num_nulls, 3 num_rows, 4 encoding, 5 def bytes, 6 rep bytes}: field id 8 of PageHeader
    h2 = (b"\x15" + _zz(num_values) + b"\x15" + _zz(num_nulls) + b"\x15" + _zz(max(num_values, 0)) + b"\x15" + _zz(0) +
          b"\x15" + _zz(def_bytes) + b"\x15" + _zz(rep_bytes) + b"\x00")
    return b"\x15" + _zz(3) + b"\x15" + _zz(len(body)) + b"\x15" + _zz(len(body)) + b"\x5c" + h2 + b"\x00" + body


def _validity_status(chunk_bytes: bytes, num_values: int):
    chunk = np.frombuffer(chunk_bytes, dtype=np.uint8)
    bits = np.zeros(max(1, (max(num_values, 0) + 7) // 8) + 64, dtype=np.uint8)
    return F.lib().tg_parquet_chunk_validity(chunk.ctypes.data, chunk.size, num_values, bits.ctypes.data)


def test_malformed_page_headers_are_rejected(built_lib):
    lv8, _ = _hybrid([("rle", 8, 1)])
    good = _page_v1(lv8, b"\x00" * 64, 8)
    assert _validity_status(good, 8) == 8
    # a negative value count on one page offset by a larger one on the next still sums to the chunk's count
    neg = _raw_page_v1(-8, len(lv8), lv8 + b"\x00" * 64)
    lv16, _ = _hybrid([("rle", 16, 1)])
    assert _validity_status(neg + _page_v1(lv16, b"\x00" * 128, 16), 8) < 0
    assert "negative" in F.last_error()
    # V1 levels length larger than the page body
    assert _validity_status(_raw_page_v1(8, 1 << 20, lv8), 8) < 0
    assert _validity_status(_raw_page_v1(8, 0xFFFFFFF0, lv8), 8) < 0
    # V1 page shorter than its own 4-byte length field
    dph = b"\x15" + _zz(8) + b"\x15" + _zz(0) + b"\x15" + _zz(3) + b"\x15" + _zz(3) + b"\x00"
    short = b"\x15" + _zz(0) + b"\x15" + _zz(2) + b"\x15" + _zz(2) + b"\x2c" + dph + b"\x00" + b"\x00\x00"
    assert _validity_status(short, 8) < 0
    # V2: negative / oversized level sections
    assert _validity_status(_raw_page_v2(8, 0, -4, 0, lv8 + b"\x00" * 64), 8) < 0
    assert _validity_status(_raw_page_v2(8, 0, 0, -4, lv8 + b"\x00" * 64), 8) < 0
    assert _validity_status(_raw_page_v2(8, 0, 1 << 20, 0, lv8 + b"\x00" * 64), 8) < 0
    assert _validity_status(_raw_page_v2(8, -1, len(lv8), 0, lv8 + b"\x00" * 64), 8) < 0
    assert _validity_status(_raw_page_v2(8, 0, len(lv8), 0, lv8 + b"\x00" * 64), 8) == 8
    # inspect agrees
    for bad in (neg, _raw_page_v2(8, 0, -4, 0, lv8)):
        chunk = np.frombuffer(bad, dtype=np.uint8)
        assert F.lib().tg_parquet_inspect_chunk(chunk.ctypes.data, chunk.size, None, 0) < 0
