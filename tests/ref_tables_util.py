"""Cell-by-cell comparison of a pyarrow table with an arrow-rs `pretty_format_batches` table (the form in which the
reference's tests/basic/main.rs holds its expected values; tests/golden/ref_basic_tables.json).  Cells are formatted
the way arrow-rs' ArrayFormatter prints them: nulls empty, floats in shortest round-trip form, binary as hex, decimals
with their scale, timestamps ISO-8601 with 0 / 3 / 6 / 9 fractional digits (a zone prints as `Z` for UTC), structs
`{a: 1, b: }`, lists `[1, , 3]`, maps `{k: v}`."""
import datetime
import math

import numpy as np
import pyarrow as pa


def parse_table(lines):
    """-> (header cells, rows of cells).  A cell is printed as one space, the value, padding: trailing blanks of a value
    cannot be told from padding, so both sides are compared without them."""
    rows = [[c[1:].rstrip() for c in ln.split("|")[1:-1]] for ln in lines if ln.startswith("|")]
    return rows[0], rows[1:]


def _float(v, bits):
    if math.isnan(v):
        return "NaN"
    if math.isinf(v):
        return "inf" if v > 0 else "-inf"
    if bits == 32:
        return np.format_float_positional(np.float32(v), unique=True, trim="0")
    return np.format_float_positional(np.float64(v), unique=True, trim="0")


def _timestamp(v, unit, tz):
    per = {"s": 1, "ms": 10**3, "us": 10**6, "ns": 10**9}[unit]
    secs, frac = divmod(int(v), per)
    frac *= 10**9 // per
    t = datetime.datetime(1970, 1, 1) + datetime.timedelta(seconds=secs)
    s = t.strftime("%Y-%m-%dT%H:%M:%S")
    s = f"{t.year:04d}" + s[s.index("-"):]  # strftime does not pad years below 1000 everywhere
    if frac:
        s += "." + (f"{frac // 10**6:03d}" if frac % 10**6 == 0 else f"{frac // 10**3:06d}" if frac % 10**3 == 0 else f"{frac:09d}")
    if tz is not None:
        s += "Z" if tz in ("UTC", "+00:00") else tz
    return s


def fmt(t: pa.DataType, v) -> str:
    """`v` is what Array.to_pylist() gives for type `t`, except timestamps, which are passed as integers."""
    if v is None:
        return ""
    if pa.types.is_boolean(t):
        return "true" if v else "false"
    if pa.types.is_float32(t):
        return _float(v, 32)
    if pa.types.is_float64(t):
        return _float(v, 64)
    if pa.types.is_binary(t) or pa.types.is_large_binary(t):
        return bytes(v).hex()
    if pa.types.is_decimal(t):
        return format(v, "f")
    if pa.types.is_date32(t):
        return v.isoformat()
    if pa.types.is_timestamp(t):
        return _timestamp(v, t.unit, t.tz)
    if pa.types.is_struct(t):
        return "{" + ", ".join(f"{t.field(i).name}: {fmt(t.field(i).type, v[t.field(i).name])}" for i in range(t.num_fields)) + "}"
    if pa.types.is_map(t):
        return "{" + ", ".join(f"{fmt(t.key_type, k)}: {fmt(t.item_type, x)}" for k, x in v) + "}"
    if pa.types.is_list(t):
        return "[" + ", ".join(fmt(t.value_type, x) for x in v) + "]"
    return str(v)


def column_cells(col: pa.ChunkedArray):
    t = col.type
    if pa.types.is_timestamp(t):
        return [fmt(t, v) for v in col.cast(pa.int64()).to_pylist()]
    return [fmt(t, v) for v in col.to_pylist()]


def assert_table_matches(table: pa.Table, lines, what=""):
    header, rows = parse_table(lines)
    assert table.column_names == header, f"{what}: columns {table.column_names} != {header}"
    assert table.num_rows == len(rows), f"{what}: {table.num_rows} rows, the reference's test expects {len(rows)}"
    for ci, name in enumerate(header):
        got = [g.rstrip() for g in column_cells(table.column(ci))]
        exp = [r[ci] for r in rows]
        assert got == exp, f"{what}: column {name}: {[(i, g, e) for i, (g, e) in enumerate(zip(got, exp)) if g != e][:5]}"
