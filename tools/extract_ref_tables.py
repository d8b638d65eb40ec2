"""Golden vectors held by the reference's own end-to-end tests: the expected tables of tests/basic/main.rs (arrow-rs
`pretty_format_batches` output) and tests/basic/misc.rs, written to tests/golden/ref_basic_tables.json together with
what each test reads (file, projection, options) and the counts it asserts.  Run in the build container, where
/root/reference exists; the JSON is committed, the tests never read the reference tree.

    python tools/extract_ref_tables.py [/root/reference]
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_basic_tables.json")

ALL_TEST_ORC = ["a", "b", "str_direct", "d", "e", "f"]
BASIC_2 = ["int_short_repeated", "int_neg_short_repeated", "int_delta", "int_neg_delta", "int_direct", "int_neg_direct",
           "bigint_direct", "bigint_neg_direct", "bigint_other", "utf8_increase", "utf8_decrease"]

# test function -> what it reads.  `fields` None = new_arrow_reader_root.  `tail` = compare the last N rows of the last
# batch; `batch_lens` / `first_batch_len` / `total_rows` / `n_batches` = the counts the test asserts; `const` = the
# expected table is a constant of misc.rs.
SPECS = {
    "test_read_long_bool": dict(file="long_bool.orc", fields=["long"], const="LONG_BOOL_EXPECTED", first_batch_len=32),
    "test_read_long_bool_gzip": dict(file="long_bool_gzip.orc", fields=["long"], const="LONG_BOOL_EXPECTED", first_batch_len=32),
    "test_read_long_string": dict(file="string_long.orc", fields=["dict"], const="LONG_STRING_EXPECTED", first_batch_len=64),
    "test_read_string_dirt": dict(file="string_dict.orc", fields=["dict"], const="LONG_STRING_DICT_EXPECTED", first_batch_len=64),
    "test_read_string_dirt_gzip": dict(file="string_dict_gzip.orc", fields=["dict"], const="LONG_STRING_DICT_EXPECTED", first_batch_len=64),
    "test_read_string_long_long": dict(file="string_long_long.orc", fields=["dict"], batch_lens=[8192, 10_000 - 8192]),
    "test_read_string_long_long_gzip": dict(file="string_long_long_gzip.orc", fields=["dict"], batch_lens=[8192, 10_000 - 8192]),
    "basic_test": dict(file="test.orc", fields=ALL_TEST_ORC),
    "basic_test_2": dict(file="test.orc", fields=BASIC_2),
    "basic_test_3": dict(file="test.orc", fields=["timestamp_simple", "date_simple"]),
    "basic_test_bigint": dict(file="test_bigint.orc", fields=["id", "appl_no"], tail=3),
    "basic_test_patched_int": dict(file="patched_int.orc", fields=["c1"], tail=3, total_rows=999596),
    "basic_test_nested_struct": dict(file="nested_struct.orc"),
    "basic_test_nested_array": dict(file="nested_array.orc"),
    "basic_test_nested_array_float": dict(file="nested_array_float.orc"),
    "basic_test_nested_array_struct": dict(file="nested_array_struct.orc"),
    "basic_test_nested_map_struct": dict(file="nested_map_struct.orc"),
    "basic_test_nested_map": dict(file="nested_map.orc"),
    "basic_test_0": dict(file="test.orc"),
    "basic_test_with_range": dict(file="test.orc", byte_range=[0, 2000], first_batch_len=5),
    "basic_test_with_range_without_data": dict(file="test.orc", byte_range=[100, 2000], n_batches=0),
    "v0_file_test": dict(file="demo-11-zlib.orc", total_rows_is_total_row_count=True),
    "v1_file_test": dict(file="demo-12-zlib.orc", total_rows_is_total_row_count=True),
    "v0_file_test_async": dict(file="demo-11-zlib.orc", total_rows=1_920_800),
    "timestamps_test": dict(file="pyarrow_timestamps.orc",
                            schema={"timestamp_notz": "timestamp[ns]", "timestamp_utc": "timestamp[ns, tz=UTC]"}),
    "overflowing_timestamps_test": dict(file="overflowing_timestamps.orc", is_err=True),
    "rlev2_test": dict(file="pyorc_rlev2_patchedbase.orc"),
}
for c in ("none", "snappy", "zlib", "lzo", "zstd", "lz4"):
    SPECS[f"alltypes_test[{c}]"] = dict(file=f"alltypes.{c}.orc", fn="alltypes_test")


def fn_body(src: str, name: str) -> str:
    m = re.search(r"\bfn " + re.escape(name) + r"\(\)[^{]*\{", src)
    assert m, name
    depth, i = 1, m.end()
    while depth:
        depth += {"{": 1, "}": -1}.get(src[i], 0)
        i += 1
    return src[m.end():i]


def expected_lines(body: str):
    m = re.search(r"let expected = \[(.*?)\];", body, re.S)
    if not m:
        return None
    return [s.encode().decode("unicode_escape").encode("latin-1").decode("utf-8") if "\\" in s else s
            for s in re.findall(r'"((?:[^"\\]|\\.)*)"', m.group(1))]


def main() -> None:
    main_rs = open(os.path.join(REF, "tests", "basic", "main.rs"), encoding="utf-8").read()
    misc_rs = open(os.path.join(REF, "tests", "basic", "misc.rs"), encoding="utf-8").read()
    consts = {m.group(1): m.group(2).split("\n") for m in re.finditer(r'pub const (\w+): &str = r#"(.*?)"#;', misc_rs, re.S)}
    out = {}
    for name, spec in SPECS.items():
        spec = dict(spec)
        fn = spec.pop("fn", name)
        body = fn_body(main_rs, fn)
        const = spec.pop("const", None)
        lines = consts[const] if const else expected_lines(body)
        if lines is not None:
            spec["expected"] = lines
        spec["source"] = f"tests/basic/main.rs: fn {fn}" + (f" (tests/basic/misc.rs: {const})" if const else "")
        out[name] = spec
    with open(OUT, "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False, indent=1)
    print(f"{len(out)} tests, {sum('expected' in v for v in out.values())} with tables -> {os.path.relpath(OUT)}")


if __name__ == "__main__":
    main()
