"""Shared helpers of the parity tests: random CSV construction and GPU-vs-oracle comparison."""
from __future__ import annotations

import random

import numpy as np

from oracle import oracle as orc


def gpu_ctx():
    import csvplus_b200 as cp
    return cp.Context.default()


def assert_table_equals_oracle(table, rows: "orc.Rows", columns=None):
    """bit-exact comparison of every column (Arrow offsets + bytes) of a device table with oracle rows"""
    assert len(table) == len(rows), f"row count {len(table)} != oracle {len(rows)}"
    cols = columns if columns is not None else table.columns
    if len(rows):
        assert sorted(c.encode() for c in cols) == sorted(rows.header(0)), (cols, rows.header(0))
    for c in cols:
        go, gd = table.column(c)
        oo, od, pres = rows.column(c)
        assert pres.all() if len(pres) else True
        assert np.array_equal(go, oo), f"offsets differ in column {c}"
        assert np.array_equal(gd, od), f"bytes differ in column {c}"


def run_both(data: bytes, *, opts: "orc.Opts | None" = None, select=None, expect=None, assume=None, like=None):
    """runs Take(FromBytes(data)...)[.Filter(Like(like))] on the GPU and in the oracle; returns (table, gerr, orows)"""
    import csvplus_b200 as cp
    o = opts or orc.Opts()
    spec = None
    hdr = True
    if select is not None:
        spec = [(n, -1) for n in select]
    elif expect is not None:
        spec = list(expect.items())
    elif assume is not None:
        spec = list(assume.items()); hdr = False
    pred = cp.Like(like) if like else None
    t, gerr = cp.parse_csv(gpu_ctx(), data, delimiter=o.comma, comment=o.comment, num_fields=o.fields_per_record,
                           lazy_quotes=o.lazy_quotes, trim_leading_space=o.trim_leading_space,
                           header_from_first_row=hdr and o.header_from_first_row, spec=spec, pred=pred)
    orows = orc.reader_rows(data, o, select=select, expect=expect, assume=assume, pred=orc.Like(like) if like else None)
    return t, gerr, orows


def check_parity(data: bytes, **kw):
    import csvplus_b200 as cp
    try:
        t, gerr, orows = run_both(data, **kw)
    except cp.DataSourceError as e:  # errors raised before any table exists
        orows = orc.reader_rows(data, kw.get("opts") or orc.Opts(), select=kw.get("select"), expect=kw.get("expect"),
                                assume=kw.get("assume"))
        assert orows.error == str(e), (orows.error, str(e))
        return None, orows
    oerr = orows.error
    assert (str(gerr) if gerr else None) == oerr, f"error mismatch: gpu={gerr} oracle={oerr}"
    assert_table_equals_oracle(t, orows, columns=None if len(orows) else [])
    return t, orows


# ------------------------------------------------------------------ adversarial CSV generator (parity, not timing)
def random_field(rng: random.Random, quoted_p=0.3, long_p=0.0, maxlen=12) -> tuple[bytes, bytes]:
    """-> (raw csv encoding, value)"""
    n = rng.randrange(0, maxlen)
    if rng.random() < long_p:
        n = rng.choice([2000, 2100, 5000, 33000, 70000])
    alphabet = b"abcXYZ019 _-.;#'\t\xc3\xa9"
    if rng.random() < quoted_p:
        specials = [b'""', b",", b"\n", b"\r\n", b"\r", b" "]
        parts, val = [], []
        for _ in range(n):
            if rng.random() < 0.25:
                s = rng.choice(specials)
                parts.append(s)
                val.append(b'"' if s == b'""' else (b"\n" if s == b"\r\n" else s))
            else:
                c = bytes([rng.choice(alphabet)])
                parts.append(c); val.append(c)
        return b'"' + b"".join(parts) + b'"', b"".join(val)
    v = bytes(rng.choice(alphabet) for _ in range(n))
    if rng.random() < 0.05:
        v = v + b"\r" + bytes([rng.choice(alphabet)])  # lone CR inside an unquoted field is data
    return v, v


def random_csv(seed: int, nrows: int, ncols: int = 5, quoted_p=0.3, long_p=0.0, crlf_p=0.3, blank_p=0.1, ragged_p=0.0,
               trailing_newline=True, header=True) -> bytes:
    rng = random.Random(seed)
    lines = []
    if header:
        lines.append(b",".join(b"c%d" % i for i in range(ncols)))
    for _ in range(nrows):
        k = ncols
        if rng.random() < ragged_p:
            k = rng.randrange(1, ncols + 3)
        fields = [random_field(rng, quoted_p, long_p)[0] for _ in range(k)]
        if k == 1 and fields[0] == b"":
            fields[0] = b"x"  # a single empty unquoted field would be a blank line
        lines.append(b",".join(fields))
    out = []
    for i, ln in enumerate(lines):
        out.append(ln)
        last = i == len(lines) - 1
        if not last or trailing_newline:
            out.append(b"\r\n" if rng.random() < crlf_p else b"\n")
        if rng.random() < blank_p:
            out.append(rng.choice([b"\n", b"\r\n"]))
    return b"".join(out)


def people_csv(n: int, seed: int = 1) -> bytes:
    """CPU-side small fixture in the shape of the reference's people.csv (csvplus_test.go:1207-1253)"""
    names = [b"Amelia", b"Olivia", b"Emily", b"Ava", b"Isla", b"Oliver", b"Jack", b"Harry", b"Jacob", b"Charlie"]
    surnames = [b"Smith", b"Jones", b"Taylor", b"Williams", b"Brown", b"Davies", b"Evans", b"Wilson", b"Thomas", b"Roberts",
                b"Johnson", b"Lewis"]
    rng = random.Random(seed)
    rows = [b"id,name,surname,born"]
    for i in range(n):
        rows.append(b"%d,%s,%s,%d" % (i, names[i // 12 % 10] if n <= 120 else rng.choice(names),
                                       surnames[i % 12] if n <= 120 else rng.choice(surnames), 1916 + rng.randrange(90)))
    return b"\n".join(rows) + b"\n"
