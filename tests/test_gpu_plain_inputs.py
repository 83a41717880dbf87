"""Quote-free inputs spanning many 32 KiB tiles of csv_scan against the oracle, mixed with what makes single tiles take the
rare paths mid-file (ragged lines, dense short lines that overflow the flat index, fields longer than the look-ahead, one
quote, one error): all tiles share the two look-back chains, so totals, offsets and error ordinals must stay exact across
any mixture.  (Written while a separate lean path for regular tiles was tried — DESIGN.md §7; kept as parity cases.)"""
import os
import random

import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import assert_table_equals_oracle, check_parity, gpu_ctx

pytestmark = pytest.mark.gpu


def _plain_csv(seed, nrows, ncols=6, crlf_p=0.0, blank_p=0.0, short_p=0.0, long_p=0.0, eol_last=True, width=(0, 14)):
    rng = random.Random(seed)
    alphabet = b"abcdefXYZ0123456789 _-.;:#'+/"
    out = [b",".join(b"c%d" % i for i in range(ncols)) + b"\n"]
    for r in range(nrows):
        fields = []
        for c in range(ncols):
            n = rng.randrange(*width)
            if rng.random() < short_p:
                n = rng.randrange(0, 2)
            if rng.random() < long_p:
                n = rng.choice([700, 1500, 2500, 3300, 5000])
            fields.append(bytes(rng.choice(alphabet) for _ in range(n)))
        if all(len(f) == 0 for f in fields) and ncols == 1:
            fields[0] = b"x"
        line = b",".join(fields)
        last = r == nrows - 1
        out.append(line + ((b"\r\n" if rng.random() < crlf_p else b"\n") if (eol_last or not last) else b""))
        if rng.random() < blank_p:
            out.append(rng.choice([b"\n", b"\r\n"]))
    return b"".join(out)


@pytest.mark.parametrize("seed,kw", [
    (1, dict()),
    (2, dict(crlf_p=1.0)),
    (3, dict(crlf_p=0.5, blank_p=0.2)),
    (4, dict(short_p=0.9, width=(0, 3))),          # dense lines: tiles whose structurals overflow the flat index
    (5, dict(long_p=0.01)),                         # fields longer than the 2 KiB look-ahead
    (6, dict(eol_last=False)),
    (7, dict(ncols=3, width=(0, 6))),
    (8, dict(ncols=12, width=(1, 5))),
    (9, dict(ncols=40, width=(0, 4))),             # more than 16 columns: the scan runs once per group of columns
])
def test_plain_inputs_many_tiles(seed, kw):
    data = _plain_csv(seed, 40_000, **kw)
    assert len(data) > 8 * 32768
    ncols = kw.get("ncols", 6)
    sel = ["c%d" % i for i in sorted(random.Random(seed).sample(range(ncols), min(ncols, 1 + seed % 4)))]
    check_parity(data, select=sel)
    check_parity(data)  # every column
    if ncols > 16:  # every column of a wide file, through a filter, and with an error in the middle
        vals = orc.reader_rows(data, select=["c1"]).values("c1")
        common = max(set(vals), key=vals.count)
        check_parity(data, like={"c1": common.decode(), "c39": ""})
        lines = data.split(b"\n")
        lines[len(lines) // 2] = b"short,line"
        t, orows = check_parity(b"\n".join(lines))
        assert orows.error is not None and "wrong number of fields" in orows.error
    if ncols >= 2:
        # a predicate that keeps ~ a quarter of the rows: compare against the most frequent value of c1
        vals = orc.reader_rows(data, select=["c1"]).values("c1")
        common = max(set(vals), key=vals.count)
        check_parity(data, select=sel + (["c1"] if "c1" not in sel else []), like={"c1": common.decode()})


def test_irregular_lines_in_the_middle():
    """one ragged line / one quoted field / one bare quote deep inside a regular file: rows, offsets and the error
    ordinal stay exact"""
    base = _plain_csv(11, 30_000).split(b"\n")
    for patch, expect_err in (
        (b"a,b,c", "wrong number of fields"),
        (b'q,"x,y",r,s,t,u', None),
        (b'1,2,3,4,5,6"7', 'bare " in non-quoted-field'),
        (b"1,2,3,4,5,6,7,8", "wrong number of fields"),
    ):
        lines = list(base)
        lines[17_000] = patch
        data = b"\n".join(lines)
        t, orows = check_parity(data, select=["c0", "c3", "c5"])
        if expect_err:
            assert orows.error is not None and expect_err in orows.error
            assert len(t) == 16_999  # the rows before the failing record are delivered
        else:
            assert orows.error is None and len(t) == 30_000


def test_any_field_count_pads_short_lines():
    """NumFieldsAny: short lines are padded with empty values"""
    lines = _plain_csv(12, 20_000).split(b"\n")
    lines[9_000] = b"only,two"
    data = b"\n".join(lines)
    check_parity(data, opts=orc.Opts(fields_per_record=-1), select=["c0", "c1", "c4"])


def test_bench_shapes_vs_oracle():
    """people / orders / products of the synthetic generator vs the oracle"""
    import csvplus_b200 as cp
    ctx = gpu_ctx()
    for kind, spec, kw in (("people", ["name", "surname", "id"], {}), ("orders", ["cust_id", "prod_id", "qty", "ts"], dict(n_cust=1000, n_prod=50)),
                           ("products", ["prod_id", "product", "price"], dict(n_prod=300_000))):
        buf = ctx.gen_csv(kind, (0, 300_000), **kw)
        t, err = cp.parse_csv(ctx, buf, spec=[(c, -1) for c in spec])
        assert err is None
        assert_table_equals_oracle(t, orc.reader_rows(buf.to_host(), select=spec))


def test_pool_reserve_then_parse():
    """cpb_pool_reserve maps pool memory ahead of time (bench.py reserves its working set); results are unaffected"""
    import csvplus_b200 as cp
    from tests.helpers import people_csv
    ctx = cp.Context(int(os.environ.get("LOCAL_RANK", "0")))
    try:
        assert ctx.reserve(256 << 20)
        assert ctx.reserve(0)
        data = people_csv(5000)
        t, err = cp.parse_csv(ctx, data, spec=[("name", -1), ("id", -1)])
        assert err is None and len(t) == 5000
        orows = orc.reader_rows(data, orc.Opts(), select=["name", "id"])
        assert_table_equals_oracle(t, orows)
        del t
    finally:
        ctx.close()
