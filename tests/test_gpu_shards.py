"""Byte-range shards of ONE file: cpb_csv_quote_parity + cpb_parse_csv_shard through csvplus_b200.dist.ShardedParse,
run here as N simulated ranks on one GPU (the exchange steps are plain Python lists), compared with the 1-shard parse and
with the oracle: quoted newlines straddling the cuts, cuts inside quoted fields, CRLF, blank lines, the error ordinal."""
import random

import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import assert_table_equals_oracle, gpu_ctx, random_csv

pytestmark = pytest.mark.gpu


def _run_sharded(data: bytes, world: int, **kw):
    import csvplus_b200 as cp
    from csvplus_b200.dist import ShardedParse
    ctx = gpu_ctx()
    ranks = [ShardedParse(ctx, r, world, len(data), lambda lo, hi: data[lo:hi], lookahead=kw.pop("lookahead", 1 << 20)) for r in range(world)]
    all_q = [sp.step1_parity() for sp in ranks]
    all_info = [sp.step2_parse(all_q, **kw) for sp in ranks]
    outs = [sp.step3_finish(all_info) for sp in ranks]
    tables = [t for t, _ in outs]
    errs = [e for _, e in outs]
    whole = cp.Table.concat(tables)
    return whole, errs, all_info


def _check(data: bytes, world: int, select=None, like=None):
    import csvplus_b200 as cp
    whole, errs, info = _run_sharded(data, world, spec=[(c, -1) for c in select] if select else None, pred=cp.Like(like) if like else None)
    o = orc.reader_rows(data, select=select, pred=orc.Like(like) if like else None)
    assert_table_equals_oracle(whole, o, columns=None if len(o) else [])
    if o.error is None:
        assert all(e is None for e in errs)
    else:
        assert all(e is not None for e in errs)
        assert len({e.Line for e in errs}) == 1 and errs[0].Line == o.error_line, (errs[0].Line, o.error_line)
        raised = [e for e in errs if not str(e.Err).startswith("(error raised")]
        assert len(raised) == 1 and str(raised[0]) == o.error
    # the records counted by the shards add up to the file's records (header excluded on shard 0 only through data_start)
    return info


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_adversarial_file_by_byte_ranges(seed, world):
    data = random_csv(500 + seed, nrows=6000, ncols=5, quoted_p=0.4, crlf_p=0.3, blank_p=0.1)
    assert len(data) > 150_000
    _check(data, world)
    _check(data, world, select=["c3", "c0"])
    _check(data, world, select=["c1", "c2"], like={"c1": ""})


def test_cut_inside_quoted_newlines():
    """cuts placed exactly inside quoted fields that contain newlines, and right at record boundaries"""
    rec = b'k%d,"multi\nline ""quoted"" \r\n field",tail\n'
    rows = [b"a,b,c\n"] + [rec % i for i in range(4000)]
    data = b"".join(rows)
    for world in (2, 5, 7):
        _check(data, world)
    # a cut that falls exactly on the first byte of a record / right after a newline
    pos = [i for i in range(len(data)) if data[i:i + 1] == b"\n"][1000] + 1
    for size in (2 * pos, 2 * pos + 1, 2 * pos - 1):
        _check(data[:size] if size <= len(data) else data, 2)


def test_error_ordinal_is_global():
    lines = [b"a,b,c"] + [b"%d,x,y" % i for i in range(30000)]
    lines[20001] = b"oops,too,many,fields"      # record ordinal 20002 (header = 1)
    lines[25000] = b'bare"quote,x,y'            # a later error in a later shard must lose
    data = b"\n".join(lines) + b"\n"
    info = _check(data, 4)
    assert sum(1 for r in info if r[1]) >= 1
    whole, errs, _ = _run_sharded(data, 4)
    assert errs[0].Line == 20002 and len(whole) == 20000
    _check(data, 4, select=["b"])


def test_plain_file_many_tiles_takes_lean_tiles_per_shard():
    """no quotes at all: every interior tile of every shard takes the lean path; shards cut mid-line"""
    import csvplus_b200 as cp
    ctx = gpu_ctx()
    buf = ctx.gen_csv("orders", (0, 200_000), n_cust=5000, n_prod=100)
    data = buf.to_host().tobytes()
    for world in (2, 3):
        _check(data, world, select=["cust_id", "qty", "ts"])


def test_lookahead_too_small_is_reported():
    import csvplus_b200 as cp
    data = b"a,b\n" + b"".join(b"%d,%s\n" % (i, b"x" * 300) for i in range(2000))
    with pytest.raises(cp.CsvPlusError, match="look-ahead too small"):
        _run_sharded(data, 2, lookahead=16)
