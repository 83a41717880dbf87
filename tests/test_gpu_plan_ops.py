"""The rest of the recognised-plan vocabulary: Top / Drop / TakeWhile / DropWhile / DropColumns on the device
against the oracle's restatement of csvplus.go:313-374 / :493-507, ToJSON's exact text (csvplus.go:446-474),
ResolveDuplicates returning a row that is not one of the group's rows (csvplus.go:838-848), and the lazily sorted
unique index (a join does not sort; iteration / Find / SubIndex / errors behave exactly as if it had been sorted)."""
import io
import random

import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import assert_table_equals_oracle, gpu_ctx, people_csv

pytestmark = pytest.mark.gpu


def _rows_equal(got, orows):
    assert got == [{k.decode(): v.decode() for k, v in r.items()} for r in orows.to_dicts()]


def test_cuts_vs_oracle():
    import csvplus_b200 as cp
    data = people_csv(3000, seed=3)
    src = lambda: cp.Take(cp.FromBytes(data).SelectColumns("id", "name", "surname"))
    o = orc.reader_rows(data, select=["id", "name", "surname"])
    first_name = o.values("name")[0].decode()
    for n in (0, 1, 7, 2999, 3000, 5000):
        _rows_equal(src().Top(n).ToRows(), o.top(n))
        _rows_equal(src().Drop(n).ToRows(), o.drop(n))
    for pred_g, pred_o in ((cp.Like({"name": first_name}), orc.Like({"name": first_name})),
                           (cp.Not(cp.Like({"surname": "Smith"})), orc.Not(orc.Like({"surname": "Smith"}))),
                           (cp.Like({"name": "nobody"}), orc.Like({"name": "nobody"})),
                           (cp.Any(cp.Like({"name": first_name}), cp.Not(cp.Like({"id": "17"}))), orc.Any(orc.Like({"name": first_name}), orc.Not(orc.Like({"id": "17"}))))):
        _rows_equal(src().TakeWhile(pred_g).ToRows(), o.take_while(pred_o))
        _rows_equal(src().DropWhile(pred_g).ToRows(), o.drop_while(pred_o))
        _rows_equal(src().DropWhile(pred_g).Top(5).ToRows(), o.drop_while(pred_o).top(5))
    _rows_equal(src().DropColumns("name").Drop(10).Top(3).ToRows(), o.drop_columns("name").drop(10).top(3))
    # a cut that stops early never meets an error further down the input; one that does not, does
    bad = data + b"too,few\n"
    ob = orc.reader_rows(bad, select=["id", "name", "surname"])
    assert ob.error is not None
    got = cp.Take(cp.FromBytes(bad).SelectColumns("id", "name", "surname")).Top(10).ToRows()
    _rows_equal(got, ob.top(10))
    assert ob.top(10).error is None
    with pytest.raises(cp.DataSourceError) as ei:
        cp.Take(cp.FromBytes(bad).SelectColumns("id", "name", "surname")).Drop(10).ToRows()
    assert str(ei.value) == ob.drop(10).error
    with pytest.raises(cp.DataSourceError):
        cp.Take(cp.FromBytes(bad).SelectColumns("id", "name", "surname")).DropWhile(cp.Like({"name": first_name})).ToRows()
    assert ob.drop_while(orc.Like({"name": first_name})).error == ob.error


def test_to_json_text():
    import csvplus_b200 as cp
    rows = [{"b": 'x"y\\z', "a": "<&> é"}, {"a": "line\nbreak\ttab\x01", "b": ""}]
    out = io.BytesIO()
    cp.TakeRows(rows).ToJSON(out)
    assert out.getvalue() == ('[{"a":"<&>\\u2028é","b":"x\\"y\\\\z"}\n,{"a":"line\\nbreak\\ttab\\u0001","b":""}\n]').encode()
    out = io.BytesIO()
    cp.TakeRows([]).ToJSON(out)
    assert out.getvalue() == b"[]"
    # through the device path: parse -> filter -> json
    data = b"id,name\n1,Ann\n2,Bob\n3,Ann\n"
    out = io.BytesIO()
    cp.Take(cp.FromBytes(data)).Filter(cp.Like({"name": "Ann"})).ToJSON(out)
    assert out.getvalue() == b'[{"id":"1","name":"Ann"}\n,{"id":"3","name":"Ann"}\n]'


def test_resolve_duplicates_with_a_new_row():
    """csvplus.go:838-848: whatever row the resolver returns takes the group's place (no re-sort); here a merged row"""
    import csvplus_b200 as cp
    rng = random.Random(4)
    rows = [{"k": "key%02d" % rng.randrange(30), "v": str(i)} for i in range(400)]
    ix = cp.TakeRows(rows).IndexOn("k")

    def merge(group):
        return {"k": group[0]["k"], "v": "+".join(sorted(r["v"] for r in group))}
    ix.ResolveDuplicates(merge)
    got = cp.Take(ix).ToRows()
    by_key = {}
    for r in rows:
        by_key.setdefault(r["k"], []).append(r["v"])
    want = [{"k": k, "v": "+".join(sorted(v))} for k, v in sorted(by_key.items())]
    # (§Q1: a trailing singleton is lost when any group exists — every key here is in a group, so nothing is lost)
    assert all(len(v) > 1 for v in by_key.values())
    assert got == want
    assert ix.Find("key07").ToRows() == [w for w in want if w["k"] == "key07"]
    j = cp.TakeRows([{"k": "key07"}, {"k": "nope"}]).Join(ix).ToRows()
    assert j == [w for w in want if w["k"] == "key07"]
    # a longer key than any existing one: the key image is rebuilt
    ix2 = cp.TakeRows(rows).IndexOn("k")
    ix2.ResolveDuplicates(lambda g: {"k": g[0]["k"] + "-merged-with-a-long-suffix", "v": str(len(g))})
    got2 = cp.Take(ix2).ToRows()
    assert [r["k"] for r in got2] == [k + "-merged-with-a-long-suffix" for k in sorted(by_key)]
    assert ix2.Find("key03-merged-with-a-long-suffix").ToRows() == [{"k": "key03-merged-with-a-long-suffix", "v": str(len(by_key["key03"]))}]


def test_lazily_sorted_unique_index():
    import csvplus_b200 as cp
    ctx = gpu_ctx()
    n = 60_000
    raw = ctx.gen_csv("customers", (0, n), n_cust=n, permute=True)
    host = raw.to_host()
    t, _ = cp.parse_csv(ctx, raw, spec=[("id", -1), ("name", -1), ("surname", -1)])
    oi = orc.reader_rows(host, select=["id", "name", "surname"]).unique_index_on("id")
    # 1. join first (no sort has happened), then everything that needs the order
    ix = t.index_on("id", unique=True)
    orders = ctx.gen_csv("orders", (0, 100_000), n_cust=n + 500, n_prod=10)  # some orders have no customer
    to, _ = cp.parse_csv(ctx, orders, spec=[("cust_id", -1), ("qty", -1)])
    oo = orc.reader_rows(orders.to_host(), select=["cust_id", "qty"])
    assert_table_equals_oracle(to.join(ix, "cust_id"), oo.join(oi, "cust_id"))
    assert_table_equals_oracle(to.join(ix, "cust_id", anti=True), oo.except_(oi, "cust_id"))
    assert_table_equals_oracle(ix.table(), oi.rows())          # iteration order = sorted order
    assert ix.Find("4242").ToRows() == [{k.decode(): v.decode() for k, v in r.items()} for r in oi.find("4242").to_dicts()]
    assert_table_equals_oracle(to.join(ix, "cust_id"), oo.join(oi, "cust_id"))  # and the join still works afterwards
    # 2. order first, join second
    ix2 = t.index_on("id", unique=True)
    assert_table_equals_oracle(ix2.table(), oi.rows())
    assert_table_equals_oracle(to.join(ix2, "cust_id"), oo.join(oi, "cust_id"))
    # 3. a duplicate is still reported with the reference's text (lowest key in sort order)
    dup = np.concatenate([host, np.frombuffer(b"77,Dup,Licate,1950,X,1.00\n9,Dup,Licate,1950,X,1.00\n", np.uint8)])
    td, _ = cp.parse_csv(ctx, dup, spec=[("id", -1), ("name", -1)])
    with pytest.raises(cp.CsvPlusError) as ei:
        td.index_on("id", unique=True)
    with pytest.raises(orc.OracleError) as eo:
        orc.reader_rows(dup, select=["id", "name"]).unique_index_on("id")
    assert str(ei.value) == str(eo.value)
    # 4. composite unique key, prefix join needs the order
    t2, _ = cp.parse_csv(ctx, raw, spec=[("surname", -1), ("id", -1), ("name", -1)])
    ix3 = t2.index_on("surname", "id", unique=True)
    oi3 = orc.reader_rows(host, select=["surname", "id", "name"]).unique_index_on("surname", "id")
    probe = cp.Table.from_rows(ctx, [{"surname": "Smith"}, {"surname": "Nobody"}, {"surname": "Lewis"}])
    oprobe = orc.take_rows([{"surname": "Smith"}, {"surname": "Nobody"}, {"surname": "Lewis"}])
    assert_table_equals_oracle(probe.join(ix3, "surname"), oprobe.join(oi3, "surname"))
    sub = ix3.SubIndex("Lewis")
    assert len(sub) == len(oi3.find("Lewis"))
