"""CUDA path against the committed golden fixtures (tests/golden/): the same five pipelines as
tests/test_golden_cpu.py, through the public API and the C ABI, compared byte for byte with the committed files
(no oracle involved at run time).  Named test_zz_* so that it runs after the oracle-parity GPU tests."""
import io
import json
import os

import pytest

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rd(name):
    with open(os.path.join(G, name), "rb") as f:
        return f.read()


def csv_of(src, *cols):
    buf = io.BytesIO()
    src.ToCsv(buf, *cols)
    return buf.getvalue()


def test_cuda_path_reproduces_goldens():
    import csvplus_b200 as cp
    people, orders, nasty = rd("people.csv"), rd("orders.csv"), rd("nasty.csv")
    meta = json.load(open(os.path.join(G, "meta.json")))
    amelia = cp.Take(cp.FromBytes(people).SelectColumns("name", "surname", "id")).Filter(cp.Like({"name": "Amelia"}))
    assert csv_of(amelia, "name", "surname", "id") == rd("people_amelia.csv")
    idx = cp.Take(cp.FromBytes(people).SelectColumns("id", "name", "surname")).UniqueIndexOn("id")
    probe = cp.Take(cp.FromBytes(orders).SelectColumns("order_id", "cust_id", "qty"))
    joined = probe.Join(idx, "cust_id")
    assert csv_of(joined, "order_id", "cust_id", "qty", "id", "name", "surname") == rd("orders_join_people.csv")
    assert len(joined.ToRows()) == meta["join_rows"]
    oidx = cp.Take(cp.FromBytes(orders)).IndexOn("prod_id", "qty")
    assert csv_of(cp.Take(oidx), "prod_id", "qty", "order_id") == rd("orders_sorted_prod_qty.csv")
    probe2 = cp.Take(cp.FromBytes(orders).SelectColumns("order_id", "cust_id"))
    assert csv_of(probe2.Except(idx, "cust_id"), "order_id", "cust_id") == rd("orders_without_customer.csv")
    assert csv_of(cp.Take(cp.FromBytes(nasty)), "c0", "c1", "c2", "c3") == rd("nasty_roundtrip.csv")
