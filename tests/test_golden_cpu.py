"""The committed golden fixtures (tests/golden/, made by tests/golden/make_golden.py) pin the oracle: it must reproduce
them byte for byte, and — for the inputs without quoting — an independent plain-Python restatement of the pipeline
(str.split, dict lookups, sorted) must produce the same files."""
import json
import os

from oracle import oracle as orc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rd(name):
    with open(os.path.join(G, name), "rb") as f:
        return f.read()


def table(data: bytes):
    lines = data.decode().split("\n")
    assert lines[-1] == ""
    hdr = lines[0].split(",")
    return [dict(zip(hdr, ln.split(","))) for ln in lines[1:-1]]


def dump(rows, cols):
    return ("\n".join([",".join(cols)] + [",".join(r[c] for c in cols) for r in rows]) + "\n").encode()


def test_oracle_reproduces_goldens():
    people, orders, nasty = rd("people.csv"), rd("orders.csv"), rd("nasty.csv")
    meta = json.load(open(os.path.join(G, "meta.json")))
    r = orc.reader_rows(people, select=["name", "surname", "id"], pred=orc.Like({"name": "Amelia"}))
    assert r.to_csv("name", "surname", "id")[0] == rd("people_amelia.csv") and len(r) == meta["people_amelia_rows"]
    idx = orc.reader_rows(people, select=["id", "name", "surname"]).unique_index_on("id")
    j = orc.reader_rows(orders, select=["order_id", "cust_id", "qty"]).join(idx, "cust_id")
    assert j.to_csv("order_id", "cust_id", "qty", "id", "name", "surname")[0] == rd("orders_join_people.csv")
    assert len(j) == meta["join_rows"]
    assert orc.reader_rows(orders).index_on("prod_id", "qty").rows().to_csv("prod_id", "qty", "order_id")[0] == rd("orders_sorted_prod_qty.csv")
    ex = orc.reader_rows(orders, select=["order_id", "cust_id"]).except_(idx, "cust_id")
    assert ex.to_csv("order_id", "cust_id")[0] == rd("orders_without_customer.csv") and len(ex) == meta["except_rows"]
    rn = orc.reader_rows(nasty)
    assert rn.error is None and len(rn) == meta["nasty_rows"]
    assert rn.to_csv("c0", "c1", "c2", "c3")[0] == rd("nasty_roundtrip.csv")


def test_goldens_match_an_independent_python_restatement():
    people, orders = table(rd("people.csv")), table(rd("orders.csv"))
    # Filter(Like{name: Amelia}) + SelectColumns (csvplus.go:276-286, :493-512, :1276-1293)
    assert dump([p for p in people if p["name"] == "Amelia"], ["name", "surname", "id"]) == rd("people_amelia.csv")
    # UniqueIndexOn(id) + Join(cust_id): probe order, index row merged under the probe row (csvplus.go:545-583)
    by_id = {p["id"]: p for p in people}
    joined = [{**{k: by_id[o["cust_id"]][k] for k in ("id", "name", "surname")}, **o} for o in orders if o["cust_id"] in by_id]
    assert dump(joined, ["order_id", "cust_id", "qty", "id", "name", "surname"]) == rd("orders_join_people.csv")
    # Except (csvplus.go:586-608)
    assert dump([o for o in orders if o["cust_id"] not in by_id], ["order_id", "cust_id"]) == rd("orders_without_customer.csv")
    # IndexOn(prod_id, qty): bytewise string comparison column by column (csvplus.go:794-807); ties keep input order
    srt = sorted(orders, key=lambda o: (o["prod_id"].encode(), o["qty"].encode()))
    assert dump(srt, ["prod_id", "qty", "order_id"]) == rd("orders_sorted_prod_qty.csv")
