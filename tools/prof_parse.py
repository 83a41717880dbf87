"""Small driver for ncu captures: `python tools/prof_parse.py [rows] [mode]`
mode: filter (configs[1] shape), orders (4 of 5 columns, no filter), join (small join pipeline)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csvplus_b200 as cp

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
mode = sys.argv[2] if len(sys.argv) > 2 else "filter"
ctx = cp.Context(0)
if mode == "filter":
    buf = ctx.gen_csv("people", (0, rows))
    for _ in range(3):
        t, e = cp.parse_csv(ctx, buf, spec=[("name", -1), ("surname", -1), ("id", -1)], pred=cp.Like({"name": "Amelia"}))
    print(len(t))
elif mode == "orders":
    buf = ctx.gen_csv("orders", (0, rows), n_cust=rows // 10, n_prod=1000)
    for _ in range(3):
        t, e = cp.parse_csv(ctx, buf, spec=[("cust_id", -1), ("prod_id", -1), ("qty", -1), ("ts", -1)])
    print(len(t))
else:
    ncust = rows // 10
    cust = ctx.gen_csv("customers", (0, ncust), n_cust=ncust, permute=True)
    orders = ctx.gen_csv("orders", (0, rows), n_cust=ncust, n_prod=1000)
    for _ in range(int(os.environ.get("PROF_ITERS", "2"))):
        tc, _ = cp.parse_csv(ctx, cust, spec=[("id", -1), ("name", -1), ("surname", -1)])
        idx = tc.index_on("id", unique=True)
        to, _ = cp.parse_csv(ctx, orders, spec=[("cust_id", -1), ("prod_id", -1), ("qty", -1), ("ts", -1)])
        j = to.join(idx, "cust_id")
    print(len(j))
