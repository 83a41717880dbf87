"""Driver for ncu captures of the index / join kernels at the headline sizes:
    python tools/prof_join.py [customers] [orders] [mode]
mode: join3 (UniqueIndexOn + two joins, the bench step), sort (IndexOn forcing the radix sort + the sorted rows)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csvplus_b200 as cp

ncust = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
nord = int(sys.argv[2]) if len(sys.argv) > 2 else 125_000_000
mode = sys.argv[3] if len(sys.argv) > 3 else "join3"
ctx = cp.Context(0)
cust = ctx.gen_csv("customers", (0, ncust), n_cust=ncust, permute=True)
tc, _ = cp.parse_csv(ctx, cust, spec=[("id", -1), ("name", -1), ("surname", -1)])
if mode == "sort":
    for it in range(int(os.environ.get("PROF_ITERS", "2"))):
        ctx.sync(); t0 = time.perf_counter()
        ix = tc.index_on("id")       # non-unique: always sorts
        ix.table()                   # and materialises the sorted rows
        ctx.sync(); print("IndexOn(%d rows) + sorted rows: %.2f ms" % (ncust, (time.perf_counter() - t0) * 1e3), flush=True)
        del ix
else:
    prod = ctx.gen_csv("products", (0, 1_000_000), n_prod=1_000_000, permute=True)
    orders = ctx.gen_csv("orders", (0, nord), n_cust=ncust, n_prod=1_000_000)
    tp, _ = cp.parse_csv(ctx, prod, spec=[("prod_id", -1), ("product", -1), ("price", -1)])
    to, _ = cp.parse_csv(ctx, orders, spec=[("cust_id", -1), ("prod_id", -1), ("qty", -1), ("ts", -1)])
    for it in range(int(os.environ.get("PROF_ITERS", "2"))):
        ctx.sync(); t0 = time.perf_counter()
        cidx = tc.index_on("id", unique=True)
        pidx = tp.index_on("prod_id", unique=True)
        j = to.join(cidx, "cust_id").join(pidx)
        ctx.sync(); print("index + join3: %.2f ms, rows %d" % ((time.perf_counter() - t0) * 1e3, len(j)), flush=True)
        del j, cidx, pidx
