"""Per-kernel timing of the join pipeline: python tools/time_join.py [orders_rows] [customers_rows]
(parse both sides once, then index + join repeatedly; prints ms per repetition from the library's CUDA-event stats)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csvplus_b200 as cp
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
ncust = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
ctx = cp.Context(0)
cust = ctx.gen_csv("customers", (0, ncust), n_cust=ncust, permute=True)
orders = ctx.gen_csv("orders", (0, rows), n_cust=ncust, n_prod=1_000_000)
tc, _ = cp.parse_csv(ctx, cust, spec=[("id", -1), ("name", -1), ("surname", -1)])
to, _ = cp.parse_csv(ctx, orders, spec=[("cust_id", -1), ("prod_id", -1), ("qty", -1), ("ts", -1)])
def step():
    idx = tc.index_on("id", unique=True)
    return to.join(idx, "cust_id")
for _ in range(2): j = step(); del j
ctx.sync()
reps = 5
ctx.stats(enable=True, reset=True)
t0 = time.perf_counter()
for _ in range(reps): j = step(); n = len(j); del j
ctx.sync()
dt = (time.perf_counter() - t0) / reps * 1e3
st = ctx.stats(); ctx.stats(enable=False)
print("rows", n, "wall ms/rep %.2f" % dt, {k: round(v["ms"] / reps, 3) for k, v in st.items()}, flush=True)
