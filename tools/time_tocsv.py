"""ToCsv timing: python tools/time_tocsv.py [rows] -- orders parse (4 columns) then ToCsv of the table, bytes left in HBM"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csvplus_b200 as cp
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
ctx = cp.Context(0)
orders = ctx.gen_csv("orders", (0, rows), n_cust=rows // 10, n_prod=1000)
t, _ = cp.parse_csv(ctx, orders, spec=[("cust_id", -1), ("prod_id", -1), ("qty", -1), ("ts", -1)])
cols = ["cust_id", "prod_id", "qty", "ts"]
for _ in range(2): b = t.to_csv_device(*cols); n = b.nbytes; b.free()
reps = 5
ctx.stats(enable=True, reset=True)
for _ in range(reps): b = t.to_csv_device(*cols); b.free()
st = ctx.stats(); ctx.stats(enable=False)
ms = {k: round(v["ms"] / reps, 3) for k, v in st.items()}
tot = sum(ms.values())
print("ToCsv", rows, "rows ->", n, "bytes;", ms, "GB/s of output: %.1f" % (n / (tot * 1e-3) / 1e9), flush=True)
