"""Quick kernel timing: python tools/time_parse.py [rows] — prints per-kernel ms from the library's CUDA-event stats."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csvplus_b200 as cp
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
ctx = cp.Context(0)
def run(name, fn, reps=5):
    for _ in range(2): r = fn(); del r
    ctx.stats(enable=True, reset=True)
    for _ in range(reps): r = fn(); del r
    st = ctx.stats(); ctx.stats(enable=False)
    print(name, {k: (round(v["ms"] / reps, 3), v["launches"] // reps) for k, v in st.items()}, flush=True)
    return st
people = ctx.gen_csv("people", (0, rows))
orders = ctx.gen_csv("orders", (0, rows), n_cust=rows // 10, n_prod=1000)
st = run("filter", lambda: cp.parse_csv(ctx, people, spec=[("name", -1), ("surname", -1), ("id", -1)], pred=cp.Like({"name": "Amelia"}))[0])
s = st["csv_scan"]; print("  filter csv_scan GB/s (S_in):", people.nbytes / (s["ms"] / s["launches"] * 1e-3) / 1e9)
st = run("orders", lambda: cp.parse_csv(ctx, orders, spec=[("cust_id", -1), ("prod_id", -1), ("qty", -1), ("ts", -1)])[0])
s = st["csv_scan"]; print("  orders csv_scan GB/s (S_in):", orders.nbytes / (s["ms"] / s["launches"] * 1e-3) / 1e9)
st = run("all6", lambda: cp.parse_csv(ctx, people)[0])
s = st["csv_scan"]; print("  all6 csv_scan GB/s (S_in):", people.nbytes / (s["ms"] / s["launches"] * 1e-3) / 1e9)
st = run("general(comment=#)", lambda: cp.parse_csv(ctx, people, spec=[("name", -1), ("surname", -1), ("id", -1)], pred=cp.Like({"name": "Amelia"}), comment="#")[0])
tot = sum(v["ms"] for v in st.values()) / 5
print("  general path GB/s (S_in, all kernels):", people.nbytes / (tot * 1e-3) / 1e9)
