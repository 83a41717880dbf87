"""Pipelined e2e join experiment with per-phase wall clock: python tools/time_e2e_join.py [orders] [batches]"""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import csvplus_b200 as cp
ctx = cp.Context(0)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ncust = rows // 10
dc = ctx.gen_csv("customers", (0, ncust), n_cust=ncust, permute=True)
d = ctx.gen_csv("orders", (0, rows), n_cust=ncust, n_prod=1000)
h = ctx.host_alloc(d.nbytes); ctx.lib.cpb_memcpy_d2h(ctx.h, h.ptr, d.ptr, d.nbytes)
hc = ctx.host_alloc(dc.nbytes); ctx.lib.cpb_memcpy_d2h(ctx.h, hc.ptr, dc.ptr, dc.nbytes)
SPEC = [("cust_id", 1), ("prod_id", 2), ("qty", 3), ("ts", 4)]
v = h.array()
bounds = [0]
for b in range(1, nb):
    pos = b * h.nbytes // nb
    bounds.append(pos + int(np.flatnonzero(v[pos:pos + 4096] == 10)[0]) + 1)
bounds.append(h.nbytes)
workers = [cp.Context(0), cp.Context(0)]
def step(log=False):
    t0 = time.perf_counter()
    tc, _ = cp.parse_csv(ctx, hc, spec=[("id", -1), ("name", -1), ("surname", -1)])
    idx = tc.index_on("id", unique=True)
    warm, _ = cp.parse_csv(ctx, b"cust_id\n0\n"); warm.join(idx, "cust_id"); ctx.sync()
    t1 = time.perf_counter()
    res = [None] * nb; tl = [[], []]
    def work(wi):
        w = workers[wi]
        for b in range(wi, nb, 2):
            a = time.perf_counter()
            lo, hi = bounds[b], bounds[b + 1]
            if b == 0: t, e = cp.parse_csv(w, h.ptr, nbytes=hi, spec=[(c, -1) for c, _ in SPEC])
            else: t, e = cp.parse_csv(w, h.ptr + lo, nbytes=hi - lo, spec=SPEC, header_from_first_row=False, num_fields=5)
            m = time.perf_counter()
            res[b] = t.join(idx, "cust_id")
            tl[wi].append((round((m - a) * 1e3, 1), round((time.perf_counter() - m) * 1e3, 1)))
    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]; [t.join() for t in th]
    for w in workers: w.sync()
    t2 = time.perf_counter()
    n = sum(len(r) for r in res)
    del res
    if log: print("build %.1f ms, probe side %.1f ms, total %.1f ms, rows %d" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t2 - t0) * 1e3, n), tl)
for i in range(5): step(i >= 2)
