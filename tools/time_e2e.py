"""H2D floor and pipelined e2e experiments: python tools/time_e2e.py"""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import csvplus_b200 as cp
ctx = cp.Context(0)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
d = ctx.gen_csv("orders", (0, rows), n_cust=1_000_000, n_prod=1000)
h = ctx.host_alloc(d.nbytes)
ctx.lib.cpb_memcpy_d2h(ctx.h, h.ptr, d.ptr, d.nbytes)
dst = ctx.device_alloc(d.nbytes)
for _ in range(2):
    t0 = time.perf_counter(); ctx.lib.cpb_memcpy_h2d(ctx.h, dst.ptr, h.ptr, d.nbytes); dt = time.perf_counter() - t0
print("H2D", d.nbytes / dt / 1e9, "GB/s", dt * 1e3, "ms")
SPEC = [("cust_id", 1), ("prod_id", 2), ("qty", 3), ("ts", 4)]
v = h.array()
for nb in (4, 8, 16):
    bounds = [0]
    for b in range(1, nb):
        pos = b * h.nbytes // nb
        bounds.append(pos + int(np.flatnonzero(v[pos:pos + 4096] == 10)[0]) + 1)
    bounds.append(h.nbytes)
    workers = [cp.Context(0), cp.Context(0)]
    def work(wi, res):
        w = workers[wi]
        for b in range(wi, nb, 2):
            lo, hi = bounds[b], bounds[b + 1]
            if b == 0: t, e = cp.parse_csv(w, h.ptr, nbytes=hi, spec=[(c, -1) for c, _ in SPEC])
            else: t, e = cp.parse_csv(w, h.ptr + lo, nbytes=hi - lo, spec=SPEC, header_from_first_row=False, num_fields=5)
            res.append(len(t))
    for rep in range(3):
        res = []
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(i, res)) for i in range(2)]
        [t.start() for t in th]; [t.join() for t in th]
        for w in workers: w.sync()
        dt = time.perf_counter() - t0
    print("batches", nb, "2 ctx parse-only e2e", dt * 1e3, "ms", sum(res))
    res = []
    t0 = time.perf_counter(); work(0, res); work(1, res); [w.sync() for w in workers]; dt = time.perf_counter() - t0
    print("batches", nb, "serial", dt * 1e3, "ms")
