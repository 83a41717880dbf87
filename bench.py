#!/usr/bin/env python
"""bench.py — the measurement contract of the csvplus_b200 hot path.

Headline workload (BASELINE.json configs[2], the "rows/sec end-to-end Join" half of the metric):
    customers (10 M rows, 6 cols)  -> parse + SelectColumns(id,name,surname) -> UniqueIndexOn(id)
    orders    (100 M rows per GPU) -> parse + SelectColumns(cust_id,prod_id,qty,ts) -> Join(idx, "cust_id")
one "step" = one full pass of that pipeline over synthetic CSV (SURVEY §8d shapes).
  value = probe rows/s with the CSV bytes already resident in HBM (whole job, all ranks);
  e2e   = the same through the public API with HOST (pinned) CSV buffers: H2D inside the timed region and
          a D2H read of the result summary (row count + bytes per column).
The "CSV parse GB/s" half of the metric (BASELINE.json configs[1]: people 100 M rows x 6 cols, parse +
SelectColumns(name,surname,id) + Filter(Like{name: Amelia})) is timed in the same run and reported under
"csv_parse", together with its own roofline.  "roofline" describes the dominant kernel of the step (csv_scan).

Multi-GPU (torchrun, one rank per GPU): the probe stream is sharded by row range (100 M orders per rank, weak
scaling); each rank parses 1/N of the build side and the build-side columns are all-gathered with NCCL.

--impl reference: the reference (pure Go) cannot be built here (no Go toolchain), so the arm times the CPU
oracle port of the same pipeline (oracle/, kind "port") on a bounded sample, single-threaded because the
reference is strictly single-threaded (csvplus.go has no goroutines).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0xC5B200
CUST_COLS = [("id", -1), ("name", -1), ("surname", -1)]
ORDER_COLS = [("cust_id", -1), ("prod_id", -1), ("qty", -1), ("ts", -1)]
PEOPLE_COLS = [("name", -1), ("surname", -1), ("id", -1)]


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi polled every 100 ms in the background (a 20 ms poll measurably slows the step: its NVML queries
    contend with the launches, 21.3 -> 23.5 ms); stop(t0, t1) keeps the samples whose timestamps fall inside the
    timed region [t0, t1] (wall clock), so that it may be started early (nvidia-smi needs ~0.1 s to start)"""

    def __init__(self, index: int):
        self.p = None
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--id={index}",
                 "--query-gpu=timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self, t0: float | None = None, t1: float | None = None):
        import datetime
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill(); out = ""
        rows = []
        for ln in out.splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), f[3:7]))
            except ValueError:
                continue
        inside = [r for r in rows if t0 is not None and t1 is not None and t0 - 0.01 <= r[0] <= t1 + 0.01]
        window = "timed region" if inside else "whole run (no sample fell inside the timed region)"
        use = inside or rows
        reasons = set()
        for r in use:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median([r[1] for r in use]) if use else None,
                "sm_max_mhz": max(r[2] for r in use) if use else None,
                "reasons": sorted(reasons), "samples": len(use), "window": window}


# ------------------------------------------------------------------ reference arm / cpu baseline (oracle port)
def cpu_join_sample(n_orders: int, n_cust: int, ctx=None, data=None):
    """the same pipeline in the CPU oracle, single thread; returns (rows/s, seconds, sample text)"""
    from oracle import oracle as orc
    if data is None:
        cust = ctx.gen_csv("customers", (0, n_cust), seed=SEED, n_cust=n_cust, permute=True).to_host()
        orders = ctx.gen_csv("orders", (0, n_orders), seed=SEED, n_cust=n_cust, n_prod=1_000_000).to_host()
    else:
        cust, orders = data
    t0 = time.perf_counter()
    idx = orc.reader_rows(cust, select=[c for c, _ in CUST_COLS]).unique_index_on("id", stable=False)
    joined = orc.reader_rows(orders, select=[c for c, _ in ORDER_COLS]).join(idx, "cust_id")
    n = len(joined)
    dt = time.perf_counter() - t0
    assert n == n_orders
    return n_orders / dt, dt, f"orders {n_orders} rows x customers {n_cust} rows, same generator, 1 thread"


def cpu_parse_sample(n_rows: int, ctx):
    from oracle import oracle as orc
    people = ctx.gen_csv("people", (0, n_rows), seed=SEED).to_host()
    t0 = time.perf_counter()
    r = orc.reader_rows(people, select=[c for c, _ in PEOPLE_COLS], pred=orc.Like({"name": "Amelia"}))
    n = len(r)
    dt = time.perf_counter() - t0
    return people.size / dt / 1e9, n_rows / dt, dt, n


def run_reference(args, rank):
    if rank != 0:
        return
    import csvplus_b200 as cp
    ctx = cp.Context(int(os.environ.get("LOCAL_RANK", "0")))
    n_orders, n_cust = args.ref_orders, args.ref_customers
    cust = ctx.gen_csv("customers", (0, n_cust), seed=SEED, n_cust=n_cust, permute=True).to_host()
    orders = ctx.gen_csv("orders", (0, n_orders), seed=SEED, n_cust=n_cust, n_prod=1_000_000).to_host()
    times = []
    for i in range(args.warmup + args.steps):
        _, dt, sample = cpu_join_sample(n_orders, n_cust, data=(cust, orders))
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = n_orders / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "rows/sec end-to-end Join", "value": value, "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": 1, "kind": "port",
                         "sample": sample + " (oracle/ C++ restatement of csvplus; the Go reference has no toolchain here)"},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(world):
    return {"workload": "orders(100 M rows/GPU) x customers(10 M): parse+SelectColumns, UniqueIndexOn(id), Join(cust_id) "
                        "[BASELINE configs[2]; configs[1] parse+filter reported under csv_parse]",
            "orders_rows_per_gpu": ORD_ROWS, "customers_rows": CUST_ROWS, "people_rows": PEOPLE_ROWS,
            "parallelism": f"row-range shards x{world}, build side all-gathered (NCCL)" if world > 1 else "single GPU",
            "l2": "inputs (>= 0.44 GB per pass) exceed the 126 MB L2; no flush needed"}


ORD_ROWS = 100_000_000
CUST_ROWS = 10_000_000
PEOPLE_ROWS = 100_000_000


def ncu_traffic(path: str):
    """mean dram__bytes_read.sum + dram__bytes_write.sum per csv_scan launch from the committed ncu metrics pass of this
    bench command's join step (profiles/r1_traffic_csv_scan.csv: the customers and the orders parse); None if absent"""
    import csv
    try:
        per_id = {}
        with open(path) as f:
            rows = [r for r in csv.reader(f) if len(r) > 14]
        hdr = rows[0]
        ki, mi, ui, vi, ii = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
        for r in rows[1:]:
            if "csv_scan" in r[ki] and r[mi] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[ui]]
                per_id[r[ii]] = per_id.get(r[ii], 0.0) + float(r[vi].replace(",", "")) * scale
        return sum(per_id.values()) / len(per_id) if per_id else None
    except Exception:
        return None


def bind_to_gpu_numa_node(local: int):
    """Pin this process (and the pinned host buffers it allocates afterwards) to the NUMA node the GPU hangs off,
    so that H2D copies do not cross the socket interconnect.  Best effort: silently skipped where /sys is not
    available.  Returns the node id or None."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def main():
    global ORD_ROWS, CUST_ROWS, PEOPLE_ROWS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--orders", type=int, default=ORD_ROWS, help="probe rows per GPU")
    ap.add_argument("--customers", type=int, default=CUST_ROWS)
    ap.add_argument("--people", type=int, default=PEOPLE_ROWS)
    ap.add_argument("--ref-orders", type=int, default=2_000_000)
    ap.add_argument("--ref-customers", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-batches", type=int, default=8)
    args = ap.parse_args()
    ORD_ROWS, CUST_ROWS, PEOPLE_ROWS = args.orders, args.customers, args.people
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    import csvplus_b200 as cp
    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = cp.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))

    # ---------------- synthetic inputs (device resident; pinned host copies for the e2e leg)
    n_prod = 1_000_000
    cust_lo, cust_hi = rank * CUST_ROWS // world, (rank + 1) * CUST_ROWS // world
    d_cust = ctx.gen_csv("customers", (cust_lo, cust_hi), seed=SEED, n_cust=CUST_ROWS, permute=True, header=True)
    d_orders = ctx.gen_csv("orders", (rank * ORD_ROWS, (rank + 1) * ORD_ROWS), seed=SEED, n_cust=CUST_ROWS, n_prod=n_prod, header=True)
    d_people = ctx.gen_csv("people", (rank * PEOPLE_ROWS, (rank + 1) * PEOPLE_ROWS), seed=SEED, header=True)
    ctx.sync()

    from csvplus_b200.dist import allgather_table, allgather_table_async

    dbg = bool(os.environ.get("BENCH_DEBUG")) and rank == 0

    def join_step(cust_src, orders_src):
        marks = []

        def mark(name):
            if dbg:
                ctx.sync(); torch.cuda.synchronize()
                marks.append((name, time.perf_counter()))
        mark("start")
        tc, err = cp.parse_csv(ctx, cust_src, spec=CUST_COLS)
        assert err is None
        mark("parse_cust")
        pending = None
        if world > 1:  # the build-side all-gather (NCCL, torch's stream) runs under the parse of the probe shard
            pending = allgather_table_async(ctx, tc, dist)
            mark("allgather_start")
        to, err = cp.parse_csv(ctx, orders_src, spec=ORDER_COLS)
        assert err is None
        mark("parse_orders")
        if pending is not None:
            tc = pending.wait()
            mark("allgather_finish")
        idx = tc.index_on("id", unique=True)
        mark("index")
        j = to.join(idx, "cust_id")
        mark("join")
        if dbg:
            print("phases(ms): " + " ".join("%s=%.2f" % (marks[i][0], (marks[i][1] - marks[i - 1][1]) * 1e3)
                                           for i in range(1, len(marks))), file=sys.stderr)
        return j

    def parse_step(people_src):
        t, err = cp.parse_csv(ctx, people_src, spec=PEOPLE_COLS, pred=cp.Like({"name": "Amelia"}))
        assert err is None
        return t

    def timed(fn, steps, warmup, sampler_dev=None):
        sampler = ClockSampler(sampler_dev) if sampler_dev is not None else None  # polls through warm-up + timed steps
        for _ in range(warmup):
            r = fn(); del r
        ctx.sync(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.stats(enable=True, reset=True)
        l0 = ctx.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin = time.time()
        e0.record(stream)
        rows = 0
        for _ in range(steps):
            r = fn(); rows = len(r); del r
        e1.record(stream)
        ctx.sync(); torch.cuda.synchronize()
        t_end = time.time()
        ms = e0.elapsed_time(e1)
        if world > 1:
            dist.barrier()
            tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        stats = ctx.stats()
        ctx.stats(enable=False)
        clocks = sampler.stop(t_begin, t_end) if sampler else None
        return ms / steps, rows, stats, ctx.kernel_launches() - l0, clocks

    # ---------------- device-resident timing (value)
    ms_join, out_rows, st_join, launches, clocks = timed(lambda: join_step(d_cust, d_orders), args.steps, args.warmup, local)
    ms_parse, parse_rows, st_parse, _, _ = timed(lambda: parse_step(d_people), args.steps, args.warmup)
    peak, peak_kind = hbm_peak()

    def roof(st, traffic=None):
        s = st.get("csv_scan")
        if not s or not s["launches"]:
            return None
        per_launch_bytes = s["algo_bytes"] / s["launches"]
        per_launch_ms = s["ms"] / s["launches"]
        ach = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": "csv_scan", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write, profiles/r1_traffic_csv_scan.csv)", "peak_kind": f"of {peak_kind}", "launches": s["launches"], "ms_per_launch": per_launch_ms,
                "algo_bytes_per_launch": per_launch_bytes}

    # ---------------- end-to-end timing from pinned host buffers
    e2e = None
    parse_e2e = None
    if not args.no_e2e:
        h_cust, h_orders, h_people = ctx.host_alloc(d_cust.nbytes), ctx.host_alloc(d_orders.nbytes), ctx.host_alloc(d_people.nbytes)
        for h, d in ((h_cust, d_cust), (h_orders, d_orders), (h_people, d_people)):
            ctx.lib.cpb_memcpy_d2h(ctx.h, h.ptr, d.ptr, d.nbytes)
        # e2e is a streaming pipeline, as a csvplus user would run a large file: the probe CSV is handed to the API in
        # batches of complete records; two contexts (two CUDA streams) alternate batches so that the H2D copy of
        # batch i+1 overlaps the parse+join of batch i.  The build side (customers) goes first on the main context.
        import threading

        import numpy as np
        nbatch = max(2, args.e2e_batches)
        oview = h_orders.array()
        bounds = [0]
        for b in range(1, nbatch):
            pos = b * h_orders.nbytes // nbatch
            nl = int(np.flatnonzero(oview[pos:pos + 4096] == 10)[0])  # synthetic rows hold no quoted newlines
            bounds.append(pos + nl + 1)
        bounds.append(h_orders.nbytes)
        workers = [cp.Context(local), cp.Context(local)]
        wstreams = [torch.cuda.ExternalStream(w.stream, device=torch.device("cuda", local)) for w in workers]
        ORDER_ASSUME = [("cust_id", 1), ("prod_id", 2), ("qty", 3), ("ts", 4)]

        def join_e2e():
            t_a = time.perf_counter()
            results = [None] * nbatch
            ready = threading.Event()
            box = {}

            def work(wi):
                w = workers[wi]
                for b in range(wi, nbatch, 2):
                    lo, hi = bounds[b], bounds[b + 1]
                    if b == 0:
                        t, e = cp.parse_csv(w, h_orders.ptr, nbytes=hi, spec=ORDER_COLS)
                    else:
                        t, e = cp.parse_csv(w, h_orders.ptr + lo, nbytes=hi - lo, spec=ORDER_ASSUME, header_from_first_row=False, num_fields=5)
                    assert e is None
                    ready.wait()  # the build side is parsed / indexed concurrently on the main context
                    results[b] = t.join(box["idx"], "cust_id")
            th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
            for t in th:
                t.start()
            tc, err = cp.parse_csv(ctx, h_cust, spec=CUST_COLS)
            assert err is None
            if world > 1:
                tc = allgather_table(ctx, tc, dist)
            idx = tc.index_on("id", unique=True)
            warm, _ = cp.parse_csv(ctx, b"cust_id\n0\n")
            warm.join(idx, "cust_id")  # builds the probe hash table once, before the workers share the index
            ctx.sync()
            box["idx"] = idx
            ready.set()
            t_b = time.perf_counter()
            for t in th:
                t.join()
            t_c = time.perf_counter()
            # D2H read of the step's result: row count + byte total of every output column of every batch
            import ctypes as C
            nb = C.c_uint64(); rows = 0
            for wi, w in enumerate(workers):
                for b in range(wi, nbatch, 2):
                    j = results[b]; rows += len(j)
                    for i in range(len(j.columns)):
                        w.lib.cpb_table_col_bytes(w.h, j.h, i, 0, len(j), C.byref(nb))
            summary_bytes[0] = (8 * 7 + 8) * nbatch
            if os.environ.get("BENCH_DEBUG"):
                print("e2e step: build %.1f ms, probe %.1f ms, summary %.1f ms" % ((t_b - t_a) * 1e3, (t_c - t_b) * 1e3, (time.perf_counter() - t_c) * 1e3), file=sys.stderr)
            return results, rows

        def timed_multi(fn, steps, warmup):
            for _ in range(warmup):
                r = fn(); del r
            for w in [ctx] + workers:
                w.sync()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0 = torch.cuda.Event(enable_timing=True)
            ends = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e0.record(stream)
            rows = 0
            for _ in range(steps):
                r, rows = fn(); del r
            ends[0].record(stream); ends[1].record(wstreams[0]); ends[2].record(wstreams[1])
            for w in [ctx] + workers:
                w.sync()
            torch.cuda.synchronize()
            ms = max(e0.elapsed_time(e) for e in ends)
            if world > 1:
                dist.barrier()
                tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
                dist.all_reduce(tms, op=dist.ReduceOp.MAX)
                ms = float(tms.item())
            return ms / steps, rows

        summary_bytes = [0]
        ms_e2e, e2e_rows = timed_multi(join_e2e, args.steps, args.warmup)
        assert e2e_rows == out_rows, (e2e_rows, out_rows)
        e2e = {"value": world * ORD_ROWS / (ms_e2e * 1e-3), "unit": "rows/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": h_cust.nbytes + h_orders.nbytes, "d2h_bytes_per_step": summary_bytes[0],
               "batches": nbatch, "host_numa_node": numa_node,
               "note": "pinned host CSV -> H2D -> parse/index/join on the GPU through the public API; the probe file is streamed in "
                       "%d batches of complete records over two contexts so H2D overlaps compute; results stay in HBM, their "
                       "summaries are read back" % nbatch}
        ms_pe2e, _, _, _, _ = timed(lambda: parse_step(h_people), args.steps, args.warmup)
        parse_e2e = {"value": world * h_people.nbytes / (ms_pe2e * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms_pe2e,
                     "h2d_bytes_per_step": h_people.nbytes}

    # ---------------- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, sample = cpu_join_sample(args.ref_orders, args.ref_customers, ctx=ctx)
        pg, pr, pdt, _ = cpu_parse_sample(4_000_000, ctx)
        cpu = {"value": v, "unit": "rows/s", "cores": 1, "kind": "port", "seconds": dt,
               "sample": sample + "; C++ restatement proxy of the Go reference (single-threaded like it), host cores: %d" % (os.cpu_count() or 0),
               "csv_parse": {"value": pg, "unit": "GB/s", "rows_per_s": pr, "sample": "people 4 M rows, parse+select+filter, 1 thread"}}

    if rank == 0:
        line = {
            "metric": "rows/sec end-to-end Join", "value": world * ORD_ROWS / (ms_join * 1e-3), "unit": "rows/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_join, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(world),
            "clocks": clocks, "gpu_launches": launches, "out_rows_per_gpu": out_rows,
            "e2e": e2e,
            "roofline": roof(st_join, ncu_traffic(os.path.join(ROOT, "profiles", "r1_traffic_csv_scan.csv")) if world == 1 else None),
            "csv_parse": {"metric": "CSV parse GB/s (configs[1]: parse+SelectColumns(name,surname,id)+Filter(Like name=Amelia))",
                          "value": world * d_people.nbytes / (ms_parse * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms_parse,
                          "rows_per_s": world * PEOPLE_ROWS / (ms_parse * 1e-3), "rows_out_per_gpu": parse_rows,
                          "input_bytes_per_gpu": d_people.nbytes, "roofline": roof(st_parse), "e2e": parse_e2e},
            "cpu_baseline": cpu,
            "kernels": {k: {"launches": v["launches"], "ms": round(v["ms"], 4)} for k, v in sorted(st_join.items())},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
