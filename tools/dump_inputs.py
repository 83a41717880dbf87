"""Writes the synthetic bench inputs as files for baseline/go: python tools/dump_inputs.py OUTDIR [people] [customers] [orders]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csvplus_b200 as cp
out = sys.argv[1]; os.makedirs(out, exist_ok=True)
n_people = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
n_cust = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
n_orders = int(sys.argv[4]) if len(sys.argv) > 4 else 2_000_000
ctx = cp.Context(0)
ctx.gen_csv("people", (0, n_people), seed=0xC5B200).to_host().tofile(os.path.join(out, "people.csv"))
ctx.gen_csv("customers", (0, n_cust), seed=0xC5B200, n_cust=n_cust, permute=True).to_host().tofile(os.path.join(out, "customers.csv"))
ctx.gen_csv("orders", (0, n_orders), seed=0xC5B200, n_cust=n_cust, n_prod=1_000_000).to_host().tofile(os.path.join(out, "orders.csv"))
