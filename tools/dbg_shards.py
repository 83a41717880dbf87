import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csvplus_b200 as cp
from csvplus_b200.dist import ShardedParse
from oracle import oracle as orc
lines = [b"a,b,c"] + [b"%d,x,y" % i for i in range(30000)]
lines[20001] = b"oops,too,many,fields"
lines[25000] = b'bare"quote,x,y'
data = b"\n".join(lines) + b"\n"
ctx = cp.Context(0)
world = 4
ranks = [ShardedParse(ctx, r, world, len(data), lambda lo, hi: data[lo:hi]) for r in range(world)]
q = [sp.step1_parity() for sp in ranks]
info = [sp.step2_parse(q) for sp in ranks]
print("parity", q, "info", info)
for sp in ranks:
    n = len(sp.table)
    print(sp.rank, sp.lo, sp.hi, n, sp.records, sp.err, data[sp.lo:sp.lo + 12], sp.table.values("a", 0, min(2, n)), sp.table.values("a", max(0, n - 2), n))
o = orc.reader_rows(data)
print("oracle", len(o), o.error)
