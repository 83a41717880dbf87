"""Aggregate an ncu report's source page per CUDA line: python tools/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None; fname = ""
agg = collections.OrderedDict()
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; ii = hdr.index("Instructions Executed"); ti = hdr.index("Thread Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] != "":
        cur = (fname, int(r[0]), r[1].strip()); agg.setdefault(cur, [0, 0, 0, 0])
    else:
        f = lambda x: float(x) if x not in ('', '-') else 0.0
        a = agg[cur]; a[0] += f(r[ii]); a[1] += f(r[ti]); a[2] += f(r[si]); a[3] += 1
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
print(f"total warp-instr {tot:.3g}, samples {tots:.0f}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:4d} inst {100*a[0]/tot:5.1f}% thr/inst {a[1]/max(a[0],1):5.1f} samples {100*a[2]/max(tots,1):5.1f}% sass {a[3]:4d} | {k[2][:90]}")
