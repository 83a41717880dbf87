"""Bucket tools/ncu_lines.py output by kernel region: python tools/ncu_regions.py lines.txt"""
import re, sys, collections
regions=[(0,76,'misc-top'),(77,145,'bytesrc/seq'),(239,390,'helpers'),(391,454,'flat_line/finish'),(455,613,'lookback fns'),(614,656,'tile setup/TMA'),(657,683,'classify'),(684,723,'chain1+IQ'),(724,812,'index build'),(813,885,'pass1 loop'),(886,911,'block scan'),(912,942,'chain2'),(943,1040,'pass2 staged'),(1041,1200,'pass2 direct')]
agg=collections.Counter(); samp=collections.Counter()
for ln in open(sys.argv[1]):
    m=re.match(r'(\S+):\s*(\d+) inst\s+([\d.]+)% thr/inst\s+([\d.]+) samples\s+([\d.]+)%',ln)
    if not m: continue
    f,line,pi,thr,ps=m.group(1),int(m.group(2)),float(m.group(3)),float(m.group(4)),float(m.group(5))
    if f=='parse_kernels.cuh':
        name=[n for a,b,n in regions if a<=line<=b]; name=name[0] if name else 'other-pk'
    else: name=f
    agg[name]+=pi; samp[name]+=ps
for k,v in agg.most_common(14): print(f'{k:32s} inst {v:5.1f}%  samples {samp[k]:5.1f}%')
