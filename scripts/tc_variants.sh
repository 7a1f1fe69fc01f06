#!/bin/bash
# kernel-only durations (ncu) of the tcgen05 GEMM kernel under debug switches
# XG_TC_DEBUG bits: 1 no MMA, 2 no promotion (one long chain), 4 no lo loads
for shape in 1792x2048x512 64x10000x512 64x2048x512; do
  for dbg in 0 2 1 5 7; do
    XG_TC_DEBUG=$dbg ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm_tc_kernel -s 3 -c 10 --csv python scripts/gemm_microbench.py $shape 12 2>/dev/null | python -c "
import csv,sys
v=[float(r[-1].replace(',','')) for r in csv.reader(l for l in sys.stdin if l.startswith('\"')) if r and r[-3]=='gpu__time_duration.sum']
u=sorted(v); print('shape $shape dbg=$dbg  kernel us: median %.1f min %.1f (n=%d)'%(u[len(u)//2]/1000 if u and u[0]>1000 else (u[len(u)//2] if u else -1), (u[0]/1000 if u and u[0]>1000 else (u[0] if u else -1)), len(u)))"
  done
done
