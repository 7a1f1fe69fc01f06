#!/bin/bash
mkdir -p gpurun_out
for s in 1792x512x512 1792x2048x512 1792x512x2048; do python scripts/gemm_microbench.py $s 50 2>&1 | grep "engine 2"; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 1 -f -o /tmp/tc64 python scripts/gemm_microbench.py 1792x512x512 8 > /dev/null 2>&1
ncu -i /tmp/tc64.ncu-rep --page details > gpurun_out/r2_gemm_tc64_ncu_details.txt 2>&1
ncu -i /tmp/tc64.ncu-rep --page source --csv > /tmp/tc64_source.csv 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open('/tmp/tc64_source.csv')))
hdr, data = rows[1], rows[2:]
iS, iN, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iN] or 0) for r in data if len(r) > iN)
with open('gpurun_out/r2_gemm_tc64_source_top.txt', 'w') as f:
    f.write("gemm_tc_kernel<64>, 1792x512x512: warp-stall samples by SASS instruction (top 30 of %d samples)\n" % tot)
    for r in sorted([r for r in data if len(r) > iN], key=lambda r: -int(r[iN] or 0))[:30]:
        st = sorted(((hdr[i], int(r[i] or 0)) for i in stall if (r[i] or '0') != '0'), key=lambda kv: -kv[1])[:2]
        f.write("%6s samples  %8s executed  %-70s %s\n" % (r[iN], r[iE], r[iS][:70], st))
PY
grep -i "duration\|Registers Per\|Grid Size" gpurun_out/r2_gemm_tc64_ncu_details.txt | head -5
head -20 gpurun_out/r2_gemm_tc64_source_top.txt
