#!/bin/bash
# Round-1 evidence refresh after the GEMM epilogue work (the persistent kernels are unchanged since r1m, their
# ncu --set full captures stay): bench line, ncu launch lists of the three paths, ncu details + source hot spots of
# the tcgen05 GEMM.  Usage (through gpurun): bash scripts/capture_profiles_r1n.sh
set -u
tag=r1n
out=gpurun_out
mkdir -p $out
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 300 $NCU -c 400 --log-file $out/${tag}_launches_greedy.csv python bench.py --steps 2 --warmup 3 --skip-extra --skip-cpu > /dev/null 2>&1
timeout 300 $NCU -c 400 --log-file $out/${tag}_launches_train.csv python scripts/profile_path.py train > /dev/null 2>&1
timeout 300 $NCU -c 400 --log-file $out/${tag}_launches_beam.csv python scripts/profile_path.py beam > /dev/null 2>&1
for w in greedy train beam; do python scripts/profile_path.py $w 5 > $out/${tag}_events_$w.txt 2>/dev/null; done
bash scripts/tc_probe.sh 1792x2048x512
mv $out/r1n_gemm_tc128_source.csv $out/r1n_gemm_tc128_source_full.csv
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r1n_gemm_tc128_source_full.csv')))
hdr, data = rows[1], rows[2:]
iS, iN, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iN] or 0) for r in data if len(r) > iN)
with open('gpurun_out/r1n_gemm_tc128_source_top.txt', 'w') as f:
    f.write("gemm_tc_kernel<128>, 1792x2048x512: warp-stall samples by SASS instruction (top 25 of %d samples)\n" % tot)
    for r in sorted([r for r in data if len(r) > iN], key=lambda r: -int(r[iN] or 0))[:25]:
        st = sorted(((hdr[i], int(r[i] or 0)) for i in stall if (r[i] or '0') != '0'), key=lambda kv: -kv[1])[:2]
        f.write("%6s samples  %8s executed  %-70s %s\n" % (r[iN], r[iE], r[iS][:70], st))
PY
rm -f $out/r1n_gemm_tc128_source_full.csv
ls -la $out
