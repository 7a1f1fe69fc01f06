#!/bin/bash
# Capture the round's evidence on one B200 (run through gpurun; everything lands in gpurun_out/):
#   bench line, ncu launch lists (durations only, no clock control) of the greedy / train / beam paths,
#   ncu --set full details of the persistent kernels.  Usage: bash scripts/capture_profiles.sh <tag>
set -u
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 300 $NCU -c 400 --log-file $out/${tag}_launches_greedy.csv python bench.py --steps 2 --warmup 3 --skip-extra > /dev/null 2>&1
timeout 300 $NCU -c 400 --log-file $out/${tag}_launches_train.csv python scripts/profile_path.py train > /dev/null 2>&1
timeout 300 $NCU -c 400 --log-file $out/${tag}_launches_beam.csv python scripts/profile_path.py beam > /dev/null 2>&1
full() {  # name regex skip script args
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $out/${tag}_$1 ${@:4} > /dev/null 2>&1
  ncu -i $out/${tag}_$1.ncu-rep --page details > $out/${tag}_$1_ncu_details.txt 2>&1
  ncu -i $out/${tag}_$1.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
if len(rows)>=3:
    h=rows[0]; v=rows[-1]
    for k in ('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','lts__t_sector_hit_rate.pct','sm__inst_executed_pipe_tensor.sum'):
        if k in h: print(k, rows[1][h.index(k)], v[h.index(k)])
" > $out/${tag}_$1_ncu_raw.txt 2>&1
  rm -f $out/${tag}_$1.ncu-rep
}
full decode_persistent 'decode_persistent_kernel' 2 python scripts/greedy_once.py 4
full decode_step_persistent 'decode_step_persistent_kernel' 5 python scripts/profile_path.py beam
full train_decode_persistent 'decode_persistent_kernel' 2 python scripts/profile_path.py train
full decode_bwd_persistent 'decode_bwd_persistent_kernel' 2 python scripts/profile_path.py train
ls -la $out
