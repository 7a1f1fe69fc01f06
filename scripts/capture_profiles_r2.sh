#!/bin/bash
# Round-2 evidence: bench line, ncu launch lists and CUDA-event tables of the three paths, ncu --set full of the grouped
# word-loop kernels (details + raw counters as text; the .ncu-rep files stay on the box).
# Usage (through gpurun): bash scripts/capture_profiles_r2.sh
set -u
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 300 $NCU -c 400 --log-file $out/${tag}_launches_greedy.csv python bench.py --steps 2 --warmup 3 --skip-extra --skip-cpu > /dev/null 2>&1
timeout 300 $NCU -c 400 --log-file $out/${tag}_launches_train.csv python scripts/profile_path.py train > /dev/null 2>&1
timeout 300 $NCU -c 400 --log-file $out/${tag}_launches_beam.csv python scripts/profile_path.py beam > /dev/null 2>&1
for w in greedy train beam; do python scripts/profile_path.py $w 5 > $out/${tag}_events_$w.txt 2>/dev/null; done
cap() {   # name, kernel regex, launches to skip, workload
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o /tmp/${tag}_$1 -f python scripts/profile_path.py $4 1 > /dev/null 2>&1
  ncu -i /tmp/${tag}_$1.ncu-rep --page details > $out/${tag}_$1_ncu_details.txt 2>/dev/null
  ncu -i /tmp/${tag}_$1.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin)); hdr, units, vals = rows[0], rows[1], rows[2]
keep = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio')
for h, u, v in zip(hdr, units, vals):
    if h in keep: print('%-80s %-12s %s' % (h, u, v))
" > $out/${tag}_$1_ncu_raw.txt
}
cap decode_grouped decode_grouped_kernel 2 greedy
cap train_grouped train_grouped 2 train
cap decode_step_grouped decode_step_grouped 2 beam      # (one launch per search since the merge runs in the kernel)
cap decode_bwd decode_bwd_persistent 2 train
ls -la $out | tail -30
