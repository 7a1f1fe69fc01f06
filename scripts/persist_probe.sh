#!/bin/bash
# ncu --set full of one decode_persistent launch with the source page (stall samples per SASS instruction)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_persistent_kernel -s 2 -c 1 -f -o gpurun_out/r1n_decode_persistent python scripts/greedy_once.py 4 > /dev/null 2>&1
ncu -i gpurun_out/r1n_decode_persistent.ncu-rep --page source --csv > gpurun_out/r1n_decode_persistent_source.csv 2>&1
ncu -i gpurun_out/r1n_decode_persistent.ncu-rep --page details > gpurun_out/r1n_decode_persistent_ncu_details.txt 2>&1
rm -f gpurun_out/r1n_decode_persistent.ncu-rep
