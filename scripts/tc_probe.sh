#!/bin/bash
# ncu --set full of one tcgen05 GEMM launch (128 x 128 tiles) with the source page: where do the warps wait?
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 1 -f -o gpurun_out/r1n_gemm_tc128 python scripts/gemm_microbench.py ${1:-1792x2048x512} 8 > /dev/null 2>&1
ncu -i gpurun_out/r1n_gemm_tc128.ncu-rep --page details > gpurun_out/r1n_gemm_tc128_ncu_details.txt 2>&1
ncu -i gpurun_out/r1n_gemm_tc128.ncu-rep --page source --csv > gpurun_out/r1n_gemm_tc128_source.csv 2>&1
rm -f gpurun_out/r1n_gemm_tc128.ncu-rep
