#!/bin/bash
# ncu --set full captures of the round-2 kernels (one launch each), reports back in gpurun_out/
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_step_grouped -s 8 -c 1 -o gpurun_out/r2_step -f python scripts/profile_path.py beam 1 > gpurun_out/ncu_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:train_grouped -s 2 -c 1 -o gpurun_out/r2_train -f python scripts/profile_path.py train 1 > gpurun_out/ncu_train.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_grouped_kernel -s 2 -c 1 -o gpurun_out/r2_greedy -f python scripts/profile_path.py greedy 1 > gpurun_out/ncu_greedy.log 2>&1
ls -la gpurun_out/*.ncu-rep
