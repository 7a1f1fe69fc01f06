#!/bin/bash
run() { env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/microbench/allreduce_bench.py 2>&1 | grep "world"; }
run XG_DUMMY=1
run NCCL_ALGO=NVLS
run NCCL_MIN_CTAS=32
run NCCL_ALGO=Ring
run NCCL_ALGO=Tree
