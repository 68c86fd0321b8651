#!/bin/bash
# r05c: 4-GPU bench at HEAD, launched as the driver does
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r05c_bench_n4.json 2> gpurun_out/r05c_bench_n4.err; echo "exit $?"
tail -c 600 gpurun_out/r05c_bench_n4.json; tail -3 gpurun_out/r05c_bench_n4.err
