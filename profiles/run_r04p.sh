#!/bin/bash
# r04p: 8-GPU bench at HEAD, launched as the driver does
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r04p_bench_n8.json 2> gpurun_out/r04p_bench_n8.err; echo "exit $?"
tail -c 600 gpurun_out/r04p_bench_n8.json; tail -3 gpurun_out/r04p_bench_n8.err
