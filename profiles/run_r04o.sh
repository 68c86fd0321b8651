#!/bin/bash
# r04o: 2-GPU bench at HEAD (launched as the driver does) + the multi-GPU GPU tests
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r04o_bench_n2.json 2> gpurun_out/r04o_bench_n2.err; echo "exit $?"
tail -c 1500 gpurun_out/r04o_bench_n2.json; tail -5 gpurun_out/r04o_bench_n2.err
