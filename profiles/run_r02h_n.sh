#!/bin/bash
# r02h: N-GPU run at HEAD as the driver launches it -- bench line, NCCL shard-invariance check, reference arm
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
   > gpurun_out/r02h_bench_n$N.json 2> gpurun_out/r02h_bench_n$N.err; echo "bench exit $?"; tail -c 400 gpurun_out/r02h_bench_n$N.err; head -c 600 gpurun_out/r02h_bench_n$N.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/multi_gpu_check.py \
   > gpurun_out/r02h_multi_gpu_check_n$N.json 2> gpurun_out/r02h_check_n$N.err; echo "check exit $?"; tail -c 600 gpurun_out/r02h_multi_gpu_check_n$N.json
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 \
   > gpurun_out/r02h_bench_reference_n$N.json 2>> gpurun_out/r02h_bench_n$N.err; echo "ref exit $?"
