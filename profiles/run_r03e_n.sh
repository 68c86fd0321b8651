#!/bin/bash
# r03e: N-GPU run as the driver launches it -- bench line (cfg 3 one-launch loop, cfg 4 DDP legs incl. the graph-captured
# all-reduce, cfg 5 SE(3) frames), NCCL shard-invariance check, reference arm
mkdir -p gpurun_out
N=${1:-2}
T=r03e
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
   > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; echo "bench exit $?"; tail -c 800 gpurun_out/${T}_bench_n$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r03e_bench_n{n}.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["pcie_floor"])
    for k in ("reverse_loop_1000_steps", "ddp_train_rows_per_sec", "se3_frames_cfg5", "mmd_pairs_per_sec", "reverse_particle_steps_per_sec"):
        print(k, json.dumps(d["extra"][k])[:1500])
except Exception as e:
    print("no bench line:", e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/multi_gpu_check.py \
   > gpurun_out/${T}_multi_gpu_check_n$N.json 2> gpurun_out/${T}_check_n$N.err; echo "check exit $?"; tail -c 900 gpurun_out/${T}_multi_gpu_check_n$N.json; tail -c 300 gpurun_out/${T}_check_n$N.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 \
   > gpurun_out/${T}_bench_reference_n$N.json 2>> gpurun_out/${T}_bench_n$N.err; echo "ref exit $?"
