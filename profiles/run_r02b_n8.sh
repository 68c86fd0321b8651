#!/bin/bash
# r02b: 8-GPU end-to-end (host-buffer) leg with and without NUMA binding of the ranks
mkdir -p gpurun_out
N=${1:-8}
( nvidia-smi topo -m; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ "$(cat $d/class)" = "0x030200" ]; then echo "$d numa_node=$(cat $d/numa_node)"; fi; done ) > gpurun_out/r02b_topology.txt 2>&1
tail -12 gpurun_out/r02b_topology.txt
for mode in bind nobind bind; do
  flag=""; [ $mode = nobind ] && flag="--no-numa-bind"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --no-extra $flag \
     2>> gpurun_out/r02b.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps({'mode':'$mode','value':d['value'],'ms':d['ms_per_step'],'e2e':d['e2e']}))
" | tee -a gpurun_out/r02b_e2e_n$N.jsonl
done
tail -c 500 gpurun_out/r02b.err
