#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tests/tools/probe_train.py --cpu > gpurun_out/r01z_probe_train.jsonl 2> gpurun_out/r01z.err
cat gpurun_out/r01z_probe_train.jsonl; tail -c 600 gpurun_out/r01z.err
