#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/quick_shard.py > gpurun_out/r2e_shard_auto.json 2> gpurun_out/r2e_shard.err
MRMP_CROSS_SPLIT=1 python tools/quick_shard.py > gpurun_out/r2e_shard_s1.json 2>> gpurun_out/r2e_shard.err
MRMP_CROSS_SPLIT=4 python tools/quick_shard.py > gpurun_out/r2e_shard_s4.json 2>> gpurun_out/r2e_shard.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2e_shard_launches.csv python tools/quick_shard.py > gpurun_out/r2e_ncu.log 2>&1
cat gpurun_out/r2e_shard_*.json; tail -3 gpurun_out/r2e_shard.err
