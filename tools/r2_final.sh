#!/bin/bash
# final measurements of round 2 (8-GPU box): scaling at both sizes, the single-GPU line with all sections,
# reference arm, launch list and the full ncu capture of the table kernel
set -x
mkdir -p gpurun_out
bash tools/r2_scale.sh r2p
timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/r2p_bench_n1_full.json 2> gpurun_out/r2p_bench_n1_full.err; echo rc=$?
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2p_bench_reference.json 2> gpurun_out/r2p_bench_reference.err; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2p_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cross_collide -s 2 -c 1 -o gpurun_out/r2p_cross -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2p_ncu_full.log 2>&1
ls -la gpurun_out/r2p_* | head -40
