#!/bin/bash
# round-2 GPU run 1: full GPU tests, bench (RPL 1 / 2), launch list, full ncu of the cross kernel
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.txt
tail -5 gpurun_out/r2a_pytest.txt
MRMP_CROSS_RPL=1 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench_rpl1.json 2> gpurun_out/r2a_bench_rpl1.err; echo rc=$?
MRMP_CROSS_RPL=2 timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2a_bench_rpl2.json 2> gpurun_out/r2a_bench_rpl2.err; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2a_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cross_collide -s 2 -c 1 -o gpurun_out/r2a_cross -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2a_ncu_full.log 2>&1
MRMP_CROSS_RPL=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cross_collide -s 2 -c 1 -o gpurun_out/r2a_cross_rpl2 -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2a_ncu_full2.log 2>&1
ls -la gpurun_out | tail -12
