#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.txt
tail -5 gpurun_out/r2i_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; echo rc=$?
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2i_bench_reference.json 2> gpurun_out/r2i_bench_reference.err; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2i_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cross_collide -s 2 -c 1 -o gpurun_out/r2i_cross -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2i_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_prm_adjacency -s 2 -c 1 -o gpurun_out/r2i_adj -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2i_ncu_adj.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_arm_collide -s 1 -c 1 -o gpurun_out/r2i_arm -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_ncu_arm.log 2>&1
ls -la gpurun_out/r2i_*
for f in gpurun_out/r2i_*.err; do echo "== $f"; tail -n 4 $f | cut -c1-300; done
