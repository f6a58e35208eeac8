#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -X faulthandler -m pytest tests -m gpu -q > gpurun_out/r2m_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.txt
tail -5 gpurun_out/r2m_pytest.txt | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2m_ncu_list.log 2>&1
for f in gpurun_out/r2m_*.err; do echo "== $f"; tail -n 3 $f | cut -c1-300; done
