#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cross.py tests/test_gpu_cross_reduce.py tests/test_gpu_strong.py tests/test_gpu_pp.py tests/test_gpu_smoother.py tests/test_gpu_prm.py -m gpu -q > gpurun_out/r2n_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.txt
tail -4 gpurun_out/r2n_pytest.txt | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2n_bench_base.json 2> gpurun_out/r2n_bench_base.err
for v in m4u8 m4u32 m3u8; do
MRMP_B200_LIB=$PWD/sssp_b200/_lib/libmrmp_b200_$v.so timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2n_bench_$v.json 2> gpurun_out/r2n_bench_$v.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2n_ncu_list.log 2>&1
for f in gpurun_out/r2n_*.err; do echo "== $f"; tail -n 3 $f | cut -c1-300; done
