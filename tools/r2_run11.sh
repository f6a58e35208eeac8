#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cross.py tests/test_gpu_cross_reduce.py tests/test_gpu_prm.py tests/test_gpu_strong.py tests/test_gpu_pp.py tests/test_gpu_smoother.py -m gpu -q > gpurun_out/r2l_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.txt
tail -4 gpurun_out/r2l_pytest.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2l_ncu_list.log 2>&1
for N in 2 8; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-extras > gpurun_out/r2l_bench_n$N.json 2> gpurun_out/r2l_bench_n$N.err; echo rc=$?
done
for f in gpurun_out/r2l_*.err; do echo "== $f"; tail -n 3 $f | cut -c1-300; done
