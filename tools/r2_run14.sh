#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cross.py tests/test_gpu_strong.py tests/test_gpu_sssp.py -m gpu -q > gpurun_out/r2o_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.txt
tail -4 gpurun_out/r2o_pytest.txt | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.err; echo rc=$?
for N in 2 4 8; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2o_bench_n$N.json 2> gpurun_out/r2o_bench_n$N.err; echo rc=$?
done
for f in gpurun_out/r2o_*.err; do echo "== $f"; tail -n 3 $f | cut -c1-300; done
