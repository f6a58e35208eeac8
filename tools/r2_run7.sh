#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_strong.py tests/test_gpu_sssp.py tests/test_gpu_cross.py -m gpu -x -q > gpurun_out/r2g_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.txt
tail -8 gpurun_out/r2g_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; echo rc=$?
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-extras > gpurun_out/r2g_bench_n2_p2p.json 2> gpurun_out/r2g_bench_n2_p2p.err; echo rc=$?
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-extras --exchange nccl > gpurun_out/r2g_bench_n2_nccl.json 2> gpurun_out/r2g_bench_n2_nccl.err; echo rc=$?
for f in gpurun_out/r2g_*.err; do echo "== $f"; tail -n 6 $f | cut -c1-300; done
