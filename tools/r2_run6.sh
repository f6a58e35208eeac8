#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_strong.py tests/test_gpu_cross.py tests/test_gpu_cross_reduce.py tests/test_gpu_prm.py tests/test_gpu_pp.py -m gpu -x -q > gpurun_out/r2f_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.txt
tail -15 gpurun_out/r2f_pytest.txt
python tools/quick_shard.py > gpurun_out/r2f_shard_auto.json 2> gpurun_out/r2f_shard.err
MRMP_CROSS_SPLIT=1 python tools/quick_shard.py > gpurun_out/r2f_shard_s1.json 2>> gpurun_out/r2f_shard.err
MRMP_CROSS_SPLIT=4 python tools/quick_shard.py > gpurun_out/r2f_shard_s4.json 2>> gpurun_out/r2f_shard.err
MRMP_CROSS_SPLIT=8 python tools/quick_shard.py > gpurun_out/r2f_shard_s8.json 2>> gpurun_out/r2f_shard.err
cat gpurun_out/r2f_shard_*.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo rc=$?
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-extras > gpurun_out/r2f_bench_n2_p2p.json 2> gpurun_out/r2f_bench_n2_p2p.err; echo rc=$?
for f in gpurun_out/r2f_*.err; do echo "== $f"; tail -n 6 $f; done
