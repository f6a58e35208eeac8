#!/bin/bash
# round-2 GPU run 2 (2 GPUs): new tests, N=1 bench (FD kernel, MINB=4 variant), strong scaling at N=2
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2b_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_strong.py tests/test_gpu_cross.py tests/test_gpu_cross_reduce.py -m gpu -x -q > gpurun_out/r2b_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.txt
tail -15 gpurun_out/r2b_pytest.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo rc=$?
MRMP_B200_LIB=$PWD/sssp_b200/_lib/libmrmp_b200_minb4.so timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2b_bench_n1_minb4.json 2> gpurun_out/r2b_bench_n1_minb4.err; echo rc=$?
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-extras > gpurun_out/r2b_bench_n2_p2p.json 2> gpurun_out/r2b_bench_n2_p2p.err; echo rc=$?
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-extras --exchange nccl > gpurun_out/r2b_bench_n2_nccl.json 2> gpurun_out/r2b_bench_n2_nccl.err; echo rc=$?
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-extras --scaling weak > gpurun_out/r2b_bench_n2_weak.json 2> gpurun_out/r2b_bench_n2_weak.err; echo rc=$?
timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --num-vertices 2000 > gpurun_out/r2b_bench_nv2000_n1.json 2> gpurun_out/r2b_bench_nv2000_n1.err; echo rc=$?
timeout 300 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-extras --num-vertices 2000 > gpurun_out/r2b_bench_nv2000_n2.json 2> gpurun_out/r2b_bench_nv2000_n2.err; echo rc=$?
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2b_ref_n2.json 2> gpurun_out/r2b_ref_n2.err; echo rc=$?
tail -3 gpurun_out/*.err | tail -60
