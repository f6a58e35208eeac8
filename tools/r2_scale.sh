#!/bin/bash
# strong-scaling lines at N = 1, 2, 4, 8 on one box: the BASELINE instance (500 vertices) and a 16x larger one
set -x
mkdir -p gpurun_out
TAG=${1:-r2h}
for NV in 500 2000; do
  ST=20; [ $NV = 2000 ] && ST=5
  timeout 300 python bench.py --steps $ST --warmup 3 --no-extras --no-cpu-baseline --num-vertices $NV > gpurun_out/${TAG}_nv${NV}_n1.json 2> gpurun_out/${TAG}_nv${NV}_n1.err; echo rc=$?
  for N in 2 4 8; do
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
    timeout 300 $TR bench.py --gpus $N --steps $ST --warmup 3 --no-extras --num-vertices $NV > gpurun_out/${TAG}_nv${NV}_n$N.json 2> gpurun_out/${TAG}_nv${NV}_n$N.err; echo rc=$?
  done
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528"
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 3 --no-extras --exchange nccl > gpurun_out/${TAG}_nv500_n8_nccl.json 2> gpurun_out/${TAG}_nv500_n8_nccl.err; echo rc=$?
timeout 300 $TR bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/${TAG}_ref_n8.json 2> gpurun_out/${TAG}_ref_n8.err; echo rc=$?
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
for f in gpurun_out/${TAG}_*.err; do echo "== $f"; tail -n 3 $f | cut -c1-300; done
