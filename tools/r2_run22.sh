#!/bin/bash
mkdir -p gpurun_out
TAG=r2w
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29623"
BENCH_E2E_TRACE=1 timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-extras > gpurun_out/${TAG}_one.json 2> gpurun_out/${TAG}_one.err; echo rc=$?
BENCH_E2E_TRACE=1 timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-extras --separate-calls > gpurun_out/${TAG}_separate.json 2> gpurun_out/${TAG}_separate.err; echo rc=$?
grep "e2e trace" gpurun_out/${TAG}_*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2w_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, j["ms_per_step"], j["e2e"]["ms_per_step"], j["multi_gpu"]["check"])
    except Exception as e: print(f, "ERR", e)
PY
