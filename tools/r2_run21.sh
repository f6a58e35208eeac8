#!/bin/bash
# N=2: where does the end-to-end time of the one-call strong step go (adoption on the second stream or not)
mkdir -p gpurun_out
TAG=r2v
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-extras > gpurun_out/${TAG}_aux.json 2> gpurun_out/${TAG}_aux.err; echo rc=$?
MRMP_STRONG_ADOPT_AUX=0 timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-extras > gpurun_out/${TAG}_main.json 2> gpurun_out/${TAG}_main.err; echo rc=$?
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-extras --separate-calls > gpurun_out/${TAG}_separate.json 2> gpurun_out/${TAG}_separate.err; echo rc=$?
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-extras > gpurun_out/${TAG}_aux2.json 2> gpurun_out/${TAG}_aux2.err; echo rc=$?
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2v_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, j["ms_per_step"], j["e2e"]["ms_per_step"], j["multi_gpu"]["check"])
    except Exception as e: print(f, "ERR", e)
PY
