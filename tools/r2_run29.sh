#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642"
timeout 150 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2ac_n2.json 2> gpurun_out/r2ac_n2.err; echo rc=$?
tail -n 4 gpurun_out/r2ac_n2.err | cut -c1-300
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2ac_n2.json").read().strip().splitlines()[-1])
print(j["ms_per_step"], j["value"], j["e2e"]["ms_per_step"], j["multi_gpu"]["check"])
PY
