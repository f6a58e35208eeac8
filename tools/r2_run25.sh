#!/bin/bash
# strong scaling after the stream fix; usage: r2_run25.sh "<N list>"
mkdir -p gpurun_out
TAG=r2z
for N in $1; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2963$N"
  timeout 200 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-extras > gpurun_out/${TAG}_nv500_n$N.json 2> gpurun_out/${TAG}_nv500_n$N.err; echo rc=$?
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2z_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, j["ms_per_step"], j["value"], j["e2e"]["ms_per_step"], j["multi_gpu"]["check"])
    except Exception as e: print(f, "ERR", e)
PY
