#!/bin/bash
# strong step as one library call per rank (mrmp_roadmaps_strong_step) against one call per stage
mkdir -p gpurun_out
TAG=r2u
run() { # N extra-args tag
  N=$1; shift; T=$1; shift
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N"
  timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-extras "$@" > gpurun_out/${TAG}_${T}.json 2> gpurun_out/${TAG}_${T}.err; echo rc=$?
}
run 8 nv500_n8
run 8 nv500_n8_separate --separate-calls
run 4 nv500_n4
run 2 nv500_n2
run 8 nv2000_n8 --num-vertices 2000 --steps 5
for f in gpurun_out/${TAG}_*.err; do echo "== $f"; tail -n 3 $f | cut -c1-300; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2u_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, j["ms_per_step"], j["value"], j["e2e"]["ms_per_step"], j["multi_gpu"]["check"])
    except Exception as e: print(f, "ERR", e)
PY
