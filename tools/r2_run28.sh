#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/r2ab_bench_default.json 2> gpurun_out/r2ab_bench_default.err; echo rc=$?
tail -n 5 gpurun_out/r2ab_bench_default.err | cut -c1-300
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2ab_bench_default.json").read().strip().splitlines()[-1])
print(j["ms_per_step"], j["value"], j["e2e"])
PY
