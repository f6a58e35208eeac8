#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest.txt
tail -3 gpurun_out/r2s_pytest.txt | cut -c1-200
timeout 300 python bench.py --steps 50 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2s_ncu_list.log 2>&1
tail -n 3 gpurun_out/r2s_bench.err | cut -c1-300
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2s_bench.json").read().strip().splitlines()[-1])
print(j["ms_per_step"], j["value"], j.get("e2e"))
PY
