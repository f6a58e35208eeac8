#!/bin/bash
# roadmaps read back beside the table kernel (read_async behind build_async): tests, e2e with and without
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2aa_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2aa_pytest.txt
tail -3 gpurun_out/r2aa_pytest.txt | cut -c1-300
timeout 200 python bench.py --steps 50 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2aa_bench_early.json 2> gpurun_out/r2aa_bench_early.err; echo rc=$?
BENCH_E2E_LATE_READ=1 timeout 200 python bench.py --steps 50 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2aa_bench_late.json 2> gpurun_out/r2aa_bench_late.err; echo rc=$?
tail -n 3 gpurun_out/r2aa_bench_early.err | cut -c1-300
python - <<'PY'
import json
for t in ("early","late"):
    try:
        j=json.loads(open(f"gpurun_out/r2aa_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, j["ms_per_step"], j["value"], j["e2e"]["ms_per_step"], j["e2e"]["note"])
    except Exception as e: print(t, "ERR", e)
PY
