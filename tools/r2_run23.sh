#!/bin/bash
# after putting the ctx and its consumers on ONE stream (default stream handle 0 -> cudaStreamLegacy): tests, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.txt
tail -5 gpurun_out/r2x_pytest.txt | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2x_smoke.txt 2>&1; echo rc=$?
timeout 900 python bench.py > gpurun_out/r2x_bench_default.json 2> gpurun_out/r2x_bench_default.err; echo rc=$?
MRMP_TABLE_OVERLAP=0 timeout 300 python bench.py --steps 50 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2x_bench_serial.json 2> gpurun_out/r2x_bench_serial.err; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2x_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2x_ncu_list.log 2>&1
tail -n 3 gpurun_out/r2x_bench_default.err | cut -c1-300
tail -2 gpurun_out/r2x_smoke.txt
python - <<'PY'
import json
for t in ("default","serial"):
    try:
        j=json.loads(open(f"gpurun_out/r2x_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, j["ms_per_step"], j["value"], j["e2e"]["ms_per_step"], j["roofline"]["kernel_ms"])
    except Exception as e: print(t, "ERR", e)
PY
