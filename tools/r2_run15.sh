#!/bin/bash
# 8-lane CSR row sort + column tiles prepared on a second stream: tests, then the step with and without the overlap
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.txt
tail -4 gpurun_out/r2q_pytest.txt | cut -c1-200
timeout 300 python bench.py --steps 50 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2q_bench_overlap.json 2> gpurun_out/r2q_bench_overlap.err; echo rc=$?
MRMP_TABLE_OVERLAP=0 timeout 300 python bench.py --steps 50 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2q_bench_serial.json 2> gpurun_out/r2q_bench_serial.err; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2q_ncu_list.log 2>&1
for f in gpurun_out/r2q_*.err; do echo "== $f"; tail -n 3 $f | cut -c1-300; done
python - <<'PY'
import json
for t in ("overlap","serial"):
    try:
        j=json.loads(open(f"gpurun_out/r2q_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, j["ms_per_step"], j["value"], j.get("e2e"))
    except Exception as e: print(t, "ERR", e)
PY
