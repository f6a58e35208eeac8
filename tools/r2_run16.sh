#!/bin/bash
# final single-GPU evidence of round 2: full GPU suite, smoke, the default bench line, the reference arm,
# launch list and full ncu captures (table kernel, row scatter)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.txt
tail -3 gpurun_out/r2r_pytest.txt | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2r_smoke.txt 2>&1; echo rc=$?
timeout 900 python bench.py > gpurun_out/r2r_bench_default.json 2> gpurun_out/r2r_bench_default.err; echo rc=$?
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2r_bench_reference.json 2> gpurun_out/r2r_bench_reference.err; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2r_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2r_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cross_collide -s 3 -c 1 -o gpurun_out/r2r_cross -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2r_ncu_cross.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_csr_rows_scatter -s 3 -c 1 -o gpurun_out/r2r_scatter -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2r_ncu_scatter.log 2>&1
for f in gpurun_out/r2r_*.err; do echo "== $f"; tail -n 3 $f | cut -c1-300; done
tail -2 gpurun_out/r2r_smoke.txt
ls -la gpurun_out/r2r_*
