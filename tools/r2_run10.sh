#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -X faulthandler -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.txt
tail -6 gpurun_out/r2k_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.txt 2>&1; tail -2 gpurun_out/r2k_smoke.txt
