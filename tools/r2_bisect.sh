#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests/test_gpu_strong.py tests/test_harness.py -m gpu -v -x > gpurun_out/r2j_bisect3.txt 2>&1
echo "rc=$?" >> gpurun_out/r2j_bisect3.txt
timeout 600 python -X faulthandler -m pytest "tests/test_gpu_strong.py::test_exchange_wait_gives_up_instead_of_hanging" "tests/test_gpu_strong.py::test_async_build_reports_capacity_overflow" tests/test_harness.py -m gpu -v -x > gpurun_out/r2j_bisect4.txt 2>&1
echo "rc=$?" >> gpurun_out/r2j_bisect4.txt
head -c 6000 gpurun_out/r2j_bisect3.txt; echo ......; tail -c 1500 gpurun_out/r2j_bisect4.txt
