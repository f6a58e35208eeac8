#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_strong.py -m gpu -x -q > gpurun_out/r2y_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.txt
tail -4 gpurun_out/r2y_pytest.txt | cut -c1-300
