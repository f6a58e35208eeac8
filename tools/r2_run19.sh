#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_strong.py -m gpu -x -q > gpurun_out/r2t_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.txt
tail -15 gpurun_out/r2t_pytest.txt | cut -c1-300
