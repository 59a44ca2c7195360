#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_scan_solver_gpu.py tests/test_full_size_gpu.py -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_r02e.txt
for s2 in 1 0; do echo "IBS_SCAN2=$s2"; IBS_SCAN2=$s2 timeout 300 python tools/time_stages.py d3d 37 2>&1 | grep -E "solve|mean"; done 2>&1 | tee gpurun_out/ab_scan2_r02.txt
for s2 in 1 0; do echo "IBS_SCAN2=$s2 ncsx"; IBS_SCAN2=$s2 timeout 300 python tools/time_stages.py ncsx 1 2>&1 | grep -E "solve|mean"; done 2>&1 | tee -a gpurun_out/ab_scan2_r02.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_r02f.txt
