#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_r02d.txt
python tools/time_stages.py d3d 37 2>&1 | tee gpurun_out/time_stages_r02d.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-single > gpurun_out/r2d_plain.json 2> gpurun_out/r2d_plain.err && ncu --set full --clock-control none --import-source on -k regex:scan_solve -s 3 -c 1 -o gpurun_out/prof_scan_r02b python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-single > gpurun_out/r2d_ncu.log 2>&1
ls -la gpurun_out/prof_scan_r02b.ncu-rep
