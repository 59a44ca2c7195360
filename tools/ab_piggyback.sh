#!/bin/bash
cd "$(dirname "$0")/.."
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-single > gpurun_out/r2i_plain.json 2> gpurun_out/r2i_plain.err && ncu --set full --clock-control none --import-source on -k regex:scan2_solve -s 3 -c 1 -o gpurun_out/prof_scan2_r02a python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-single > gpurun_out/r2i_ncu.log 2>&1
ls -la gpurun_out/prof_scan2_r02a.ncu-rep
