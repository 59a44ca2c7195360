#!/bin/bash
# GPU box: sampler diagnostics, then one ncu --set full capture of the solver kernel and the launch list of the default bench
cd "$(dirname "$0")/.."
IBS_BENCH_NO_SAMPLER=1 python bench.py --steps 100 --no-cpu-baseline --no-e2e-full --no-single 2>/dev/null | python tools/bench_brief.py - nosampler
python bench.py --steps 100 --no-cpu-baseline --no-e2e-full --no-single 2>/dev/null | python tools/bench_brief.py - sampler
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-single > gpurun_out/r2p_plain.json 2> gpurun_out/r2p_plain.err && ncu --set full --clock-control none --import-source on -k regex:scan2_solve -s 3 -c 1 -o gpurun_out/prof_scan2_r02b python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-single > gpurun_out/r2p_ncu.log 2>&1
ls -la gpurun_out/prof_scan2_r02b.ncu-rep
