#!/bin/bash
# GPU box: the round's final bench lines for every workload, the reference arm, and the launch list of the default bench
cd "$(dirname "$0")/.."
python -m pytest tests -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_r02_d3d.json 2> gpurun_out/bench_r02_d3d.err; python tools/bench_brief.py gpurun_out/bench_r02_d3d.json d3d
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference_arm.json 2> gpurun_out/bench_r02_reference_arm.err; tail -c 600 gpurun_out/bench_r02_reference_arm.json
for w in ncsx hberg salpha; do python bench.py --workload $w --steps 30 > gpurun_out/bench_r02_$w.json 2> gpurun_out/bench_r02_$w.err; python tools/bench_brief.py gpurun_out/bench_r02_$w.json $w; done
python bench.py --workload adjoint --points 32768 --steps 5 > gpurun_out/bench_r02_adjoint_32k.json 2> gpurun_out/bench_r02_adjoint_32k.err; python tools/bench_brief.py gpurun_out/bench_r02_adjoint_32k.json adjoint
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-single > gpurun_out/r2q_plain.json 2> gpurun_out/r2q_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_d3d.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-single > gpurun_out/r2q_ncu.log 2>&1
tail -3 gpurun_out/launches_r02_d3d.csv
ncu --set full --clock-control none --import-source on -k regex:scan2_solve -s 3 -c 1 -o gpurun_out/prof_scan2_r02c python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-single > gpurun_out/r2q_ncu2.log 2>&1
ls -la gpurun_out/prof_scan2_r02c.ncu-rep
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
