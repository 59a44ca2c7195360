#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_r02c.txt
for v in u2; do for wl in ncsx hberg; do echo "variant $v $wl"; IBS_LIB=$PWD/ideal-ballooning-solver_b200/lib/variants/libibs_$v.so python tools/time_stages.py $wl 1 2>&1 | grep -E "geometry"; done; done 2>&1 | tee gpurun_out/ab_geo_r02c.txt
python bench.py --steps 50 > gpurun_out/bench_r02_d3d.json 2> gpurun_out/bench_r02_d3d.err
IBS_GEO3D=0 python bench.py --workload ncsx --steps 30 > gpurun_out/bench_r02_ncsx.json 2> gpurun_out/bench_r02_ncsx.err
IBS_GEO3D=0 python bench.py --workload hberg --steps 10 > gpurun_out/bench_r02_hberg.json 2> gpurun_out/bench_r02_hberg.err
python bench.py --workload salpha --steps 30 > gpurun_out/bench_r02_salpha.json 2> gpurun_out/bench_r02_salpha.err
IBS_GEO3D=0 python bench.py --workload adjoint --steps 5 --points 32768 > gpurun_out/bench_r02_adjoint32k.json 2> gpurun_out/bench_r02_adjoint32k.err
tail -2 gpurun_out/bench_r02_*.err
