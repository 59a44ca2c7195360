#!/bin/bash
# GPU box: does the solver execute more instructions, or the same instructions more slowly, under the two K1 epilogues?
cd "$(dirname "$0")/.."
V=$PWD/ideal-ballooning-solver_b200/lib/variants
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,sm__warps_active.avg.per_cycle_active
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-single > gpurun_out/r2s_plain.json 2> gpurun_out/r2s_plain.err || exit 1
IBS_BENCH_RAMP_S=0 ncu --metrics $M --clock-control none -k regex:scan2_solve -s 3 -c 2 --csv --log-file gpurun_out/r2s_default.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-single > /dev/null 2>&1
IBS_LIB=$V/libibs_slowepi.so IBS_BENCH_RAMP_S=0 ncu --metrics $M --clock-control none -k regex:scan2_solve -s 3 -c 2 --csv --log-file gpurun_out/r2s_slowepi.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-single > /dev/null 2>&1
python - <<'PY'
import csv
for n in ("default", "slowepi"):
    rows = [r for r in csv.reader(open(f"gpurun_out/r2s_{n}.csv")) if len(r) > 10 and r[0].isdigit()]
    for r in rows:
        print(n, r[0], r[-3], r[-1])
PY
