#!/bin/bash
# GPU box: full stall breakdown of the solver kernel on the original and on the bit-truncated base arrays
cd "$(dirname "$0")/.."
python tools/ds_one.py > /dev/null 2>&1 || exit 1
ncu --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:scan2_solve -s 2 -c 1 --csv --page raw --log-file gpurun_out/ds_orig.csv python tools/ds_one.py > /dev/null 2>&1
DS_TRUNC=1 ncu --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:scan2_solve -s 2 -c 1 --csv --page raw --log-file gpurun_out/ds_trunc.csv python tools/ds_one.py > /dev/null 2>&1
python - <<'PY'
import csv
def load(n):
    rows = list(csv.reader(open(f"gpurun_out/{n}.csv")))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    return dict(zip(rows[hdr], rows[hdr + 2]))
a, b = load("ds_orig"), load("ds_trunc")
for k in a:
    try:
        x, y = float(a[k].replace(",", "")), float(b[k].replace(",", ""))
    except Exception:
        continue
    if x != 0 and abs(y - x) / abs(x) > 0.03 or "duration" in k or "inst_executed.sum" in k:
        print(f"{k:90s} {x:14.4f} {y:14.4f}  {100 * (y - x) / x if x else 0:+.1f}%")
PY
