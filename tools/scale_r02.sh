#!/bin/bash
# GPU box with 8 GPUs: weak and strong scaling of the default workload (final code of round 2)
cd "$(dirname "$0")/.."
run() { n=$1; shift; out=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" > gpurun_out/$out 2> gpurun_out/${out%.json}.err; tail -c 300 gpurun_out/$out | head -c 300; echo; }
python bench.py --steps 50 --no-cpu-baseline --no-e2e-full > gpurun_out/scale_r02b_weak_1.json 2> gpurun_out/scale_r02b_weak_1.err
for n in 2 4 8; do run $n scale_r02b_weak_$n.json --steps 50 --no-cpu-baseline --no-e2e-full --no-single; done
python bench.py --steps 50 --strong --equilibria 37 --no-cpu-baseline --no-e2e-full > gpurun_out/scale_r02b_strong_1.json 2> gpurun_out/scale_r02b_strong_1.err
for n in 2 4 8; do run $n scale_r02b_strong_$n.json --steps 50 --strong --equilibria 37 --no-cpu-baseline --no-e2e-full; done
run 8 scale_r02b_strong296_8.json --steps 50 --strong --equilibria 296 --no-cpu-baseline --no-e2e-full
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/scale_r02b_*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f.split('/')[-1], d['n_gpus'], '%.4g' % d['value'], 'ms/step %.3f' % d['ms_per_step'], d['scaling'], 'e2e %.4g' % d.get('e2e', {}).get('value', float('nan')))
    except Exception as e:
        print(f, 'ERR', e)
PY
