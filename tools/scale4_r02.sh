#!/bin/bash
# GPU box with 4 GPUs: weak scaling 1/2/4 of the default workload with the final code (asynchronous exchange, power-state ramp)
cd "$(dirname "$0")/.."
run() { n=$1; shift; out=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n "$@" > gpurun_out/$out 2> gpurun_out/${out%.json}.err; python tools/bench_brief.py gpurun_out/$out $out | cut -c1-220; }
python bench.py --steps 100 --no-cpu-baseline --no-e2e-full --no-single > gpurun_out/scale_r02c_weak_1.json 2> gpurun_out/scale_r02c_weak_1.err; python tools/bench_brief.py gpurun_out/scale_r02c_weak_1.json weak1 | cut -c1-220
for n in 2 4; do run $n scale_r02c_weak_$n.json --steps 100 --no-cpu-baseline --no-e2e-full --no-single; done
run 4 scale_r02c_strong148_4.json --steps 100 --strong --equilibria 148 --no-cpu-baseline --no-e2e-full
