#!/bin/bash
# quick GPU check: parity tests + C2/C3 bench lines (tag = $1)
tag=${1:-x}
/usr/local/graft/bin/gpurun --timeout 900 -- "python -m pytest tests -m gpu -x -q 2>&1 | tail -15; python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_C2.json; python bench.py --config C3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_C3.json; python - <<PYEOF
import json
for c in ('C2','C3'):
    d=json.load(open('gpurun_out/${tag}_bench_'+c+'.json'))
    print(c, round(d['value'],1), 'Mcs/s', round(d['ms_per_step'],3), 'ms; step_frac', round(d['roofline']['step_frac'],3), 'iters', d['config']['krylov_iterations_mean'], 'e2e', round(d['e2e']['value'],1))
    print(d['roofline']['kernels_ms_per_step'])
    print(d['phase_ms'])
PYEOF" 2>&1 | grep -v "^\[gpurun\] sending\|merged"
