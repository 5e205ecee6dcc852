#!/bin/bash
# C2/C3 bench lines only (no tests) — for A/B runs of a kernel variant (tag = $1)
tag=${1:-x}
/usr/local/graft/bin/gpurun --timeout 600 -- "python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_C2.json; python bench.py --config C3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_C3.json; python - <<PYEOF
import json
for c in ('C2','C3'):
    d=json.load(open('gpurun_out/${tag}_bench_'+c+'.json'))
    k=d['roofline']['kernels_ms_per_step']
    print(c, round(d['value'],1), 'Mcs/s', round(d['ms_per_step'],3), 'ms; step_frac', round(d['roofline']['step_frac'],3), 'flux', [v for n,v in k.items() if 'flux' in n], 'src', k.get('k_cell_source2'))
PYEOF" 2>&1 | grep -v "^\[gpurun\] sending\|merged\|^import\|^for c\|^    \|^PYEOF"
