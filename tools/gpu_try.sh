#!/bin/bash
# one GPU call: parity tests, then bench lines of the configs given (tag = $1, configs = $2.., e.g. "C3 C2")
tag=$1; shift
cfgs="$*"
cmd="python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_pytest.log;"
for c in $cfgs; do
  cmd="$cmd RHEO_BENCH_VERBOSE=1 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err; tail -3 gpurun_out/${tag}_bench_$c.err;"
done
cmd="$cmd python - <<PYEOF
import json
for c in '$cfgs'.split():
    try:
        d=json.load(open('gpurun_out/${tag}_bench_'+c+'.json'))
        print(c, round(d['value'],1), 'Mcs/s', round(d['ms_per_step'],3), 'ms; step_frac', round(d['roofline']['step_frac'],3), 'iters', d['config']['krylov_iterations_mean'], 'e2e', round(d['e2e']['value'],1))
        print(d['roofline']['kernels_ms_per_step'])
        print(d['phase_ms'])
    except Exception as e:
        print(c, 'failed', e)
PYEOF"
/usr/local/graft/bin/gpurun --timeout 900 -- "$cmd" 2>&1 | grep -v "^\[gpurun\] sending\|merged"
