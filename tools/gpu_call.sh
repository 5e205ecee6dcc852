# scratch GPU call (tag r3g): div_tau parity + bench lines with the new e2e
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
$B --config C3 > gpurun_out/r3g_bench_C3.json 2> gpurun_out/r3g_bench_C3.err; tail -3 gpurun_out/r3g_bench_C3.err
$B --config C2 > gpurun_out/r3g_bench_C2.json 2> gpurun_out/r3g_bench_C2.err
python - <<PYEOF
import json
for c in ('C3','C2'):
    try:
        d=json.load(open('gpurun_out/r3g_bench_'+c+'.json'))
        print(c, round(d['value'],1), 'Mcs/s', round(d['ms_per_step'],3), 'ms; step_frac', round(d['roofline']['step_frac'],3), 'e2e', d['e2e'], 'e2e_tau', d['e2e_tau_download']['value'])
        print({k:v for k,v in d['roofline']['kernels_ms_per_step'].items()})
    except Exception as e:
        print(c, 'failed', e)
PYEOF
