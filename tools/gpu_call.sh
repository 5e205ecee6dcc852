# scratch GPU call (tag r3l): whole GPU suite with the PDL policy + BMPLog, C5 N = 1 default bench line
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r3l_pytest.log; cat gpurun_out/r3l_pytest.log
(time python bench.py --steps 10 --warmup 3) > gpurun_out/r3l_bench_C5.json 2> gpurun_out/r3l_bench_C5.err; tail -4 gpurun_out/r3l_bench_C5.err
python - <<PYEOF
import json
d=json.load(open('gpurun_out/r3l_bench_C5.json'))
print('C5', round(d['value'],1), 'Mcs/s', round(d['ms_per_step'],3), 'ms; step_frac', round(d['roofline']['step_frac'],3), d['config']['ms_per_timed_step'], 'e2e', d['e2e']['value'], 'e2e_tau', d['e2e_tau_download']['value'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], 'parity', d['parity'] and d['parity']['ok'])
print(d['roofline']['kernels_ms_per_step']); print(d['phase_ms'])
PYEOF
