# final verification call (tag r2z): the whole GPU suite on one B200, then the default bench line and its reference arm
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2z_pytest.log; cat gpurun_out/r2z_pytest.log
(time python bench.py) > gpurun_out/r2z_bench_C5.json 2> gpurun_out/r2z_bench_C5.err; tail -4 gpurun_out/r2z_bench_C5.err
python - <<PYEOF
import json
d=json.load(open('gpurun_out/r2z_bench_C5.json'))
print('C5', round(d['value'],1), 'Mcs/s', round(d['ms_per_step'],3), 'ms; step_frac', round(d['roofline']['step_frac'],3), 'frac', round(d['roofline']['frac'],3), 'traffic', d['roofline']['traffic'], 'e2e', round(d['e2e']['value'],1), 'e2e_tau', round(d['e2e_tau_download']['value'],1), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],2), 'parity', d['parity'] and d['parity']['ok'], 'launches', d['gpu_launches'], 'steps', d['steps'], d['warmup'])
print(d['config']['krylov_iterations_per_step']); print(d['roofline']['kernels_ms_per_step'])
PYEOF
