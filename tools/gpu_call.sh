# scratch GPU call (tag r3f): k_flux3 occupancy variants
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity"
$B --config C3 > gpurun_out/r3f_bench_C3.json 2> gpurun_out/r3f_bench_C3.err
for v in m16d1 m16d2; do RHEO_LIB_PATH=$PWD/build/variants/librheo_$v.so $B --config C3 > gpurun_out/r3f_bench_C3$v.json 2> gpurun_out/r3f_bench_C3$v.err; done
RHEO_LIB_PATH=$PWD/build/variants/librheo_m16d1.so $B --config C2 > gpurun_out/r3f_bench_C2m16d1.json 2> gpurun_out/r3f_bench_C2m16d1.err
python - <<PYEOF
import json
for c in ('C3','C3m16d1','C3m16d2','C2m16d1'):
    try:
        d=json.load(open('gpurun_out/r3f_bench_'+c+'.json'))
        print(c, round(d['value'],1), 'Mcs/s', round(d['ms_per_step'],3), 'ms; step_frac', round(d['roofline']['step_frac'],3), 'iters', d['config']['krylov_iterations_mean'], 'e2e', round(d['e2e']['value'],1))
        print({k:v for k,v in d['roofline']['kernels_ms_per_step'].items() if 'flux' in k or 'source' in k})
    except Exception as e:
        print(c, 'failed', e)
PYEOF
