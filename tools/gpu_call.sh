# scratch GPU call of the current experiment (tag r3d): parity tests, then A/B bench lines
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r3d_pytest.log; cat gpurun_out/r3d_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
$B --config C3 > gpurun_out/r3d_bench_C3.json 2> gpurun_out/r3d_bench_C3.err
$B --config C2 > gpurun_out/r3d_bench_C2.json 2> gpurun_out/r3d_bench_C2.err
RHEO_TILE_ORDER=0 $B --config C3 --no-parity > gpurun_out/r3d_bench_C3noord.json 2> gpurun_out/r3d_bench_C3noord.err
RHEO_LIB_PATH=$PWD/build/variants/librheo_d1.so $B --config C3 --no-parity > gpurun_out/r3d_bench_C3d1.json 2> gpurun_out/r3d_bench_C3d1.err
$B --config C5 --scale 0.5 --no-parity > gpurun_out/r3d_bench_C5h.json 2> gpurun_out/r3d_bench_C5h.err
RHEO_FLUX=1 $B --config C5 --scale 0.5 --no-parity > gpurun_out/r3d_bench_C5hv1.json 2> gpurun_out/r3d_bench_C5hv1.err
python - <<PYEOF
import json
for c in ('C3','C2','C3noord','C3d1','C5h','C5hv1'):
    try:
        d=json.load(open('gpurun_out/r3d_bench_'+c+'.json'))
        print(c, round(d['value'],1), 'Mcs/s', round(d['ms_per_step'],3), 'ms; step_frac', round(d['roofline']['step_frac'],3), 'iters', d['config']['krylov_iterations_mean'], 'e2e', round(d['e2e']['value'],1), 'parity', d['parity'] and d['parity']['relL2_theta'])
        print(d['roofline']['kernels_ms_per_step'])
    except Exception as e:
        print(c, 'failed', e)
PYEOF
