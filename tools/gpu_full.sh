#!/bin/bash
# GPU round-trip: parity tests, C2/C3 bench lines, ncu launch list + full-metric capture (tag = $1)
tag=${1:-x}
/usr/local/graft/bin/gpurun --timeout 1500 -- "python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_C2.json 2> gpurun_out/${tag}_bench_C2.err
python bench.py --config C3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_C3.json 2> gpurun_out/${tag}_bench_C3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_C2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_l_C2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_C3.csv python bench.py --config C3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_l_C3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_convect|k_grad_theta|k_cell_source|k_eig_tau|k_spmv|k_sweep|k_update|k_make_s|k_krylov_init' -s 60 -c 40 -f -o gpurun_out/${tag}_full_C3 python bench.py --config C3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_f_C3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_convect|k_grad_theta|k_cell_source|k_eig_tau|k_spmv|k_sweep|k_update|k_make_s|k_krylov_init' -s 60 -c 40 -f -o gpurun_out/${tag}_full_C2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_f_C2.log 2>&1
python - <<PYEOF
import json
for c in ('C2','C3'):
    d=json.load(open('gpurun_out/${tag}_bench_'+c+'.json'))
    print(c, round(d['value'],1), 'Mcs/s', round(d['ms_per_step'],3), 'ms; step_frac', round(d['roofline']['step_frac'],3), 'iters', d['config']['krylov_iterations_mean'], 'e2e', round(d['e2e']['value'],1))
    print(d['roofline']['kernels_ms_per_step'])
    print(d['phase_ms'])
PYEOF" 2>&1 | grep -v "^\[gpurun\] sending\|merged"
