#!/bin/bash
# (round 1's evidence script, kernel list updated; round 2's evidence call is tools/gpu_call.sh tag r2f, see profiles/README.md)
# evidence for profiles/: bench lines (C2 with cpu_baseline, C3), ncu launch lists of the same commands, one ncu --set full
# capture per config of one launch of every hot kernel (tag = $1).  Keeps gpurun_out small (raw CSV pages, no big reps).
tag=${1:-x}
KR='k_flux3|k_source_init|k_flux_assemble|k_cell_source2|k_eig_tau|k_krylov_init|k_bsweep|k_bspmv0|k_update_p|k_sweep|k_spmv|k_make_s|k_update_x_r'
/usr/local/graft/bin/gpurun --timeout 1500 -- "python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_C2.json 2> gpurun_out/${tag}_bench_C2.err
python bench.py --config C3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_C3.json 2> gpurun_out/${tag}_bench_C3.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_C2_reference.json 2> gpurun_out/${tag}_bench_C2_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_C2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_l_C2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_C3.csv python bench.py --config C3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_l_C3.log 2>&1
ncu --set full --clock-control none -k regex:'${KR}' -s 40 -c 24 -f -o gpurun_out/${tag}_full_C3 python bench.py --config C3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_f_C3.log 2>&1
ncu -i gpurun_out/${tag}_full_C3.ncu-rep --page raw --csv > gpurun_out/${tag}_full_C3_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:'${KR}' -s 40 -c 24 -f -o gpurun_out/${tag}_full_C2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_f_C2.log 2>&1
ncu -i gpurun_out/${tag}_full_C2.ncu-rep --page raw --csv > gpurun_out/${tag}_full_C2_raw.csv 2>/dev/null
find gpurun_out -name '*.ncu-rep' -size +12M -delete
ls -la gpurun_out | head -40
python - <<PYEOF
import json
for c in ('C2','C3'):
    d=json.load(open('gpurun_out/${tag}_bench_'+c+'.json'))
    print(c, round(d['value'],1), 'Mcs/s', round(d['ms_per_step'],3), 'ms; step_frac', round(d['roofline']['step_frac'],3), 'roof', d['roofline']['kernel'], round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
PYEOF
tail -2 gpurun_out/${tag}_bench_C2_reference.json" 2>&1 | grep -v "^\[gpurun\] sending\|merged"
