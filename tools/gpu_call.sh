# GPU evidence call (tag r2f): (1) ncu launch list of the default bench command (C5, N = 1) with the DRAM byte counters of every
# launch in the same pass; (2) ncu --set full of one whole step of the half-scale C5 mesh (8 M cells, same kernels / colours)
KR='k_flux3|k_source_init|k_bsweep|k_bspmv0|k_spmv|k_update_x_r|k_eig_tau'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_C5.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2f_ncu_l_C5.log 2>&1
tail -2 gpurun_out/r2f_ncu_l_C5.log | cut -c1-300; wc -l gpurun_out/r2f_launches_C5.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$KR" -s 70 -c 30 -f -o gpurun_out/r2f_full_C5h \
  python bench.py --config C5 --scale 0.5 --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2f_ncu_f_C5h.log 2>&1
ncu -i gpurun_out/r2f_full_C5h.ncu-rep --page raw --csv > gpurun_out/r2f_full_C5h_raw.csv 2>/dev/null
ls -la gpurun_out/r2f_*
# (3) chunk kernels at 3 CTAs per SM (RK_BLK_MINB=3) against the default 2, half-scale C5
B="python bench.py --config C5 --scale 0.5 --steps 10 --warmup 3 --no-cpu-baseline --no-parity"
$B > gpurun_out/r2f_bench_C5h.json 2> gpurun_out/r2f_bench_C5h.err
RHEO_LIB_PATH=$PWD/build/variants/librheo_blk3.so $B > gpurun_out/r2f_bench_C5h_blk3.json 2> gpurun_out/r2f_bench_C5h_blk3.err
python - <<PYEOF
import json
for c in ('C5h','C5h_blk3'):
    try:
        d=json.load(open('gpurun_out/r2f_bench_'+c+'.json'))
        print(c, round(d['value'],1), 'Mcs/s', round(d['ms_per_step'],3), d['config']['ms_per_timed_step'][:4])
        print({k:v for k,v in d['roofline']['kernels_ms_per_step'].items() if 'bs' in k or 'spmv' in k})
    except Exception as e:
        print(c, 'failed', e)
PYEOF
