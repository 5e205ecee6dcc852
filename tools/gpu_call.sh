# GPU evidence call (tag r2f): (1) ncu launch list of the default bench command (C5, N = 1) with the DRAM byte counters of every
# launch in the same pass; (2) ncu --set full of one whole step of the half-scale C5 mesh (8 M cells, same kernels / colours).
# Only CSV pages come back (gpurun_out is limited to 64 MiB): the .ncu-rep stays in /tmp on the box.
KR='k_flux3|k_source_init|k_bsweep|k_bspmv0|k_spmv|k_update_x_r|k_eig_tau'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_C5.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2f_ncu_l_C5.log 2>&1
tail -2 gpurun_out/r2f_ncu_l_C5.log | cut -c1-300; wc -l gpurun_out/r2f_launches_C5.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$KR" -s 42 -c 14 -f -o /tmp/r2f_full_C5h \
  python bench.py --config C5 --scale 0.5 --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2f_ncu_f_C5h.log 2>&1
ncu -i /tmp/r2f_full_C5h.ncu-rep --page raw --csv > gpurun_out/r2f_full_C5h_raw.csv 2>/dev/null
ncu -i /tmp/r2f_full_C5h.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:k_flux3 > /tmp/r2f_src_flux3.csv 2>/dev/null
python tools/ncu_lines.py /tmp/r2f_full_C5h.ncu-rep k_flux3 40 > gpurun_out/r2f_flux3_lines.txt 2>&1
ls -la gpurun_out/ /tmp/r2f_full_C5h.ncu-rep; du -sh gpurun_out
