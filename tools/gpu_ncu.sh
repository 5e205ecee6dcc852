#!/bin/bash
# targeted ncu capture on one config: tools/gpu_ncu.sh <tag> <config> <kernel-regex> [count] [extra bench args]
# keeps the .ncu-rep only if small; always leaves the raw CSV page in gpurun_out/
tag=$1; cfg=$2; rx=$3; cnt=${4:-8}
/usr/local/graft/bin/gpurun --timeout 900 -- "ncu --set full --clock-control none --import-source on -k regex:'${rx}' -s 1 -c ${cnt} -f -o gpurun_out/${tag} python bench.py --config ${cfg} --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv --kernel-name regex:k_flux > gpurun_out/${tag}_src_convect.csv 2>/dev/null
ls -la gpurun_out/
find gpurun_out -name '*.ncu-rep' -size +24M -delete
tail -3 gpurun_out/${tag}.log" 2>&1 | grep -v "^\[gpurun\] sending\|merged"
