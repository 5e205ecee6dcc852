#!/bin/bash
# Multi-GPU evidence in one gpurun call (N GPUs of one box):  gpurun --gpus 8 --timeout 2400 -- 'bash tools/gpu_scaling_run.sh 8'
#   1. tests/test_multi_gpu.py on every rank count the box allows (parity of the decomposed GPU run with the 1-rank oracle);
#   2. bench.py (C5 64 M cells, strong scaling) on N, N/2, ... GPUs, then C4 and C3 on N.
# Logs land in gpurun_out/ (copy the pytest log and the JSON lines to profiles/).
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 900 -rs -v 2>&1 | tail -25 > gpurun_out/r2_pytest_multi_gpu_${N}gpus.log
port=29600
run() {   # n config extra...
  n=$1; cfg=$2; shift 2
  port=$((port + 1))
  (time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --config $cfg "$@") > gpurun_out/r2_scale_${cfg}_n$n.log 2>&1
}
n=$N
while [ $n -ge 2 ]; do run $n C5 --steps 20 --warmup 5; n=$((n / 2)); done
run $N C4 --steps 10 --warmup 3
run $N C3 --steps 20 --warmup 5
