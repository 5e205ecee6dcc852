# scratch GPU call (tag r3m): multi-GPU parity + C5 strong-scaling line on the GPUs of this box
N=${1:-2}
python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 600 -rs -v 2>&1 | tail -25 > gpurun_out/r3m_pytest_multi_gpu_${N}gpus.log; tail -8 gpurun_out/r3m_pytest_multi_gpu_${N}gpus.log
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --config C5 --steps 10 --warmup 3) > gpurun_out/r3m_scale_C5_n$N.log 2>&1
grep '^{' gpurun_out/r3m_scale_C5_n$N.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), d['config']['ms_per_timed_step'], d['parity'] and (d['parity']['ok'], d['parity']['relL2_theta']), d['comm']['peer_kernels_share_of_step'], 'e2e', round(d['e2e']['value'],1))
    print(d['roofline']['kernels_ms_per_step'])
"
