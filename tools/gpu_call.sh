# scratch GPU call (tag r3p): multi-rank PBiCG parity (2 GPUs) + the other multi-GPU cases + 1-GPU PBiCG tests
python -m pytest tests/test_multi_gpu.py tests/test_gpu_pbicg.py -m gpu -q --timeout 600 -rs -v 2>&1 | tail -30 > gpurun_out/r3p_pytest_multi_gpu_2gpus.log; tail -22 gpurun_out/r3p_pytest_multi_gpu_2gpus.log
