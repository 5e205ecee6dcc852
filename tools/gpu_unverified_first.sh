#!/bin/bash
# First GPU call of the next round: the device paths written after the previous round's GPU budget was spent
# (PBiCG, SaramitoLog, CrankNicolson) and the reference-fixture GPU tests.  Usage:
#   gpurun --timeout 900 -- 'bash tools/gpu_unverified_first.sh'
# Every case is marked xfail(strict=False): read XPASS / XFAIL per test in gpurun_out/unverified.log.
mkdir -p gpurun_out
python -m pytest tests/test_zz_gpu_not_yet_run.py -m gpu -q -rxX --timeout 600 2>&1 | tee gpurun_out/unverified.log
compute-sanitizer --tool memcheck python -m pytest tests/test_zz_gpu_not_yet_run.py -m gpu -q -k "pbicg and (C1 or fixture)" --timeout 900 > gpurun_out/unverified_memcheck.log 2>&1
tail -5 gpurun_out/unverified_memcheck.log
