# scratch GPU call (tag r3q): BMPLog on two ranks + the 1-GPU BMPLog / divTau tests with the corrected oracle
python -m pytest tests/test_multi_gpu.py tests/test_bmp_log.py tests/test_grad_u.py -m gpu -q --timeout 300 -k "bmp or grad_u or gradient" -rs 2>&1 | tail -12
