# last GPU call of the round (tag r2zzz): smoke() and the quick GPU tests on the rebuilt library
python __graft_entry__.py --smoke 2>&1 | tail -1
python -m pytest tests/test_grad_u.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
