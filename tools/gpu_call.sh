# scratch GPU call (tag r3s): the CUDA path against the reference's BMPLog numbers
python -m pytest tests/test_bmp_log.py -m gpu -q --timeout 300 2>&1 | tail -12 | cut -c1-300
