# scratch GPU call (tag r3t): the CUDA path against the reference's BMPLog numbers (fluidity equation from the reference text)
python -m pytest tests/test_bmp_log.py -m gpu -q --timeout 300 -k "fixture" 2>&1 | tail -12 | cut -c1-300
