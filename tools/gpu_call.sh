# scratch GPU call (tag r3r): BMPLog through the other code paths
python -m pytest tests/test_bmp_log.py -m gpu -q --timeout 300 -k "other_code_paths" 2>&1 | tail -12 | cut -c1-300
