# last GPU call of the round (tag r2zz): smoke() and the whole GPU suite on the committed state
python __graft_entry__.py --smoke 2>&1 | tail -2
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2zz_pytest.log; cat gpurun_out/r2zz_pytest.log
