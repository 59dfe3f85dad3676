timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_zz_dropin.py -m gpu -x -q 2>&1 | tail -6
