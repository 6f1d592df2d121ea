timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_bench_contract.py -x -q -m gpu 2>&1 | tail -3
