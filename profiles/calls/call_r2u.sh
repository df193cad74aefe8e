timeout 600 python profiles/tcw_ab.py 2>&1 | tail -7
timeout 900 python -m pytest tests/test_gpu_parity_full_size.py tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "wide or c5 or sharded" 2>&1 | tail -4
