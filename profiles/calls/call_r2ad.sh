timeout 1200 python -m pytest tests/test_gpu_api.py -m gpu -q --timeout 600 -x 2>&1 | tail -15
