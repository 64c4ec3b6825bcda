timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_timestep.py tests/test_adaptor.py 2>&1 | tail -15
