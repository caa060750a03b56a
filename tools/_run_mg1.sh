python tools/small_configs.py 2>&1 | head -3
PFEM_PRECOND=block python tools/small_configs.py 2>&1 | head -2
python -m pytest tests/test_gpu_mg.py -m gpu -x -q 2>&1 | tail -2
