python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/solve_n.py 69 mg block 2>&1
python bench.py --steps 3 --warmup 1 2>/dev/null | tail -1
