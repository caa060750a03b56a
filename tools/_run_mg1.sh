python tools/solve_n.py 69 mg 2>&1 | grep -v "MG smoother"
PFEM_MG_NUC=2 python tools/solve_n.py 69 mg:1 2>&1 | grep -v "^n=" | head -1
PFEM_MG_NUC=3 python tools/solve_n.py 69 mg:1 2>&1 | grep -v "^n=" | head -1
PFEM_MG_NUC=3 python tools/solve_n.py 69 mg:2 2>&1 | grep -v "^n=" | head -1
PFEM_MG_NUC=4 python tools/solve_n.py 69 mg:1 2>&1 | grep -v "^n=" | head -1
PFEM_MG_NUC=3 python tools/solve_n.py 50 mg:1 mg:2 --cloud 2>&1 | grep -v "^n=\|^    "
PFEM_MG_NUC=3 python tools/solve_n.py 700 mg:1 mg:2 --dim 2 2>&1 | grep -v "^n=\|^    "
