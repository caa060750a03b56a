"""Wall clock of pfem_pspg_solve alone (system assembled once) at a given size; env knobs select the solver variants.
Usage: python tools/solve_bench.py [cells] [tol] [dim] [cloud]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfem_b200 import meshgen as mg
from pfem_b200.capi import PfemContext

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 69
tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-12
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mesh = mg.delaunay_cloud(dim, cells) if "cloud" in sys.argv else mg.kuhn_box(dim, cells)   # cloud: cells = number of points
q, qp = mg.pspg_state(mesh)
P = mg.PSPG_PARAMS
with PfemContext(dim, 0) as ctx:
    ctx.set_mesh(mesh)
    ctx.set_states(0, q)
    par = ctx.pspg_params(P["rho"], P["mu"], P["dt"], mg.gravity(dim))
    ctx.pspg_assemble(par, qp)
    ts = []
    reps = int(os.environ.get("SOLVE_REPS", "6"))
    for it in range(reps):
        if it == reps - 1:
            ctx.profile_enable(True); ctx.profile_reset()
        t0 = time.perf_counter()
        s = ctx.pspg_solve(tol, 40000, fetch=False)
        ts.append(1e3 * (time.perf_counter() - t0))
    print(f"cells {cells} dim {dim} tol {tol:g}: {s['iters']} its, status {s['status']}, rel_res {s['rel_res']:.2e}, "
          f"solve ms {['%.2f' % t for t in ts]}  env { {k: v for k, v in os.environ.items() if k.startswith('PFEM_')} }")
    for ph in ("Preconditioner setup", "Preconditioner apply", "SpMV", "GMRES orthogonalisation", "Solve system"):
        ms, n = ctx.profile_get(ph)
        if n: print(f"      device phase {ph}: {ms:.2f} ms x{n}")
