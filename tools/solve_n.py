"""Solve the bench PSPG system at size n with a chosen preconditioner and report iterations / time / phases.

usage: python tools/solve_n.py n [kind[:sweeps[:damping]] ...] [--cloud] [--tol 1e-12] [--dim 3]
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from pfem_b200 import meshgen as mg            # noqa: E402
from pfem_b200.capi import PfemContext         # noqa: E402


def main():
    args, skip = [], False
    for a in sys.argv[1:]:
        if skip:
            skip = False
        elif a in ("--dim", "--tol", "--maxit"):
            skip = True
        elif not a.startswith("--"):
            args.append(a)
    n = int(args[0])
    kinds = args[1:] or ["mg", "block"]
    dim = int(sys.argv[sys.argv.index("--dim") + 1]) if "--dim" in sys.argv else 3
    tol = float(sys.argv[sys.argv.index("--tol") + 1]) if "--tol" in sys.argv else 1e-12
    maxit = int(sys.argv[sys.argv.index("--maxit") + 1]) if "--maxit" in sys.argv else 40000
    mesh = mg.delaunay_cloud(dim, (n + 1) ** dim) if "--cloud" in sys.argv else mg.kuhn_box(dim, n)
    q, qp = mg.pspg_state(mesh)
    P = mg.PSPG_PARAMS
    print(f"n={n} dim={dim} nodes={mesh.n_nodes} elems={mesh.n_elems}", flush=True)
    ref = None
    with PfemContext(dim, 0) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, q)
        ctx.pspg_set_qprev(qp)
        par = ctx.pspg_params(P["rho"], P["mu"], P["dt"], mg.gravity(dim))
        for kind in kinds:
            a = kind.split(":")
            ctx.pspg_set_preconditioner(a[0], int(a[1]) if len(a) > 1 else 0, float(a[2]) if len(a) > 2 else 0.0)
            for rep in range(2):
                ctx.pspg_assemble_resident(par)
                ctx.profile_reset()
                ctx.profile_enable((2 if '--detail' in sys.argv else 1) if rep == 1 else 0)
                t0 = time.perf_counter()
                s = ctx.pspg_solve(tol, maxit, fetch=True)
                dt = time.perf_counter() - t0
                ctx.profile_enable(False)
            used, lv = ctx.pspg_get_preconditioner()
            if ref is None:
                ref = s["q"]
            dq = np.abs(s["q"] - ref).max() / np.abs(ref).max()
            print(f"{kind:14s} used={used} levels={lv} status={s['status']} iters={s['iters']} rel={s['rel_res']:.2e} "
                  f"wall={dt * 1e3:.1f} ms  |q-q0|/|q0|={dq:.1e}", flush=True)
            names = ["Solve system", "SpMV", "Preconditioner pattern", "Preconditioner setup", "Preconditioner apply",
                     "MG smoother setup L0", "MG smoother setup coarse", "MG Galerkin L0", "MG Galerkin coarse",
                     "MG coarsest inverse", "MG smooth L0", "MG resid L0"]
            ph = {k: ctx.profile_get(k) for k in names}
            print("    " + "  ".join(f"{k}={v[0]:.2f}ms/{v[1]}" for k, v in ph.items()), flush=True)


if __name__ == "__main__":
    main()
