"""Iteration counts of the multigrid- vs node-block-Jacobi-preconditioned solve over material / time-step regimes and mesh
families (robustness of the l1 damping and of the AUTO hand-over).  usage: python tools/mg_robustness.py [n3d] [n2d]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from pfem_b200 import meshgen as mg            # noqa: E402
from pfem_b200.capi import PfemContext         # noqa: E402

REGIMES = [  # rho, mu, dt
    (1000.0, 1e-3, 1e-3),   # water, dam break (examples/3D/damBreakKoshizuka)
    (100.0, 1.0, 1e-3),     # viscous drop (examples/2D/squareToDisk)
    (1000.0, 1e-3, 1e-5),   # very small dt: mass dominated
    (1000.0, 1e-3, 1e-1),   # large dt
    (1000.0, 10.0, 1e-2),   # viscosity dominated
    (1.0, 1e-5, 1e-3),      # gas-like
]


def run(mesh, label):
    dim = mesh.dim
    q, qp = mg.pspg_state(mesh)
    with PfemContext(dim, 0) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, q)
        ctx.pspg_set_qprev(qp)
        for rho, mu, dt in REGIMES:
            par = ctx.pspg_params(rho, mu, dt, mg.gravity(dim))
            out = []
            ref = None
            for kind in (("auto",) if "--auto" in sys.argv else ("mg", "block", "auto")):
                ctx.pspg_set_preconditioner(kind)
                ctx.pspg_assemble_resident(par)
                s = ctx.pspg_solve(1e-10, 3000 if kind != "auto" else 20000, fetch=True)
                if ref is None and s["status"] == 0:
                    ref = s["q"]
                dq = float(np.abs(s["q"] - ref).max() / max(np.abs(ref).max(), 1e-300)) if ref is not None else float("nan")
                out.append(f"{kind}: st={s['status']} it={s['iters']:5d} rel={s['rel_res']:.1e} dq={dq:.1e}")
            print(f"{label:12s} rho={rho:<7g} mu={mu:<6g} dt={dt:<6g} | " + " | ".join(out), flush=True)


def main():
    a = [v for v in sys.argv[1:] if not v.startswith("--")]
    n3 = int(a[0]) if len(a) > 0 else 24
    n2 = int(a[1]) if len(a) > 1 else 160
    run(mg.kuhn_box(3, n3), f"kuhn3d n={n3}")
    run(mg.delaunay_cloud(3, (n3 + 1) ** 3), "cloud3d")
    run(mg.kuhn_box(2, n2), f"kuhn2d n={n2}")
    run(mg.delaunay_cloud(2, (n2 + 1) ** 2), "cloud2d")


if __name__ == "__main__":
    main()
