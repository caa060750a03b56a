"""Short driver for ncu captures: a few launches of each hot kernel on a Kuhn box (default C4, n=69)."""
import argparse
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfem_b200 import meshgen as mg
from pfem_b200.capi import PfemContext

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=69)
ap.add_argument("--what", default="pspg,wc")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--permute", action="store_true")
args = ap.parse_args()

mesh = mg.kuhn_box(3, args.cells, permute=args.permute)
g = mg.gravity(3)
if "pspg" in args.what:
    q, qp = mg.pspg_state(mesh)
    P = mg.PSPG_PARAMS
    with PfemContext(3, 0) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, q)
        ctx.pspg_set_qprev(qp)
        par = ctx.pspg_params(P["rho"], P["mu"], P["dt"], g)
        for _ in range(args.iters):
            ctx.pspg_assemble_resident(par)
        sol = ctx.pspg_solve(1e-10, 8, fetch=False)
        print("pspg ok", sol["iters"], sol["rel_res"])
if "wc" in args.what:
    W = mg.WC_PARAMS
    st = mg.wc_state(mesh)
    with PfemContext(3, 0) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, np.concatenate([st["v"], st["p"], st["rho"], st["acc"]]))
        wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, True)
        for _ in range(args.iters):
            dt = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
            ctx.wc_step(wp, dt)
        print("wc ok", dt)
