"""A few explicit steps on a Kuhn box (default C5) for ncu launch lists: python tools/wc_steps.py [cells] [steps] [variant]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfem_b200 import meshgen as mg
from pfem_b200.capi import PfemContext
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 150
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
mesh = mg.kuhn_box(3, cells)
st = mg.wc_state(mesh)
W = mg.WC_PARAMS
with PfemContext(3, 0) as ctx:
    ctx.set_mesh(mesh)
    ctx.set_states(0, np.concatenate([st["v"], st["p"], st["rho"], st["acc"]]))
    if variant: ctx.wc_set_variant(variant)
    wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], mg.gravity(3), True)
    dt = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
    for _ in range(steps):
        ctx.wc_step(wp, dt)
        dt = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
    print("dt", dt)
