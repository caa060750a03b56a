"""Time assembly / WC kernels with CUDA events via the context's phase timers (quick A/B harness)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfem_b200 import meshgen as mg
from pfem_b200.capi import PfemContext

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 69
what = sys.argv[2] if len(sys.argv) > 2 else "pspg,wc"
permute = len(sys.argv) > 3 and sys.argv[3] == "permute"
mesh = mg.kuhn_box(3, cells, permute=permute)
g = mg.gravity(3)
if "pspg" in what:
    q, qp = mg.pspg_state(mesh)
    P = mg.PSPG_PARAMS
    with PfemContext(3, 0) as ctx:
        t0 = time.perf_counter(); ctx.set_topology(mesh.conn, mesh.flags); t1 = time.perf_counter()
        ctx.set_positions(mesh.x); ctx.set_dirichlet(mesh.dir_mask, mesh.dir_val)
        ctx.set_states(0, q); ctx.pspg_set_qprev(qp)
        par = ctx.pspg_params(P["rho"], P["mu"], P["dt"], g)
        for _ in range(3): ctx.pspg_assemble_resident(par)
        ctx.profile_enable(True); ctx.profile_reset()
        for _ in range(10): ctx.pspg_assemble_resident(par)
        ms, n = ctx.profile_get("Assemble system")
        print(f"ASM cfg={os.environ.get('PFEM_ASM_CFG','0')} cells={cells} permute={permute}: {ms/n*1e3:.1f} us/assemble  ({mesh.n_elems/(ms/n*1e-3)/1e6:.0f} Melem/s)  topo {1e3*(t1-t0):.1f} ms")
if "wc" in what:
    W = mg.WC_PARAMS
    st = mg.wc_state(mesh)
    with PfemContext(3, 0) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, np.concatenate([st["v"], st["p"], st["rho"], st["acc"]]))
        wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, True)
        for _ in range(3): ctx.wc_step(wp, 1e-7)
        ctx.profile_enable(True); ctx.profile_reset()
        for _ in range(10):
            ctx.wc_step(wp, 1e-7); ctx.wc_next_dt(wp, 0.1, 1e-3)
        tot = 0
        for ph in ("Update solutions", "Solving continuity eq", "Solving momentum eq", "Compute next dt"):
            ms, n = ctx.profile_get(ph); tot += ms / n
            print(f"  WC {ph}: {ms/n*1e3:.1f} us")
        for ph in ("CFL nodal pass", "CFL element pass", "Build tiles"):
            ms, n = ctx.profile_get(ph)
            if n: print(f"    ({ph}: {ms/n*1e3:.1f} us x {n})")
        print(f"WC cells={cells}: {tot*1e3:.1f} us/step ({mesh.n_elems/(tot*1e-3)/1e6:.0f} Melem/s)")
