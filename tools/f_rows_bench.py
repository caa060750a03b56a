"""Timings of the SURVEY 8(f) rows at C4 size (general, untuned kernels): Bingham / Boussinesq PSPG assembly, implicit heat
system + CG, BoussinesqWC explicit step, fractional-step systems + CG.  Usage: python tools/f_rows_bench.py [cells] [wc_cells]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfem_b200 import meshgen as mg
from pfem_b200.capi import PfemContext

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 69
wc_cells = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dim = 3
eps = np.finfo(float).eps


def timed(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(1e3 * (time.perf_counter() - t0))
    return min(ts), out


mesh = mg.kuhn_box(dim, cells)
nn, ne = mesh.n_nodes, mesh.n_elems
q, qp = mg.pspg_state(mesh)
P = mg.PSPG_PARAMS
c = mesh.coords()
T0 = 300.0 + 10.0 * c[:, 0]
bound = (mesh.flags & mg.F_BOUND) != 0
t_mask = (bound & ((np.abs(c[:, 0]) < 1e-12) | (np.abs(c[:, 0] - 1.0) < 1e-12))).astype(np.uint8)
t_val = np.where(c[:, 0] < 0.5, 310.0, 290.0)
print(f"mesh: Kuhn box n={cells}: {ne} tets, {nn} nodes (wall clock of the C-ABI calls incl. their host copies, best of 5)")
with PfemContext(dim, 0) as ctx:
    ctx.set_mesh(mesh)
    ctx.set_states(0, q)
    par = ctx.pspg_params(P["rho"], P["mu"], P["dt"], mg.gravity(dim))
    ctx.pspg_set_qprev(qp)
    t, _ = timed(lambda: ctx.pspg_assemble_resident(par))
    print(f"PSPG assembly, tuned kernel (IncompNewtonNoT):      {t:7.2f} ms  {ne / t / 1e3:7.0f} Melem/s")
    ctx.set_bingham(50.0, 100.0)
    t, _ = timed(lambda: ctx.pspg_assemble_resident(par))
    print(f"PSPG assembly, general kernels, Bingham:            {t:7.2f} ms  {ne / t / 1e3:7.0f} Melem/s")
    ctx.set_bingham(None)
    ctx.set_thermal(0.6, 4.186, 6.9e-3, 300.0)
    ctx.set_temperature(T0)
    t, _ = timed(lambda: ctx.pspg_assemble_resident(par))
    print(f"PSPG assembly, general kernels, Boussinesq:         {t:7.2f} ms  {ne / t / 1e3:7.0f} Melem/s")
    s = ctx.pspg_solve(1e-12, 5000, fetch=False)
    print(f"   its multigrid-FGMRES solve: {s['iters']} iterations, status {s['status']}")
    ctx.set_temperature_bc(t_mask, t_val)
    t, _ = timed(lambda: ctx.heat_assemble(P["rho"], 4.186, 6.0e3, P["dt"], T0))
    print(f"implicit heat system M(cv rho) + dt L(k):           {t:7.2f} ms  {ne / t / 1e3:7.0f} Melem/s")

    def heat():
        ctx.set_temperature(T0)
        return ctx.heat_solve(1e-12, 10000, fetch=False)
    t, s = timed(heat)
    print(f"heat solve, Jacobi-CG to 1e-12:                     {t:7.2f} ms  {s['iters']} iterations ({1e3 * t / max(s['iters'], 1):.0f} us each), status {s['status']}")
    ctx.set_thermal(None)
    # fractional step
    t, _ = timed(lambda: ctx.fs_assemble_vapp(par, 1.0, qp))
    print(f"FracStep velocity prediction (M/dt + K, BC):        {t:7.2f} ms  {ne / t / 1e3:7.0f} Melem/s")
    t, s = timed(lambda: ctx.fs_solve(0, eps, 10000), 3)
    print(f"   Eigen-style CG to eps:                           {t:7.2f} ms  {s['iters']} iterations ({1e3 * t / max(s['iters'], 1):.0f} us each), status {s['status']}")
    vt = s["x"]
    t, _ = timed(lambda: ctx.fs_assemble_pcorr(P["rho"], P["dt"], 1.0, vt, qp[dim * nn:]))
    print(f"FracStep pressure system (L, rhs):                  {t:7.2f} ms  {ne / t / 1e3:7.0f} Melem/s")
    dp = 100.0 * c[:, 2]
    t, _ = timed(lambda: ctx.fs_assemble_vcorr(P["rho"], P["dt"], dp))
    print(f"FracStep velocity correction (M, rhs):              {t:7.2f} ms  {ne / t / 1e3:7.0f} Melem/s")
    t, s = timed(lambda: ctx.fs_solve(2, eps, 10000), 3)
    print(f"   Eigen-style CG to eps:                           {t:7.2f} ms  {s['iters']} iterations ({1e3 * t / max(s['iters'], 1):.0f} us each), status {s['status']}")

# BoussinesqWC explicit step (gather kernels + explicit heat pass) against the plain explicit step
mesh = mg.kuhn_box(dim, wc_cells)
nn, ne = mesh.n_nodes, mesh.n_elems
st = mg.wc_state(mesh)
W = mg.WC_PARAMS
packed = np.concatenate([st["v"], st["p"], st["rho"], st["acc"]])
for thermal in (False, True):
    with PfemContext(dim, 0) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, packed)
        wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], mg.gravity(dim), True)
        if thermal:
            ctx.set_thermal(6.0e3, 4.186, 6.9e-3, 300.0)
            ctx.set_temperature(300.0 + 10.0 * mesh.coords()[:, 0])
        dt = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
        dt, _ = ctx.wc_run(wp, 5, W["securityCoeff"], 1e-3, dt)
        t0 = time.perf_counter()
        dt, _ = ctx.wc_run(wp, 20, W["securityCoeff"], 1e-3, dt)
        t = 1e3 * (time.perf_counter() - t0) / 20
        print(f"explicit step n={wc_cells} ({ne} tets) {'BoussinesqWC (two-pass + heat pass)     ' if thermal else 'WCompNewtonNoT (two-pass kernels)      '}: {t:6.3f} ms/step  {ne / t / 1e3:7.0f} Melem/s")
