"""Timing of the small reference configurations (C1/C2/C3-sized stand-ins): launch-latency-bound regime."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfem_b200 import meshgen as mg
from pfem_b200.capi import PfemContext

def pspg(dim, n, label):
    mesh = mg.kuhn_box(dim, n); q, qp = mg.pspg_state(mesh); P = mg.PSPG_PARAMS
    with PfemContext(dim, 0) as ctx:
        ctx.set_mesh(mesh); ctx.set_states(0, q)
        par = ctx.pspg_params(P["rho"], P["mu"], P["dt"], mg.gravity(dim))
        ctx.snapshot_positions(); ctx.pspg_assemble(par, qp)
        ctx.pspg_solve(1e-10, 20000, fetch=False)
        t0 = time.perf_counter(); ctx.pspg_assemble(par, qp); t1 = time.perf_counter()
        s = ctx.pspg_solve(1e-10, 20000, fetch=True); t2 = time.perf_counter()
        print(f"{label}: {mesh.n_elems} elems, assemble {1e3*(t1-t0):.2f} ms, solve {1e3*(t2-t1):.1f} ms, {s['iters']} iters, {1e6*(t2-t1)/max(s['iters'],1):.1f} us/iter, status {s['status']}")

def wc(dim, n, label, steps=200):
    mesh = mg.kuhn_box(dim, n); st = mg.wc_state(mesh); W = mg.WC_PARAMS
    with PfemContext(dim, 0) as ctx:
        ctx.set_mesh(mesh); ctx.set_states(0, np.concatenate([st["v"], st["p"], st["rho"], st["acc"]]))
        wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], mg.gravity(dim), True)
        dt = ctx.wc_next_dt(wp, 0.1, 1e-3)
        for _ in range(10): ctx.wc_step(wp, dt); dt = ctx.wc_next_dt(wp, 0.1, 1e-3)
        t0 = time.perf_counter()
        for _ in range(steps): ctx.wc_step(wp, dt); dt = ctx.wc_next_dt(wp, 0.1, 1e-3)
        t1 = time.perf_counter()
        print(f"{label}: {mesh.n_elems} elems, {1e6*(t1-t0)/steps:.1f} us/step (step + CFL dt, host-synchronous)")

pspg(2, 28, "C1-size 2D PSPG")
pspg(3, 13, "C2-size 3D PSPG")
wc(2, 28, "C3-size 2D WC")


def wc_graph(dim, n, label, steps=2000):
    mesh = mg.kuhn_box(dim, n); st = mg.wc_state(mesh); W = mg.WC_PARAMS
    with PfemContext(dim, 0) as ctx:
        ctx.set_mesh(mesh); ctx.set_states(0, np.concatenate([st["v"], st["p"], st["rho"], st["acc"]]))
        wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], mg.gravity(dim), True)
        dt = ctx.wc_next_dt(wp, 0.1, 1e-3)
        dt, _ = ctx.wc_run(wp, 50, 0.1, 1e-3, dt)
        t0 = time.perf_counter(); dt, el = ctx.wc_run(wp, steps, 0.1, 1e-3, dt); t1 = time.perf_counter()
        print(f"{label}: {mesh.n_elems} elems, {1e6*(t1-t0)/steps:.1f} us/step (pfem_wc_run: CUDA graph, dt chained on the device), simulated {el:.3e} s")

wc_graph(2, 28, "C3-size 2D WC graph")
