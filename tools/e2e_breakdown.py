"""Wall-clock breakdown of one remeshed PSPG step through the host-buffer ABI (what bench.py's e2e leg times)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfem_b200 import meshgen as mg
from pfem_b200.capi import PfemContext

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 69
mesh = mg.kuhn_box(3, cells)
q, qp = mg.pspg_state(mesh)
P = mg.PSPG_PARAMS
def pin(a, dt=None):
    a = np.ascontiguousarray(a, dtype=dt)
    t = torch.empty(max(a.nbytes, 1), dtype=torch.uint8, pin_memory=True)
    v = t.numpy()[: a.nbytes].view(a.dtype).reshape(a.shape); v[...] = a
    pin.keep.append(t); return v
pin.keep = []
conn, flags, x, qh, qph = pin(mesh.conn, np.uint64), pin(mesh.flags, np.uint8), pin(mesh.x), pin(q), pin(qp)
dm, dv = pin(mesh.dir_mask, np.uint8), pin(mesh.dir_val)
with PfemContext(3, 0) as ctx:
    par = ctx.pspg_params(P["rho"], P["mu"], P["dt"], mg.gravity(3))
    for it in range(4):
        ctx.profile_enable(True); ctx.profile_reset()
        T = [time.perf_counter()]
        ctx.set_topology(conn, flags); T.append(time.perf_counter())
        ctx.set_positions(x); ctx.set_dirichlet(dm, dv); ctx.set_states(0, qh); T.append(time.perf_counter())
        ctx.pspg_assemble(par, qph); T.append(time.perf_counter())
        s = ctx.pspg_solve(1e-12, 40000, fetch=True); T.append(time.perf_counter())
        names = ["set_topology", "fields H2D", "assemble (+qPrev H2D)", "solve + D2H"]
        print(f"iter {it}: " + ", ".join(f"{n} {1e3*(b-a):.1f} ms" for n, a, b in zip(names, T[:-1], T[1:])) + f" | total {1e3*(T[-1]-T[0]):.1f} ms, {s['iters']} its")
        for ph in ("Build pattern", "Preconditioner pattern", "Preconditioner setup", "Preconditioner apply", "Solve system", "Assemble system", "SpMV"):
            ms, n = ctx.profile_get(ph)
            if n: print(f"      device phase {ph}: {ms:.2f} ms x{n}")
