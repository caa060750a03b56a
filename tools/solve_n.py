import sys, numpy as np
sys.path.insert(0,'/root/repo')
from pfem_b200 import meshgen as mg
from pfem_b200.capi import PfemContext
n=int(sys.argv[1])
mesh=mg.kuhn_box(3,n); q,qp=mg.pspg_state(mesh); P=mg.PSPG_PARAMS
with PfemContext(3,0) as ctx:
    ctx.set_mesh(mesh); ctx.set_states(0,q); ctx.pspg_set_qprev(qp)
    par=ctx.pspg_params(P["rho"],P["mu"],P["dt"],mg.gravity(3))
    for t in range(3):
        ctx.pspg_assemble_resident(par)
        s=ctx.pspg_solve(1e-10,40000,fetch=False)
        print(n,'status',s['status'],'iters',s['iters'],'rel',s['rel_res'])
