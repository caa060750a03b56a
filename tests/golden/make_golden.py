"""Generates tests/golden/*.npz from the REFERENCE'S OWN code (oracle/_ref/libpfem_ref.so, see oracle/refbuild/).

Run in the development container, where /root/reference exists:   python tests/golden/make_golden.py
The GPU box has no /root/reference; the fixtures written here are what travels (together with the prebuilt .so).

Every fixture stores the complete input (mesh arrays, states, parameters) next to the reference output, so the tests
never depend on the mesh generator reproducing the same numbers.  Outputs come from: MomContEqIncompNewton::
m_buildAbPSPG / m_applyBCPSPG / m_computeTauPSPG / solve (Picard), SolverWCompNewton::computeNextDT /
m_solveWCompNewtonNoT, MatrixBuilder get{M,K,D,L,C,F,H}, Element::computeJ/DetJ/InvJ/getRin, Mesh quadrature tables.
The direct solver behind the stand-in Eigen::SparseLU is SciPy SuperLU (COLAMD) for the Picard fixtures.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from oracle import ref  # noqa: E402
from pfem_b200 import meshgen as mg  # noqa: E402


def mesh_arrays(mesh):
    return dict(dim=np.int64(mesh.dim), x=mesh.x, conn=mesh.conn, flags=mesh.flags, dir_mask=mesh.dir_mask, dir_val=mesh.dir_val)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.0f} KiB")


def pspg_fixture(name, mesh, q, q_prev, par, facets=None, gamma=0.0):
    dim, nn = mesh.dim, mesh.n_nodes
    extra = {} if facets is None else dict(facets=facets, gamma=np.float64(gamma))
    with ref.RefCase(mesh, "pspg", par, facets=facets, gamma=gamma) as rc:
        rc.set_states(q)
        detJ, J, invJ, rin = rc.element_geometry()
        em = rc.element_matrices()
        Ae, be, tau = rc.pspg_elements(q_prev)
        A0, b0 = rc.pspg_build(q_prev, False)
        A1, b1 = rc.pspg_build(q_prev, True)
    assert (A0.indptr == A1.indptr).all() and (A0.indices == A1.indices).all()
    keep = np.arange(0, mesh.n_elems, max(1, mesh.n_elems // 24))   # a sample of the element-local systems
    save(name, **mesh_arrays(mesh), q=q, q_prev=q_prev, par=par, detJ=detJ, invJ=invJ, rin=rin, tau=tau,
         elem_ids=keep, Ae=Ae[keep], be=be[keep], **{"el_" + k: v[keep] for k, v in em.items()},
         indptr=A1.indptr.astype(np.int64), indices=A1.indices.astype(np.int32), A_nobc=A0.data, b_nobc=b0, A=A1.data, b=b1,
         **extra)


def picard_fixture(name, mesh, q_prev, par, max_iter=10, min_res=1e-6):
    ref.use_scipy_direct_solver(True)
    with ref.RefCase(mesh, "pspg", np.concatenate([par, [max_iter, min_res]])) as rc:
        rc.set_states(q_prev)
        ok, iters = rc.pspg_solve()
        q, x = rc.get_states(), rc.positions()
    ref.use_scipy_direct_solver(False)
    assert ok
    save(name, **mesh_arrays(mesh), q_prev=q_prev, par=par, max_iter=np.int64(max_iter), min_res=np.float64(min_res),
         ok=np.int64(ok), iters=np.int64(iters), q=q, x_new=x)


def wc_fixture(name, mesh, st, eq, meduri, n_steps=3, max_dt=1e-3, facets=None, gamma=0.0):
    W = mg.WC_PARAMS
    g = mg.gravity(mesh.dim)
    wpar = orc.wc_param_array(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, meduri, eq)
    q0 = np.concatenate([st["v"], st["p"], st["rho"], st["acc"]])
    dts, states, xs = [], [], []
    extra = {} if facets is None else dict(facets=facets, gamma=np.float64(gamma))
    with ref.RefCase(mesh, "wc", np.concatenate([wpar, [1e-6, max_dt, W["securityCoeff"]]]), facets=facets, gamma=gamma) as rc:
        rc.set_states(q0)
        for _ in range(n_steps):
            dt = rc.wc_next_dt()          # SolverWCompNewton::computeNextDT on the current state
            assert rc.wc_step(dt)         # m_solveWCompNewtonNoT
            dts.append(dt)
            states.append(rc.get_states())
            xs.append(rc.positions())
    save(name, **mesh_arrays(mesh), q0=q0, wpar=wpar, security_coeff=np.float64(W["securityCoeff"]), max_dt=np.float64(max_dt),
         dts=np.array(dts), states=np.array(states), xs=np.array(xs), **extra)


def boussinesq_fixtures():
    """BoussinesqWC: SolverWCompNewton::m_solveBoussinesqWC = explicit heat equation, continuity,
    momentum with the buoyancy factor, CFL with the thermal diffusivity; temperature Dirichlet data on two walls.
    Conductivity and expansion coefficient are exaggerated so that three steps show them."""
    W = mg.WC_PARAMS
    for dim, n in ((2, 8), (3, 4)):
        mesh = mg.kuhn_box(dim, n, free_fraction=0.02, permute=True)
        nn = mesh.n_nodes
        st = mg.wc_state(mesh)
        st["acc"] = 0.3 * np.random.default_rng(3).standard_normal(st["acc"].shape)
        c = mesh.coords()
        T0 = 300.0 + 10.0 * c[:, 0] + 2.0 * np.random.default_rng(5).standard_normal(nn)
        bound = (mesh.flags & mg.F_BOUND) != 0
        t_mask = (bound & ((np.abs(c[:, 0]) < 1e-12) | (np.abs(c[:, 0] - 1.0) < 1e-12))).astype(np.uint8)
        t_val = np.where(c[:, 0] < 0.5, 310.0, 290.0)
        th = dict(k=6.0e3, cv=4.186, alpha=6.9e-3, Tr=300.0, t_mask=t_mask, t_val=t_val)
        wpar = orc.wc_param_array(W["mu"], W["K0"], W["K0p"], W["rhoStar"], mg.gravity(dim), True, "CDS_dpdt")
        q0 = np.concatenate([st["v"], st["p"], st["rho"], st["acc"], T0])
        dts, states, xs = [], [], []
        with ref.RefCase(mesh, "wc", np.concatenate([wpar, [1e-6, 1e-3, W["securityCoeff"]]]), thermal=th) as rc:
            rc.set_states(q0)
            for _ in range(3):
                dt = rc.wc_next_dt()
                assert rc.wc_step(dt)
                dts.append(dt)
                states.append(rc.get_states())
                xs.append(rc.positions())
        save(f"wcb_{dim}d_boussinesq", **mesh_arrays(mesh), q0=q0, wpar=wpar, security_coeff=np.float64(W["securityCoeff"]),
             max_dt=np.float64(1e-3), dts=np.array(dts), states=np.array(states), xs=np.array(xs),
             thermal=np.array([th["k"], th["cv"], th["alpha"], th["Tr"]]), t_mask=t_mask, t_val=t_val)


def bingham_fixtures():
    """Problem id "Bingham": shear-rate dependent K factor of MomContEquation.inl:102-119 (regularised yield stress)."""
    tau0, m_reg = 50.0, 100.0
    for dim, n in ((2, 6), (3, 4)):
        mesh, q, q_prev, par = H.pspg_case(dim, n, free_fraction=0.02, permute=True)
        with ref.RefCase(mesh, "pspg", par, bingham=(tau0, m_reg)) as rc:
            rc.set_states(q)
            Ae, be, tau = rc.pspg_elements(q_prev)
            A1, b1 = rc.pspg_build(q_prev, True)
        save(f"pspgb_{dim}d_bingham", **mesh_arrays(mesh), q=q, q_prev=q_prev, par=par, bingham=np.array([tau0, m_reg]), tau=tau,
             indptr=A1.indptr.astype(np.int64), indices=A1.indices.astype(np.int32), A=A1.data, b=b1)


def boussinesq_incompressible_fixtures():
    """Problem id "Boussinesq": PSPG system with the buoyancy factors of MomContEquation.inl:166-199 and the implicit heat
    system HeatEqIncompNewton::m_buildAb + m_applyBC (IncompNewton/HeatEquation.inl:227-412), temperature Dirichlet data
    on two walls, no flux terms."""
    for dim, n in ((2, 6), (3, 4)):
        mesh, q, q_prev, par = H.pspg_case(dim, n, free_fraction=0.02, permute=True)
        nn = mesh.n_nodes
        c = mesh.coords()
        T = 300.0 + 10.0 * c[:, 0] + 2.0 * np.random.default_rng(5).standard_normal(nn)
        bound = (mesh.flags & mg.F_BOUND) != 0
        t_mask = (bound & ((np.abs(c[:, 0]) < 1e-12) | (np.abs(c[:, 0] - 1.0) < 1e-12))).astype(np.uint8)
        t_val = np.where(c[:, 0] < 0.5, 310.0, 290.0)
        th = dict(alpha=6.9e-3, Tr=300.0, k=0.6, cv=4.186, t_mask=t_mask, t_val=t_val)
        with ref.RefCase(mesh, "pspg", par, thermal=th) as rc:
            rc.set_states(np.concatenate([q, T]))
            A, b = rc.pspg_build(q_prev, True)
            Ah, bh = rc.in_heat_build(T, True)
        save(f"inb_{dim}d_boussinesq", **mesh_arrays(mesh), q=q, q_prev=q_prev, par=par, T=T, t_mask=t_mask, t_val=t_val,
             thermal=np.array([th["alpha"], th["Tr"], th["k"], th["cv"]]),
             indptr=A.indptr.astype(np.int64), indices=A.indices.astype(np.int32), A=A.data, b=b,
             h_indptr=Ah.indptr.astype(np.int64), h_indices=Ah.indices.astype(np.int32), h_A=Ah.data, h_b=bh)


def dam_break_fixtures():
    """The reference's first configurations in their real proportions and size (C1/C3: 1001 nodes, 1600 triangles at
    ncol = 20; a small 3-D tank): water column + tank walls whose DRY nodes are isBound and isFree at once."""
    P, W = mg.PSPG_PARAMS, mg.WC_PARAMS
    for dim, ncol in ((2, 20), (3, 4)):
        mesh = mg.dam_break(dim, ncol, permute=True)
        nn, L = mesh.n_nodes, mesh.meta["L"]
        c = mesh.coords()
        wet = (mesh.flags & mg.F_FREE) == 0
        p = np.where(wet, P["rho"] * 9.81 * (2 * L - c[:, dim - 1]), 0.0)
        v = 0.05 * np.random.default_rng(3).standard_normal((dim, nn))
        v[:, (mesh.flags & mg.F_BOUND) != 0] = 0.0
        q = np.concatenate([v.reshape(-1), p])
        par = orc.pspg_param_array(P["rho"], P["mu"], P["dt"], mg.gravity(dim))
        pspg_fixture(f"pspg_{dim}d_dambreak", mesh, q, q, par)
        picard_fixture(f"picard_{dim}d_dambreak", mesh, q, par)
        st = dict(v=v.reshape(-1).copy(), p=p.copy(), rho=mg.tait_density(p, W["K0"], W["K0p"], W["rhoStar"]), acc=np.zeros(dim * nn))
        wc_fixture(f"wc_{dim}d_dambreak", mesh, st, "CDS_dpdt", True)


def tables_fixture():
    mesh = mg.kuhn_box(3, 2)
    par = orc.pspg_param_array(1000.0, 1e-3, 1e-3, mg.gravity(3))
    out = {}
    with ref.RefCase(mesh, "pspg", par) as rc:
        for dimension, n_gp in ((1, 3), (2, 3), (3, 4)):   # facet + element rules used by the equations (MomContEquation.inl:13-23)
            gp, w, sf, ref_size = rc.tables(dimension, n_gp)
            out.update({f"gp{dimension}": gp, f"w{dimension}": w, f"sf{dimension}": sf, f"ref{dimension}": np.float64(ref_size)})
    save("tables", **out)


def main():
    if not ref.available():
        raise SystemExit("oracle/_ref not built and /root/reference absent")
    tables_fixture()
    for dim, n, kw in ((2, 5, dict(free_fraction=0.03, permute=True)), (3, 3, dict(free_fraction=0.03, permute=True))):
        mesh, q, q_prev, par = H.pspg_case(dim, n, **kw)
        pspg_fixture(f"pspg_{dim}d_kuhn", mesh, q, q_prev, par)
    for dim, npts in ((2, 60), (3, 90)):
        mesh = mg.delaunay_cloud(dim, npts, free_fraction=0.02)
        q, q_prev = mg.pspg_state(mesh)
        q_prev = q_prev + 0.02 * np.random.default_rng(8).standard_normal(q_prev.shape)
        P = mg.PSPG_PARAMS
        par = orc.pspg_param_array(P["rho"], P["mu"], P["dt"], mg.gravity(dim))
        pspg_fixture(f"pspg_{dim}d_delaunay", mesh, q, q_prev, par)
    for dim, n in ((2, 8), (3, 4)):
        mesh = mg.kuhn_box(dim, n)
        _, q_prev = mg.pspg_state(mesh)
        P = mg.PSPG_PARAMS
        picard_fixture(f"picard_{dim}d", mesh, q_prev, orc.pspg_param_array(P["rho"], P["mu"], P["dt"], mg.gravity(dim)))
    for dim, n, kw in ((2, 6, dict(free_fraction=0.03, permute=True)), (3, 3, dict(free_fraction=0.03, permute=True))):
        mesh = mg.kuhn_box(dim, n, **kw)
        for eq in ("CDS_dpdt", "CDS_drhodt", "CDS_rho"):
            for meduri in (True, False):
                st = mg.wc_state(mesh)
                st["acc"] = 0.5 * np.random.default_rng(4).standard_normal(st["acc"].shape)
                wc_fixture(f"wc_{dim}d_{eq}_{'meduri' if meduri else 'none'}", mesh, st, eq, meduri)
    fst_fixtures()
    boussinesq_fixtures()
    bingham_fixtures()
    boussinesq_incompressible_fixtures()
    dam_break_fixtures()


def fst_fixtures():
    """gamma > 0 on a wavy free surface (a flat one has zero net force): MatrixBuilder::getFST through the facet loops of
    m_applyBCPSPG (PSPG.inl:155-187) and MomEqWCompNewton::m_applyBC (MomEquation.inl:312-336)."""
    gamma = 7.28   # 100x water-air: makes the term ~1e-3 of the gravity load on these coarse meshes
    for dim, n in ((2, 6), (3, 4)):
        mesh, q, q_prev, par = H.pspg_case(dim, n, permute=True)
        H.wavy_free_surface(mesh)
        fac = mg.boundary_facets(mesh)
        pspg_fixture(f"pspg_{dim}d_fst", mesh, q, q_prev, par, facets=fac, gamma=gamma)
        for eq, meduri in (("CDS_dpdt", True), ("CDS_rho", False)):
            st = mg.wc_state(mesh)
            st["acc"] = 0.5 * np.random.default_rng(4).standard_normal(st["acc"].shape)
            wc_fixture(f"wc_{dim}d_fst_{eq}", mesh, st, eq, meduri, facets=fac, gamma=gamma)


if __name__ == "__main__":
    main()
