"""Pins the CPU oracle against the REFERENCE'S OWN code.

tests/golden/*.npz hold outputs of the reference's unmodified hot-path sources (MatricesBuilder.inl, Element.cpp,
MomContEquationPSPG.inl, PicardAlgo.cpp, WCompNewton/{Cont,Mom}Equation.inl, WCompNewton/Solver.cpp, ...) compiled in
place from /root/reference against stand-in Eigen/sol2/gmsh headers (oracle/refbuild/, generator
tests/golden/make_golden.py).  These tests check oracle/pfem_oracle.cpp and oracle/literal_numpy.py against them; where
the library itself is present (development container) further randomised cases are compared live.
Tolerance 1e-12 relative per block type (the stand-in evaluates dense products in the order they are written, real
Eigen may reassociate small products; in practice the oracle reproduces these fixtures bit for bit).
"""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import ref
from pfem_b200 import meshgen as mg

from helpers import (block_errors, golden_csc, golden_names, load_golden, pspg_case, rel_err, split_wc,
                     vec_block_errors)

TOL = 1e-12


def test_fixtures_present():
    assert len(golden_names("pspg_")) >= 4 and len(golden_names("wc_")) >= 12 and len(golden_names("picard_")) >= 2


def test_quadrature_tables_match_reference():
    """Mesh::getGaussPoints/getGaussWeight/getShapeFunctions/getRefElementSize (Mesh.cpp:342-530) vs SURVEY appendix A literals."""
    _, z = load_golden("tables")
    a, b = 0.585410196624968, 0.138196601125011
    assert np.array_equal(z["gp2"][:, :2], np.array([[1 / 6, 1 / 6], [1 / 6, 2 / 3], [2 / 3, 1 / 6]]))
    assert np.array_equal(z["w2"], np.full(3, 1 / 3)) and z["ref2"] == 0.5
    assert np.array_equal(z["gp3"], np.array([[a, b, b], [b, a, b], [b, b, a], [b, b, b]]))
    assert np.array_equal(z["w3"], np.full(4, 0.25)) and z["ref3"] == 1 / 6
    for d in (2, 3):                                   # N = [1 - sum(xi), xi...]
        gp, sf = z[f"gp{d}"][:, :d], z[f"sf{d}"]
        assert np.abs(sf[:, 0] - (1 - gp.sum(1))).max() < 3e-16 and np.array_equal(sf[:, 1:], gp)


@pytest.mark.parametrize("name", golden_names("pspg_"))
def test_pspg_oracle_matches_reference_fixture(name):
    mesh, z = load_golden(name)
    dim, nn = mesh.dim, mesh.n_nodes
    vcur = z["q"][: dim * nn].copy()
    orc.set_facets(dim, z.get("facets"), float(z.get("gamma", 0.0)))     # gamma > 0 fixtures: facet terms on
    Ae, be, tau = orc.pspg_elements(mesh, vcur, z["q_prev"], z["par"])
    assert rel_err(tau, z["tau"]) < TOL
    ids = z["elem_ids"]
    assert rel_err(Ae[ids], z["Ae"]) < TOL and rel_err(be[ids], z["be"]) < TOL
    assert rel_err(mg.det_j(mesh), z["detJ"]) < TOL
    for bc, (ka, kb) in ((False, ("A_nobc", "b_nobc")), (True, ("A", "b"))):
        A, b = orc.pspg_build(mesh, vcur, z["q_prev"], z["par"], bc)
        A_ref = golden_csc(z, ka)
        errs = block_errors(A, A_ref, nn, dim)         # asserts the identical CSC pattern
        assert max(errs.values()) < TOL, (bc, errs)
        berr = vec_block_errors(b, z[kb], nn, dim)
        assert max(berr.values()) < TOL, (bc, berr)
    orc.set_facets(dim)
    if "facets" in z:                                    # the fixture must actually exercise the facet terms
        _, b0 = orc.pspg_build(mesh, vcur, z["q_prev"], z["par"], True)
        assert np.abs(b0 - z["b"]).max() > 1e-6 * np.abs(z["b"]).max()


@pytest.mark.parametrize("name", golden_names("pspg_"))
def test_literal_numpy_matches_reference_fixture(name):
    """The second restatement (dense products in numpy) against the reference's element systems."""
    from oracle import literal_numpy as lit
    mesh, z = load_golden(name)
    dim, nn = mesh.dim, mesh.n_nodes
    ids, P = z["elem_ids"], z["par"]
    Ae, be, tau = lit.pspg_elements(mesh, z["q"][: dim * nn], z["q_prev"], P[0], P[1], P[2], P[3:6])
    assert rel_err(Ae[ids], z["Ae"]) < 1e-11 and rel_err(be[ids], z["be"]) < 1e-11 and rel_err(tau, z["tau"]) < 1e-12


@pytest.mark.parametrize("name", golden_names("picard_"))
def test_picard_oracle_matches_reference_fixture(name):
    """MomContEqIncompNewton::solve -> PicardAlgo::solve (PicardAlgo.cpp:31-94, PSPG.inl:262-373): same iteration
    count, fields to 1e-8 (north_star), moved positions to 1e-12."""
    mesh, z = load_golden(name)
    dim, nn = mesh.dim, mesh.n_nodes
    out = orc.pspg_picard(mesh, z["q_prev"], z["q_prev"], z["par"], max_iter=int(z["max_iter"]), min_res=float(z["min_res"]))
    assert out["ok"] and out["iters"] == int(z["iters"])
    assert rel_err(out["q"][: dim * nn], z["q"][: dim * nn]) < 1e-8
    assert rel_err(out["q"][dim * nn:], z["q"][dim * nn:]) < 1e-8
    assert np.abs(out["x"] - z["x_new"]).max() < 1e-12


@pytest.mark.parametrize("name", golden_names("wc_"))
def test_wc_oracle_matches_reference_fixture(name):
    """SolverWCompNewton::computeNextDT + m_solveWCompNewtonNoT (WC/Solver.cpp:192-276) over three steps."""
    mesh, z = load_golden(name)
    dim, nn = mesh.dim, mesh.n_nodes
    st, x = split_wc(z["q0"], dim, nn), mesh.x
    st = {k: np.ascontiguousarray(v) for k, v in st.items()}
    orc.set_facets(dim, z.get("facets"), float(z.get("gamma", 0.0)))
    for step in range(z["dts"].shape[0]):
        dt = orc.wc_next_dt(mesh, x, st, z["wpar"], float(z["security_coeff"]), float(z["max_dt"]))
        assert abs(dt - z["dts"][step]) <= 1e-13 * z["dts"][step]
        x, st = orc.wc_step(mesh, x, st, z["wpar"], float(z["dts"][step]))
        want = split_wc(z["states"][step], dim, nn)
        for k in ("v", "p", "rho", "acc"):
            assert rel_err(st[k], want[k]) < TOL * 10 ** step, (k, step)
        assert np.abs(x - z["xs"][step]).max() < 1e-13
    orc.set_facets(dim)
    if "facets" in z:                                    # the fixture must actually exercise the facet terms
        st0 = {k: np.ascontiguousarray(v) for k, v in split_wc(z["q0"], dim, nn).items()}
        _, st1 = orc.wc_step(mesh, mesh.x, st0, z["wpar"], float(z["dts"][0]))
        assert rel_err(st1["acc"], split_wc(z["states"][0], dim, nn)["acc"]) > 1e-7


@pytest.mark.parametrize("name", golden_names("wcb_"))
def test_boussinesq_wc_oracle_matches_reference_fixture(name):
    """BoussinesqWC: HeatEqWCompNewton (explicit, lumped), buoyancy factor in the momentum body force, thermal diffusivity
    in computeNextDT, order heat -> continuity -> momentum of m_solveBoussinesqWC (WC/Solver.cpp:278-320).  Oracle level
    only so far: the CUDA path for this row is the next round's work."""
    mesh, z = load_golden(name)
    dim, nn = mesh.dim, mesh.n_nodes
    k, cv, alpha, Tr = z["thermal"]
    th = dict(k=k, cv=cv, alpha=alpha, Tr=Tr, t_mask=np.ascontiguousarray(z["t_mask"]), t_val=np.ascontiguousarray(z["t_val"]))
    q0 = z["q0"]
    st = {kk: np.ascontiguousarray(v) for kk, v in split_wc(q0[: (2 * dim + 2) * nn], dim, nn).items()}
    st["T"] = np.ascontiguousarray(q0[(2 * dim + 2) * nn:])
    x = mesh.x
    for step in range(z["dts"].shape[0]):
        dt = orc.wc_next_dt(mesh, x, st, z["wpar"], float(z["security_coeff"]), float(z["max_dt"]), thermal=th)
        assert abs(dt - z["dts"][step]) <= 1e-13 * z["dts"][step]
        x, st = orc.wc_step(mesh, x, st, z["wpar"], float(z["dts"][step]), thermal=th)
        q = z["states"][step]
        want = split_wc(q[: (2 * dim + 2) * nn], dim, nn)
        for kk in ("v", "p", "rho", "acc"):
            assert rel_err(st[kk], want[kk]) < TOL * 10 ** step, (kk, step)
        assert rel_err(st["T"], q[(2 * dim + 2) * nn:]) < TOL
        assert np.abs(x - z["xs"][step]).max() < 1e-13
    # the fixture exercises both couplings: conduction changes interior temperatures, buoyancy changes the acceleration
    st0 = {kk: np.ascontiguousarray(v) for kk, v in split_wc(q0[: (2 * dim + 2) * nn], dim, nn).items()}
    st0["T"] = np.ascontiguousarray(q0[(2 * dim + 2) * nn:])
    _, s_nob = orc.wc_step(mesh, mesh.x, st0, z["wpar"], float(z["dts"][0]), thermal=dict(th, alpha=0.0))
    _, s_b = orc.wc_step(mesh, mesh.x, st0, z["wpar"], float(z["dts"][0]), thermal=th)
    assert rel_err(s_nob["acc"], s_b["acc"]) > 1e-3
    interior = (z["t_mask"] == 0) & ((mesh.flags & mg.F_FREE) == 0)
    assert np.abs(s_b["T"][interior] - st0["T"][interior]).max() > 1e-3


@pytest.mark.parametrize("name", golden_names("pspgb_"))
def test_bingham_oracle_matches_reference_fixture(name):
    """Problem id "Bingham" (MomContEquation.inl:102-119): oracle level only so far, CUDA path next round."""
    mesh, z = load_golden(name)
    dim, nn = mesh.dim, mesh.n_nodes
    vcur = z["q"][: dim * nn].copy()
    orc.set_bingham(*z["bingham"])
    try:
        A, b = orc.pspg_build(mesh, vcur, z["q_prev"], z["par"], True)
    finally:
        orc.set_bingham()
    A_ref = golden_csc(z, "A")
    assert max(block_errors(A, A_ref, nn, dim).values()) < TOL
    assert max(vec_block_errors(b, z["b"], nn, dim).values()) < TOL
    A0, _ = orc.pspg_build(mesh, vcur, z["q_prev"], z["par"], True)       # Newtonian: the vv blocks must differ visibly
    assert block_errors(A0, A_ref, nn, dim)["vv"] > 1e-3


@pytest.mark.parametrize("name", golden_names("inb_"))
def test_incompressible_boussinesq_oracle_matches_reference_fixture(name):
    """Problem id "Boussinesq": buoyancy factors in the PSPG right-hand side and the implicit heat system (oracle level)."""
    import scipy.sparse as sp
    mesh, z = load_golden(name)
    dim, nn = mesh.dim, mesh.n_nodes
    alpha, Tr, k, cv = z["thermal"]
    vcur = z["q"][: dim * nn].copy()
    orc.set_pspg_thermal(alpha, Tr, z["T"])
    try:
        A, b = orc.pspg_build(mesh, vcur, z["q_prev"], z["par"], True)
    finally:
        orc.set_pspg_thermal()
    assert max(block_errors(A, golden_csc(z, "A"), nn, dim).values()) < TOL
    assert max(vec_block_errors(b, z["b"], nn, dim).values()) < TOL
    _, b0 = orc.pspg_build(mesh, vcur, z["q_prev"], z["par"], True)
    assert max(vec_block_errors(b0, z["b"], nn, dim).values()) > 1e-4        # buoyancy is visible
    Ah, bh = orc.in_heat_build(mesh, z["T"], z["par"][0], cv, k, z["par"][2], z["t_mask"], z["t_val"], True)
    Ah_ref = sp.csc_matrix((z["h_A"], z["h_indices"], z["h_indptr"]), shape=(nn, nn))
    assert (Ah.indptr == Ah_ref.indptr).all() and (Ah.indices == Ah_ref.indices).all()
    assert rel_err(Ah.data, Ah_ref.data) < TOL and rel_err(bh, z["h_b"]) < TOL


# ---- live comparisons (development container only: needs oracle/_ref/libpfem_ref.so) ----------------------------------
needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (no /root/reference here)")


@needs_ref
@pytest.mark.parametrize("dim,n,kw", [(2, 14, dict(free_fraction=0.02, permute=True)), (3, 6, dict(free_fraction=0.01)),
                                      (3, 5, dict(permute=True)), (2, 9, dict())])
def test_live_pspg_assembly(dim, n, kw):
    mesh, q, q_prev, par = pspg_case(dim, n, **kw)
    nn = mesh.n_nodes
    with ref.RefCase(mesh, "pspg", par) as rc:
        rc.set_states(q)
        A_ref, b_ref = rc.pspg_build(q_prev, True)
        em = rc.element_matrices()
    A, b = orc.pspg_build(mesh, q[: dim * nn].copy(), q_prev, par, True)
    assert max(block_errors(A, A_ref, nn, dim).values()) < TOL
    assert max(vec_block_errors(b, b_ref, nn, dim).values()) < TOL
    # algebraic invariants on the REFERENCE'S element matrices: sum(M) = rho*V, K.translation = 0, L.1 = 0
    V = mg.det_j(mesh) * (0.5 if dim == 2 else 1 / 6)
    assert np.abs(em["M"].sum((1, 2)) - par[0] * V).max() < 1e-12 * np.abs(par[0] * V).max()
    t = np.zeros(dim * (dim + 1)); t[: dim + 1] = 1.0
    assert np.abs(em["K"] @ t).max() < 1e-9 * np.abs(em["K"]).max()
    assert np.abs(em["L"].sum(2)).max() < 1e-9 * np.abs(em["L"]).max()


@needs_ref
@pytest.mark.parametrize("dim,npts", [(2, 150), (3, 200)])
def test_live_pspg_unstructured(dim, npts):
    mesh = mg.delaunay_cloud(dim, npts, free_fraction=0.02, seed=21)
    q, q_prev = mg.pspg_state(mesh)
    P = mg.PSPG_PARAMS
    par = orc.pspg_param_array(P["rho"], P["mu"], P["dt"], mg.gravity(dim))
    nn = mesh.n_nodes
    with ref.RefCase(mesh, "pspg", par) as rc:
        rc.set_states(q)
        A_ref, b_ref = rc.pspg_build(q_prev, True)
    A, b = orc.pspg_build(mesh, q[: dim * nn].copy(), q_prev, par, True)
    assert max(block_errors(A, A_ref, nn, dim).values()) < TOL
    assert max(vec_block_errors(b, b_ref, nn, dim).values()) < TOL


@needs_ref
@pytest.mark.parametrize("dim,n,eq,meduri", [(2, 12, "CDS_dpdt", True), (3, 5, "CDS_dpdt", False), (3, 5, "CDS_drhodt", True),
                                             (2, 12, "CDS_rho", False)])
def test_live_wc_steps(dim, n, eq, meduri):
    mesh = mg.kuhn_box(dim, n, free_fraction=0.02, permute=True)
    st = mg.wc_state(mesh)
    st["acc"] = 0.3 * np.random.default_rng(12).standard_normal(st["acc"].shape)
    W = mg.WC_PARAMS
    wpar = orc.wc_param_array(W["mu"], W["K0"], W["K0p"], W["rhoStar"], mg.gravity(dim), meduri, eq)
    nn, x = mesh.n_nodes, mesh.x
    with ref.RefCase(mesh, "wc", np.concatenate([wpar, [1e-6, 1e-3, W["securityCoeff"]]])) as rc:
        rc.set_states(np.concatenate([st["v"], st["p"], st["rho"], st["acc"]]))
        for step in range(4):
            dt_ref = rc.wc_next_dt()
            dt = orc.wc_next_dt(mesh, x, st, wpar, W["securityCoeff"], 1e-3)
            assert abs(dt - dt_ref) <= 1e-13 * dt_ref
            assert rc.wc_step(dt_ref)
            x, st = orc.wc_step(mesh, x, st, wpar, dt_ref)
            want = split_wc(rc.get_states(), dim, nn)
            for k in ("v", "p", "rho", "acc"):
                assert rel_err(st[k], want[k]) < TOL * 10 ** step, (k, step)
            assert np.abs(x - rc.positions()).max() < 1e-13


@needs_ref
def test_live_wc_forty_steps_do_not_drift():
    """40 explicit steps with the CFL time step of each side: the oracle and the reference's own code stay bit-identical
    (states, positions, every dt), so the per-step agreement is not an accident of the first steps."""
    dim = 2
    mesh = mg.kuhn_box(dim, 10, free_fraction=0.01, permute=True)
    st = mg.wc_state(mesh)
    W = mg.WC_PARAMS
    wpar = orc.wc_param_array(W["mu"], W["K0"], W["K0p"], W["rhoStar"], mg.gravity(dim), True, "CDS_dpdt")
    nn, x = mesh.n_nodes, mesh.x
    with ref.RefCase(mesh, "wc", np.concatenate([wpar, [1e-6, 1e-3, W["securityCoeff"]]])) as rc:
        rc.set_states(np.concatenate([st["v"], st["p"], st["rho"], st["acc"]]))
        for _ in range(40):
            dt_ref = rc.wc_next_dt()
            dt = orc.wc_next_dt(mesh, x, st, wpar, W["securityCoeff"], 1e-3)
            assert dt == dt_ref
            assert rc.wc_step(dt_ref)
            x, st = orc.wc_step(mesh, x, st, wpar, dt)
        want = split_wc(rc.get_states(), dim, nn)
        for k in ("v", "p", "rho", "acc"):
            assert np.array_equal(st[k], want[k]), k
        assert np.array_equal(x, rc.positions())
        assert np.abs(want["v"]).max() > 1e-3          # the column has started to move


@needs_ref
def test_live_pspg_five_time_steps():
    """Five consecutive PSPG time steps (MomContEqIncompNewton::solve = Picard loop each, states and moved positions carried
    over, no remeshing) on a C1-like 2-D column: oracle loop vs the reference's own code, same direct solver behind both."""
    dim = 2
    mesh = mg.kuhn_box(dim, 10)
    _, q = mg.pspg_state(mesh)
    P = mg.PSPG_PARAMS
    par = orc.pspg_param_array(P["rho"], P["mu"], P["dt"], mg.gravity(dim))
    nn = mesh.n_nodes
    ref.use_scipy_direct_solver(True)
    try:
        with ref.RefCase(mesh, "pspg", np.concatenate([par, [10, 1e-6]])) as rc:
            rc.set_states(q)
            x = mesh.x.copy()
            for step in range(5):
                ok, iters = rc.pspg_solve()
                moved = mg.Mesh(dim=dim, x=x, conn=mesh.conn, flags=mesh.flags, dir_mask=mesh.dir_mask, dir_val=mesh.dir_val)
                out = orc.pspg_picard(moved, q, q, par, max_iter=10, min_res=1e-6)
                assert ok and out["ok"] and iters == out["iters"], step
                q, x = out["q"], out["x"]
                assert rel_err(q[: dim * nn], rc.get_states()[: dim * nn]) < 1e-8, step
                assert rel_err(q[dim * nn:], rc.get_states()[dim * nn:]) < 1e-8, step
                assert np.abs(x - rc.positions()).max() < 1e-11, step
            assert np.abs(x - mesh.x).max() > 1e-7          # the mesh has moved over the five steps
    finally:
        ref.use_scipy_direct_solver(False)


@needs_ref
def test_live_picard_dense_lu():
    """Same Picard loop with the stand-in's own dense LU instead of SuperLU: the direct solver does not matter at 1e-8."""
    mesh = mg.kuhn_box(2, 8)
    _, q_prev = mg.pspg_state(mesh)
    P = mg.PSPG_PARAMS
    par = orc.pspg_param_array(P["rho"], P["mu"], P["dt"], mg.gravity(2))
    ref.use_scipy_direct_solver(False)
    with ref.RefCase(mesh, "pspg", np.concatenate([par, [10, 1e-6]])) as rc:
        rc.set_states(q_prev)
        ok, iters = rc.pspg_solve()
        q = rc.get_states()
    out = orc.pspg_picard(mesh, q_prev, q_prev, par, max_iter=10, min_res=1e-6)
    assert ok and out["ok"] and iters == out["iters"]
    assert rel_err(out["q"], q) < 1e-8


@needs_ref
@pytest.mark.parametrize("dim,n", [(2, 12), (3, 5)])
def test_live_fracstep_pressure_solve_does_not_converge(dim, n):
    """Evidence for DESIGN.md section 7: the reference's own fractional-step code (MomContEquationFracStep.inl, run here
    through oracle/_ref) hands the pressure-correction matrix m_L to ConjugateGradient without any Dirichlet row for the
    free-surface nodes (only isFree rows are masked, :124-130, 349-376), so the system is singular with an inconsistent
    right-hand side: the stand-in CG (Jacobi-preconditioned, tolerance eps, 2n iterations -- Eigen's defaults as recalled)
    converges on the two velocity systems and hits the iteration cap with a large residual on every pressure solve."""
    mesh = mg.kuhn_box(dim, n)
    _, q_prev = mg.pspg_state(mesh)
    P = mg.PSPG_PARAMS
    par = orc.pspg_param_array(P["rho"], P["mu"], P["dt"], mg.gravity(dim))
    ref.RefCase.cg_log()
    with ref.RefCase(mesh, "pspg", np.concatenate([par, [10, 1e-6, 1.0, 2.0]]), solver_id="FracStep") as rc:
        rc.set_states(q_prev)
        rc.pspg_solve()
    log = ref.RefCase.cg_log()
    vel = log[log[:, 0] == dim * mesh.n_nodes]
    prs = log[log[:, 0] == mesh.n_nodes]
    assert len(vel) >= 2 and len(prs) >= 1
    assert (vel[:2, 3] == 0).all() and (vel[:2, 2] < 1e-14).all()          # first velocity systems: converged to eps
    assert (prs[:, 3] == 2).all() and (prs[:, 1] == 2 * mesh.n_nodes).all() # pressure: NoConvergence at 2n iterations
    assert np.nanmax(prs[:, 2]) > 1e-6 and prs[0, 2] > 1e-6


@needs_ref
@pytest.mark.parametrize("seed", range(10))
def test_live_randomised_cases(seed):
    """Seeded random cases: dimension, mesh family, size, free-node fraction, numbering, Dirichlet mask and data, material
    constants, continuity variant and stabilisation all drawn at random; assembly (+BC) and two explicit steps must agree
    with the reference's own code to 1e-12."""
    rng = np.random.default_rng(1000 + seed)
    dim = int(rng.integers(2, 4))
    if rng.random() < 0.5:
        mesh = mg.kuhn_box(dim, int(rng.integers(3, 7 if dim == 3 else 12)), free_fraction=float(rng.choice([0.0, 0.02, 0.1])),
                           permute=bool(rng.integers(0, 2)), jitter=float(rng.uniform(0.0, 0.25)), seed=int(rng.integers(1, 10 ** 6)))
    else:
        mesh = mg.delaunay_cloud(dim, int(rng.integers(60, 220)), seed=int(rng.integers(1, 10 ** 6)),
                                 free_fraction=float(rng.choice([0.0, 0.05])))
    nn = mesh.n_nodes
    bound = np.flatnonzero(mesh.flags & mg.F_BOUND)
    mesh.dir_mask[:] = 0
    mesh.dir_mask[bound[rng.random(bound.size) < 0.8]] = 1          # some bound nodes carry no velocity BC
    mesh.dir_val = np.ascontiguousarray((0.2 * rng.standard_normal((dim, nn)) * (mesh.dir_mask != 0)).reshape(-1))
    rho, mu, dt = float(10 ** rng.uniform(2, 3.3)), float(10 ** rng.uniform(-4, 0)), float(10 ** rng.uniform(-4, -2))
    g = np.zeros(3); g[:dim] = rng.uniform(-10, 10, dim)
    par = orc.pspg_param_array(rho, mu, dt, g)
    v = rng.standard_normal(dim * nn)
    q = np.concatenate([v, 1e3 * rng.standard_normal(nn)])
    q_prev = q + 0.1 * rng.standard_normal(q.shape)
    with ref.RefCase(mesh, "pspg", par) as rc:
        rc.set_states(q)
        A_ref, b_ref = rc.pspg_build(q_prev, True)
    A, b = orc.pspg_build(mesh, q[: dim * nn].copy(), q_prev, par, True)
    assert max(block_errors(A, A_ref, nn, dim).values()) < TOL
    assert max(vec_block_errors(b, b_ref, nn, dim).values()) < TOL

    eq = str(rng.choice(["CDS_dpdt", "CDS_drhodt", "CDS_rho"]))
    meduri = bool(rng.integers(0, 2))
    W = mg.WC_PARAMS
    wpar = orc.wc_param_array(mu, W["K0"] * float(rng.uniform(0.5, 2)), W["K0p"], W["rhoStar"], g, meduri, eq)
    st = mg.wc_state(mesh)
    st["v"] = 0.3 * rng.standard_normal(dim * nn)
    st["acc"] = rng.standard_normal(dim * nn)
    x = mesh.x
    with ref.RefCase(mesh, "wc", np.concatenate([wpar, [1e-6, 1e-3, 0.1]])) as rc:
        rc.set_states(np.concatenate([st["v"], st["p"], st["rho"], st["acc"]]))
        for step in range(2):
            dt_ref = rc.wc_next_dt()
            assert abs(orc.wc_next_dt(mesh, x, st, wpar, 0.1, 1e-3) - dt_ref) <= 1e-13 * dt_ref
            assert rc.wc_step(dt_ref)
            x, st = orc.wc_step(mesh, x, st, wpar, dt_ref)
            want = split_wc(rc.get_states(), dim, nn)
            for k in ("v", "p", "rho", "acc"):
                assert rel_err(st[k], want[k]) < TOL * 10 ** step, (k, step, eq, meduri)
