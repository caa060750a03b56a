"""Drop-in test of the boundary: the REFERENCE'S OWN host code -- Mesh, SolverIncompNewton /
SolverWCompNewton, PicardAlgo, compiled from /root/reference into oracle/_ref/libpfem_ref_dropin.so (oracle/refbuild) --
runs one time step twice: once with its own equation classes (CPU, stand-in Eigen, SciPy SuperLU behind SparseLU) and
once with the shim classes of shim/pfem_b200_equations.hpp swapped in at the REGISTER_EQ seam, which call
libpfem_b200.so through the C ABI.  Node states and positions on the reference's Mesh object must agree: 1e-8 for the
PSPG step (north_star field tolerance), 1e-12 for the explicit steps."""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import ref
from pfem_b200 import meshgen as mg

from helpers import rel_err, split_wc, wavy_free_surface

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _dropin_library():
    if not ref.dropin_available():
        pytest.skip("oracle/_ref/libpfem_ref_dropin.so was not built (needs /root/reference at build time)")
    ref.use_dropin_library()
    ref.use_scipy_direct_solver(True)
    yield
    ref.use_scipy_direct_solver(False)


@pytest.mark.parametrize("dim,n,gamma,permute", [(2, 12, 0.0, False), (3, 5, 0.0, True), (3, 4, 7.28, True)])
def test_pspg_time_step_through_the_shim(dim, n, gamma, permute):
    """permute: random node/element numbering on the reference side; the shim uploads in Morton order (Renumbering)."""
    mesh = mg.kuhn_box(dim, n, permute=permute)
    facets = None
    if gamma > 0:
        wavy_free_surface(mesh)
        facets = mg.boundary_facets(mesh)
    _, q_prev = mg.pspg_state(mesh)
    P = mg.PSPG_PARAMS
    par = np.concatenate([orc.pspg_param_array(P["rho"], P["mu"], P["dt"], mg.gravity(dim)), [10, 1e-6]])
    nn = mesh.n_nodes
    out = {}
    for which in ("reference", "b200"):
        with ref.RefCase(mesh, "pspg", par, facets=facets, gamma=gamma) as rc:
            if which == "b200":
                rc.use_b200_equation()
            rc.set_states(q_prev)
            ok, _ = rc.pspg_solve()             # Equation::solve() of whichever class sits in m_pEquations[0]
            assert ok
            out[which] = (rc.get_states(), rc.positions())
    (q_ref, x_ref), (q, x) = out["reference"], out["b200"]
    assert rel_err(q[: dim * nn], q_ref[: dim * nn]) < 1e-8
    assert rel_err(q[dim * nn:], q_ref[dim * nn:]) < 1e-8
    assert np.abs(x - x_ref).max() < 1e-9
    assert np.abs(x_ref - mesh.x).max() > 1e-6  # the step moved the mesh


@pytest.mark.parametrize("dim,n,eq,gamma,permute", [(2, 10, "CDS_dpdt", 0.0, True), (3, 5, "CDS_drhodt", 0.0, False),
                                                     (3, 4, "CDS_dpdt", 7.28, True)])
def test_wc_steps_through_the_shim(dim, n, eq, gamma, permute):
    mesh = mg.kuhn_box(dim, n, free_fraction=0.02, permute=permute)
    facets = None
    if gamma > 0:
        wavy_free_surface(mesh)
        facets = mg.boundary_facets(mesh)
    st = mg.wc_state(mesh)
    st["acc"] = 0.3 * np.random.default_rng(2).standard_normal(st["acc"].shape)
    W = mg.WC_PARAMS
    wpar = np.concatenate([orc.wc_param_array(W["mu"], W["K0"], W["K0p"], W["rhoStar"], mg.gravity(dim), True, eq),
                           [1e-6, 1e-3, W["securityCoeff"]]])
    q0 = np.concatenate([st["v"], st["p"], st["rho"], st["acc"]])
    nn = mesh.n_nodes
    with ref.RefCase(mesh, "wc", wpar, facets=facets, gamma=gamma) as a, ref.RefCase(mesh, "wc", wpar, facets=facets, gamma=gamma) as b:
        a.set_states(q0)
        b.set_states(q0)
        dt = a.wc_next_dt()                      # SolverWCompNewton::computeNextDT
        for step in range(3):
            assert a.wc_step(dt)                 # reference: m_solveWCompNewtonNoT
            assert b.wc_step_b200(dt)            # shim: WCompNewtonStepB200::step + download to the reference Mesh
            want, got = split_wc(a.get_states(), dim, nn), split_wc(b.get_states(), dim, nn)
            for k in ("v", "p", "rho", "acc"):
                assert rel_err(got[k], want[k]) < 1e-12 * 10 ** step, (k, step)
            assert np.abs(a.positions() - b.positions()).max() < 1e-13
            dt_ref, dt = a.wc_next_dt(), b.wc_next_dt_b200()
            assert abs(dt - dt_ref) <= 1e-13 * dt_ref


def _wc_pair(dim, n, eq, steps, check):
    mesh = mg.kuhn_box(dim, n, free_fraction=0.02, permute=True)
    st = mg.wc_state(mesh)
    st["acc"] = 0.3 * np.random.default_rng(2).standard_normal(st["acc"].shape)
    # non-zero Dirichlet data on part of the walls (hazard 10: it becomes the boundary nodes' "acceleration")
    vals = mesh.dir_val.reshape(dim, mesh.n_nodes)
    sel = (mesh.dir_mask != 0) & (np.arange(mesh.n_nodes) % 2 == 0)
    vals[:, sel] = 0.05 * np.random.default_rng(8).standard_normal((dim, int(sel.sum())))
    mesh.dir_val = np.ascontiguousarray(vals.reshape(-1))
    W = mg.WC_PARAMS
    wpar = np.concatenate([orc.wc_param_array(W["mu"], W["K0"], W["K0p"], W["rhoStar"], mg.gravity(dim), True, eq),
                           [1e-6, 1e-3, W["securityCoeff"]]])
    q0 = np.concatenate([st["v"], st["p"], st["rho"], st["acc"]])
    nn = mesh.n_nodes
    with ref.RefCase(mesh, "wc", wpar) as a, ref.RefCase(mesh, "wc", wpar) as b:
        a.set_states(q0)
        b.set_states(q0)
        dt = a.wc_next_dt()
        for step in range(steps):
            assert a.wc_step(dt) and b.wc_step_b200(dt)
            want, got = split_wc(a.get_states(), dim, nn), split_wc(b.get_states(), dim, nn)
            for k in ("v", "p", "rho", "acc"):
                assert rel_err(got[k], want[k]) < 1e-12 * 10 ** step, (k, step)
            assert np.abs(a.positions() - b.positions()).max() < 1e-13
            dt_ref, dt = a.wc_next_dt(), b.wc_next_dt_b200()
            assert abs(dt - dt_ref) <= 1e-13 * dt_ref
            check(step, want)


def test_wc_time_dependent_dirichlet_table_through_the_shim():
    """The reference evaluates "<type>V"(pos, t + dt) for every bound node on EVERY explicit step
    (WCompNewton/MomEquation.inl:355-371); the shim must refresh the device table per step, not per remesh."""
    seen = []
    ref.set_bc_ramp(2.0e5)  # dt ~ 1e-6: the table changes by tens of percent from step to step
    try:
        _wc_pair(3, 5, "CDS_dpdt", 4, lambda step, want: seen.append(np.abs(want["acc"]).max()))
    finally:
        ref.set_bc_ramp(0.0)
    assert len(set(seen)) > 1


@pytest.mark.parametrize("devices", ["0,0", "0,0,0"])
def test_wc_steps_through_the_shim_on_several_ranks(devices, monkeypatch):
    """PFEM_DEVICES lists several devices: the shim partitions the reference's mesh with pfem_partition_*, drives one
    context per entry from its own thread (pfem_comm_local_*) and gathers the owned nodes back -- the reference's single
    host process on N GPUs.  Here the entries name the same GPU."""
    monkeypatch.setenv("PFEM_DEVICES", devices)
    _wc_pair(3, 6, "CDS_dpdt", 3, lambda step, want: None)


@pytest.mark.parametrize("dim,n", [(2, 8), (3, 4)])
def test_bingham_time_step_through_the_shim(dim, n):
    """Problem id "Bingham" (SURVEY 8f rank 3): the shim reads tau0 / mReg from the Material table (MomContEquation.inl:54-58)
    and the library assembles with the shear-rate dependent viscosity; one PicardAlgo time step on both sides."""
    mesh = mg.kuhn_box(dim, n, free_fraction=0.02, permute=True)
    _, q_prev = mg.pspg_state(mesh)
    P = mg.PSPG_PARAMS
    par = np.concatenate([orc.pspg_param_array(P["rho"], P["mu"], P["dt"], mg.gravity(dim)), [10, 1e-6]])
    nn = mesh.n_nodes
    out = {}
    for which, bingham in (("newtonian", None), ("reference", (50.0, 100.0)), ("b200", (50.0, 100.0))):
        with ref.RefCase(mesh, "pspg", par, bingham=bingham) as rc:
            if which == "b200":
                rc.use_b200_equation()
            rc.set_states(q_prev)
            ok, _ = rc.pspg_solve()
            assert ok
            out[which] = rc.get_states()
    assert rel_err(out["b200"][: dim * nn], out["reference"][: dim * nn]) < 1e-8
    assert rel_err(out["b200"][dim * nn:], out["reference"][dim * nn:]) < 1e-8
    assert rel_err(out["reference"][: dim * nn], out["newtonian"][: dim * nn]) > 1e-4  # the yield stress is live


@pytest.mark.parametrize("dim,n", [(2, 8), (3, 4)])
def test_boussinesq_wc_steps_through_the_shim(dim, n):
    """Problem id "BoussinesqWC": the shim sets the thermal factors from the Material table, uploads the temperature
    state 2 dim + 2 and evaluates the heat equation's "<type>T" table (HeatEquation.inl:226-244); the reference runs
    m_solveBoussinesqWC (WCompNewton/Solver.cpp:278-320)."""
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
    W = mg.WC_PARAMS
    wpar = np.concatenate([orc.wc_param_array(W["mu"], W["K0"], W["K0p"], W["rhoStar"], mg.gravity(dim), True, "CDS_dpdt"),
                           [1e-6, 1e-3, W["securityCoeff"]]])
    q0 = np.concatenate([st["v"], st["p"], st["rho"], st["acc"], T0])
    with ref.RefCase(mesh, "wc", wpar, thermal=th) as a, ref.RefCase(mesh, "wc", wpar, thermal=th) as b:
        a.set_states(q0)
        b.set_states(q0)
        dt = a.wc_next_dt()
        for step in range(3):
            assert a.wc_step(dt) and b.wc_step_b200(dt)
            qa, qb = a.get_states(), b.get_states()
            want, got = split_wc(qa[: (2 * dim + 2) * nn], dim, nn), split_wc(qb[: (2 * dim + 2) * nn], dim, nn)
            for k in ("v", "p", "rho", "acc"):
                assert rel_err(got[k], want[k]) < 1e-12 * 10 ** step, (k, step)
            Ta, Tb = qa[(2 * dim + 2) * nn:], qb[(2 * dim + 2) * nn:]
            assert rel_err(Tb, Ta) < 1e-13
            assert np.array_equal(Ta[t_mask != 0], t_val[t_mask != 0])
            assert np.abs(a.positions() - b.positions()).max() < 1e-13
            dt_ref, dt = a.wc_next_dt(), b.wc_next_dt_b200()
            assert abs(dt - dt_ref) <= 1e-13 * dt_ref
        assert np.abs(Ta - T0).max() > 1e-3  # conduction moved the temperature


@pytest.mark.parametrize("dim,n", [(2, 8), (3, 4)])
def test_boussinesq_time_step_through_the_shim(dim, n):
    """Problem id "Boussinesq" (incompressible): SolverIncompNewton::m_solveBoussinesq with solveHeatFirst = the heat
    equation's solve() then the momentum-continuity solve() (IN/Solver.cpp:249-264), once with the reference's
    HeatEqIncompNewton + MomContEqIncompNewton and once with HeatEqIncompNewtonB200 + MomContEqIncompNewtonB200."""
    mesh = mg.kuhn_box(dim, n, free_fraction=0.02, permute=True)
    nn = mesh.n_nodes
    _, q_prev = mg.pspg_state(mesh)
    c = mesh.coords()
    T0 = 300.0 + 10.0 * c[:, 0] + 2.0 * np.random.default_rng(5).standard_normal(nn)
    bound = (mesh.flags & mg.F_BOUND) != 0
    t_mask = (bound & ((np.abs(c[:, 0]) < 1e-12) | (np.abs(c[:, 0] - 1.0) < 1e-12))).astype(np.uint8)
    t_val = np.where(c[:, 0] < 0.5, 310.0, 290.0)
    P = mg.PSPG_PARAMS
    par = np.concatenate([orc.pspg_param_array(P["rho"], P["mu"], P["dt"], mg.gravity(dim)), [10, 1e-6]])
    out = {}
    for which, alpha in (("no_buoyancy", 0.0), ("reference", 6.9e-3), ("b200", 6.9e-3)):
        th = dict(alpha=alpha, Tr=300.0, k=6.0e4, cv=4.186, t_mask=t_mask, t_val=t_val)  # conduction exaggerated: one step shows it
        with ref.RefCase(mesh, "pspg", par, thermal=th) as rc:
            if which == "b200":
                rc.use_b200_equation()
            rc.set_states(np.concatenate([q_prev, T0]))
            assert rc.heat_solve()
            ok, _ = rc.pspg_solve()
            assert ok
            out[which] = (rc.get_states(0, dim + 2), rc.positions())
    (q_ref, x_ref), (q, x) = out["reference"], out["b200"]
    assert rel_err(q[: dim * nn], q_ref[: dim * nn]) < 1e-8
    assert rel_err(q[dim * nn: (dim + 1) * nn], q_ref[dim * nn: (dim + 1) * nn]) < 1e-8
    T_ref, T = q_ref[(dim + 1) * nn:], q[(dim + 1) * nn:]
    assert rel_err(T, T_ref) < 1e-9
    assert np.allclose(T[t_mask != 0], t_val[t_mask != 0], rtol=1e-12, atol=0)   # rows reduced to their diagonal, CG to 1e-12
    assert np.abs(T_ref - T0).max() > 1e-2                       # the heat equation moved the temperature
    assert np.abs(x - x_ref).max() < 1e-9
    q_nob = out["no_buoyancy"][0]
    assert rel_err(q_ref[: dim * nn], q_nob[: dim * nn]) > 1e-6  # the buoyancy factor is live
