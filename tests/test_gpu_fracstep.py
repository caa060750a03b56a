"""SURVEY 8(f) rank 1 on the GPU: the three linear systems of one Picard body of the fractional-step solver
(MomContEquationFracStep.inl) against fixtures made by the reference's own code (tests/golden/make_fracstep.py), and Eigen's
Jacobi-preconditioned conjugate gradients on the two velocity systems against the stand-in ConjugateGradient's solution and
iteration count.  The pressure system is singular as the reference builds it: matrix and right-hand
side are compared, and its solve must stop at the iteration cap like the reference's."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import golden_names, load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _csc(z, prefix):
    n = z[prefix + "_indptr"].shape[0] - 1
    return sp.csc_matrix((z[prefix + "_data"], z[prefix + "_indices"], z[prefix + "_indptr"]), shape=(n, n))


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _mat_rel(A, B):
    D = (A - B).tocoo()
    return float(np.abs(D.data).max() / np.abs(B.data).max()) if D.nnz else 0.0


@pytest.mark.parametrize("name", golden_names("fs_"))
def test_fractional_step_systems_match_reference_fixture(gpu_ctx_factory, name):
    mesh, z = load_golden(name)
    dim, nn = mesh.dim, mesh.n_nodes
    rho, mu, dt = [float(v) for v in z["par"][:3]]
    gfs = float(z["gamma_fs"])
    eps = np.finfo(float).eps
    with gpu_ctx_factory(dim) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, z["q_prev"])
        par = ctx.pspg_params(rho, mu, dt, z["par"][3:6])

        # 1: velocity prediction (m_buildMatFracStep + m_applyBCVAppStep)
        ctx.fs_assemble_vapp(par, gfs, z["q_prev"])
        A, b = ctx.pspg_export_csc()
        A = A.tocsr()
        nv = dim * nn
        assert _mat_rel(A[:nv, :nv].tocsc(), _csc(z, "A0")) < TOL
        assert _rel(b[:nv], z["b0"]) < TOL
        assert abs(A[:nv, nv:]).sum() == 0 and abs(A[nv:, :nv]).sum() == 0                     # no (v, p) coupling in this system
        assert (abs(A[nv:, nv:] - sp.identity(nn)).sum() == 0) and np.all(b[nv:] == 0)           # identity pressure rows
        s0 = ctx.fs_solve(0, eps, 2 * nv)
        assert s0["status"] == 0 and abs(s0["iters"] - int(z["cg0"][0])) <= 2, (s0["iters"], z["cg0"])
        assert _rel(s0["x"], z["v_tilde"]) < 1e-10

        # 2: pressure (m_buildMatPcorrStep + m_applyBCPCorrStep) from the reference's vTilde
        ctx.fs_assemble_pcorr(rho, dt, gfs, z["v_tilde"], z["q_prev"][nv:])
        L, bp = ctx.heat_export_csc()
        assert _mat_rel(L, _csc(z, "A1")) < TOL
        assert _rel(ctx.fs_get_rhs(1), z["b1"]) < TOL and np.array_equal(bp, ctx.fs_get_rhs(1))
        s1 = ctx.fs_solve(1, eps, 2 * nn)
        assert s1["status"] == 1 and s1["iters"] == 2 * nn and int(z["cg1"][2]) == 2             # NoConvergence on both sides

        # 3: velocity correction (m_buildMatVStep + m_applyBCVStep)
        ctx.fs_assemble_vcorr(rho, dt, z["delta_p"])
        M, _ = ctx.heat_export_csc()
        assert _mat_rel(sp.kron(sp.identity(dim), M).tocsc(), _csc(z, "A2")) < TOL
        assert _rel(ctx.fs_get_rhs(2), z["b2"]) < TOL
        s2 = ctx.fs_solve(2, eps, 2 * nv)
        assert s2["status"] == 0 and abs(s2["iters"] - int(z["cg2"][0])) <= 2, (s2["iters"], z["cg2"])
        assert _rel(s2["x"], z["dv"]) < 1e-10


def test_fractional_step_velocity_systems_at_size(gpu_ctx_factory):
    """The velocity systems at 41 k tets: CG converges to machine epsilon in a mesh-independent number of iterations (mass-
    dominated at the dam-break time step), the solutions satisfy their systems, and the block system also goes through the
    multigrid-preconditioned solver of the PSPG path (same storage)."""
    from pfem_b200 import meshgen as mg
    mesh = mg.kuhn_box(3, 19, free_fraction=0.002, permute=True)
    dim, nn = 3, mesh.n_nodes
    _, q_prev = mg.pspg_state(mesh)
    P = mg.PSPG_PARAMS
    eps = np.finfo(float).eps
    with gpu_ctx_factory(dim) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, q_prev)
        par = ctx.pspg_params(P["rho"], P["mu"], P["dt"], mg.gravity(dim))
        ctx.fs_assemble_vapp(par, 1.0, q_prev)
        A, b = ctx.pspg_export_csc()
        s0 = ctx.fs_solve(0, eps, 6 * nn)
        assert s0["status"] == 0 and s0["iters"] < 200
        x = np.concatenate([s0["x"], np.zeros(nn)])
        assert np.linalg.norm(A @ x - b) <= 1e-13 * np.linalg.norm(b)
        sol = ctx.pspg_solve(1e-13, 2000)
        assert sol["status"] == 0 and _rel(sol["q"][: dim * nn], s0["x"]) < 1e-9
        dp = 100.0 * mesh.coords()[:, 2]
        ctx.fs_assemble_vcorr(P["rho"], P["dt"], dp)
        M, _ = ctx.heat_export_csc()
        s2 = ctx.fs_solve(2, eps, 6 * nn)
        assert s2["status"] == 0 and s2["iters"] < 100
        r = sp.kron(sp.identity(dim), M) @ s2["x"] - ctx.fs_get_rhs(2)
        assert np.linalg.norm(r) <= 1e-13 * max(np.linalg.norm(ctx.fs_get_rhs(2)), 1e-300)


@pytest.mark.parametrize("seed", range(4))
def test_fractional_step_systems_match_oracle_on_seeded_meshes(gpu_ctx_factory, seed):
    """CUDA path against the numpy restatement (oracle/fracstep_numpy.py, pinned to the reference by tests/test_oracle_fracstep.py)
    on seeded Kuhn / Delaunay meshes with random numbering, Dirichlet data, viscosity, time step and gammaFS."""
    from oracle import fracstep_numpy as fs
    from pfem_b200 import meshgen as mg
    rng = np.random.default_rng(900 + seed)
    dim = 2 + seed % 2
    if seed < 2:
        mesh = mg.kuhn_box(dim, int(rng.integers(5, 9)), free_fraction=0.03, permute=True, jitter=0.15, seed=int(rng.integers(1, 10 ** 6)))
    else:
        mesh = mg.delaunay_cloud(dim, int(rng.integers(150, 300)), seed=int(rng.integers(1, 10 ** 6)), free_fraction=0.05)
    nn = mesh.n_nodes
    mesh.dir_val = np.ascontiguousarray((0.1 * rng.standard_normal((dim, nn)) * (mesh.dir_mask != 0)).reshape(-1))
    _, q_prev = mg.pspg_state(mesh)
    q_prev = q_prev + 0.05 * rng.standard_normal(q_prev.shape)
    rho, mu, dt = 1000.0, float(rng.choice([1e-3, 1.0])), float(rng.choice([1e-3, 1e-2]))
    body, g = mg.gravity(dim), float(rng.choice([0.0, 0.5, 1.0]))
    dp = 30.0 * rng.standard_normal(nn)
    nv = dim * nn
    eps = np.finfo(float).eps
    with gpu_ctx_factory(dim) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, q_prev)
        par = ctx.pspg_params(rho, mu, dt, body)
        ctx.fs_assemble_vapp(par, g, q_prev)
        A, b = ctx.pspg_export_csc()
        O0, c0 = fs.velocity_prediction(mesh, q_prev[:nv], q_prev[nv:], rho, mu, dt, body, g)
        assert _rel(A.toarray()[:nv, :nv], O0) < TOL and _rel(b[:nv], c0) < TOL
        x_ref, it_ref, _, ok = fs.conjugate_gradient(O0, c0)
        s0 = ctx.fs_solve(0, eps, 2 * nv)
        assert ok and s0["status"] == 0 and abs(s0["iters"] - it_ref) <= 2 and _rel(s0["x"], x_ref) < 1e-10
        ctx.fs_assemble_pcorr(rho, dt, g, x_ref, q_prev[nv:])
        L, bp = ctx.heat_export_csc()
        O1, c1 = fs.pressure(mesh, x_ref, q_prev[nv:], rho, mu, dt, body, g)
        assert _rel(L.toarray(), O1) < TOL and _rel(bp, c1) < TOL
        ctx.fs_assemble_vcorr(rho, dt, dp)
        M, _ = ctx.heat_export_csc()
        O2, c2 = fs.velocity_correction(mesh, dp, rho, mu, dt, body)
        assert _rel(sp.kron(sp.identity(dim), M).toarray(), O2) < TOL and _rel(ctx.fs_get_rhs(2), c2) < TOL
        x_ref, it_ref, _, ok = fs.conjugate_gradient(O2, c2)
        s2 = ctx.fs_solve(2, eps, 2 * nv)
        assert ok and s2["status"] == 0 and abs(s2["iters"] - it_ref) <= 2 and _rel(s2["x"], x_ref) < 1e-10
