"""SURVEY 8(f) rank 3 on the GPU, against fixtures made by the reference's own code (tests/golden/make_golden.py):
Bingham viscosity (pspgb_*), incompressible Boussinesq factors + implicit heat system (inb_*), BoussinesqWC explicit
steps with the heat equation, the buoyancy factor and the thermal CFL (wcb_*), and the Jacobi-CG heat solve."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from helpers import block_errors, golden_csc, golden_names, load_golden, vec_block_errors

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _pspg_ctx(gpu_ctx_factory, mesh, z):
    ctx = gpu_ctx_factory(mesh.dim)
    ctx.set_mesh(mesh)
    ctx.set_states(0, z["q"])
    rho, mu, dt = z["par"][:3]
    return ctx, ctx.pspg_params(rho, mu, dt, z["par"][3:6])


@pytest.mark.parametrize("name", golden_names("pspgb_"))
def test_bingham_assembly_matches_reference_fixture(gpu_ctx_factory, name):
    mesh, z = load_golden(name)
    ctx, par = _pspg_ctx(gpu_ctx_factory, mesh, z)
    with ctx:
        ctx.set_bingham(float(z["bingham"][0]), float(z["bingham"][1]))
        ctx.pspg_assemble(par, z["q_prev"])
        A, b = ctx.pspg_export_csc()
        A_ref = golden_csc(z)
        for k, e in block_errors(A, A_ref, mesh.n_nodes, mesh.dim).items():
            assert e < TOL, (k, e)
        for k, e in vec_block_errors(b, z["b"], mesh.n_nodes, mesh.dim).items():
            assert e < TOL, (k, e)
        # the general path feeds the same solver: fields of a direct solve to 1e-8
        sol = ctx.pspg_solve(1e-13, 5000)
        x_ref = spla.splu(A_ref.tocsc()).solve(z["b"])
        assert sol["status"] == 0
        nn, dim = mesh.n_nodes, mesh.dim
        for sl in (slice(0, dim * nn), slice(dim * nn, None)):
            assert np.abs(sol["q"][sl] - x_ref[sl]).max() <= 1e-8 * np.abs(x_ref[sl]).max()
        # switching the factor off restores the Newtonian matrix (different values)
        ctx.set_bingham(None)
        ctx.pspg_assemble(par, z["q_prev"])
        A0, _ = ctx.pspg_export_csc()
        assert np.abs(A0.data - A.data).max() > 1e-6 * np.abs(A.data).max()


@pytest.mark.parametrize("name", golden_names("inb_"))
def test_boussinesq_pspg_and_heat_system_match_reference_fixture(gpu_ctx_factory, name):
    mesh, z = load_golden(name)
    alpha, Tr, k, cv = [float(v) for v in z["thermal"]]
    ctx, par = _pspg_ctx(gpu_ctx_factory, mesh, z)
    rho, dt = float(z["par"][0]), float(z["par"][2])
    with ctx:
        ctx.set_thermal(k, cv, alpha, Tr)
        ctx.set_temperature(z["T"])
        ctx.set_temperature_bc(z["t_mask"], z["t_val"])
        ctx.pspg_assemble(par, z["q_prev"])
        A, b = ctx.pspg_export_csc()
        for kk, e in block_errors(A, golden_csc(z), mesh.n_nodes, mesh.dim).items():
            assert e < TOL, (kk, e)
        for kk, e in vec_block_errors(b, z["b"], mesh.n_nodes, mesh.dim).items():
            assert e < TOL, (kk, e)
        # implicit heat system: pattern identical, values and RHS to 1e-12
        ctx.heat_assemble(rho, cv, k, dt, z["T"])
        Ah, bh = ctx.heat_export_csc()
        assert (Ah.indptr == z["h_indptr"]).all() and (Ah.indices == z["h_indices"]).all()
        assert np.abs(Ah.data - z["h_A"]).max() <= TOL * np.abs(z["h_A"]).max()
        assert np.abs(bh - z["h_b"]).max() <= TOL * np.abs(z["h_b"]).max()
        # Jacobi-CG with the previous temperature as initial guess vs a direct solve of the reference's system
        import scipy.sparse as sp
        Aref = sp.csc_matrix((z["h_A"], z["h_indices"], z["h_indptr"]), shape=Ah.shape)
        T_ref = spla.splu(Aref).solve(z["h_b"])
        out = ctx.heat_solve(1e-14, 10 * mesh.n_nodes)
        assert out["status"] == 0, out
        assert np.abs(out["T"] - T_ref).max() <= 1e-10 * np.abs(T_ref).max()
        assert np.array_equal(ctx.get_temperature(), out["T"])  # the solution becomes the device temperature
        # started from the solution, CG stops at once (solveWithGuess)
        again = ctx.heat_solve(1e-12, 10 * mesh.n_nodes)
        assert again["status"] == 0 and again["iters"] <= 1


@pytest.mark.parametrize("variant", [0, 6, 11, 12])
@pytest.mark.parametrize("name", golden_names("wcb_"))
def test_boussinesq_wc_steps_match_reference_fixture(gpu_ctx_factory, name, variant):
    """variant: 0 default by size (gather kernels on these small meshes), 6 gather, 11 two-pass element records (what a C5-size
    BoussinesqWC mesh takes: k_wc_mom_elem<.., TH>), 12 two-pass continuity + gather momentum."""
    mesh, z = load_golden(name)
    dim, nn = mesh.dim, mesh.n_nodes
    k, cv, alpha, Tr = [float(v) for v in z["thermal"]]
    nst = 2 * dim + 2
    with gpu_ctx_factory(dim) as ctx:
        ctx.set_mesh(mesh)
        if variant:
            ctx.wc_set_variant(variant)
        ctx.set_states(0, z["q0"][: nst * nn])
        ctx.set_thermal(k, cv, alpha, Tr)
        ctx.set_temperature(z["q0"][nst * nn:])
        ctx.set_temperature_bc(z["t_mask"], z["t_val"])
        w = z["wpar"]
        wp = ctx.wc_params(w[0], w[1], w[2], w[3], w[4:7], bool(w[7]), {0: "CDS_dpdt", 1: "CDS_drhodt", 2: "CDS_rho"}[int(w[8])])
        for step in range(z["dts"].shape[0]):
            dt = ctx.wc_next_dt(wp, float(z["security_coeff"]), float(z["max_dt"]))
            assert abs(dt - z["dts"][step]) <= 1e-13 * z["dts"][step], (step, dt, z["dts"][step])
            ctx.wc_step(wp, float(z["dts"][step]))
            got = np.concatenate([ctx.get_states(0, nst), ctx.get_temperature()])
            want = z["states"][step]
            for s in range(nst + 1):
                a, r = got[s * nn:(s + 1) * nn], want[s * nn:(s + 1) * nn]
                m = np.abs(r).max()
                assert np.abs(a - r).max() <= TOL * 10 ** step * (m if m > 0 else 1.0), (step, s)
            assert np.abs(ctx.get_positions() - z["xs"][step]).max() < 1e-13
        # without the thermal factors the same call sequence gives different velocities (the buoyancy term is live)
        ctx.set_thermal(None)
        ctx.set_states(0, z["q0"][: nst * nn])
        ctx.set_positions(mesh.x)
        ctx.wc_step(wp, float(z["dts"][0]))
        assert np.abs(ctx.get_states(0, dim) - z["states"][0][: dim * nn]).max() > 0


@pytest.mark.parametrize("name", golden_names("wcb_"))
def test_boussinesq_wc_kernel_variants_agree_through_the_chained_run(gpu_ctx_factory, name):
    """Gather kernels against the two-pass kernels with the buoyancy factor (k_wc_mom_elem<.., TH>), through pfem_wc_run
    (chained CFL steps with the thermal diffusivity term): same per-node summation order, fields within rounding of each
    other (the two kernels contract their multiply-adds differently, so not bit for bit)."""
    mesh, z = load_golden(name)
    dim, nn = mesh.dim, mesh.n_nodes
    k, cv, alpha, Tr = [float(v) for v in z["thermal"]]
    nst = 2 * dim + 2
    out = {}
    for variant in (6, 11):
        with gpu_ctx_factory(dim) as ctx:
            ctx.set_mesh(mesh)
            ctx.wc_set_variant(variant)
            ctx.set_states(0, z["q0"][: nst * nn])
            ctx.set_thermal(k, cv, alpha, Tr)
            ctx.set_temperature(z["q0"][nst * nn:])
            ctx.set_temperature_bc(z["t_mask"], z["t_val"])
            w = z["wpar"]
            wp = ctx.wc_params(w[0], w[1], w[2], w[3], w[4:7], bool(w[7]), "CDS_dpdt")
            dt0 = ctx.wc_next_dt(wp, float(z["security_coeff"]), float(z["max_dt"]))
            dt, el = ctx.wc_run(wp, 5, float(z["security_coeff"]), float(z["max_dt"]), dt0)
            out[variant] = (dt0, dt, el, ctx.get_states(0, nst), ctx.get_temperature(), ctx.get_positions())
    a, b = out[6], out[11]
    assert a[0] == b[0] and abs(a[1] - b[1]) <= 1e-12 * a[1] and abs(a[2] - b[2]) <= 1e-12 * a[2]
    for s in range(nst):
        u, v = a[3][s * nn:(s + 1) * nn], b[3][s * nn:(s + 1) * nn]
        assert np.abs(u - v).max() <= 1e-11 * max(np.abs(u).max(), 1e-300), s
    assert np.abs(a[4] - b[4]).max() <= 1e-12 * np.abs(a[4]).max() and np.abs(a[5] - b[5]).max() < 1e-13
