"""Parity at scale and solver robustness.

* the assembled PSPG system of a 413 k-tet mesh against the oracle's Eigen-style assembly of the SAME mesh, every entry
  (the oracle is C++/OpenMP: a few seconds at this size);
* three explicit weakly-compressible steps on a 1.05 M-tet mesh against the oracle, default kernels at that size
  (two-pass element records) -- the small-mesh tests never reach them through the size rule;
* the AUTO preconditioner over mesh families x material / time-step regimes: every case must converge, within a stated
  iteration bound, to the same fields as node-block Jacobi.
"""
import numpy as np
import pytest

from helpers import TOL_AB, block_errors, rel_err, vec_block_errors
from oracle import oracle as orc
from pfem_b200 import meshgen as mg

pytestmark = pytest.mark.gpu


def test_assembly_matches_oracle_at_400k_tets(gpu_ctx_factory):
    dim, n = 3, 41
    mesh = mg.kuhn_box(dim, n, free_fraction=0.001)
    assert mesh.n_elems >= 400_000
    q, q_prev = mg.pspg_state(mesh)
    q_prev = q_prev + 0.02 * np.random.default_rng(5).standard_normal(q_prev.shape)
    P = mg.PSPG_PARAMS
    g = mg.gravity(dim)
    par = orc.pspg_param_array(P["rho"], P["mu"], P["dt"], g)
    with gpu_ctx_factory(dim) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, q)
        ctx.pspg_assemble(ctx.pspg_params(P["rho"], P["mu"], P["dt"], g), q_prev)
        A, b = ctx.pspg_export_csc()
    A_ref, b_ref = orc.pspg_build(mesh, q[: dim * mesh.n_nodes].copy(), q_prev, par, True)
    for k, e in block_errors(A, A_ref, mesh.n_nodes, dim).items():   # also asserts the CSC pattern is identical
        assert e < TOL_AB, (k, e)
    for k, e in vec_block_errors(b, b_ref, mesh.n_nodes, dim).items():
        assert e < TOL_AB, (k, e)


def test_wc_steps_match_oracle_at_1m_tets(gpu_ctx_factory):
    dim, n = 3, 56
    mesh = mg.kuhn_box(dim, n)
    assert mesh.n_elems >= 1_000_000
    st = mg.wc_state(mesh)
    st["acc"] = 0.3 * np.random.default_rng(4).standard_normal(st["acc"].shape)
    W = mg.WC_PARAMS
    g = mg.gravity(dim)
    wpar = orc.wc_param_array(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, True, "CDS_dpdt")
    x_ref, st_ref = mesh.x, st
    nn = mesh.n_nodes
    with gpu_ctx_factory(dim) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, np.concatenate([st["v"], st["p"], st["rho"], st["acc"]]))
        wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, True)
        for step in range(3):
            dt = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
            dt_ref = orc.wc_next_dt(mesh, x_ref, st_ref, wpar, W["securityCoeff"], 1e-3)
            assert abs(dt - dt_ref) <= 1e-13 * dt_ref
            ctx.wc_step(wp, dt_ref)
            x_ref, st_ref = orc.wc_step(mesh, x_ref, st_ref, wpar, dt_ref)
            got = ctx.get_states(0, 2 * dim + 2)
            for k, sl in (("v", slice(0, dim * nn)), ("p", slice(dim * nn, (dim + 1) * nn)),
                          ("rho", slice((dim + 1) * nn, (dim + 2) * nn)), ("acc", slice((dim + 2) * nn, None))):
                assert rel_err(got[sl], st_ref[k]) < 1e-12 * 10 ** step, (k, step)
            assert np.abs(ctx.get_positions() - x_ref).max() < 1e-13


REGIMES = {  # rho, mu, dt
    "water_dambreak": (1000.0, 1e-3, 1e-3),     # examples/3D/damBreakKoshizuka
    "viscous_drop": (100.0, 1.0, 1e-3),         # examples/2D/squareToDisk
    "small_dt": (1000.0, 1e-3, 1e-5),
    "large_dt": (1000.0, 1e-3, 1e-1),
    "viscosity_dominated": (1000.0, 10.0, 1e-2),
    "gas_like": (1.0, 1e-5, 1e-3),
}
MESHES = {
    "kuhn3d": lambda: mg.kuhn_box(3, 16),
    "cloud3d": lambda: mg.delaunay_cloud(3, 4000, seed=3),
    "kuhn2d": lambda: mg.kuhn_box(2, 96),
}


@pytest.mark.parametrize("regime", sorted(REGIMES))
@pytest.mark.parametrize("family", sorted(MESHES))
def test_auto_preconditioner_converges_everywhere(gpu_ctx_factory, family, regime):
    """AUTO = multigrid with the hand-over to node-block Jacobi when the cycle is not a contraction.  Bound: whichever
    preconditioner finishes the solve, AUTO may not need more than the node-block Jacobi count + 400 iterations."""
    mesh = MESHES[family]()
    dim = mesh.dim
    rho, mu, dt = REGIMES[regime]
    q, q_prev = mg.pspg_state(mesh)
    g = mg.gravity(dim)
    with gpu_ctx_factory(dim) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, q)
        ctx.pspg_set_qprev(q_prev)
        par = ctx.pspg_params(rho, mu, dt, g)
        out = {}
        for kind in ("block", "auto"):
            ctx.pspg_set_preconditioner(kind)
            ctx.pspg_assemble_resident(par)
            out[kind] = ctx.pspg_solve(1e-11, 40000)
            out[kind]["used"] = ctx.pspg_get_preconditioner()[0]
    blk, aut = out["block"], out["auto"]
    assert blk["status"] == 0 and aut["status"] == 0, (family, regime, blk["status"], aut["status"], aut["iters"])
    assert aut["iters"] <= blk["iters"] + 400, (family, regime, aut["iters"], blk["iters"], aut["used"])
    nn = mesh.n_nodes
    for sl in (slice(0, dim * nn), slice(dim * nn, None)):
        assert rel_err(aut["q"][sl], blk["q"][sl]) < 1e-7, (family, regime)
