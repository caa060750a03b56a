"""Unstructured (Delaunay + alpha criterion) meshes: node valence up to ~40 tets exercises the multi-chunk paths
(more than 32 incident elements / more than 28 neighbour blocks per node) of the gather kernels."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import oracle as orc
from pfem_b200 import meshgen as mg
from pfem_b200.capi import PfemContext

from helpers import TOL_AB, TOL_Q, block_errors, rel_err, vec_block_errors

pytestmark = pytest.mark.gpu


def _case(dim, n_points, seed):
    mesh = mg.delaunay_cloud(dim, n_points, seed=seed, free_fraction=0.01)
    q, q_prev = mg.pspg_state(mesh)
    rng = np.random.default_rng(seed)
    q_prev = q_prev + 0.02 * rng.standard_normal(q_prev.shape)
    vals = mesh.dir_val.reshape(dim, mesh.n_nodes)
    sel = (mesh.dir_mask != 0) & (np.arange(mesh.n_nodes) % 2 == 0)
    vals[:, sel] = 0.1 * rng.standard_normal((dim, int(sel.sum())))
    mesh.dir_val = vals.reshape(-1)
    return mesh, q, q_prev


@pytest.mark.parametrize("dim,n_points,seed", [(2, 1200, 11), (3, 3000, 12), (3, 9000, 13)])
def test_pspg_on_delaunay_mesh(dim, n_points, seed):
    mesh, q, q_prev = _case(dim, n_points, seed)
    valence = np.bincount(mesh.conn.ravel(), minlength=mesh.n_nodes).max()
    if dim == 3:
        assert valence > 32, valence                  # the case must actually reach the second 32-element chunk
    P = mg.PSPG_PARAMS
    par = orc.pspg_param_array(P["rho"], P["mu"], P["dt"], mg.gravity(dim))
    A_ref, b_ref = orc.pspg_build(mesh, q[: dim * mesh.n_nodes].copy(), q_prev, par, True)
    with PfemContext(dim, 0) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, q)
        ctx.pspg_assemble(ctx.pspg_params(par[0], par[1], par[2], par[3:6]), q_prev)
        A, b = ctx.pspg_export_csc()
        sol = ctx.pspg_solve(1e-13, 40000)
    errs = block_errors(A, A_ref, mesh.n_nodes, dim)
    assert max(errs.values()) < TOL_AB, errs
    berr = vec_block_errors(b, b_ref, mesh.n_nodes, dim)
    assert max(berr.values()) < TOL_AB, berr
    x_ref = spla.splu(A_ref.tocsc(), permc_spec="COLAMD").solve(b_ref)
    nn = mesh.n_nodes
    assert sol["status"] == 0, sol
    assert rel_err(sol["q"][: dim * nn], x_ref[: dim * nn]) < TOL_Q
    assert rel_err(sol["q"][dim * nn:], x_ref[dim * nn:]) < TOL_Q


@pytest.mark.parametrize("dim,n_points,seed", [(2, 1200, 21), (3, 3000, 22)])
def test_wc_on_delaunay_mesh(dim, n_points, seed):
    mesh = mg.delaunay_cloud(dim, n_points, seed=seed, free_fraction=0.01)
    st = mg.wc_state(mesh)
    st["acc"] = 0.5 * np.random.default_rng(seed).standard_normal(st["acc"].shape)
    W = mg.WC_PARAMS
    g = mg.gravity(dim)
    wp_ref = orc.wc_param_array(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, True)
    with PfemContext(dim, 0) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, np.concatenate([st["v"], st["p"], st["rho"], st["acc"]]))
        wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, True)
        x_ref, st_ref = mesh.x, st
        for _ in range(2):
            dt_ref = orc.wc_next_dt(mesh, x_ref, st_ref, wp_ref, W["securityCoeff"], 1e-3)
            dt = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
            assert abs(dt - dt_ref) <= 1e-12 * dt_ref
            ctx.wc_step(wp, dt_ref)
            x_ref, st_ref = orc.wc_step(mesh, x_ref, st_ref, wp_ref, dt_ref)
        out = ctx.get_states(0, 2 * dim + 2)
        x = ctx.get_positions()
    nn = mesh.n_nodes
    got = dict(v=out[: dim * nn], p=out[dim * nn:(dim + 1) * nn], rho=out[(dim + 1) * nn:(dim + 2) * nn], acc=out[(dim + 2) * nn:])
    for k in ("v", "p", "rho", "acc"):
        assert rel_err(got[k], st_ref[k]) < 1e-11, k
    assert np.abs(x - x_ref).max() < 1e-13
