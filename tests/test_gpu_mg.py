"""Aggregation-multigrid preconditioner of the device Krylov solve (csrc/mg.cu, csrc/krylov.cu): same solution as node-block Jacobi and as a
direct solve of the oracle's matrix, far fewer iterations, bit-reproducible."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import oracle as orc
from pfem_b200 import meshgen as mg
from pfem_b200.capi import PfemContext

from helpers import TOL_Q, pspg_case, rel_err

pytestmark = pytest.mark.gpu


def _solve(mesh, q, q_prev, par, kind, sweeps=0, damping=0.0, tol=1e-13):
    with PfemContext(mesh.dim, 0) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, q)
        ctx.pspg_assemble(ctx.pspg_params(par[0], par[1], par[2], par[3:6]), q_prev)
        ctx.pspg_set_preconditioner(kind, sweeps, damping)
        sol = ctx.pspg_solve(tol, 20000)
        sol["used"], sol["levels"] = ctx.pspg_get_preconditioner()
    return sol


@pytest.mark.parametrize("dim,n,kw", [(2, 40, dict(free_fraction=0.01)), (3, 12, dict()), (3, 10, dict(permute=True, free_fraction=0.01))])
def test_mg_matches_direct_solve(dim, n, kw):
    mesh, q, q_prev, par = pspg_case(dim, n, **kw)
    A_ref, b_ref = orc.pspg_build(mesh, q[: dim * mesh.n_nodes].copy(), q_prev, par, True)
    x_ref = spla.splu(A_ref.tocsc(), permc_spec="COLAMD").solve(b_ref)
    sol = _solve(mesh, q, q_prev, par, "mg")
    assert sol["used"] == "mg" and sol["levels"] >= 2
    assert sol["status"] == 0 and sol["rel_res"] <= 1e-13 * 1.01, sol
    nn = mesh.n_nodes
    assert rel_err(sol["q"][: dim * nn], x_ref[: dim * nn]) < TOL_Q
    assert rel_err(sol["q"][dim * nn:], x_ref[dim * nn:]) < TOL_Q


@pytest.mark.parametrize("dim,n", [(2, 64), (3, 20)])
def test_mg_needs_fewer_iterations_than_block_jacobi(dim, n):
    mesh = mg.kuhn_box(dim, n)
    q, q_prev = mg.pspg_state(mesh)
    P = mg.PSPG_PARAMS
    par = orc.pspg_param_array(P["rho"], P["mu"], P["dt"], mg.gravity(dim))
    blk = _solve(mesh, q, q_prev, par, "block", tol=1e-12)
    mgs = _solve(mesh, q, q_prev, par, "mg", tol=1e-12)
    assert blk["status"] == 0 and mgs["status"] == 0
    assert blk["used"] == "block" and mgs["used"] == "mg"
    assert mgs["iters"] * 3 < blk["iters"], (mgs["iters"], blk["iters"])
    assert rel_err(mgs["q"], blk["q"]) < TOL_Q


def test_mg_solve_is_bit_reproducible():
    mesh, q, q_prev, par = pspg_case(3, 9, permute=True)
    a = _solve(mesh, q, q_prev, par, "mg", 1, 1.5)
    b = _solve(mesh, q, q_prev, par, "mg", 1, 1.5)
    assert a["iters"] == b["iters"] and (a["q"] == b["q"]).all()


def test_mg_on_delaunay_cloud():
    mesh = mg.delaunay_cloud(3, 4000, seed=3)
    q, q_prev = mg.pspg_state(mesh)
    P = mg.PSPG_PARAMS
    par = orc.pspg_param_array(P["rho"], P["mu"], P["dt"], mg.gravity(3))
    blk = _solve(mesh, q, q_prev, par, "block", tol=1e-12)
    mgs = _solve(mesh, q, q_prev, par, "mg", tol=1e-12)
    assert mgs["status"] == 0 and mgs["used"] == "mg"
    assert mgs["iters"] < blk["iters"]
    assert rel_err(mgs["q"], blk["q"]) < TOL_Q


def test_tiny_mesh_keeps_block_jacobi():
    mesh, q, q_prev, par = pspg_case(2, 4)          # 25 nodes: no second level
    sol = _solve(mesh, q, q_prev, par, "auto")
    assert sol["status"] == 0 and sol["used"] == "block" and sol["levels"] == 1


def test_auto_hands_over_when_the_cycle_is_not_a_contraction():
    """Viscosity-dominated regime: node-block Jacobi is no smoother for the coupled system, the multigrid-preconditioned
    iteration diverges; AUTO must detect it early and finish with node-block Jacobi (same answer as BLOCK alone)."""
    mesh = mg.kuhn_box(3, 24)     # measured: the cycle diverges here (tools/mg_robustness.py), converges on smaller boxes
    q, q_prev = mg.pspg_state(mesh)
    par = orc.pspg_param_array(1000.0, 10.0, 1e-2, mg.gravity(3))
    blk = _solve(mesh, q, q_prev, par, "block", tol=1e-11)
    aut = _solve(mesh, q, q_prev, par, "auto", tol=1e-11)
    assert blk["status"] == 0 and aut["status"] == 0, (blk, aut)
    assert aut["used"] == "block"
    assert aut["iters"] < 2 * blk["iters"] + 100
    assert rel_err(aut["q"], blk["q"]) < TOL_Q


@pytest.mark.parametrize("dim,n", [(2, 48), (3, 16)])
def test_krylov_variants_agree(dim, n, monkeypatch):
    """Flexible GMRES (default under multigrid) against BiCGSTAB with the same cycle (PFEM_GMRES_M=0), a restart length that
    forces several restarts (PFEM_GMRES_M=4), polling after every iteration, and the fp64-vector cycle (PFEM_MG_FP32V=0):
    same fields; GMRES spends at most as many cycles as BiCGSTAB (which runs two per iteration)."""
    mesh = mg.kuhn_box(dim, n, free_fraction=0.01, permute=True)
    q, q_prev = mg.pspg_state(mesh)
    P = mg.PSPG_PARAMS
    par = orc.pspg_param_array(P["rho"], P["mu"], P["dt"], mg.gravity(dim))
    out = {}
    for name, env in (("gmres", {}), ("bicgstab", {"PFEM_GMRES_M": "0"}), ("restarts", {"PFEM_GMRES_M": "4"}),
                      ("poll_every", {"PFEM_GMRES_POLL_EVERY": "1"}), ("fp64_vectors", {"PFEM_MG_FP32V": "0"})):
        for k in ("PFEM_GMRES_M", "PFEM_GMRES_POLL_EVERY", "PFEM_MG_FP32V"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        out[name] = _solve(mesh, q, q_prev, par, "mg", tol=1e-12)
        assert out[name]["status"] == 0 and out[name]["used"] == "mg" and out[name]["rel_res"] <= 1e-12 * 1.01, (name, out[name])
    ref_q = out["gmres"]["q"]
    for name in ("bicgstab", "restarts", "poll_every", "fp64_vectors"):
        assert rel_err(out[name]["q"], ref_q) < TOL_Q, name
    assert out["gmres"]["iters"] <= 2 * out["bicgstab"]["iters"]
    assert out["restarts"]["iters"] >= out["gmres"]["iters"]
    assert out["poll_every"]["iters"] == out["gmres"]["iters"] and (out["poll_every"]["q"] == ref_q).all()
    assert abs(out["fp64_vectors"]["iters"] - out["gmres"]["iters"]) <= 1
