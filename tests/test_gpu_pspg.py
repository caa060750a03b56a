"""CUDA path (through the C ABI) vs the CPU oracle: PSPG assembly, BC, export, SpMV, Krylov solve, Picard."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import oracle as orc
from pfem_b200 import meshgen as mg
from pfem_b200.capi import PfemContext, PfemError

from helpers import TOL_AB, TOL_Q, block_errors, pspg_case, rel_err, vec_block_errors

pytestmark = pytest.mark.gpu

CASES = [
    (2, 4, dict()),
    (2, 24, dict(free_fraction=0.01)),                 # ~C1 size: 1.2 k triangles
    (2, 16, dict(permute=True, free_fraction=0.02)),
    (3, 3, dict()),
    (3, 13, dict(free_fraction=0.005)),                # ~C2 size: 13 k tets
    (3, 8, dict(permute=True, free_fraction=0.01)),
]


def _ctx_with_system(mesh, q, q_prev, par):
    dim = mesh.dim
    ctx = PfemContext(dim, 0)
    ctx.set_mesh(mesh)
    ctx.set_states(0, q)
    p = ctx.pspg_params(par[0], par[1], par[2], par[3:6])
    ctx.pspg_assemble(p, q_prev)
    return ctx, p


@pytest.mark.parametrize("dim,n,kw", CASES)
def test_assembly_matches_oracle(dim, n, kw):
    mesh, q, q_prev, par = pspg_case(dim, n, **kw)
    A_ref, b_ref = orc.pspg_build(mesh, q[: dim * mesh.n_nodes].copy(), q_prev, par, True)
    ctx, _ = _ctx_with_system(mesh, q, q_prev, par)
    with ctx:
        A, b = ctx.pspg_export_csc()
        info = ctx.info()
    assert A.nnz == A_ref.nnz == info.nnzReference
    errs = block_errors(A, A_ref, mesh.n_nodes, dim)          # also asserts identical CSC pattern
    assert max(errs.values()) < TOL_AB, errs
    berr = vec_block_errors(b, b_ref, mesh.n_nodes, dim)
    assert max(berr.values()) < TOL_AB, berr


@pytest.mark.parametrize("dim,n", [(2, 12), (3, 6)])
def test_assembly_is_bit_reproducible(dim, n):
    mesh, q, q_prev, par = pspg_case(dim, n, permute=True)
    outs = []
    for _ in range(2):
        ctx, _ = _ctx_with_system(mesh, q, q_prev, par)
        with ctx:
            A, b = ctx.pspg_export_csc()
        outs.append((A.data.copy(), b.copy()))
    assert (outs[0][0] == outs[1][0]).all() and (outs[0][1] == outs[1][1]).all()


@pytest.mark.parametrize("dim,n", [(2, 10), (3, 5)])
def test_matvec_and_residual(dim, n):
    mesh, q, q_prev, par = pspg_case(dim, n, free_fraction=0.01)
    A_ref, b_ref = orc.pspg_build(mesh, q[: dim * mesh.n_nodes].copy(), q_prev, par, True)
    x = np.random.default_rng(3).standard_normal(A_ref.shape[0])
    ctx, _ = _ctx_with_system(mesh, q, q_prev, par)
    with ctx:
        y = ctx.pspg_matvec(x)
        res = ctx.pspg_residual(x)
    y_ref = A_ref @ x
    assert rel_err(y, y_ref) < 1e-12
    assert abs(res - np.linalg.norm(y_ref - b_ref)) < 1e-12 * np.linalg.norm(y_ref - b_ref)


@pytest.mark.parametrize("dim,n,kw", [(2, 24, dict(free_fraction=0.01)), (3, 10, dict()), (3, 8, dict(permute=True))])
def test_bicgstab_matches_direct_solve(dim, n, kw):
    """Fields within 1e-8 of a direct (SuperLU/COLAMD) solve of the oracle's matrix: the stand-in for Eigen::SparseLU."""
    mesh, q, q_prev, par = pspg_case(dim, n, **kw)
    A_ref, b_ref = orc.pspg_build(mesh, q[: dim * mesh.n_nodes].copy(), q_prev, par, True)
    x_ref = spla.splu(A_ref.tocsc(), permc_spec="COLAMD").solve(b_ref)
    ctx, _ = _ctx_with_system(mesh, q, q_prev, par)
    with ctx:
        sol = ctx.pspg_solve(1e-13, 20000)
    assert sol["status"] == 0, sol
    assert sol["rel_res"] <= 1e-13 * 1.01
    nn = mesh.n_nodes
    assert rel_err(sol["q"][: dim * nn], x_ref[: dim * nn]) < TOL_Q
    assert rel_err(sol["q"][dim * nn:], x_ref[dim * nn:]) < TOL_Q


def test_solve_reports_non_convergence():
    mesh, q, q_prev, par = pspg_case(3, 6)
    ctx, _ = _ctx_with_system(mesh, q, q_prev, par)
    with ctx:
        sol = ctx.pspg_solve(1e-14, 3)
    assert sol["status"] == 1 and sol["iters"] <= 3           # PFEM_NOT_CONVERGED -> shim returns false -> dt halving


@pytest.mark.parametrize("dim,n", [(2, 12), (3, 6)])
def test_picard_iterations_match_oracle(dim, n):
    """PicardAlgo::solve (PicardAlgo.cpp:31-94) driven through pfem_pspg_picard_iter vs the oracle loop with SuperLU."""
    mesh = mg.kuhn_box(dim, n)
    q, q_prev = mg.pspg_state(mesh)
    P = mg.PSPG_PARAMS
    par = orc.pspg_param_array(P["rho"], P["mu"], P["dt"], mg.gravity(dim))
    ref = orc.pspg_picard(mesh, q, q_prev, par, max_iter=10, min_res=1e-6)
    assert ref["ok"]
    with PfemContext(dim, 0) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, q)
        p = ctx.pspg_params(par[0], par[1], par[2], par[3:6])
        ctx.snapshot_positions()
        ctx.pspg_assemble(p, q_prev)
        res, it, hist = np.finfo(float).max, 0, []
        while res > 1e-6:
            assert it <= 10
            out = ctx.pspg_picard_iter(p, q_prev, 1e-13, 20000)
            assert out["status"] == 0
            res = out["res"]
            hist.append(res)
            it += 1
        x = ctx.get_positions()
        states = ctx.get_states(0, dim + 1)
    assert it == ref["iters"]
    nn = mesh.n_nodes
    assert rel_err(out["q"][: dim * nn], ref["q"][: dim * nn]) < TOL_Q
    assert rel_err(out["q"][dim * nn:], ref["q"][dim * nn:]) < TOL_Q
    assert rel_err(states, ref["q"]) < TOL_Q
    assert np.abs(x - ref["x"]).max() < 1e-12


def test_snapshot_restore_and_move():
    mesh = mg.kuhn_box(3, 4)
    delta = 1e-3 * np.random.default_rng(0).standard_normal(3 * mesh.n_nodes)
    with PfemContext(3, 0) as ctx:
        ctx.set_mesh(mesh)
        ctx.snapshot_positions()
        ctx.move_positions(delta)
        x1 = ctx.get_positions()
        ctx.move_positions(2 * delta, from_snapshot=True)
        x2 = ctx.get_positions()
        ctx.restore_positions()
        x3 = ctx.get_positions()
    assert (x1 == orc.move_positions(mesh, delta, mesh.x)).all()
    assert (x2 == orc.move_positions(mesh, 2 * delta, mesh.x)).all()
    assert (x3 == mesh.x).all()


def test_error_paths():
    with PfemContext(3, 0) as ctx:
        with pytest.raises(PfemError):
            ctx.n_nodes = 4
            ctx.set_positions(np.zeros(12))                   # before set_topology
        bad = np.array([[0, 1, 2, 99]], dtype=np.uint64)
        with pytest.raises(PfemError):
            ctx.set_topology(bad, np.zeros(4, dtype=np.uint8))
    mesh = mg.kuhn_box(2, 2)
    with PfemContext(2, 0) as ctx:
        ctx.set_mesh(mesh)
        with pytest.raises(PfemError):
            ctx.pspg_solve()                                   # no assembled system


def test_empty_and_degenerate_topologies():
    """Only isolated nodes (every row an identity row) and a single element."""
    with PfemContext(2, 0) as ctx:
        ctx.set_topology(np.zeros((0, 3), dtype=np.uint64), np.zeros(5, dtype=np.uint8))
        ctx.set_positions(np.random.default_rng(0).random(10))
        ctx.set_dirichlet(np.zeros(5, dtype=np.uint8), np.zeros(10))
        qp = np.arange(15, dtype=float)
        p = ctx.pspg_params(1000.0, 1e-3, 1e-3, [0, -9.81, 0])
        ctx.pspg_assemble(p, qp)
        A, b = ctx.pspg_export_csc()
        sol = ctx.pspg_solve(1e-12, 10)
    assert (A.toarray() == np.eye(15)).all()
    assert np.allclose(b[:5], qp[:5]) and np.allclose(b[5:10], qp[5:10] - 9.81e-3) and (b[10:] == 0).all()
    assert sol["status"] == 0 and np.allclose(sol["q"], b)
