"""Partitioned mesh vs one GPU, on a SINGLE-GPU box: the ranks are contexts of this process driven by threads
(pfem_comm_local_*, pfem_b200/localranks.py), all on device 0.  Same library code path as the NCCL runs except for the
transport underneath commHalo / commAllReduce / commAllGather.

Bars: explicit weakly-compressible steps bit-identical to the 1-GPU run (same per-node summation order); PSPG fields
within 1e-8 (north_star) with the multigrid-preconditioned Krylov solve taking at most 1.5x (+2) the single-GPU iterations.
"""
import os

import numpy as np
import pytest

from pfem_b200 import meshgen as mg
from pfem_b200.capi import PfemContext
from pfem_b200.localranks import run_ranks
from pfem_b200.partition import gather_owned

pytestmark = pytest.mark.gpu


def _gather(results, parts, n_comp, nn):
    return gather_owned([r for r in results], parts, n_comp, nn)


@pytest.mark.parametrize("n_ranks", [2, 3, 4])
@pytest.mark.parametrize("variant", [6, 11, 13])
def test_wc_steps_bit_identical(n_ranks, variant):
    dim = 3
    mesh = mg.kuhn_box(dim, 10, free_fraction=0.002, permute=True)
    nn = mesh.n_nodes
    W = mg.WC_PARAMS
    g = mg.gravity(dim)
    st = mg.wc_state(mesh)
    st["acc"] = 0.3 * np.random.default_rng(4).standard_normal(st["acc"].shape)
    packed = np.concatenate([st["v"], st["p"], st["rho"], st["acc"]])
    nst = 2 * dim + 2

    def steps(ctx, q):
        ctx.wc_set_variant(variant)
        ctx.set_states(0, q)
        wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, True)
        dts = []
        for _ in range(3):
            dt = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
            dts.append(dt)
            ctx.wc_step(wp, dt)
        return dts, ctx.get_states(0, nst), ctx.get_positions()

    with PfemContext(dim, 0) as one:
        one.set_mesh(mesh)
        dts1, q1, x1 = steps(one, packed)

    parts = [None] * n_ranks

    def fn(r, ctx, part):
        parts[r] = part
        return steps(ctx, part.scatter_nodal(packed, nst, nn))

    res = run_ranks(mesh, n_ranks, fn)
    for r in range(n_ranks):
        assert res[r][0] == dts1, (res[r][0], dts1)
    q = gather_owned([r[1] for r in res], parts, nst, nn)
    x = gather_owned([r[2] for r in res], parts, dim, nn)
    assert np.array_equal(q, q1) and np.array_equal(x, x1)
    # ghost copies equal their owners' values after the last exchange (velocity, pressure, density)
    for r, part in enumerate(parts):
        nl = part.l2g_nodes.size
        loc = res[r][1].reshape(nst, nl)[: dim + 2, part.n_owned:]
        assert np.array_equal(loc, q1.reshape(nst, nn)[: dim + 2, part.l2g_nodes[part.n_owned:]])


@pytest.mark.parametrize("n_ranks,repl,factor", [(2, 50000, 1.5), (4, 50000, 1.5), (2, 0, 2.5), (4, 0, 2.5), (3, 300, 1.5)])
def test_pspg_solve_and_picard_match_single_gpu(n_ranks, repl, factor, monkeypatch):
    """repl = PFEM_MG_REPL_NODES: 50000 replicates level 1 at once (the small-mesh default path), 0 keeps every level
    distributed down to a smoothing-only coarsest level (no exact coarse solve: a looser iteration bound), 300 mixes the
    two kinds (level 1 distributed with ghost aggregates, level 2 replicated)."""
    dim = 3
    monkeypatch.setenv("PFEM_MG_REPL_NODES", str(repl))
    mesh = mg.kuhn_box(dim, 14, free_fraction=0.002, permute=True)
    nn = mesh.n_nodes
    P = mg.PSPG_PARAMS
    g = mg.gravity(dim)
    q, q_prev = mg.pspg_state(mesh)

    def solve(ctx, ql, qpl):
        ctx.set_states(0, ql)
        par = ctx.pspg_params(P["rho"], P["mu"], P["dt"], g)
        ctx.snapshot_positions()
        ctx.pspg_assemble(par, qpl)
        s = ctx.pspg_solve(1e-12, 5000)
        kind, levels = ctx.pspg_get_preconditioner()
        o = ctx.pspg_picard_iter(par, None, 1e-12, 5000)
        return s, o, kind, levels

    with PfemContext(dim, 0) as one:
        one.set_mesh(mesh)
        s1, o1, kind1, lev1 = solve(one, q, q_prev)
    assert s1["status"] == 0 and kind1 == "mg"

    parts = [None] * n_ranks

    def fn(r, ctx, part):
        parts[r] = part
        return solve(ctx, part.scatter_nodal(q, dim + 1, nn), part.scatter_nodal(q_prev, dim + 1, nn))

    res = run_ranks(mesh, n_ranks, fn)
    for r in range(n_ranks):
        s, o, kind, levels = res[r]
        assert s["status"] == 0 and o["status"] == 0 and kind == "mg", (s["status"], o["status"], kind)
        assert s["iters"] == res[0][0]["iters"]  # every rank ran the same iteration
        assert s["iters"] <= factor * s1["iters"] + 2, (s["iters"], s1["iters"], levels)
    for name, idx, ref in (("solve", 0, s1["q"]), ("picard", 1, o1["q"])):
        qq = gather_owned([r[idx]["q"] for r in res], parts, dim + 1, nn)
        ev = np.abs(qq[: dim * nn] - ref[: dim * nn]).max() / np.abs(ref[: dim * nn]).max()
        ep = np.abs(qq[dim * nn:] - ref[dim * nn:]).max() / np.abs(ref[dim * nn:]).max()
        assert ev < 1e-8 and ep < 1e-8, (name, ev, ep)
    assert abs(res[0][1]["res"] - o1["res"]) <= 1e-6 * max(o1["res"], 1e-30) + 1e-9


def test_failing_rank_releases_the_others():
    """A rank that fails outside the library must not leave the others waiting in a barrier for ever."""
    mesh = mg.kuhn_box(3, 6)

    def fn(r, ctx, part):
        if r == 1:
            raise RuntimeError("boom")
        W = mg.WC_PARAMS
        wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], mg.gravity(3), True)
        return ctx.wc_next_dt(wp, 0.1, 1e-3)  # all-reduce: would block without the abort

    with pytest.raises(Exception):
        run_ranks(mesh, 2, fn)


@pytest.mark.parametrize("n_ranks", [2, 4])
def test_wc_run_chained_dt_on_partitioned_mesh(n_ranks):
    """pfem_wc_run on a partitioned mesh (CFL minimum all-reduced on the device, no host round trip between steps) ==
    the single-GPU chain, bit for bit."""
    dim = 3
    mesh = mg.kuhn_box(dim, 10, free_fraction=0.002, permute=True)
    nn = mesh.n_nodes
    W = mg.WC_PARAMS
    g = mg.gravity(dim)
    st = mg.wc_state(mesh)
    packed = np.concatenate([st["v"], st["p"], st["rho"], st["acc"]])
    nst = 2 * dim + 2

    def chain(ctx, q):
        ctx.wc_set_variant(11)
        ctx.set_states(0, q)
        wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, True)
        dt0 = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
        dt, el = ctx.wc_run(wp, 6, W["securityCoeff"], 1e-3, dt0)
        return dt0, dt, el, ctx.get_states(0, nst), ctx.get_positions()

    with PfemContext(dim, 0) as one:
        one.set_mesh(mesh)
        ref = chain(one, packed)
    parts = [None] * n_ranks

    def fn(r, ctx, part):
        parts[r] = part
        return chain(ctx, part.scatter_nodal(packed, nst, nn))

    res = run_ranks(mesh, n_ranks, fn)
    for r in res:
        assert r[0] == ref[0] and r[1] == ref[1] and r[2] == ref[2]
    assert np.array_equal(gather_owned([r[3] for r in res], parts, nst, nn), ref[3])
    assert np.array_equal(gather_owned([r[4] for r in res], parts, dim, nn), ref[4])
