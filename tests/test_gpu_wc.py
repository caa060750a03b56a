"""CUDA explicit weakly-compressible step and CFL time step vs the CPU oracle."""
import numpy as np
import pytest

from oracle import oracle as orc
from pfem_b200 import meshgen as mg
from pfem_b200.capi import PfemContext

from helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _pack(st):
    return np.concatenate([st["v"], st["p"], st["rho"], st["acc"]])


@pytest.mark.parametrize("dim,n,kw,meduri,eq", [
    (2, 24, dict(free_fraction=0.01), True, "CDS_dpdt"),          # ~C3 size
    (2, 10, dict(permute=True), False, "CDS_dpdt"),
    (3, 4, dict(), True, "CDS_dpdt"),
    (3, 10, dict(free_fraction=0.01, permute=True), True, "CDS_dpdt"),
    (3, 8, dict(), False, "CDS_dpdt"),
    (2, 24, dict(free_fraction=0.01), True, "CDS_drhodt"),
    (3, 8, dict(permute=True), False, "CDS_drhodt"),
    (2, 10, dict(permute=True), False, "CDS_rho"),
    (3, 10, dict(free_fraction=0.01, permute=True), True, "CDS_rho"),
])
@pytest.mark.parametrize("variant", [6, 11, 12, 13])
def test_wc_steps_match_oracle(dim, n, kw, meduri, eq, variant):
    """variant: pfem_wc_set_variant -- gather kernels | two-pass element records | mixed | tiles with the element records in shared memory (all must match the oracle)."""
    mesh = mg.kuhn_box(dim, n, **kw)
    st = mg.wc_state(mesh)
    st["acc"] = 0.5 * np.random.default_rng(4).standard_normal(st["acc"].shape)
    W = mg.WC_PARAMS
    g = mg.gravity(dim)
    wp_ref = orc.wc_param_array(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, meduri, eq)
    with PfemContext(dim, 0) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, _pack(st))
        ctx.wc_set_variant(variant)
        wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, meduri, eq)
        x_ref, st_ref = mesh.x, st
        for step in range(3):
            dt_ref = orc.wc_next_dt(mesh, x_ref, st_ref, wp_ref, W["securityCoeff"], 1e-3)
            dt = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
            assert abs(dt - dt_ref) <= 1e-13 * dt_ref
            ctx.wc_step(wp, dt_ref)
            x_ref, st_ref = orc.wc_step(mesh, x_ref, st_ref, wp_ref, dt_ref)
            out = ctx.get_states(0, 2 * dim + 2)
            nn = mesh.n_nodes
            got = dict(v=out[: dim * nn], p=out[dim * nn:(dim + 1) * nn], rho=out[(dim + 1) * nn:(dim + 2) * nn],
                       acc=out[(dim + 2) * nn:])
            for k in ("v", "p", "rho", "acc"):
                assert rel_err(got[k], st_ref[k]) < TOL * (10 ** step), (k, step)
            assert np.abs(ctx.get_positions() - x_ref).max() < 1e-13


def test_wc_bit_reproducible():
    mesh = mg.kuhn_box(3, 6, permute=True)
    st = mg.wc_state(mesh)
    W = mg.WC_PARAMS
    outs = []
    for _ in range(2):
        with PfemContext(3, 0) as ctx:
            ctx.set_mesh(mesh)
            ctx.set_states(0, _pack(st))
            wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], mg.gravity(3), True)
            for _s in range(2):
                ctx.wc_step(wp, 1e-6)
            outs.append(ctx.get_states(0, 8))
    assert (outs[0] == outs[1]).all()


def test_state_roundtrip():
    mesh = mg.kuhn_box(2, 3)
    q = np.random.default_rng(1).standard_normal(6 * mesh.n_nodes)
    with PfemContext(2, 0) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, q)
        assert (ctx.get_states(0, 6) == q).all()
        assert (ctx.get_states(2, 2) == q[2 * mesh.n_nodes: 4 * mesh.n_nodes]).all()
        assert (ctx.get_positions() == mesh.x).all()


@pytest.mark.parametrize("dim,n,eq", [(2, 20, "CDS_dpdt"), (3, 6, "CDS_dpdt"), (3, 5, "CDS_rho")])
@pytest.mark.parametrize("variant", [6, 11, 13])
def test_wc_run_equals_step_loop(dim, n, eq, variant):
    """pfem_wc_run (device-chained CFL dt, one CUDA graph per step) == the host loop of wc_step + wc_next_dt, bit for bit."""
    mesh = mg.kuhn_box(dim, n, free_fraction=0.01)
    st = mg.wc_state(mesh)
    W = mg.WC_PARAMS
    g = mg.gravity(dim)
    n_steps = 7
    with PfemContext(dim, 0) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, _pack(st))
        ctx.wc_set_variant(variant)
        wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, True, eq)
        dt = dt0 = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
        elapsed = 0.0
        for _ in range(n_steps):
            ctx.wc_step(wp, dt)
            elapsed += dt
            dt = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
        ref_states, ref_x, ref_dt = ctx.get_states(0, 2 * dim + 2), ctx.get_positions(), dt
    with PfemContext(dim, 0) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, _pack(st))
        ctx.wc_set_variant(variant)
        wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, True, eq)
        dt_next, el = ctx.wc_run(wp, n_steps, W["securityCoeff"], 1e-3, dt0)
        assert (ctx.get_states(0, 2 * dim + 2) == ref_states).all()
        assert (ctx.get_positions() == ref_x).all()
    assert dt_next == ref_dt and abs(el - elapsed) <= 1e-15 * elapsed


@pytest.mark.parametrize("dim,n", [(2, 16), (3, 8)])
@pytest.mark.parametrize("variant", [11, 12, 13])
def test_wc_dt_after_step_matches_full_recomputation(dim, n, variant):
    """pfem_wc_next_dt right after pfem_wc_step may reuse what the step left on the device (element he, nodal
    max(u^2, c^2) and alpha^2 -- two-pass configuration); after any other call it recomputes everything.  Same dt."""
    mesh = mg.kuhn_box(dim, n, free_fraction=0.01, permute=True)
    st = mg.wc_state(mesh)
    st["acc"] = 0.5 * np.random.default_rng(9).standard_normal(st["acc"].shape)
    W = mg.WC_PARAMS
    g = mg.gravity(dim)
    wp_ref = orc.wc_param_array(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, True, "CDS_dpdt")
    with PfemContext(dim, 0) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, _pack(st))
        ctx.wc_set_variant(variant)
        wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, True, "CDS_dpdt")
        x_ref, st_ref = mesh.x, st
        dt = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
        for _ in range(3):
            ctx.wc_step(wp, dt)
            x_ref, st_ref = orc.wc_step(mesh, x_ref, st_ref, wp_ref, dt)
            dt_after_step = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
            ctx.get_positions()                                   # any other call drops the step's leftovers
            dt_full = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
            dt_oracle = orc.wc_next_dt(mesh, x_ref, st_ref, wp_ref, W["securityCoeff"], 1e-3)
            assert abs(dt_after_step - dt_full) <= 1e-15 * dt_full
            assert abs(dt_full - dt_oracle) <= 1e-12 * dt_oracle
            dt = dt_full


@pytest.mark.parametrize("variant", [6, 11, 13])
@pytest.mark.parametrize("poison", ["velocity", "density"])
def test_nan_state_is_reported_not_stepped_over(variant, poison):
    """A NaN velocity or density must surface as PFEM_NAN ("NaN time step!", WCompNewton/Solver.cpp:228-232) from both
    pfem_wc_next_dt and pfem_wc_run -- fmin/fmax drop NaN operands, so the reductions must carry it explicitly."""
    from pfem_b200.capi import PfemError
    dim = 3
    mesh = mg.kuhn_box(dim, 6)
    st = mg.wc_state(mesh)
    nn = mesh.n_nodes
    node = nn // 2
    if poison == "velocity":
        st["v"][node] = np.nan
    else:
        st["rho"][node] = np.nan
    W = mg.WC_PARAMS
    g = mg.gravity(dim)
    with PfemContext(dim, 0) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, _pack(st))
        ctx.wc_set_variant(variant)
        wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, True)
        with pytest.raises(PfemError) as e1:
            ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
        assert e1.value.code == 2
    # NaN appearing DURING chained steps: start from a clean state, poison after the first dt.  (A NaN density alone is
    # overwritten by the dp/dt continuity step, which recomputes rho from p -- the reference does the same.)
    if poison == "density":
        return
    st = mg.wc_state(mesh)
    with PfemContext(dim, 0) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, _pack(st))
        ctx.wc_set_variant(variant)
        wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, True)
        dt0 = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
        bad = mg.wc_state(mesh)
        bad["v" if poison == "velocity" else "rho"][node] = np.nan
        ctx.set_states(0, _pack(bad))
        with pytest.raises(PfemError) as e2:
            ctx.wc_run(wp, 4, W["securityCoeff"], 1e-3, dt0)
        assert e2.value.code == 2
