"""CUDA path (through the C ABI) vs fixtures produced by the REFERENCE'S OWN code (tests/golden/, see
tests/golden/make_golden.py and oracle/refbuild/): assembled CSC matrix and RHS to 1e-12 per block type, Picard fields
to 1e-8 with the same iteration count, explicit weakly-compressible steps to 1e-12.  No oracle in the loop."""
import numpy as np
import pytest

from pfem_b200.capi import PfemContext

from helpers import (TOL_AB, TOL_Q, block_errors, golden_csc, golden_names, load_golden, rel_err, split_wc,
                     vec_block_errors)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names("pspg_"))
def test_assembly_matches_reference_fixture(name):
    mesh, z = load_golden(name)
    dim, nn, par = mesh.dim, mesh.n_nodes, z["par"]
    with PfemContext(dim, 0) as ctx:
        ctx.set_mesh(mesh)
        if "facets" in z:                                  # surface tension: getFST on the free-surface facets
            ctx.set_facets(z["facets"])
            ctx.set_surface_tension(float(z["gamma"]))
        ctx.set_states(0, z["q"])
        p = ctx.pspg_params(par[0], par[1], par[2], par[3:6])
        ctx.pspg_assemble(p, z["q_prev"])
        A, b = ctx.pspg_export_csc()
        if "facets" in z:
            ctx.set_surface_tension(0.0)
            ctx.pspg_assemble(p, z["q_prev"])
            _, b_off = ctx.pspg_export_csc()
            assert np.abs(b_off - z["b"]).max() > 1e-6 * np.abs(z["b"]).max()   # the fixture exercises the facet terms
    A_ref = golden_csc(z, "A")
    assert A.nnz == A_ref.nnz
    errs = block_errors(A, A_ref, nn, dim)                 # also asserts the identical CSC pattern
    assert max(errs.values()) < TOL_AB, errs
    berr = vec_block_errors(b, z["b"], nn, dim)
    assert max(berr.values()) < TOL_AB, berr


@pytest.mark.parametrize("name", golden_names("picard_"))
def test_picard_matches_reference_fixture(name):
    """The reference's MomContEqIncompNewton::solve() (Picard on the mesh position, SuperLU behind the stand-in SparseLU)
    vs pfem_pspg_picard_iter driven by the PicardAlgo.cpp:31-94 loop."""
    mesh, z = load_golden(name)
    dim, nn, par = mesh.dim, mesh.n_nodes, z["par"]
    q_prev, min_res, max_iter = z["q_prev"], float(z["min_res"]), int(z["max_iter"])
    with PfemContext(dim, 0) as ctx:
        ctx.set_mesh(mesh)
        ctx.set_states(0, q_prev)
        p = ctx.pspg_params(par[0], par[1], par[2], par[3:6])
        ctx.snapshot_positions()
        ctx.pspg_assemble(p, q_prev)
        res, it = np.finfo(float).max, 0
        while res > min_res:
            assert it <= max_iter
            out = ctx.pspg_picard_iter(p, q_prev, 1e-13, 20000)
            assert out["status"] == 0
            res = out["res"]
            it += 1
        x = ctx.get_positions()
        states = ctx.get_states(0, dim + 1)
    assert it == int(z["iters"])
    assert rel_err(states[: dim * nn], z["q"][: dim * nn]) < TOL_Q
    assert rel_err(states[dim * nn:], z["q"][dim * nn:]) < TOL_Q
    assert np.abs(x - z["x_new"]).max() < 1e-9           # positions move by dt*v: 1e-3 * (1e-8 relative of O(1))


@pytest.mark.parametrize("variant", [6, 11, 13])
@pytest.mark.parametrize("name", golden_names("wc_"))
def test_wc_steps_match_reference_fixture(name, variant):
    mesh, z = load_golden(name)
    dim, nn, wpar = mesh.dim, mesh.n_nodes, z["wpar"]
    eq = {0: "CDS_dpdt", 1: "CDS_drhodt", 2: "CDS_rho"}[int(wpar[8])]
    with PfemContext(dim, 0) as ctx:
        ctx.set_mesh(mesh)
        if "facets" in z:
            ctx.set_facets(z["facets"])
            ctx.set_surface_tension(float(z["gamma"]))
        ctx.set_states(0, z["q0"])
        ctx.wc_set_variant(variant)                      # gather kernels | two-pass element records
        wp = ctx.wc_params(wpar[0], wpar[1], wpar[2], wpar[3], wpar[4:7], bool(wpar[7]), eq)
        for step in range(z["dts"].shape[0]):
            dt = ctx.wc_next_dt(wp, float(z["security_coeff"]), float(z["max_dt"]))
            assert abs(dt - z["dts"][step]) <= 1e-13 * z["dts"][step]
            ctx.wc_step(wp, float(z["dts"][step]))
            got = split_wc(ctx.get_states(0, 2 * dim + 2), dim, nn)
            want = split_wc(z["states"][step], dim, nn)
            for k in ("v", "p", "rho", "acc"):
                assert rel_err(got[k], want[k]) < 1e-12 * 10 ** step, (k, step)
            assert np.abs(ctx.get_positions() - z["xs"][step]).max() < 1e-13
