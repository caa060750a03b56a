"""The C++ oracle against the independent numpy restatement (the pin against the reference's own code is tests/test_oracle_vs_reference.py)."""
import numpy as np
import pytest

from oracle import literal_numpy as lit
from oracle import oracle as orc
from pfem_b200 import meshgen as mg

from helpers import pspg_case, rel_err


@pytest.mark.parametrize("dim,n", [(2, 5), (3, 3)])
def test_pspg_elements(dim, n):
    mesh, q, q_prev, par = pspg_case(dim, n, free_fraction=0.02)
    vcur = q[: dim * mesh.n_nodes].copy()
    Ae, be, tau = orc.pspg_elements(mesh, vcur, q_prev, par)
    Ae2, be2, tau2 = lit.pspg_elements(mesh, vcur, q_prev, par[0], par[1], par[2], par[3:6])
    npe, nd = dim + 1, dim * (dim + 1)
    for sl in ((slice(0, nd), slice(0, nd)), (slice(0, nd), slice(nd, None)), (slice(nd, None), slice(0, nd)),
               (slice(nd, None), slice(nd, None))):
        assert rel_err(Ae[:, sl[0], sl[1]], Ae2[:, sl[0], sl[1]]) < 1e-13
    assert rel_err(be[:, :nd], be2[:, :nd]) < 1e-13
    assert rel_err(be[:, nd:], be2[:, nd:]) < 1e-13
    assert rel_err(tau, tau2) < 1e-14


@pytest.mark.parametrize("dim,n", [(2, 5), (3, 3)])
def test_pspg_assembly_and_bc(dim, n):
    mesh, q, q_prev, par = pspg_case(dim, n, free_fraction=0.02)
    vcur = q[: dim * mesh.n_nodes].copy()
    A, b = orc.pspg_build(mesh, vcur, q_prev, par, True)
    Ae2, be2, _ = lit.pspg_elements(mesh, vcur, q_prev, par[0], par[1], par[2], par[3:6])
    Ad, bd = lit.pspg_assemble_dense(mesh, Ae2, be2, q_prev, par[2], par[3:6], True)
    nn = mesh.n_nodes
    Aa = A.toarray()
    for rs in (slice(0, dim * nn), slice(dim * nn, None)):
        for cs in (slice(0, dim * nn), slice(dim * nn, None)):
            assert rel_err(Aa[rs, cs], Ad[rs, cs]) < 1e-13
    assert rel_err(b[: dim * nn], bd[: dim * nn]) < 1e-13
    assert rel_err(b[dim * nn:], bd[dim * nn:]) < 1e-13
    # pattern: every dense non-zero is stored; stored explicit zeros only in eliminated columns / masked structure
    stored = np.zeros(Ad.shape, dtype=bool)
    stored[A.indices, np.repeat(np.arange(A.shape[1]), np.diff(A.indptr))] = True
    assert ((Ad != 0) <= stored).all()


@pytest.mark.parametrize("dim,n,meduri,eq", [(2, 5, True, "CDS_dpdt"), (2, 4, False, "CDS_dpdt"), (3, 3, True, "CDS_dpdt"),
                                              (2, 5, True, "CDS_drhodt"), (3, 3, False, "CDS_drhodt"),
                                              (2, 4, False, "CDS_rho"), (3, 3, True, "CDS_rho")])
def test_wc_step(dim, n, meduri, eq):
    mesh = mg.kuhn_box(dim, n, free_fraction=0.02)
    st = mg.wc_state(mesh)
    st["acc"] = 0.1 * np.random.default_rng(2).standard_normal(st["acc"].shape)
    W = mg.WC_PARAMS
    g = mg.gravity(dim)
    wp = orc.wc_param_array(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, meduri, eq)
    x1, s1 = orc.wc_step(mesh, mesh.x, st, wp, 1e-5)
    x2, s2 = lit.wc_step(mesh, mesh.x, st, W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, 1e-5, meduri, eq)
    assert rel_err(x1, x2) < 1e-15
    for k in ("v", "p", "rho", "acc"):
        assert rel_err(s1[k], s2[k]) < 1e-12, k
