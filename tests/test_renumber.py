"""Host-side spatial renumbering (pfem_b200/renumber.py): a pure relabelling -- the oracle gives the same fields."""
import numpy as np

from oracle import oracle as orc
from pfem_b200 import meshgen as mg
from pfem_b200.renumber import spatial_renumber

from helpers import rel_err


def test_renumbering_is_a_relabelling():
    mesh = mg.kuhn_box(3, 5, permute=True, free_fraction=0.02)
    rn = spatial_renumber(mesh)
    assert sorted(rn.old_of_new.tolist()) == list(range(mesh.n_nodes))
    assert (rn.new_of_old[rn.old_of_new] == np.arange(mesh.n_nodes)).all()
    assert mg.det_j(rn.mesh).min() > 0 and rn.mesh.n_elems == mesh.n_elems
    q = np.random.default_rng(0).standard_normal(4 * mesh.n_nodes)
    assert np.array_equal(rn.to_old(rn.to_new(q)), q)
    # Morton order restores locality: the index span of an element shrinks by a large factor
    span_old = np.ptp(mesh.conn, axis=1).mean()
    span_new = np.ptp(rn.mesh.conn, axis=1).mean()
    assert span_new < 0.6 * span_old


def test_wc_step_is_invariant_under_renumbering():
    mesh = mg.kuhn_box(3, 5, permute=True, free_fraction=0.02)
    rn = spatial_renumber(mesh)
    st = mg.wc_state(mesh)
    st["acc"] = 0.3 * np.random.default_rng(1).standard_normal(st["acc"].shape)
    W = mg.WC_PARAMS
    par = orc.wc_param_array(W["mu"], W["K0"], W["K0p"], W["rhoStar"], mg.gravity(3), True)
    dt = orc.wc_next_dt(mesh, mesh.x, st, par, 0.1, 1e-3)
    x1, s1 = orc.wc_step(mesh, mesh.x, st, par, dt)
    st_new = {k: rn.to_new(v) for k, v in st.items()}
    assert abs(orc.wc_next_dt(rn.mesh, rn.mesh.x, st_new, par, 0.1, 1e-3) - dt) <= 1e-15 * dt
    x2, s2 = orc.wc_step(rn.mesh, rn.mesh.x, st_new, par, dt)
    for k in s1:
        assert rel_err(rn.to_old(s2[k]), s1[k]) < 1e-12
    assert np.abs(rn.to_old(x2) - x1).max() < 1e-14
