"""The numpy restatement of the fractional-step systems (oracle/fracstep_numpy.py) against the reference-made fixtures
(tests/golden/fs_*.npz) and, where the reference build is present, against the reference's own code on other meshes."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import fracstep_numpy as fs
from oracle import oracle as orc
from oracle import ref
from pfem_b200 import meshgen as mg

from helpers import golden_names, load_golden

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (no /root/reference here)")
TOL = 1e-12


def _dense(z, prefix):
    n = z[prefix + "_indptr"].shape[0] - 1
    return sp.csc_matrix((z[prefix + "_data"], z[prefix + "_indices"], z[prefix + "_indptr"]), shape=(n, n)).toarray()


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("name", golden_names("fs_"))
def test_oracle_matches_fracstep_fixture(name):
    mesh, z = load_golden(name)
    dim, nn = mesh.dim, mesh.n_nodes
    rho, mu, dt = z["par"][:3]
    body, g = z["par"][3:6], float(z["gamma_fs"])
    v_prev, p_prev = z["q_prev"][: dim * nn], z["q_prev"][dim * nn:]
    A0, b0 = fs.velocity_prediction(mesh, v_prev, p_prev, rho, mu, dt, body, g)
    assert _rel(A0, _dense(z, "A0")) < TOL and _rel(b0, z["b0"]) < TOL
    x, it, _, ok = fs.conjugate_gradient(A0, b0)
    assert ok and it == int(z["cg0"][0]) and _rel(x, z["v_tilde"]) < 1e-12
    A1, b1 = fs.pressure(mesh, z["v_tilde"], p_prev, rho, mu, dt, body, g)
    assert _rel(A1, _dense(z, "A1")) < TOL and _rel(b1, z["b1"]) < TOL
    _, it1, _, ok1 = fs.conjugate_gradient(A1, b1)
    assert not ok1 and it1 == 2 * nn and int(z["cg1"][2]) == 2     # singular as built: the cap on both sides
    A2, b2 = fs.velocity_correction(mesh, z["delta_p"], rho, mu, dt, body)
    assert _rel(A2, _dense(z, "A2")) < TOL and _rel(b2, z["b2"]) < TOL
    x, it, _, ok = fs.conjugate_gradient(A2, b2)
    assert ok and it == int(z["cg2"][0]) and _rel(x, z["dv"]) < 1e-12


@needs_ref
@pytest.mark.parametrize("seed", range(4))
def test_live_fracstep_systems(seed):
    """Seeded cases (dimension, mesh family, numbering, Dirichlet data, gammaFS): the restatement against the reference's own
    m_buildMat* + m_applyBC* run here through oracle/_ref."""
    rng = np.random.default_rng(500 + seed)
    dim = 2 + seed % 2
    if seed < 2:
        mesh = mg.kuhn_box(dim, int(rng.integers(4, 7)), free_fraction=0.03, permute=True, jitter=0.15, seed=int(rng.integers(1, 10 ** 6)))
    else:
        mesh = mg.delaunay_cloud(dim, int(rng.integers(60, 120)), seed=int(rng.integers(1, 10 ** 6)), free_fraction=0.05)
    nn = mesh.n_nodes
    mesh.dir_val = np.ascontiguousarray((0.1 * rng.standard_normal((dim, nn)) * (mesh.dir_mask != 0)).reshape(-1))
    _, q_prev = mg.pspg_state(mesh)
    q_prev = q_prev + 0.05 * rng.standard_normal(q_prev.shape)
    rho, mu, dt = 1000.0, float(rng.choice([1e-3, 1.0])), float(rng.choice([1e-3, 1e-2]))
    body = mg.gravity(dim)
    g = float(rng.choice([0.0, 0.5, 1.0]))
    par = orc.pspg_param_array(rho, mu, dt, body)
    dp = 30.0 * rng.standard_normal(nn)
    with ref.RefCase(mesh, "pspg", np.concatenate([par, [10, 1e-6, g, 2.0]]), solver_id="FracStep") as rc:
        rc.set_states(q_prev)
        A0, b0 = rc.fs_build(0, q_prev[: dim * nn], q_prev[dim * nn:])
        vt, it0, _, info0 = rc.fs_solve(0)
        A1, b1 = rc.fs_build(1, vt, q_prev[dim * nn:])
        A2, b2 = rc.fs_build(2, dp)
        dv, it2, _, info2 = rc.fs_solve(2)
    O0, c0 = fs.velocity_prediction(mesh, q_prev[: dim * nn], q_prev[dim * nn:], rho, mu, dt, body, g)
    O1, c1 = fs.pressure(mesh, vt, q_prev[dim * nn:], rho, mu, dt, body, g)
    O2, c2 = fs.velocity_correction(mesh, dp, rho, mu, dt, body)
    for O, c, A, b in ((O0, c0, A0, b0), (O1, c1, A1, b1), (O2, c2, A2, b2)):
        assert _rel(O, A.toarray()) < TOL and _rel(c, b) < TOL
    x, it, _, ok = fs.conjugate_gradient(O0, c0)
    assert ok and info0 == 0 and abs(it - it0) <= 1 and _rel(x, vt) < 1e-11
    x, it, _, ok = fs.conjugate_gradient(O2, c2)
    assert ok and info2 == 0 and abs(it - it2) <= 1 and _rel(x, dv) < 1e-11
