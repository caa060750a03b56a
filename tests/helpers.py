"""Shared helpers for the parity tests."""
import numpy as np

from oracle import oracle as orc
from pfem_b200 import meshgen as mg

TOL_AB = 1e-12   # north_star: assembled matrices and RHS to 1e-12 relative (per block type)
TOL_Q = 1e-8     # north_star: per-step fields to 1e-8 relative


def body_force(dim):
    return mg.gravity(dim)


def pspg_case(dim, n, **kw):
    mesh = mg.kuhn_box(dim, n, **kw)
    q, q_prev = mg.pspg_state(mesh)
    rng = np.random.default_rng(5)
    q_prev = q_prev + 0.02 * rng.standard_normal(q_prev.shape)   # v_prev != v so that both enter distinctly
    # non-zero Dirichlet data on part of the walls exercises the column elimination
    vals = mesh.dir_val.reshape(dim, mesh.n_nodes)
    sel = (mesh.dir_mask != 0) & (np.arange(mesh.n_nodes) % 3 == 0)
    vals[:, sel] = 0.1 * rng.standard_normal((dim, int(sel.sum())))
    mesh.dir_val = vals.reshape(-1)
    # a bound node whose tag carries no velocity BC (reference hazard 11)
    bound = np.flatnonzero(mesh.dir_mask)
    if bound.size:
        mesh.dir_mask[bound[::7]] = 0
    P = mg.PSPG_PARAMS
    par = orc.pspg_param_array(P["rho"], P["mu"], P["dt"], body_force(dim))
    return mesh, q, q_prev, par


def block_errors(A, A_ref, n_nodes, dim):
    """max|dA|/max|A_ref| per block type (vv, vp, pv, pp): the blocks differ in scale by up to 1e13."""
    assert (A.indptr == A_ref.indptr).all(), "CSC column pointers differ"
    assert (A.indices == A_ref.indices).all(), "CSC row indices differ"
    cols = np.repeat(np.arange(A_ref.shape[1]), np.diff(A_ref.indptr))
    rows = A_ref.indices
    rp = rows >= dim * n_nodes
    cp = cols >= dim * n_nodes
    out = {}
    for name, sel in (("vv", ~rp & ~cp), ("vp", ~rp & cp), ("pv", rp & ~cp), ("pp", rp & cp)):
        if sel.any():
            ref = np.abs(A_ref.data[sel]).max()
            out[name] = float(np.abs(A.data[sel] - A_ref.data[sel]).max() / (ref if ref > 0 else 1.0))
    return out


def vec_block_errors(b, b_ref, n_nodes, dim):
    out = {}
    for name, sl in (("v", slice(0, dim * n_nodes)), ("p", slice(dim * n_nodes, None))):
        ref = np.abs(b_ref[sl]).max()
        out[name] = float(np.abs(b[sl] - b_ref[sl]).max() / (ref if ref > 0 else 1.0))
    return out


def rel_err(a, ref):
    m = np.abs(ref).max()
    return float(np.abs(a - ref).max() / (m if m > 0 else 1.0))


# ---- golden fixtures generated from the reference's own code (tests/golden/make_golden.py) ----
import glob as _glob
import os as _os

GOLDEN_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden")


def golden_names(prefix):
    return sorted(_os.path.basename(p)[:-4] for p in _glob.glob(_os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load_golden(name):
    """Returns (mesh, dict of arrays) of one fixture."""
    z = dict(np.load(_os.path.join(GOLDEN_DIR, name + ".npz")))
    mesh = None
    if "conn" in z:
        mesh = mg.Mesh(dim=int(z["dim"]), x=np.ascontiguousarray(z["x"]), conn=np.ascontiguousarray(z["conn"]),
                       flags=np.ascontiguousarray(z["flags"]), dir_mask=np.ascontiguousarray(z["dir_mask"]),
                       dir_val=np.ascontiguousarray(z["dir_val"]))
    return mesh, z


def golden_csc(z, key="A"):
    import scipy.sparse as sp
    n = z["indptr"].shape[0] - 1
    return sp.csc_matrix((z[key], z["indices"], z["indptr"]), shape=(n, n))


def split_wc(q, dim, nn):
    return dict(v=q[: dim * nn], p=q[dim * nn:(dim + 1) * nn], rho=q[(dim + 1) * nn:(dim + 2) * nn], acc=q[(dim + 2) * nn:])


def wavy_free_surface(mesh, amp=0.2, seed=17):
    """Displace the free-surface nodes (flag bit 3) normal to the box top by U(-amp*h, amp*h): a flat surface has zero
    net surface-tension force on every interior node, so the facet terms need curvature to show up."""
    dim, nn = mesh.dim, mesh.n_nodes
    h = 1.0 / max(1, round((nn ** (1.0 / dim)) - 1))
    fs = np.flatnonzero(mesh.flags & mg.F_FS)
    x = mesh.x.reshape(dim, nn).copy()
    x[dim - 1, fs] += amp * h * np.random.default_rng(seed).uniform(-1, 1, fs.size)
    mesh.x = np.ascontiguousarray(x.reshape(-1))
    assert mg.det_j(mesh).min() > 0
    return mesh
