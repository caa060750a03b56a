"""ctypes front-end of the CPU oracle (oracle/pfem_oracle.cpp) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module.  Pinned against the reference's own sources built by oracle/refbuild (tests/golden/,
tests/test_oracle_vs_reference.py); see the header of pfem_oracle.cpp for the caveat.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpfem_oracle.so")
_lib = None

PHASES = ["Prepare matrix assembly", "Compute triplets", "Push back (n, n, 1)", "Assemble matrix",
          "Assemble vector", "Apply boundary conditions"]  # timer names of PSPG.inl:19-145, 276


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pfem_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        dp, ip, bp, i64 = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_uint8), C.c_int64
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
        L.oracle_pspg_elements.argtypes = [C.c_int, i64, i64, ip, dp, dp, dp, dp, dp, dp, dp]
        L.oracle_pspg_build.restype = C.c_void_p
        L.oracle_pspg_build.argtypes = [C.c_int, i64, i64, ip, dp, dp, dp, bp, dp, C.c_int, bp, dp, dp, dp]
        L.oracle_set_facets.argtypes = [C.c_int, i64, ip, C.c_double]
        L.oracle_set_pspg_thermal.argtypes = [C.c_int, C.c_double, C.c_double, dp]
        L.oracle_in_heat_build.restype = C.c_void_p
        L.oracle_in_heat_build.argtypes = [C.c_int, i64, i64, ip, dp, bp, bp, dp, dp, dp, C.c_int, dp]
        L.oracle_set_bingham.argtypes = [C.c_int, C.c_double, C.c_double]
        L.oracle_set_thermal.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, dp, bp, dp]
        L.oracle_csc_nnz.restype = i64
        L.oracle_csc_nnz.argtypes = [C.c_void_p]
        L.oracle_csc_copy.argtypes = [C.c_void_p, ip, C.POINTER(C.c_int32), dp]
        L.oracle_csc_free.argtypes = [C.c_void_p]
        L.oracle_move_positions.argtypes = [C.c_int, i64, bp, dp, dp, dp]
        L.oracle_wc_step.argtypes = [C.c_int, i64, i64, ip, dp, bp, bp, dp, dp, dp, dp, dp, dp, C.c_double]
        L.oracle_wc_next_dt.restype = C.c_double
        L.oracle_wc_next_dt.argtypes = [C.c_int, i64, i64, ip, dp, dp, dp, dp, dp, C.c_double, C.c_double]
        L.oracle_csc_matvec.argtypes = [i64, ip, C.POINTER(C.c_int32), dp, dp, dp]
        L.oracle_bicgstab_csr.restype = C.c_int
        L.oracle_bicgstab_csr.argtypes = [i64, ip, C.POINTER(C.c_int32), dp, dp, dp, C.c_double, C.c_int, dp]
        _lib = L
    return _lib


def _d(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    assert a.dtype == np.int64 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _i32(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _b(a):
    assert a.dtype == np.uint8 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def pspg_param_array(rho, mu, dt, body_force):
    bf = np.zeros(3)
    bf[: len(body_force)] = body_force
    return np.array([rho, mu, dt, bf[0], bf[1], bf[2]], dtype=np.float64)


EQ_TYPES = {"CDS_dpdt": 0, "CDS_drhodt": 1, "CDS_rho": 2}   # solver ids, ContEquation.inl:38-43


def wc_param_array(mu, K0, K0p, rhoStar, body_force, meduri=True, eq_type="CDS_dpdt"):
    bf = np.zeros(3)
    bf[: len(body_force)] = body_force
    return np.array([mu, K0, K0p, rhoStar, bf[0], bf[1], bf[2], 1.0 if meduri else 0.0,
                     float(EQ_TYPES[eq_type])], dtype=np.float64)


def set_facets(dim, facets=None, gamma=0.0):
    """Boundary facets (nF x (dim+2): facet nodes, out node, element) and surface tension for the following
    pspg_build(apply_bc=True) / wc_step calls; set_facets(dim) switches the facet terms off again."""
    if facets is None or gamma == 0.0:
        lib().oracle_set_facets(dim, 0, None, 0.0)
        return
    f = np.ascontiguousarray(facets, dtype=np.int64)
    lib().oracle_set_facets(dim, f.shape[0], _i(f), float(gamma))


_pspg_thermal_T = None   # keeps the array alive while the context points at it


def set_pspg_thermal(alpha=None, Tr=0.0, T=None):
    """Incompressible Boussinesq buoyancy factors (MomContEquation.inl:166-199) for the following pspg_* calls; () = off."""
    global _pspg_thermal_T
    if alpha is None:
        _pspg_thermal_T = None
        lib().oracle_set_pspg_thermal(0, 0.0, 0.0, None)
    else:
        _pspg_thermal_T = np.ascontiguousarray(T, dtype=np.float64)
        lib().oracle_set_pspg_thermal(1, float(alpha), float(Tr), _d(_pspg_thermal_T))


def in_heat_build(mesh, theta_prev, rho, cv, k, dt, t_mask, t_val, apply_bc=True, x=None):
    """HeatEqIncompNewton::m_buildAb (+ m_applyBC), IncompNewton/HeatEquation.inl:227-412: (A as CSC, b)."""
    L = lib()
    nn = mesh.n_nodes
    b = np.zeros(nn)
    x = mesh.x if x is None else x
    par = np.array([rho, cv, k, dt], dtype=np.float64)
    h = L.oracle_in_heat_build(mesh.dim, nn, mesh.n_elems, _i(mesh.conn), _d(x), _b(mesh.flags),
                               _b(np.ascontiguousarray(t_mask, dtype=np.uint8)), _d(np.ascontiguousarray(t_val, dtype=np.float64)),
                               _d(np.ascontiguousarray(theta_prev, dtype=np.float64)), _d(par), 1 if apply_bc else 0, _d(b))
    assert h
    nnz = L.oracle_csc_nnz(h)
    col_ptr, row_idx, val = np.zeros(nn + 1, dtype=np.int64), np.zeros(nnz, dtype=np.int32), np.zeros(nnz)
    L.oracle_csc_copy(h, _i(col_ptr), _i32(row_idx), _d(val))
    L.oracle_csc_free(h)
    return sp.csc_matrix((val, row_idx, col_ptr), shape=(nn, nn)), b


def set_bingham(tau0=None, m_reg=0.0):
    """Bingham regularised viscosity (MomContEquation.inl:102-119) for the following pspg_* calls; set_bingham() = off."""
    if tau0 is None:
        lib().oracle_set_bingham(0, 0.0, 0.0)
    else:
        lib().oracle_set_bingham(1, float(tau0), float(m_reg))


def pspg_elements(mesh, vcur, q_prev, params):
    """Per-element (Ae, be, tau): MB.inl + PSPG.inl:26-53."""
    nt = (mesh.dim + 1) ** 2
    ne = mesh.n_elems
    Ae = np.zeros((ne, nt, nt))
    be = np.zeros((ne, nt))
    tau = np.zeros(ne)
    rc = lib().oracle_pspg_elements(mesh.dim, mesh.n_nodes, ne, _i(mesh.conn), _d(mesh.x), _d(vcur), _d(q_prev),
                                    _d(params), _d(Ae), _d(be), _d(tau))
    assert rc == 0
    return Ae, be, tau


def pspg_build(mesh, vcur, q_prev, params, apply_bc=True, x=None, phase_sec=None):
    """m_buildAbPSPG (+ m_applyBCPSPG): returns (A as scipy CSC with explicit zeros kept, b)."""
    L = lib()
    n_dof = (mesh.dim + 1) * mesh.n_nodes
    b = np.zeros(n_dof)
    x = mesh.x if x is None else x
    ph = np.zeros(6) if phase_sec is None else phase_sec
    h = L.oracle_pspg_build(mesh.dim, mesh.n_nodes, mesh.n_elems, _i(mesh.conn), _d(x), _d(vcur), _d(q_prev),
                            _b(mesh.flags), _d(params), 1 if apply_bc else 0, _b(mesh.dir_mask), _d(mesh.dir_val),
                            _d(b), _d(ph))
    assert h
    nnz = L.oracle_csc_nnz(h)
    col_ptr = np.zeros(n_dof + 1, dtype=np.int64)
    row_idx = np.zeros(nnz, dtype=np.int32)
    val = np.zeros(nnz)
    L.oracle_csc_copy(h, _i(col_ptr), _i32(row_idx), _d(val))
    L.oracle_csc_free(h)
    A = sp.csc_matrix((val, row_idx, col_ptr), shape=(n_dof, n_dof))  # no sum_duplicates/eliminate_zeros
    return A, b


def move_positions(mesh, delta, base):
    x = base.copy()
    lib().oracle_move_positions(mesh.dim, mesh.n_nodes, _b(mesh.flags), _d(np.ascontiguousarray(delta)), _d(base), _d(x))
    return x


def _set_thermal(thermal, T):
    """thermal: None or dict(k, cv, alpha, Tr, t_mask uint8[nNodes], t_val float64[nNodes]) -- BoussinesqWC."""
    if thermal is None:
        lib().oracle_set_thermal(0, 0.0, 1.0, 0.0, 0.0, None, None, None)
    else:
        lib().oracle_set_thermal(1, float(thermal["k"]), float(thermal["cv"]), float(thermal["alpha"]), float(thermal["Tr"]),
                                 _d(T), _b(thermal["t_mask"]), _d(thermal["t_val"]))


def wc_step(mesh, x, state, params, dt, thermal=None):
    """One explicit step (WC/Solver.cpp:236-263; with `thermal` and state["T"]: m_solveBoussinesqWC, :278-320).
    Returns new (x, state) copies."""
    x = x.copy()
    st = {k: v.copy() for k, v in state.items()}
    _set_thermal(thermal, st.get("T"))
    rc = lib().oracle_wc_step(mesh.dim, mesh.n_nodes, mesh.n_elems, _i(mesh.conn), _d(x), _b(mesh.flags),
                              _b(mesh.dir_mask), _d(mesh.dir_val), _d(st["v"]), _d(st["acc"]), _d(st["p"]),
                              _d(st["rho"]), _d(params), float(dt))
    _set_thermal(None, None)
    assert rc == 0
    return x, st


def wc_next_dt(mesh, x, state, params, security_coeff, max_dt, thermal=None):
    if thermal is not None:
        _set_thermal(thermal, state["T"])
        try:
            return lib().oracle_wc_next_dt(mesh.dim, mesh.n_nodes, mesh.n_elems, _i(mesh.conn), _d(x), _d(state["v"]),
                                           _d(state["p"]), _d(state["rho"]), _d(params), float(security_coeff), float(max_dt))
        finally:
            _set_thermal(None, None)
    return lib().oracle_wc_next_dt(mesh.dim, mesh.n_nodes, mesh.n_elems, _i(mesh.conn), _d(x), _d(state["v"]),
                                   _d(state["p"]), _d(state["rho"]), _d(params), float(security_coeff), float(max_dt))


def bicgstab(A_csr, b, tol=1e-12, max_iter=10000, x0=None):
    A_csr = A_csr.tocsr()
    n = A_csr.shape[0]
    x = np.zeros(n) if x0 is None else x0.copy()
    rr = np.zeros(1)
    it = lib().oracle_bicgstab_csr(n, _i(A_csr.indptr.astype(np.int64)), _i32(A_csr.indices.astype(np.int32)),
                                   _d(np.ascontiguousarray(A_csr.data)), _d(b), _d(x), tol, max_iter, _d(rr))
    return x, it, float(rr[0])


def num_threads():
    return lib().oracle_num_threads()


# --------------------------------------------------------------------------------------
# Picard driver (PicardAlgo.cpp:31-94 + PSPG.inl:262-373) with scipy SuperLU standing in for
# Eigen::SparseLU (COLAMD).  residual = "Ax_f" (absolute ||A q - b||_2, PSPG.inl:368).
# --------------------------------------------------------------------------------------
def pspg_picard(mesh, q_state, q_prev, params, max_iter=10, min_res=1e-6, direct_solve=None):
    import scipy.sparse.linalg as spla

    dim, nn = mesh.dim, mesh.n_nodes
    dt = params[2]
    if direct_solve is None:
        def direct_solve(A, b):
            return spla.splu(A.tocsc(), permc_spec="COLAMD").solve(b)
    x_save = mesh.x.copy()                       # Mesh::saveNodesList
    vcur = q_state[: dim * nn].copy()            # node states (used by tau)
    A, b = pspg_build(mesh, vcur, q_prev, params, True, x=x_save)   # m_prepare
    q = np.zeros_like(q_prev)
    res = np.finfo(np.float64).max
    it = 0
    x = x_save
    history = []
    while res > min_res:
        if it > max_iter:
            return dict(ok=False, q=q, x=x_save, iters=it, res=res, history=history)
        q = direct_solve(A, b)
        vcur = q[: dim * nn].copy()              # setNodesStatesfromQ (PSPG.inl:293)
        x = move_positions(mesh, q[: dim * nn] * dt, x_save)  # updateNodesPositionFromSave (PSPG.inl:294-295)
        A, b = pspg_build(mesh, vcur, q_prev, params, True, x=x)
        res = float(np.linalg.norm(A @ q - b))
        history.append(res)
        if np.isnan(res):
            return dict(ok=False, q=q, x=x_save, iters=it, res=res, history=history)
        it += 1
    return dict(ok=True, q=q, x=x, iters=it, res=res, history=history, A=A, b=b)
