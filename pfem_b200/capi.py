"""ctypes binding of include/pfem_b200.h -> pfem_b200/csrc/libpfem_b200.so.

This is the only way Python reaches the CUDA path; there is no CPU fallback.  If the shared
library is missing or no CUDA device is usable, construction raises -- loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpfem_b200.so")
_lib = None

PFEM_OK, PFEM_NOT_CONVERGED, PFEM_NAN = 0, 1, 2


class PfemError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pfem_b200 error {code}: {msg}")
        self.code = code


class PspgParams(C.Structure):
    _fields_ = [("rho", C.c_double), ("mu", C.c_double), ("dt", C.c_double), ("bodyForce", C.c_double * 3)]


class WcParams(C.Structure):
    _fields_ = [("mu", C.c_double), ("K0", C.c_double), ("K0p", C.c_double), ("rhoStar", C.c_double),
                ("bodyForce", C.c_double * 3), ("meduri", C.c_int32), ("eqType", C.c_int32)]


class ThermalParams(C.Structure):
    _fields_ = [("k", C.c_double), ("cv", C.c_double), ("alpha", C.c_double), ("Tr", C.c_double)]


class Info(C.Structure):
    _fields_ = [("dim", C.c_int32), ("device", C.c_int32), ("nRanks", C.c_int32), ("rank", C.c_int32),
                ("nNodes", C.c_int64), ("nElems", C.c_int64), ("nDof", C.c_int64), ("nnzBlocks", C.c_int64),
                ("nnzReference", C.c_int64), ("deviceBytes", C.c_int64), ("maxElemsPerNode", C.c_int32),
                ("maxNeighbours", C.c_int32)]


# name -> (restype, argtypes); every symbol include/pfem_b200.h declares
_VP, _DP, _U8P, _I32P, _I64P = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
SYMBOLS = {
    "pfem_abi_version": (C.c_int, []),
    "pfem_create": (C.c_int, [C.POINTER(_VP), C.c_int, C.c_int]),
    "pfem_destroy": (C.c_int, [_VP]),
    "pfem_last_error": (C.c_char_p, [_VP]),
    "pfem_set_stream": (C.c_int, [_VP, _VP]),
    "pfem_get_info": (C.c_int, [_VP, C.POINTER(Info)]),
    "pfem_set_topology": (C.c_int, [_VP, C.c_int64, C.c_int64, C.POINTER(C.c_uint64), _U8P]),
    "pfem_set_positions": (C.c_int, [_VP, _DP]),
    "pfem_get_positions": (C.c_int, [_VP, _DP]),
    "pfem_snapshot_positions": (C.c_int, [_VP]),
    "pfem_restore_positions": (C.c_int, [_VP]),
    "pfem_move_positions": (C.c_int, [_VP, _DP, C.c_int]),
    "pfem_set_states": (C.c_int, [_VP, C.c_int, C.c_int, _DP]),
    "pfem_get_states": (C.c_int, [_VP, C.c_int, C.c_int, _DP]),
    "pfem_set_dirichlet": (C.c_int, [_VP, _U8P, _DP]),
    "pfem_set_facets": (C.c_int, [_VP, C.c_int64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "pfem_set_surface_tension": (C.c_int, [_VP, C.c_double]),
    "pfem_set_thermal": (C.c_int, [_VP, C.POINTER(ThermalParams)]),
    "pfem_set_bingham": (C.c_int, [_VP, C.c_int, C.c_double, C.c_double]),
    "pfem_set_temperature": (C.c_int, [_VP, _DP]),
    "pfem_get_temperature": (C.c_int, [_VP, _DP]),
    "pfem_set_temperature_bc": (C.c_int, [_VP, _U8P, _DP]),
    "pfem_heat_assemble": (C.c_int, [_VP, C.c_double, C.c_double, C.c_double, C.c_double, _DP]),
    "pfem_heat_solve": (C.c_int, [_VP, C.c_double, C.c_int, _DP, C.POINTER(C.c_int), _DP]),
    "pfem_heat_export_csc": (C.c_int, [_VP, _I64P, _I32P, _I32P, _DP, _DP]),
    "pfem_fs_assemble_vapp": (C.c_int, [_VP, C.POINTER(PspgParams), C.c_double, _DP]),
    "pfem_fs_assemble_pcorr": (C.c_int, [_VP, C.c_double, C.c_double, C.c_double, _DP, _DP]),
    "pfem_fs_assemble_vcorr": (C.c_int, [_VP, C.c_double, C.c_double, _DP]),
    "pfem_fs_get_rhs": (C.c_int, [_VP, _DP]),
    "pfem_fs_solve": (C.c_int, [_VP, C.c_double, C.c_int, _DP, C.POINTER(C.c_int), _DP]),
    "pfem_pspg_set_qprev": (C.c_int, [_VP, _DP]),
    "pfem_pspg_assemble": (C.c_int, [_VP, C.POINTER(PspgParams), _DP]),
    "pfem_pspg_assemble_resident": (C.c_int, [_VP, C.POINTER(PspgParams)]),
    "pfem_pspg_solve": (C.c_int, [_VP, C.c_double, C.c_int, _DP, C.POINTER(C.c_int), _DP]),
    "pfem_pspg_set_preconditioner": (C.c_int, [_VP, C.c_int, C.c_int, C.c_double]),
    "pfem_pspg_get_preconditioner": (C.c_int, [_VP, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pfem_pspg_residual": (C.c_int, [_VP, _DP, _DP]),
    "pfem_pspg_picard_iter": (C.c_int, [_VP, C.POINTER(PspgParams), _DP, C.c_double, C.c_int, _DP, _DP, C.POINTER(C.c_int)]),
    "pfem_pspg_export_csc": (C.c_int, [_VP, _I64P, _I32P, _I32P, _DP, _DP]),
    "pfem_pspg_matvec": (C.c_int, [_VP, _DP, _DP]),
    "pfem_wc_step": (C.c_int, [_VP, C.POINTER(WcParams), C.c_double]),
    "pfem_wc_set_variant": (C.c_int, [_VP, C.c_int]),
    "pfem_wc_next_dt": (C.c_int, [_VP, C.POINTER(WcParams), C.c_double, C.c_double, _DP]),
    "pfem_wc_run": (C.c_int, [_VP, C.POINTER(WcParams), C.c_int, C.c_double, C.c_double, _DP, _DP]),
    "pfem_comm_unique_id": (C.c_int, [_VP]),
    "pfem_comm_init": (C.c_int, [_VP, C.c_int, C.c_int, _VP]),
    "pfem_set_partition": (C.c_int, [_VP, C.c_int64, C.c_int, _I32P, _I64P, _I32P, _I64P, _I64P]),
    "pfem_comm_local_create": (C.c_int, [C.c_int, C.POINTER(_VP)]),
    "pfem_comm_local_destroy": (C.c_int, [_VP]),
    "pfem_comm_init_local": (C.c_int, [_VP, _VP, C.c_int]),
    "pfem_comm_abort": (C.c_int, [_VP]),
    "pfem_partition_create": (C.c_int, [C.POINTER(_VP), C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_uint64), _DP, C.c_int]),
    "pfem_partition_destroy": (C.c_int, [_VP]),
    "pfem_partition_owner": (C.c_int, [_VP, _I32P]),
    "pfem_partition_local_sizes": (C.c_int, [_VP, C.c_int, _I64P, _I64P, _I64P, _I32P, _I64P]),
    "pfem_partition_local_get": (C.c_int, [_VP, C.c_int, _I64P, _I64P, C.POINTER(C.c_uint64), _I32P, _I64P, _I32P, _I64P, _I64P]),
    "pfem_profile_enable": (C.c_int, [_VP, C.c_int]),
    "pfem_profile_reset": (C.c_int, [_VP]),
    "pfem_profile_get": (C.c_int, [_VP, C.c_char_p, _DP, _I64P]),
    "pfem_launch_count": (C.c_int, [_VP, _I64P]),
}


def load_library(path: str | None = None):
    """dlopen libpfem_b200.so and type every exported symbol.  Raises if the library is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise PfemError(-100, f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
    L = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)  # AttributeError if the header and the library disagree
        fn.restype, fn.argtypes = res, args
    if path is None:
        _lib = L
    return L


def _dptr(a):
    return a.ctypes.data_as(_DP)


def _f64(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} doubles, got {a.size}")
    return a


class NativePartition:
    """RCB partition + ghost layer + halo plan computed by the library (pfem_partition_*, csrc/partition.cu)."""

    def __init__(self, dim, conn, x, n_ranks):
        self._L = load_library()
        self._h = _VP()
        conn = np.ascontiguousarray(conn, dtype=np.uint64)
        x = _f64(x)
        self.dim, self.n_ranks = dim, n_ranks
        self._keep = (conn, x)  # the handle reads the caller's connectivity until it is destroyed
        self.n_nodes, self.n_elems = x.size // dim, conn.shape[0]
        rc = self._L.pfem_partition_create(C.byref(self._h), dim, self.n_nodes, self.n_elems,
                                           conn.ctypes.data_as(C.POINTER(C.c_uint64)), _dptr(x), n_ranks)
        if rc != 0:
            raise PfemError(rc, "pfem_partition_create failed")

    def owner(self):
        o = np.empty(self.n_nodes, dtype=np.int32)
        self._L.pfem_partition_owner(self._h, o.ctypes.data_as(_I32P))
        return o

    def local(self, rank):
        nl, no, ne, ns = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
        npeers = C.c_int32(0)
        rc = self._L.pfem_partition_local_sizes(self._h, rank, C.byref(nl), C.byref(no), C.byref(ne), C.byref(npeers), C.byref(ns))
        if rc != 0:
            raise PfemError(rc, "pfem_partition_local_sizes failed")
        npe = self.dim + 1
        l2g_nodes = np.empty(nl.value, dtype=np.int64)
        l2g_elems = np.empty(ne.value, dtype=np.int64)
        conn = np.empty((ne.value, npe), dtype=np.uint64)
        peers = np.empty(npeers.value, dtype=np.int32)
        send_off = np.zeros(npeers.value + 1, dtype=np.int64)
        send_idx = np.empty(ns.value, dtype=np.int32)
        recv_start = np.empty(npeers.value, dtype=np.int64)
        recv_count = np.empty(npeers.value, dtype=np.int64)
        rc = self._L.pfem_partition_local_get(self._h, rank, l2g_nodes.ctypes.data_as(_I64P), l2g_elems.ctypes.data_as(_I64P),
                                              conn.ctypes.data_as(C.POINTER(C.c_uint64)), peers.ctypes.data_as(_I32P),
                                              send_off.ctypes.data_as(_I64P), send_idx.ctypes.data_as(_I32P),
                                              recv_start.ctypes.data_as(_I64P), recv_count.ctypes.data_as(_I64P))
        if rc != 0:
            raise PfemError(rc, "pfem_partition_local_get failed")
        return dict(n_owned=no.value, l2g_nodes=l2g_nodes, l2g_elems=l2g_elems, conn=conn, peers=peers, send_offsets=send_off,
                    send_idx=send_idx, recv_start=recv_start, recv_count=recv_count)

    def close(self):
        if getattr(self, "_h", None) and self._h:
            self._L.pfem_partition_destroy(self._h)
            self._h = _VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LocalGroup:
    """In-process communicator: n_ranks contexts of this process, each driven by its own thread (pfem_comm_local_*)."""

    def __init__(self, n_ranks):
        self._L = load_library()
        self._h = _VP()
        rc = self._L.pfem_comm_local_create(n_ranks, C.byref(self._h))
        if rc != 0:
            raise PfemError(rc, (self._L.pfem_last_error(None) or b"").decode())
        self.n_ranks = n_ranks

    def close(self):
        if getattr(self, "_h", None) and self._h:
            self._L.pfem_comm_local_destroy(self._h)
            self._h = _VP()


class PfemContext:
    """Thin RAII wrapper of pfem_ctx; method names follow the C ABI."""

    def __init__(self, dim: int, device: int = 0):
        self._L = load_library()
        self._h = _VP()
        self.dim = dim
        rc = self._L.pfem_create(C.byref(self._h), dim, device)
        if rc != 0:
            msg = self._L.pfem_last_error(self._h)  # NULL ctx -> the creation error
            raise PfemError(rc, (msg or b"").decode())
        self.n_nodes = 0
        self.n_elems = 0

    # -- plumbing -----------------------------------------------------------------
    def _chk(self, rc, allow=()):
        if rc != 0 and rc not in allow:
            raise PfemError(rc, (self._L.pfem_last_error(self._h) or b"").decode())
        return rc

    def close(self):
        if getattr(self, "_h", None) and self._h:
            self._L.pfem_destroy(self._h)
            self._h = _VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def n_dof(self):
        return (self.dim + 1) * self.n_nodes

    def set_stream(self, cuda_stream: int):
        self._chk(self._L.pfem_set_stream(self._h, _VP(cuda_stream)))

    def info(self) -> Info:
        i = Info()
        self._chk(self._L.pfem_get_info(self._h, C.byref(i)))
        return i

    # -- mesh ---------------------------------------------------------------------
    def set_topology(self, conn, flags):
        conn = np.ascontiguousarray(conn, dtype=np.uint64)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        assert conn.ndim == 2 and conn.shape[1] == self.dim + 1
        self.n_nodes, self.n_elems = int(flags.size), int(conn.shape[0])
        self._chk(self._L.pfem_set_topology(self._h, self.n_nodes, self.n_elems,
                                            conn.ctypes.data_as(C.POINTER(C.c_uint64)), flags.ctypes.data_as(_U8P)))

    def set_mesh(self, mesh):
        """Convenience: topology + positions + Dirichlet data of a meshgen.Mesh."""
        self.set_topology(mesh.conn, mesh.flags)
        self.set_positions(mesh.x)
        self.set_dirichlet(mesh.dir_mask, mesh.dir_val)

    def set_positions(self, x):
        x = _f64(x, self.dim * self.n_nodes)
        self._chk(self._L.pfem_set_positions(self._h, _dptr(x)))

    def get_positions(self):
        x = np.empty(self.dim * self.n_nodes)
        self._chk(self._L.pfem_get_positions(self._h, _dptr(x)))
        return x

    def snapshot_positions(self):
        self._chk(self._L.pfem_snapshot_positions(self._h))

    def restore_positions(self):
        self._chk(self._L.pfem_restore_positions(self._h))

    def move_positions(self, delta, from_snapshot=False):
        d = _f64(delta, self.dim * self.n_nodes)
        self._chk(self._L.pfem_move_positions(self._h, _dptr(d), 1 if from_snapshot else 0))

    def set_states(self, first, q):
        q = _f64(q)
        count, rem = divmod(q.size, self.n_nodes)
        assert rem == 0
        self._chk(self._L.pfem_set_states(self._h, first, count, _dptr(q)))

    def get_states(self, first, count):
        q = np.empty(count * self.n_nodes)
        self._chk(self._L.pfem_get_states(self._h, first, count, _dptr(q)))
        return q

    def set_dirichlet(self, mask, values):
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        values = _f64(values, self.dim * self.n_nodes)
        self._chk(self._L.pfem_set_dirichlet(self._h, mask.ctypes.data_as(_U8P), _dptr(values)))

    def set_facets(self, facets):
        """facets: (nF, dim+2) rows [facet nodes, out node, element index] (meshgen.boundary_facets); None clears."""
        if facets is None or len(facets) == 0:
            self._chk(self._L.pfem_set_facets(self._h, 0, None, None, None))
            return
        f = np.ascontiguousarray(facets, dtype=np.uint64)
        d = self.dim
        nodes = np.ascontiguousarray(f[:, :d])
        out, elem = np.ascontiguousarray(f[:, d]), np.ascontiguousarray(f[:, d + 1])
        u64 = C.POINTER(C.c_uint64)
        self._chk(self._L.pfem_set_facets(self._h, f.shape[0], nodes.ctypes.data_as(u64), out.ctypes.data_as(u64),
                                          elem.ctypes.data_as(u64)))

    def set_surface_tension(self, gamma):
        self._chk(self._L.pfem_set_surface_tension(self._h, float(gamma)))

    # -- Boussinesq / Bingham / heat ---------------------------------------------------
    def set_thermal(self, k=None, cv=1.0, alpha=0.0, Tr=0.0):
        """Boussinesq constants; k=None switches the factors off."""
        if k is None:
            self._chk(self._L.pfem_set_thermal(self._h, None))
        else:
            t = ThermalParams(k, cv, alpha, Tr)
            self._chk(self._L.pfem_set_thermal(self._h, C.byref(t)))

    def set_bingham(self, tau0=None, m_reg=0.0):
        self._chk(self._L.pfem_set_bingham(self._h, 0 if tau0 is None else 1, float(tau0 or 0.0), float(m_reg)))

    def set_temperature(self, T):
        T = _f64(T, self.n_nodes)
        self._chk(self._L.pfem_set_temperature(self._h, _dptr(T)))

    def get_temperature(self):
        T = np.empty(self.n_nodes)
        self._chk(self._L.pfem_get_temperature(self._h, _dptr(T)))
        return T

    def set_temperature_bc(self, mask, values):
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        values = _f64(values, self.n_nodes)
        self._chk(self._L.pfem_set_temperature_bc(self._h, mask.ctypes.data_as(_U8P), _dptr(values)))

    def heat_assemble(self, rho, cv, k, dt, theta_prev):
        th = _f64(theta_prev, self.n_nodes)
        self._chk(self._L.pfem_heat_assemble(self._h, rho, cv, k, dt, _dptr(th)))

    def heat_solve(self, rel_tol=1e-14, max_iter=10000, fetch=True):
        T = np.empty(self.n_nodes) if fetch else None
        it, rr = C.c_int(0), C.c_double(0)
        rc = self._chk(self._L.pfem_heat_solve(self._h, rel_tol, max_iter, _dptr(T) if fetch else None, C.byref(it), C.byref(rr)),
                       allow=(PFEM_NOT_CONVERGED, PFEM_NAN))
        return dict(status=rc, T=T, iters=it.value, rel_res=rr.value)

    def heat_export_csc(self):
        import scipy.sparse as sp
        nnz = C.c_int64(0)
        self._chk(self._L.pfem_heat_export_csc(self._h, C.byref(nnz), None, None, None, None))
        n = self.n_nodes
        col_ptr = np.empty(n + 1, dtype=np.int32)
        row_idx = np.empty(nnz.value, dtype=np.int32)
        val = np.empty(nnz.value)
        b = np.empty(n)
        self._chk(self._L.pfem_heat_export_csc(self._h, C.byref(nnz), col_ptr.ctypes.data_as(_I32P), row_idx.ctypes.data_as(_I32P),
                                               _dptr(val), _dptr(b)))
        return sp.csc_matrix((val, row_idx, col_ptr), shape=(n, n)), b

    # -- fractional-step solver (FracStep): the three systems of one Picard body -------
    def fs_assemble_vapp(self, params, gamma_fs, q_prev):
        q = _f64(q_prev, self.n_dof)
        self._chk(self._L.pfem_fs_assemble_vapp(self._h, C.byref(params), float(gamma_fs), _dptr(q)))

    def fs_assemble_pcorr(self, rho, dt, gamma_fs, v_tilde, p_prev):
        v = _f64(v_tilde, self.dim * self.n_nodes)
        pp = _f64(p_prev, self.n_nodes)
        self._chk(self._L.pfem_fs_assemble_pcorr(self._h, rho, dt, float(gamma_fs), _dptr(v), _dptr(pp)))

    def fs_assemble_vcorr(self, rho, dt, delta_p):
        dp = _f64(delta_p, self.n_nodes)
        self._chk(self._L.pfem_fs_assemble_vcorr(self._h, rho, dt, _dptr(dp)))

    def fs_get_rhs(self, which):
        b = np.empty((self.dim if which == 2 else 1) * self.n_nodes)
        self._chk(self._L.pfem_fs_get_rhs(self._h, _dptr(b)))
        return b

    def fs_solve(self, which, rel_tol=np.finfo(float).eps, max_iter=None):
        """Eigen-style Jacobi-CG on the system assembled last; which = 0 | 1 | 2 only sizes the result."""
        n = (1 if which == 1 else self.dim) * self.n_nodes
        x = np.empty(n)
        it, rr = C.c_int(0), C.c_double(0)
        rc = self._chk(self._L.pfem_fs_solve(self._h, float(rel_tol), int(max_iter or 2 * n), _dptr(x), C.byref(it), C.byref(rr)),
                       allow=(PFEM_NOT_CONVERGED, PFEM_NAN))
        return dict(status=rc, x=x, iters=it.value, rel_res=rr.value)

    # -- PSPG ---------------------------------------------------------------------
    @staticmethod
    def pspg_params(rho, mu, dt, body_force):
        p = PspgParams(rho, mu, dt)
        for i, v in enumerate(body_force[:3]):
            p.bodyForce[i] = v
        return p

    def pspg_set_qprev(self, q_prev):
        q = _f64(q_prev, self.n_dof)
        self._chk(self._L.pfem_pspg_set_qprev(self._h, _dptr(q)))

    def pspg_assemble(self, params, q_prev):
        q = _f64(q_prev, self.n_dof)
        self._chk(self._L.pfem_pspg_assemble(self._h, C.byref(params), _dptr(q)))

    def pspg_assemble_resident(self, params):
        self._chk(self._L.pfem_pspg_assemble_resident(self._h, C.byref(params)))

    def pspg_solve(self, rel_tol=1e-12, max_iter=10000, fetch=True):
        q = np.empty(self.n_dof) if fetch else None
        it, rr = C.c_int(0), C.c_double(0)
        rc = self._chk(self._L.pfem_pspg_solve(self._h, rel_tol, max_iter, _dptr(q) if fetch else None,
                                               C.byref(it), C.byref(rr)), allow=(PFEM_NOT_CONVERGED, PFEM_NAN))
        return dict(status=rc, q=q, iters=it.value, rel_res=rr.value)

    PRECOND = {"auto": 0, "point": 1, "block": 2, "mg": 3}

    def pspg_set_preconditioner(self, kind="auto", sweeps=0, damping=0.0):
        """kind: auto | point | block | mg (pfem_pspg_set_preconditioner)."""
        self._chk(self._L.pfem_pspg_set_preconditioner(self._h, self.PRECOND[kind], int(sweeps), float(damping)))

    def pspg_get_preconditioner(self):
        """(kind used by the last solve, multigrid levels)."""
        k, lv = C.c_int(0), C.c_int(0)
        self._chk(self._L.pfem_pspg_get_preconditioner(self._h, C.byref(k), C.byref(lv)))
        return {v: n for n, v in self.PRECOND.items()}[k.value], lv.value

    def pspg_residual(self, q=None):
        r = C.c_double(0)
        qq = None if q is None else _f64(q, self.n_dof)
        self._chk(self._L.pfem_pspg_residual(self._h, None if qq is None else _dptr(qq), C.byref(r)))
        return r.value

    def pspg_picard_iter(self, params, q_prev, rel_tol=1e-12, max_iter=10000, fetch=True):
        qp = None if q_prev is None else _f64(q_prev, self.n_dof)  # None: keep the qPrev already on the device
        q = np.empty(self.n_dof) if fetch else None
        res, it = C.c_double(0), C.c_int(0)
        rc = self._chk(self._L.pfem_pspg_picard_iter(self._h, C.byref(params), None if qp is None else _dptr(qp), rel_tol, max_iter,
                                                     _dptr(q) if fetch else None, C.byref(res), C.byref(it)),
                       allow=(PFEM_NOT_CONVERGED, PFEM_NAN))
        return dict(status=rc, q=q, res=res.value, iters=it.value)

    def pspg_export_csc(self):
        """(scipy.sparse.csc_matrix with the reference pattern incl. explicit zeros, b)."""
        import scipy.sparse as sp
        nnz = C.c_int64(0)
        self._chk(self._L.pfem_pspg_export_csc(self._h, C.byref(nnz), None, None, None, None))
        col_ptr = np.empty(self.n_dof + 1, dtype=np.int32)
        row_idx = np.empty(nnz.value, dtype=np.int32)
        val = np.empty(nnz.value)
        b = np.empty(self.n_dof)
        self._chk(self._L.pfem_pspg_export_csc(self._h, C.byref(nnz), col_ptr.ctypes.data_as(_I32P),
                                               row_idx.ctypes.data_as(_I32P), _dptr(val), _dptr(b)))
        A = sp.csc_matrix((val, row_idx, col_ptr), shape=(self.n_dof, self.n_dof))
        return A, b

    def pspg_reference_nnz(self) -> int:
        """nnz of the reference's m_A for the current system (count pass of the export only)."""
        nnz = C.c_int64(0)
        self._chk(self._L.pfem_pspg_export_csc(self._h, C.byref(nnz), None, None, None, None))
        return nnz.value

    def pspg_matvec(self, x):
        x = _f64(x, self.n_dof)
        y = np.empty(self.n_dof)
        self._chk(self._L.pfem_pspg_matvec(self._h, _dptr(x), _dptr(y)))
        return y

    # -- weakly compressible --------------------------------------------------------
    @staticmethod
    def wc_params(mu, K0, K0p, rhoStar, body_force, meduri=True, eq_type="CDS_dpdt"):
        p = WcParams(mu, K0, K0p, rhoStar)
        for i, v in enumerate(body_force[:3]):
            p.bodyForce[i] = v
        p.meduri = 1 if meduri else 0
        p.eqType = {"CDS_dpdt": 0, "CDS_drhodt": 1, "CDS_rho": 2}[eq_type]
        return p

    def wc_step(self, params, dt):
        self._chk(self._L.pfem_wc_step(self._h, C.byref(params), float(dt)))

    def wc_set_variant(self, variant):
        self._chk(self._L.pfem_wc_set_variant(self._h, int(variant)))

    def wc_next_dt(self, params, security_coeff, max_dt):
        dt = C.c_double(0)
        rc = self._chk(self._L.pfem_wc_next_dt(self._h, C.byref(params), security_coeff, max_dt, C.byref(dt)),
                       allow=(PFEM_NAN,))
        if rc == PFEM_NAN:
            raise PfemError(rc, "NaN time step!")  # WCompNewton/Solver.cpp:231-232
        return dt.value

    def wc_run(self, params, n_steps, security_coeff, max_dt, dt0):
        """n_steps explicit steps with the CFL dt chained on the device; returns (next dt, elapsed simulated time)."""
        dt, el = C.c_double(dt0), C.c_double(0)
        rc = self._chk(self._L.pfem_wc_run(self._h, C.byref(params), int(n_steps), security_coeff, max_dt, C.byref(dt), C.byref(el)),
                       allow=(PFEM_NAN,))
        if rc == PFEM_NAN:
            raise PfemError(rc, "NaN time step!")  # WCompNewton/Solver.cpp:231-232
        return dt.value, el.value

    # -- multi-GPU ------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._chk(self._L.pfem_comm_unique_id(C.cast(buf, _VP)))
        return buf.raw

    def comm_init(self, n_ranks, rank, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        self._chk(self._L.pfem_comm_init(self._h, n_ranks, rank, C.cast(buf, _VP)))

    def comm_init_local(self, group: "LocalGroup", rank: int):
        self._chk(self._L.pfem_comm_init_local(self._h, group._h, rank))

    def comm_abort(self):
        """Release the other ranks of a local group from their barriers after this rank failed outside the library."""
        self._L.pfem_comm_abort(self._h)

    def set_partition(self, part):
        """Halo plan of a partition.LocalPart whose local mesh was given to set_topology/set_mesh."""
        n_peers = len(part.peers)
        peer = np.asarray(part.peers, dtype=np.int32)
        counts = np.array([len(s) for s in part.send_idx], dtype=np.int64)
        send_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        send_idx = (np.concatenate(part.send_idx).astype(np.int32) if n_peers else np.zeros(0, dtype=np.int32))
        send_idx = np.ascontiguousarray(send_idx)
        recv_start = np.asarray(part.recv_start, dtype=np.int64)
        recv_count = np.asarray(part.recv_count, dtype=np.int64)
        self._chk(self._L.pfem_set_partition(self._h, part.n_owned, n_peers, peer.ctypes.data_as(_I32P),
                                             send_off.ctypes.data_as(_I64P), send_idx.ctypes.data_as(_I32P),
                                             recv_start.ctypes.data_as(_I64P), recv_count.ctypes.data_as(_I64P)))

    # -- instrumentation --------------------------------------------------------------
    def profile_enable(self, on=True):
        """on: False/0 off, True/1 phases, 2 also the per-kernel phases of the multigrid cycle (un-graphed)."""
        self._chk(self._L.pfem_profile_enable(self._h, int(on)))

    def profile_reset(self):
        self._chk(self._L.pfem_profile_reset(self._h))

    def profile_get(self, phase: str):
        ms, calls = C.c_double(0), C.c_int64(0)
        self._chk(self._L.pfem_profile_get(self._h, phase.encode(), C.byref(ms), C.byref(calls)))
        return ms.value, calls.value

    def launch_count(self) -> int:
        n = C.c_int64(0)
        self._chk(self._L.pfem_launch_count(self._h, C.byref(n)))
        return n.value
