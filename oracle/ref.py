"""ctypes front-end of oracle/_ref/libpfem_ref.so -- the REFERENCE'S OWN hot-path sources compiled in place from
/root/reference against stand-in Eigen/sol2/gmsh headers (oracle/refbuild/).  TEST INFRASTRUCTURE ONLY.

Used by tests/ and tests/golden/make_golden.py to pin oracle/pfem_oracle.cpp; never imported by pfem_b200/.
`available()` is False where neither the prebuilt library nor /root/reference exists.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libpfem_ref.so")
_SO_DROPIN = os.path.join(_HERE, "_ref", "libpfem_ref_dropin.so")   # + shim/pfem_b200_equations.hpp, linked to libpfem_b200.so
_lib = None
_use_dropin = False
_solver_cb = None   # keeps the ctypes callback alive

DP, IP, BP, I64 = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_uint8), C.c_int64
_SOLVER_FN = C.CFUNCTYPE(C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), DP, DP, DP)


def build() -> None:
    """make -C oracle/refbuild (a no-op without /root/reference)."""
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "refbuild"), "-s", "-j", str(os.cpu_count() or 4)])


def available() -> bool:
    if not os.path.exists(_SO) and os.path.isdir("/root/reference/srcs"):
        build()
    return os.path.exists(_SO)


def dropin_available() -> bool:
    if not os.path.exists(_SO_DROPIN) and os.path.isdir("/root/reference/srcs"):
        build()
    return os.path.exists(_SO_DROPIN)


def use_dropin_library() -> None:
    """Load libpfem_ref_dropin.so (reference host code + shim + libpfem_b200.so) instead of libpfem_ref.so.  Must be
    called before the first lib() of the process; GPU tests only."""
    global _use_dropin
    assert _lib is None or _use_dropin, "the plain reference library is already loaded in this process"
    _use_dropin = True


def lib():
    global _lib
    if _lib is None:
        if not (dropin_available() if _use_dropin else available()):
            raise RuntimeError("oracle/_ref library not built and /root/reference absent")
        L = C.CDLL(_SO_DROPIN if _use_dropin else _SO)
        if _use_dropin:
            L.pfem_ref_use_b200_equation.argtypes = [C.c_void_p]
            L.pfem_ref_wc_step_b200.argtypes = [C.c_void_p, C.c_double, C.c_int]
            L.pfem_ref_wc_next_dt_b200.restype = C.c_double
            L.pfem_ref_wc_next_dt_b200.argtypes = [C.c_void_p]
        L.pfem_ref_last_error.restype = C.c_char_p
        L.pfem_ref_create.restype = C.c_void_p
        L.pfem_ref_create.argtypes = [C.c_int, I64, I64, IP, DP, BP, BP, DP, I64, IP, C.c_char_p, C.c_char_p, DP]
        L.pfem_ref_destroy.argtypes = [C.c_void_p]
        for name in ("pfem_ref_set_states", "pfem_ref_get_states"):
            getattr(L, name).argtypes = [C.c_void_p, DP, C.c_int, C.c_int]
        L.pfem_ref_get_positions.argtypes = [C.c_void_p, DP]
        L.pfem_ref_set_dirichlet_values.argtypes = [C.c_void_p, DP]
        L.pfem_ref_set_time_step.argtypes = [C.c_void_p, C.c_double]
        L.pfem_ref_element_geometry.argtypes = [C.c_void_p, DP, DP, DP, DP]
        L.pfem_ref_tables.argtypes = [C.c_void_p, C.c_int, C.c_int, DP, DP, DP, DP]
        L.pfem_ref_element_matrices.argtypes = [C.c_void_p] + [DP] * 7
        L.pfem_ref_pspg_elements.argtypes = [C.c_void_p, DP, DP, DP, DP]
        L.pfem_ref_pspg_build.restype = I64
        L.pfem_ref_pspg_build.argtypes = [C.c_void_p, DP, C.c_int]
        L.pfem_ref_csc_copy.argtypes = [C.c_void_p, IP, C.POINTER(C.c_int32), DP, DP]
        L.pfem_ref_pspg_solve.argtypes = [C.c_void_p]
        L.pfem_ref_heat_solve.argtypes = [C.c_void_p]
        L.pfem_ref_fs_build.restype = I64
        L.pfem_ref_fs_build.argtypes = [C.c_void_p, C.c_int, DP, DP]
        L.pfem_ref_fs_copy.argtypes = [C.c_void_p, C.c_int, IP, C.POINTER(C.c_int32), DP, DP]
        L.pfem_ref_fs_solve.argtypes = [C.c_void_p, C.c_int, DP, DP]
        L.pfem_ref_in_heat_build.restype = I64
        L.pfem_ref_in_heat_build.argtypes = [C.c_void_p, DP, C.c_int]
        L.pfem_ref_in_heat_copy.argtypes = [C.c_void_p, IP, C.POINTER(C.c_int32), DP, DP]
        L.pfem_ref_wc_step.argtypes = [C.c_void_p, C.c_double]
        L.pfem_ref_wc_next_dt.restype = C.c_double
        L.pfem_ref_wc_next_dt.argtypes = [C.c_void_p]
        L.pfem_ref_set_direct_solver.argtypes = [C.c_void_p]
        L.pfem_ref_direct_solves.restype = C.c_long
        L.pfem_ref_set_threads.argtypes = [C.c_int]
        if hasattr(L, "pfem_ref_set_bc_ramp"):
            L.pfem_ref_set_bc_ramp.argtypes = [C.c_double]
        L.pfem_ref_set_thermal_bc.argtypes = [I64, BP, DP]
        L.pfem_ref_get_threads.restype = C.c_int
        L.pfem_ref_cg_log.restype = C.c_long
        L.pfem_ref_cg_log.argtypes = [DP, C.c_long]
        _lib = L
    return _lib


def _d(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(DP)


def set_threads(n: int = 0) -> int:
    """OpenMP threads (= Problem::m_nThreads, one parameter table each) of the cases created from now on; 0 = all cores."""
    lib().pfem_ref_set_threads(int(n))
    return lib().pfem_ref_get_threads()


def set_bc_ramp(ramp: float = 0.0) -> None:
    """Velocity Dirichlet values of the cases created from now on become dirVal (1 + ramp t): a time-dependent
    "<type>V"(pos, t) table like the reference's examples/2D/cylinder/cylinderComp.lua."""
    lib().pfem_ref_set_bc_ramp(float(ramp))


def use_scipy_direct_solver(enable: bool = True) -> None:
    """Route the stand-in Eigen::SparseLU to SciPy SuperLU (COLAMD) instead of the built-in dense LU."""
    global _solver_cb
    import scipy.sparse.linalg as spla

    if not enable:
        lib().pfem_ref_set_direct_solver(None)
        _solver_cb = None
        return

    def solve(n, col_ptr, row_idx, val, b, x):
        try:
            cp = np.ctypeslib.as_array(col_ptr, shape=(n + 1,))
            nnz = int(cp[n])
            A = sp.csc_matrix((np.ctypeslib.as_array(val, shape=(nnz,)).copy(),
                               np.ctypeslib.as_array(row_idx, shape=(nnz,)).copy(), cp.copy()), shape=(n, n))
            sol = spla.splu(A, permc_spec="COLAMD").solve(np.ctypeslib.as_array(b, shape=(n,)).copy())
            np.ctypeslib.as_array(x, shape=(n,))[:] = sol
            return 0
        except Exception:  # noqa: BLE001 - singular matrix etc. -> NumericalIssue on the C++ side
            return 1

    _solver_cb = _SOLVER_FN(solve)
    lib().pfem_ref_set_direct_solver(C.cast(_solver_cb, C.c_void_p))


class RefCase:
    """One reference Problem/Mesh/Solver/Equation set built from arrays (see refbuild/ref_driver.cpp)."""

    def __init__(self, mesh, kind: str, params, *, solver_id=None, facets=None, gamma=0.0, thermal=None, bingham=None):
        L = lib()
        self.mesh, self.kind = mesh, kind
        self.dim, self.N, self.E = mesh.dim, mesh.n_nodes, mesh.n_elems
        if kind == "pspg":      # params = oracle.pspg_param_array + (max_iter, min_res)
            rho, mu, dt, bx, by, bz = params[:6]
            max_iter, min_res = (params[6], params[7]) if len(params) >= 8 else (10, 1e-6)
            gamma_fs, residual = (params[8], params[9]) if len(params) >= 10 else (1.0, 0.0)
            p = np.array([rho, mu, dt, bx, by, bz, gamma, max_iter, min_res, gamma_fs, residual], dtype=np.float64)
            prob, sid = b"IncompNewtonNoT", (solver_id or "PSPG").encode()
            if bingham is not None:   # (tau0, mReg): Problem id "Bingham"
                prob = b"Bingham"
                p = np.concatenate([p, [bingham[0], bingham[1]]])
            elif thermal is not None:   # incompressible Boussinesq: dict(alpha, Tr, k, cv, t_mask, t_val); states gain T
                prob = b"Boussinesq"
                self.n_states = self.dim + 2
                p = np.concatenate([p, [thermal["alpha"], thermal["Tr"], thermal["k"], thermal["cv"]]])
                tm = np.ascontiguousarray(thermal["t_mask"], dtype=np.uint8)
                tv = np.ascontiguousarray(thermal["t_val"], dtype=np.float64)
                L.pfem_ref_set_thermal_bc(self.N, tm.ctypes.data_as(BP), _d(tv))
            self.n_states = self.dim + 1
        elif kind == "wc":      # params = oracle.wc_param_array + (initial_dt, max_dt, security_coeff)
            mu, K0, K0p, rho_star, bx, by, bz, meduri, eq_type = params[:9]
            dt0, max_dt, sec = (params[9], params[10], params[11]) if len(params) >= 12 else (1e-6, 1.0, 0.1)
            p = np.array([mu, K0, K0p, rho_star, bx, by, bz, meduri, gamma, dt0, max_dt, sec], dtype=np.float64)
            sid = (solver_id or {0: "CDS_dpdt", 1: "CDS_drhodt", 2: "CDS_rho"}[int(eq_type)]).encode()
            prob = b"WCompNewtonNoT"
            self.n_states = 2 * self.dim + 2
            if thermal is not None:   # BoussinesqWC: dict(k, cv, alpha, Tr, t_mask, t_val); the state vector gains T
                prob = b"BoussinesqWC"
                self.n_states = 2 * self.dim + 3
                p = np.concatenate([p, [thermal["k"], thermal["cv"], thermal["alpha"], thermal["Tr"]]])
                tm = np.ascontiguousarray(thermal["t_mask"], dtype=np.uint8)
                tv = np.ascontiguousarray(thermal["t_val"], dtype=np.float64)
                L.pfem_ref_set_thermal_bc(self.N, tm.ctypes.data_as(BP), _d(tv))
        else:
            raise ValueError(kind)
        fac = np.zeros((0, self.dim + 2), dtype=np.int64) if facets is None else np.ascontiguousarray(facets, dtype=np.int64)
        conn = np.ascontiguousarray(mesh.conn, dtype=np.int64)
        self._h = L.pfem_ref_create(self.dim, self.N, self.E, conn.ctypes.data_as(IP), _d(np.ascontiguousarray(mesh.x)),
                                    mesh.flags.ctypes.data_as(BP), mesh.dir_mask.ctypes.data_as(BP),
                                    _d(np.ascontiguousarray(mesh.dir_val)), fac.shape[0], fac.ctypes.data_as(IP),
                                    prob, sid, _d(p))
        if not self._h:
            raise RuntimeError("pfem_ref_create: " + L.pfem_ref_last_error().decode())

    def close(self):
        if self._h:
            lib().pfem_ref_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _chk(self, rc, what):
        if rc < 0:
            raise RuntimeError(f"{what}: {lib().pfem_ref_last_error().decode()}")
        return rc

    def set_states(self, q, s0=0):
        q = np.ascontiguousarray(q, dtype=np.float64)
        lib().pfem_ref_set_states(self._h, _d(q), s0, q.size // self.N)

    def get_states(self, s0=0, ns=None):
        ns = self.n_states - s0 if ns is None else ns
        q = np.zeros(ns * self.N)
        lib().pfem_ref_get_states(self._h, _d(q), s0, ns)
        return q

    def positions(self):
        x = np.zeros(self.dim * self.N)
        lib().pfem_ref_get_positions(self._h, _d(x))
        return x

    def set_time_step(self, dt):
        lib().pfem_ref_set_time_step(self._h, float(dt))

    def element_geometry(self):
        detJ, J, invJ, rin = np.zeros(self.E), np.zeros((self.E, 3, 3)), np.zeros((self.E, 3, 3)), np.zeros(self.E)
        lib().pfem_ref_element_geometry(self._h, _d(detJ), _d(J), _d(invJ), _d(rin))
        return detJ, J, invJ, rin

    def tables(self, dimension, n_gp):
        gp, w, sf, ref = np.zeros((n_gp, 3)), np.zeros(n_gp), np.zeros((n_gp, dimension + 1)), np.zeros(1)
        self._chk(lib().pfem_ref_tables(self._h, dimension, n_gp, _d(gp), _d(w), _d(sf), _d(ref)), "tables")
        return gp, w, sf, float(ref[0])

    def element_matrices(self):
        d, npe, E = self.dim, self.dim + 1, self.E
        out = dict(M=np.zeros((E, npe, npe)), K=np.zeros((E, d * npe, d * npe)), D=np.zeros((E, npe, d * npe)),
                   L=np.zeros((E, npe, npe)), C=np.zeros((E, npe, d * npe)), F=np.zeros((E, d * npe)), H=np.zeros((E, npe)))
        self._chk(lib().pfem_ref_element_matrices(self._h, *[_d(out[k]) for k in "MKDLCFH"]), "element_matrices")
        return out

    def pspg_elements(self, q_prev):
        nt = (self.dim + 1) ** 2
        Ae, be, tau = np.zeros((self.E, nt, nt)), np.zeros((self.E, nt)), np.zeros(self.E)
        self._chk(lib().pfem_ref_pspg_elements(self._h, _d(np.ascontiguousarray(q_prev)), _d(Ae), _d(be), _d(tau)), "pspg_elements")
        return Ae, be, tau

    def pspg_build(self, q_prev, apply_bc=True):
        n_dof = (self.dim + 1) * self.N
        nnz = self._chk(lib().pfem_ref_pspg_build(self._h, _d(np.ascontiguousarray(q_prev)), 1 if apply_bc else 0), "pspg_build")
        col_ptr, row_idx = np.zeros(n_dof + 1, dtype=np.int64), np.zeros(nnz, dtype=np.int32)
        val, b = np.zeros(nnz), np.zeros(n_dof)
        lib().pfem_ref_csc_copy(self._h, col_ptr.ctypes.data_as(IP), row_idx.ctypes.data_as(C.POINTER(C.c_int32)), _d(val), _d(b))
        return sp.csc_matrix((val, row_idx, col_ptr), shape=(n_dof, n_dof)), b

    def in_heat_build(self, theta_prev, apply_bc=True):
        """HeatEqIncompNewton::m_buildAb (+ m_applyBC) of a "Boussinesq" case: (A as CSC over the nodes, b)."""
        nnz = self._chk(lib().pfem_ref_in_heat_build(self._h, _d(np.ascontiguousarray(theta_prev, dtype=np.float64)), 1 if apply_bc else 0),
                        "in_heat_build")
        col_ptr, row_idx = np.zeros(self.N + 1, dtype=np.int64), np.zeros(nnz, dtype=np.int32)
        val, b = np.zeros(nnz), np.zeros(self.N)
        lib().pfem_ref_in_heat_copy(self._h, col_ptr.ctypes.data_as(IP), row_idx.ctypes.data_as(C.POINTER(C.c_int32)), _d(val), _d(b))
        return sp.csc_matrix((val, row_idx, col_ptr), shape=(self.N, self.N)), b

    def pspg_solve(self):
        """MomContEqIncompNewton::solve() (Picard loop).  Returns (ok, number of direct solves = Picard iterations)."""
        n0 = lib().pfem_ref_direct_solves()
        ok = self._chk(lib().pfem_ref_pspg_solve(self._h), "pspg_solve")
        return bool(ok), lib().pfem_ref_direct_solves() - n0

    def fs_build(self, which, a, b=None):
        """FracStep sub-system `which` (0 velocity prediction: a = v_prev, b = p_prev; 1 pressure: a = vTilde, b = p_prev;
        2 velocity correction: a = deltaP) built by the reference's own m_buildMat* + m_applyBC* (MomContEquationFracStep.inl).
        Returns (A csc, rhs)."""
        a = np.ascontiguousarray(a, dtype=np.float64)
        bb = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
        nnz = lib().pfem_ref_fs_build(self._h, int(which), _d(a), None if bb is None else _d(bb))
        if nnz < 0:
            raise RuntimeError(f"fs_build: {lib().pfem_ref_last_error().decode()}")
        n = self.N if which == 1 else self.dim * self.N
        col_ptr = np.zeros(n + 1, dtype=np.int64)
        row_idx = np.zeros(nnz, dtype=np.int32)
        val = np.zeros(nnz)
        rhs = np.zeros(n)
        lib().pfem_ref_fs_copy(self._h, int(which), col_ptr.ctypes.data_as(IP), row_idx.ctypes.data_as(C.POINTER(C.c_int32)), _d(val), _d(rhs))
        return sp.csc_matrix((val, row_idx, col_ptr), shape=(n, n)), rhs

    def fs_solve(self, which):
        """m_solverIt.compute(A); x = m_solverIt.solve(b) on the sub-system last built (stand-in ConjugateGradient: Jacobi,
        tolerance eps, 2n iterations).  Returns (x, iterations, error, info)."""
        n = self.N if which == 1 else self.dim * self.N
        x = np.zeros(n)
        out = np.zeros(3)
        self._chk(lib().pfem_ref_fs_solve(self._h, int(which), _d(x), _d(out)), "fs_solve")
        return x, int(out[0]), float(out[1]), int(out[2])

    def heat_solve(self):
        """m_pEquations[1]->solve(): the heat equation of the incompressible Boussinesq problem (IN/Solver.cpp:249-258)."""
        return bool(self._chk(lib().pfem_ref_heat_solve(self._h), "heat_solve"))

    @staticmethod
    def cg_log(max_rows=256):
        """(n, iterations, relative residual, info) of the stand-in ConjugateGradient solves since the last call."""
        out = np.zeros((max_rows, 4))
        k = lib().pfem_ref_cg_log(_d(out), max_rows)
        return out[:k]

    # ---- drop-in build only (libpfem_ref_dropin.so) ----
    def use_b200_equation(self):
        """Swap the reference's MomContEqIncompNewton<dim> for the shim's MomContEqIncompNewtonB200<dim> (= the REGISTER_EQ edit);
        problem id "Boussinesq": HeatEqIncompNewton<dim> becomes HeatEqIncompNewtonB200<dim> as well."""
        self._chk(lib().pfem_ref_use_b200_equation(self._h), "use_b200_equation")

    def wc_step_b200(self, dt, download=True):
        return bool(self._chk(lib().pfem_ref_wc_step_b200(self._h, float(dt), 1 if download else 0), "wc_step_b200"))

    def wc_next_dt_b200(self):
        return lib().pfem_ref_wc_next_dt_b200(self._h)

    def wc_step(self, dt):
        return bool(self._chk(lib().pfem_ref_wc_step(self._h, float(dt)), "wc_step"))

    def wc_next_dt(self):
        return lib().pfem_ref_wc_next_dt(self._h)
