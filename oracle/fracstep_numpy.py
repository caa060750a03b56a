"""Restatement of the fractional-step solver's three linear systems in numpy -- TEST INFRASTRUCTURE ONLY.

One Picard body of MomContEqIncompNewton with solver id "FracStep" (srcs/simulation/physics/IncompNewton/
MomContEquationFracStep.inl:464-548) builds: the velocity-prediction system (m_buildMatFracStep :8-217 + m_applyBCVAppStep
:219-298), the pressure system (m_buildMatPcorrStep :300-347 + m_applyBCPCorrStep :349-376) and the velocity-correction
system (m_buildMatVStep :378-425 + m_applyBCVStep :427-452).  Element matrices come from oracle/literal_numpy.MatrixBuilder
(factors: MomContEquation.inl:96-158: M rho, K mu, D 1, L 1, F rho).  Dense, python loop over elements: small meshes only.
Pinned against the reference's own code (oracle/_ref) by tests/test_oracle_vs_reference.py::test_live_fracstep_systems and
against tests/golden/fs_*.npz.
"""
from __future__ import annotations

import numpy as np

from .literal_numpy import MatrixBuilder, _elem_vec

F_BOUND, F_FREE, F_FREE_SURFACE = 1, 2, 8   # include/pfem_b200.h PFEM_NODE_*


def _builder(dim):
    mb = MatrixBuilder(dim)
    mb.ddev = np.diag([2.0] * dim + [1.0] * (3 if dim == 3 else 1))
    mb.m = np.array([1.0] * dim + [0.0] * (3 if dim == 3 else 1))
    return mb


def _elements(mesh, rho, mu, body):
    dim = mesh.dim
    mb = _builder(dim)
    c = mesh.coords()
    out = []
    for en in mesh.conn:
        _, detJ, invJ = mb.geometry(c[en])
        g = mb.gradN(invJ)
        B = mb.B(g)
        Ms = mb.getM(detJ, lambda N: rho)
        out.append(dict(en=en, Ms=Ms, K=mb.getK(detJ, B, lambda N: mu), DT=mb.getD(detJ, B, lambda N: 1.0).T,
                        L=mb.getL(detJ, g, lambda N: 1.0), F=mb.getF(detJ, np.asarray(body[:dim], dtype=float), lambda N: rho)))
    return out


def velocity_prediction(mesh, v_prev, p_prev, rho, mu, dt, body, gamma_fs):
    """(M/dt + K) vTilde = F + M/dt v_prev + gammaFS D^T p_prev; rows of bound / free nodes = identity (:73-121, 167-183); then
    m_applyBCVAppStep: free, not bound -> v_prev + dt g; bound with a velocity function -> its value, column eliminated (:256-294).
    Returns dense (A, b), dof order n + d N."""
    dim, npe, nn = mesh.dim, mesh.dim + 1, mesh.n_nodes
    A = np.zeros((dim * nn, dim * nn))
    b = np.zeros(dim * nn)
    bound = (mesh.flags & F_BOUND) != 0
    free = (mesh.flags & F_FREE) != 0
    masked = bound | free
    for E in _elements(mesh, rho, mu, body):
        en = E["en"]
        Me_dt = np.kron(np.eye(dim), E["Ms"] / dt)
        Ae = Me_dt + E["K"]
        be = E["F"] + Me_dt @ _elem_vec(v_prev, en, nn, dim) + gamma_fs * E["DT"] @ p_prev[en]
        for i in range(npe):
            if masked[en[i]]:
                continue
            for d in range(dim):
                r = en[i] + d * nn
                b[r] += be[i + d * npe]
                for j in range(npe):
                    for d2 in range(dim):
                        A[r, en[j] + d2 * nn] += Ae[i + d * npe, j + d2 * npe]
    for n in np.flatnonzero(masked):
        for d in range(dim):
            A[n + d * nn, n + d * nn] = 1.0
    dv = mesh.dir_val.reshape(dim, nn)
    for n in range(nn):
        if free[n] and not bound[n]:
            for d in range(dim):
                b[n + d * nn] = v_prev[n + d * nn] + dt * body[d]
        if bound[n] and mesh.dir_mask[n]:
            for d in range(dim):
                col = n + d * nn
                b[col] = dv[d, n]
                rows = np.flatnonzero(A[:, col])
                rows = rows[rows != col]
                b[rows] -= A[rows, col] * dv[d, n]
                A[rows, col] = 0.0
    return A, b


def pressure(mesh, v_tilde, p_prev, rho, mu, dt, body, gamma_fs):
    """L p = -(rho/dt) D vTilde + gammaFS L p_prev on the rows of nodes that are neither free nor on the free surface (:318-335),
    0 elsewhere; L rows of free nodes = identity (:124-130, 162-166) and their columns cleared (:359-371)."""
    dim, npe, nn = mesh.dim, mesh.dim + 1, mesh.n_nodes
    A = np.zeros((nn, nn))
    b = np.zeros(nn)
    free = (mesh.flags & F_FREE) != 0
    fs = (mesh.flags & F_FREE_SURFACE) != 0
    for E in _elements(mesh, rho, mu, body):
        en = E["en"]
        rhs = -(rho / dt) * E["DT"].T @ _elem_vec(v_tilde, en, nn, dim) + gamma_fs * E["L"] @ p_prev[en]
        for i in range(npe):
            if not free[en[i]]:
                A[en[i], en] += E["L"][i]
            if not free[en[i]] and not fs[en[i]]:
                b[en[i]] += rhs[i]
    for n in np.flatnonzero(free):
        A[n, n] = 1.0
        b[n] = 0.0
        col = A[:, n].copy()
        A[:, n] = 0.0
        A[n, n] = 1.0 if col[n] != 0 else 0.0
    return A, b


def velocity_correction(mesh, delta_p, rho, mu, dt, body):
    """M deltaV = dt D^T deltaP on the rows of nodes that are neither free nor bound (:404-416); M rows of bound / free nodes =
    identity (:76-83, 173-176), their columns cleared (:436-447), right-hand side 0 there."""
    dim, npe, nn = mesh.dim, mesh.dim + 1, mesh.n_nodes
    A = np.zeros((dim * nn, dim * nn))
    b = np.zeros(dim * nn)
    masked = (mesh.flags & (F_BOUND | F_FREE)) != 0
    for E in _elements(mesh, rho, mu, body):
        en = E["en"]
        Me = np.kron(np.eye(dim), E["Ms"])
        rhs = dt * E["DT"] @ delta_p[en]
        for i in range(npe):
            if masked[en[i]]:
                continue
            for d in range(dim):
                r = en[i] + d * nn
                b[r] += rhs[i + d * npe]
                for j in range(npe):
                    A[r, en[j] + d * nn] += Me[i + d * npe, j + d * npe]
    for n in np.flatnonzero(masked):
        for d in range(dim):
            k = n + d * nn
            A[:, k] = 0.0
            A[k, k] = 1.0
            b[k] = 0.0
    return A, b


def conjugate_gradient(A, b, tol=np.finfo(float).eps, max_iter=None):
    """Eigen's conjugate_gradient() with the diagonal preconditioner and a zero initial guess (what m_solverIt.solve(b) runs,
    FracStep.inl:472-478): returns (x, iterations as Eigen counts them, relative residual, converged)."""
    n = b.size
    max_iter = 2 * n if max_iter is None else max_iter
    d = A.diagonal() if hasattr(A, "diagonal") else np.diag(A)
    inv = np.where(d != 0, 1.0 / np.where(d != 0, d, 1.0), 1.0)
    x = np.zeros(n)
    r = b.copy()
    bb = b @ b
    if bb == 0:
        return x, 0, 0.0, True
    thr = max(tol * tol * bb, np.finfo(float).tiny)
    rr = r @ r
    if rr < thr:
        return x, 0, np.sqrt(rr / bb), True
    p = inv * r
    abs_new = r @ p
    i = 0
    while i < max_iter:
        tmp = A @ p
        alpha = abs_new / (p @ tmp)
        x += alpha * p
        r -= alpha * tmp
        rr = r @ r
        if rr < thr:
            break
        z = inv * r
        abs_old = abs_new
        abs_new = r @ z
        p = z + (abs_new / abs_old) * p
        i += 1
    return x, i, np.sqrt(rr / bb), rr < thr
