"""Second, independent restatement of the reference element arithmetic in numpy -- TEST INFRASTRUCTURE ONLY.

Where oracle/pfem_oracle.cpp unrolls the reference loops by hand in C++, this module writes the same
formulas as the dense matrix products the reference hands to Eigen (B^T ddev B, sumWNT m^T B, ...) and
lets numpy evaluate them.  It is slow (python loop over elements) and is used on small meshes only, to
pin the C++ oracle.  File:line citations are to /root/reference/srcs/...
"""
from __future__ import annotations

import numpy as np

# Mesh.cpp:375-378, 395-399 (points), :443-444, 458-459 (weights), :479-483 (ref size)
GP = {2: np.array([[1 / 6, 1 / 6], [1 / 6, 2 / 3], [2 / 3, 1 / 6]]),
      3: np.array([[0.585410196624968, 0.138196601125011, 0.138196601125011],
                   [0.138196601125011, 0.585410196624968, 0.138196601125011],
                   [0.138196601125011, 0.138196601125011, 0.585410196624968],
                   [0.138196601125011, 0.138196601125011, 0.138196601125011]])}
GW = {2: np.full(3, 1 / 3), 3: np.full(4, 0.25)}
REF = {2: 0.5, 3: 0.16666666666666666666666666666667}


class MatrixBuilder:
    """matricesBuilder/MatricesBuilder.inl, element part."""

    def __init__(self, dim):
        self.dim, self.npe = dim, dim + 1
        gp = GP[dim]
        self.w = GW[dim]
        self.N = [np.concatenate([[1 - g.sum()], g])[None, :] for g in gp]          # Mesh.cpp:509-519
        self.NtN = [n.T @ n for n in self.N]                                         # MB.inl:72-76
        self.Ntilde = [np.kron(np.eye(dim), n) for n in self.N]                      # MB.inl:28-42
        self.ref = REF[dim]
        self.ddev = None
        self.m = None

    @staticmethod
    def geometry(xe):
        """Element.cpp:15-135.  xe: (npe, dim) node coordinates."""
        J = (xe[1:] - xe[0]).T                     # columns = edge vectors
        detJ = np.linalg.det(J)
        return J, detJ, np.linalg.inv(J)

    def gradN(self, invJ):                          # MB.inl:93-127
        g = np.zeros((self.dim, self.npe))
        g[:, 1:] = invJ.T
        g[:, 0] = -invJ.sum(axis=0)
        return g

    def B(self, g):                                 # MB.inl:130-164
        dim, npe = self.dim, self.npe
        if dim == 2:
            B = np.zeros((3, 6))
            B[0, 0:3] = g[0]; B[2, 3:6] = g[0]
            B[1, 3:6] = g[1]; B[2, 0:3] = g[1]
        else:
            B = np.zeros((6, 12))
            B[0, 0:4] = g[0]; B[3, 4:8] = g[0]; B[4, 8:12] = g[0]
            B[1, 4:8] = g[1]; B[3, 0:4] = g[1]; B[5, 8:12] = g[1]
            B[2, 8:12] = g[2]; B[4, 0:4] = g[2]; B[5, 4:8] = g[2]
        return B

    def getM(self, detJ, f):                        # MB.inl:217-229
        return sum(f(N) * NtN * w for N, NtN, w in zip(self.N, self.NtN, self.w)) * detJ * self.ref

    def getK(self, detJ, B, f):                     # MB.inl:277-290
        fact = sum(f(N) * w for N, w in zip(self.N, self.w))
        return detJ * self.ref * fact * B.T @ self.ddev @ B

    def getD(self, detJ, B, f):                     # MB.inl:293-306
        sumWNT = sum(f(N) * N.T * w for N, w in zip(self.N, self.w))
        return detJ * self.ref * sumWNT @ self.m[None, :] @ B

    def getL(self, detJ, g, f):                     # MB.inl:309-322
        fact = sum(f(N) * w for N, w in zip(self.N, self.w))
        return detJ * self.ref * fact * g.T @ g

    def getC(self, detJ, g, f):                     # MB.inl:325-338
        sumNW = sum(f(N) * Nt * w for N, Nt, w in zip(self.N, self.Ntilde, self.w))
        return detJ * self.ref * g.T @ sumNW

    def getF(self, detJ, vec, f):                   # MB.inl:341-355
        return sum(f(N) * Nt.T @ vec * w for N, Nt, w in zip(self.N, self.Ntilde, self.w)) * detJ * self.ref

    def getH(self, detJ, vec, g, f):                # MB.inl:373-386
        return sum(f(N) * g.T @ vec * w for N, w in zip(self.N, self.w)) * detJ * self.ref


def _elem_vec(q, en, nn, dim):
    """getElementVecState (StatesFromToQ.hpp:136-170): [u0..u_npe-1, v0.., (w0..)]."""
    return np.concatenate([q[en + d * nn] for d in range(dim)])


def pspg_elements(mesh, vcur, q_prev, rho, mu, dt, body):
    """Ae, be, tau for every element: PSPG.inl:26-53, :238-259; factors MomContEquation.inl:73-222."""
    dim, npe, nn = mesh.dim, mesh.dim + 1, mesh.n_nodes
    mb = MatrixBuilder(dim)
    mb.ddev = np.diag([2.0] * dim + [1.0] * (3 if dim == 3 else 1))
    mb.m = np.array([1.0] * dim + [0.0] * (3 if dim == 3 else 1))
    body = np.asarray(body[:dim], dtype=float)
    c = mesh.coords()
    nt = (dim + 1) * npe
    Ae = np.zeros((mesh.n_elems, nt, nt)); be = np.zeros((mesh.n_elems, nt)); taus = np.zeros(mesh.n_elems)
    for e, en in enumerate(mesh.conn):
        _, detJ, invJ = mb.geometry(c[en])
        h = np.sqrt(mb.ref * detJ / np.pi)
        U = np.mean([np.sqrt(sum(vcur[n + d * nn] ** 2 for d in range(dim))) for n in en])
        tau = 1 / np.sqrt((2 / dt) ** 2 + (2 * U / h) ** 2 + 9 * (4 * mu / (h * h * rho)) ** 2)
        g = mb.gradN(invJ); B = mb.B(g)
        Me_dt = np.kron(np.eye(dim), (1 / dt) * mb.getM(detJ, lambda N: rho))       # diagBlock, MB.hpp:120-129
        Ke = mb.getK(detJ, B, lambda N: mu)
        De = mb.getD(detJ, B, lambda N: 1.0)
        Ce_dt = (tau / dt) * mb.getC(detJ, g, lambda N: 1.0)
        Le = tau * mb.getL(detJ, g, lambda N: 1 / rho)
        Fe = mb.getF(detJ, body, lambda N: rho)
        He = tau * mb.getH(detJ, body, g, lambda N: 1.0)
        Ae[e] = np.block([[Me_dt + Ke, -De.T], [Ce_dt + De, Le]])
        vPrev = _elem_vec(q_prev, en, nn, dim)
        be[e] = np.concatenate([Fe + Me_dt @ vPrev, He + Ce_dt @ vPrev])
        taus[e] = tau
    return Ae, be, taus


def pspg_assemble_dense(mesh, Ae, be, q_prev, dt, body, apply_bc=True):
    """Dense A, b with the row masks, identity rows and BC pass of PSPG.inl:58-128, 190-232."""
    dim, npe, nn = mesh.dim, mesh.dim + 1, mesh.n_nodes
    nd = (dim + 1) * nn
    A = np.zeros((nd, nd)); b = np.zeros(nd)
    bound = (mesh.flags & 1) != 0; free = (mesh.flags & 2) != 0
    for e, en in enumerate(mesh.conn):
        for i in range(npe):
            for j in range(npe):
                for d1 in range(dim + 1):
                    masked = (bound[en[i]] or free[en[i]]) if d1 < dim else free[en[i]]
                    if masked:
                        continue
                    for d2 in range(dim + 1):
                        A[en[i] + d1 * nn, en[j] + d2 * nn] += Ae[e][i + d1 * npe, j + d2 * npe]
            for d in range(dim + 1):
                b[en[i] + d * nn] += be[e][i + d * npe]
    for n in range(nn):
        if free[n]:
            A[n + dim * nn, n + dim * nn] += 1
        if bound[n] or free[n]:
            for d in range(dim):
                A[n + d * nn, n + d * nn] += 1
    if apply_bc:
        for n in range(nn):
            if free[n]:
                b[n + dim * nn] = 0
                if not bound[n]:
                    for d in range(dim):
                        b[n + d * nn] = q_prev[n + d * nn] + dt * body[d]
            if bound[n] and mesh.dir_mask[n]:
                for d in range(dim):
                    col = n + d * nn
                    r = mesh.dir_val[col]
                    b[col] = r
                    colv = A[:, col].copy(); colv[col] = 0
                    b -= colv * r
                    diag = A[col, col]
                    A[:, col] = 0; A[col, col] = diag
    return A, b


def wc_step(mesh, x, st, mu, K0, K0p, rho_star, body, dt, meduri=True, eq_type="CDS_dpdt"):
    """One explicit step: WC/Solver.cpp:236-263, ContEquation.inl:123-301, 334-413, MomEquation.inl:201-374."""
    dim, npe, nn = mesh.dim, mesh.dim + 1, mesh.n_nodes
    body = np.asarray(body[:dim], dtype=float)
    F0pre = None
    if eq_type == "CDS_rho":    # preCompute() on the configuration before the move, ContEquation.inl:196-232
        mb0 = MatrixBuilder(dim)
        c0 = x.reshape(dim, nn).T
        F0pre = np.zeros(nn)
        for en in mesh.conn:
            _, detJ, _ = mb0.geometry(c0[en])
            np.add.at(F0pre, en, mb0.getM(detJ, lambda N: 1.0) @ st["rho"][en])
    v = st["v"] + 0.5 * dt * st["acc"]
    x = x.copy()
    fixed = (mesh.flags & 4) != 0
    for d in range(dim):
        sl = slice(d * nn, (d + 1) * nn)
        x[sl] = np.where(fixed, x[sl], x[sl] + v[sl] * dt)
    c = x.reshape(dim, nn).T
    mb = MatrixBuilder(dim)
    mb.m = np.array([1.0] * dim + [0.0] * (3 if dim == 3 else 1))
    dd = np.eye(3 if dim == 2 else 6)
    dd[:dim, :dim] = -2 / 3; dd[np.arange(dim), np.arange(dim)] = 4 / 3
    mb.ddev = dd
    bound = (mesh.flags & 1) != 0; free = (mesh.flags & 2) != 0
    # continuity
    p = st["p"]
    invM = np.zeros(nn); F0 = np.zeros(nn)
    if eq_type == "CDS_dpdt":
        for en in mesh.conn:
            _, detJ, invJ = mb.geometry(c[en])
            Me = mb.getM(detJ, lambda N: 1.0)
            MeL = Me.sum(axis=1)
            P = p[en]; V = _elem_vec(v, en, nn, dim)
            g = mb.gradN(invJ); B = mb.B(g)
            D = mb.getD(detJ, B, lambda N: K0 + K0p * (N @ P).item())
            F0e = -dt * D @ V + (Me @ P if meduri else MeL * P)
            np.add.at(invM, en, MeL); np.add.at(F0, en, F0e)
        with np.errstate(divide="ignore"):
            invM = 1 / invM
        F0[free] = 0; invM[free] = 1
        p = invM * F0
        rho = np.power((K0p / K0) * p + 1, 1 / K0p) * rho_star
    else:
        rho0 = st["rho"]
        for en in mesh.conn:
            _, detJ, invJ = mb.geometry(c[en])
            Me = mb.getM(detJ, lambda N: 1.0)
            MeL = Me.sum(axis=1)
            np.add.at(invM, en, MeL)
            if eq_type == "CDS_drhodt":
                R = rho0[en]; V = _elem_vec(v, en, nn, dim)
                g = mb.gradN(invJ); B = mb.B(g)
                D = mb.getD(detJ, B, lambda N: (N @ R).item())
                np.add.at(F0, en, -dt * D @ V + (Me @ R if meduri else MeL * R))
        if eq_type == "CDS_rho":
            F0 = F0pre
        with np.errstate(divide="ignore"):
            invM = 1 / invM
        fs = free | ((mesh.flags & 8) != 0)
        F0[fs] = rho_star; invM[fs] = 1
        rho = invM * F0
        p = (K0 / K0p) * (np.power(rho / rho_star, K0p) - 1)
    # momentum
    Md = np.zeros(dim * nn); F = np.zeros(dim * nn)
    for en in mesh.conn:
        _, detJ, invJ = mb.geometry(c[en])
        V = _elem_vec(v, en, nn, dim); P = p[en]; R = rho[en]
        g = mb.gradN(invJ); B = mb.B(g)
        Mt = mb.getM(detJ, lambda N: (N @ R).item())
        lumped = np.kron(np.eye(dim), Mt).sum(axis=1)
        K = mb.getK(detJ, B, lambda N: mu)
        D = mb.getD(detJ, B, lambda N: 1.0)
        Fe = mb.getF(detJ, body, lambda N: (N @ R).item())
        FT = -K @ V + D.T @ P + Fe
        idx = np.concatenate([en + d * nn for d in range(dim)])
        np.add.at(Md, idx, lumped); np.add.at(F, idx, FT)
    with np.errstate(divide="ignore"):
        invMd = 1 / Md
    for n in range(nn):
        if free[n] and not bound[n]:
            for d in range(dim):
                F[n + d * nn] = body[d]; invMd[n + d * nn] = 1
        elif bound[n] and mesh.dir_mask[n]:
            for d in range(dim):
                F[n + d * nn] = mesh.dir_val[n + d * nn]; invMd[n + d * nn] = 1
    acc = invMd * F
    vnew = v + 0.5 * dt * acc
    return x, dict(v=vnew, p=p, rho=rho, acc=acc)
