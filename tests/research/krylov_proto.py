"""CPU prototype: preconditioner applications needed by BiCGSTAB vs right-preconditioned GMRES / GCR with the same
aggregation-multigrid cycle (research tooling, see precond_proto.py).  Usage: python tests/research/krylov_proto.py [n] [nu] [gs]"""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests/research")
import precond_proto as pp            # noqa: E402
from pfem_b200 import meshgen as mg   # noqa: E402


def gmres_right(A, b, M, tol, maxit=400, restart=None, cgs=False):
    """full (or restarted) right-preconditioned GMRES; returns x, number of preconditioner applications."""
    n = b.size
    x = np.zeros(n)
    bn = np.linalg.norm(b)
    napp = 0
    while True:
        r = b - A @ x
        beta = np.linalg.norm(r)
        if beta <= tol * bn or napp >= maxit:
            return x, napp, beta / bn
        m = restart or maxit
        V = [r / beta]
        Z = []
        H = np.zeros((m + 1, m))
        g = np.zeros(m + 1)
        g[0] = beta
        cs, sn = [], []
        k = 0
        for k in range(m):
            z = M(V[k])
            napp += 1
            w = A @ z
            if cgs:   # classical Gram-Schmidt: all dots against the same w, one update (what one fused kernel does)
                hh = [V[i] @ w for i in range(k + 1)]
                for i in range(k + 1):
                    H[i, k] = hh[i]
                    w = w - hh[i] * V[i]
            else:
                for i in range(k + 1):
                    H[i, k] = V[i] @ w
                    w = w - H[i, k] * V[i]
            H[k + 1, k] = np.linalg.norm(w)
            V.append(w / H[k + 1, k])
            Z.append(z)
            for i in range(k):
                t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]
                H[i, k] = t
            d = np.hypot(H[k, k], H[k + 1, k])
            cs.append(H[k, k] / d)
            sn.append(H[k + 1, k] / d)
            H[k, k] = d
            H[k + 1, k] = 0
            g[k + 1] = -sn[k] * g[k]
            g[k] = cs[k] * g[k]
            if abs(g[k + 1]) <= tol * bn or napp >= maxit:
                break
        y = np.linalg.solve(np.triu(H[: k + 1, : k + 1]), g[: k + 1])
        for i in range(k + 1):
            x = x + y[i] * Z[i]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    nu = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    tol = 1e-12
    mesh, A, b = pp.build(n)
    bs = mesh.dim + 1
    A_phys = A
    A, b, s = pp.equilibrate(A, b)
    coords = mesh.coords()
    vol = np.abs(mg.det_j(mesh)).mean() / 6
    h = (6 * vol) ** (1.0 / 3)
    gs = "gs" in sys.argv   # node-block Gauss-Seidel smoothing instead of damped Jacobi
    for over in ((1.0, 1.5) if gs else (1.5,)):
        mgp = pp.MG(A_phys, coords, h, bs, nu=nu, w=0.7, factor=2.0, cycle="V", over=over, gs=gs)
        M = lambda r, mgp=mgp: mgp(r / s) / s     # noqa: E731
        x, it, res = pp.bicgstab(A, b, M, tol=tol)
        print(f"bicgstab  over={over}: {it} iterations = {2 * it} cycles + {2 * it} spmv, res {res:.1e}")
        for restart, cgs in ((None, False), (30, True), (20, True), (10, True)):
            x, napp, res = gmres_right(A, b, M, tol, restart=restart, cgs=cgs)
            true = np.linalg.norm(b - A @ x) / np.linalg.norm(b)
            print(f"gmres({restart}, cgs={cgs}) over={over}: {napp} cycles + {napp} spmv, res {res:.1e} true {true:.1e}")


if __name__ == "__main__":
    main()
