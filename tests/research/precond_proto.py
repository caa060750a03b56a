"""CPU prototype (scipy) used to choose the device preconditioner of the PSPG BiCGSTAB solve.

Research tooling only (lives under tests/ because it builds its matrix with the oracle, which only tests may import): builds the oracle matrix of the bench case at a small size and counts BiCGSTAB iterations for
  jac   : node-block Jacobi (what round-1 csrc/krylov.cu does)
  schur : block lower-triangular  [A_vv 0; A_pv S^]  with S^ = aggregation-AMG V-cycle on a nodal Laplacian
  mono  : monolithic aggregation multigrid on the node-block matrix, block-Jacobi smoothing
Usage: python tests/research/precond_proto.py [n] [variant ...]
"""
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

sys.path.insert(0, ".")
from oracle import oracle as orc          # noqa: E402
from pfem_b200 import meshgen as mg       # noqa: E402


def build(n, dim=3, jitter=0.1, cloud=False):
    mesh = mg.delaunay_cloud(dim, (n + 1) ** dim) if cloud else mg.kuhn_box(dim, n, jitter=jitter)
    q, q_prev = mg.pspg_state(mesh)
    P = mg.PSPG_PARAMS
    par = orc.pspg_param_array(P["rho"], P["mu"], P["dt"], mg.gravity(dim))
    A, b = orc.pspg_build(mesh, q[: dim * mesh.n_nodes], q_prev, par, True)
    nn, bs = mesh.n_nodes, dim + 1
    # node-block ordering: new = node*bs + d   <-  old = node + d*nn
    old = (np.arange(nn)[:, None] + np.arange(bs)[None, :] * nn).reshape(-1)
    Pm = sp.csr_matrix((np.ones(nn * bs), (np.arange(nn * bs), old)), shape=(nn * bs, nn * bs))
    A = (Pm @ A.tocsr() @ Pm.T).tocsr()
    b = Pm @ b
    return mesh, A, b


def equilibrate(A, b):
    d = np.abs(A.diagonal())
    s = 1.0 / np.sqrt(np.where(d > 0, d, 1.0))
    S = sp.diags(s)
    return (S @ A @ S).tocsr(), S @ b, s


def block_diag_inv(A, bs):
    n = A.shape[0] // bs
    B = A.tobsr(blocksize=(bs, bs))
    B.sort_indices()
    rows = np.repeat(np.arange(n), np.diff(B.indptr))
    sel = np.flatnonzero(B.indices == rows)
    D = np.zeros((n, bs, bs))
    D[rows[sel]] = B.data[sel]
    return np.linalg.inv(D)


def apply_bdinv(Dinv, r):
    bs = Dinv.shape[1]
    return np.einsum("nij,nj->ni", Dinv, r.reshape(-1, bs)).reshape(-1)


def bicgstab(A, b, M, tol=1e-10, maxit=20000):
    x = np.zeros_like(b)
    r = b.copy()
    r0 = r.copy()
    rho = alpha = omega = 1.0
    v = np.zeros_like(b)
    p = np.zeros_like(b)
    bn = np.linalg.norm(b)
    for it in range(1, maxit + 1):
        rho_new = r0 @ r
        beta = (rho_new / rho) * (alpha / omega)
        p = r + beta * (p - omega * v)
        ph = M(p)
        v = A @ ph
        alpha = rho_new / (r0 @ v)
        s = r - alpha * v
        sh = M(s)
        t = A @ sh
        omega = (t @ s) / (t @ t)
        x += alpha * ph + omega * sh
        r = s - omega * t
        rho = rho_new
        if np.linalg.norm(r) <= tol * bn:
            break
    return x, it, np.linalg.norm(b - A @ x) / bn


# ------------------------------------------------------------------ aggregation
def aggregate_grid(coords, h):
    """aggregates = nodes in the same cell of a uniform grid of size h."""
    lo = coords.min(axis=0)
    ijk = np.floor((coords - lo) / h + 1e-9).astype(np.int64)
    dims = ijk.max(axis=0) + 1
    key = ijk[:, 0]
    for d in range(1, coords.shape[1]):
        key = key * dims[d] + ijk[:, d]
    _, agg = np.unique(key, return_inverse=True)
    return agg


def coarse_coords(coords, agg):
    nc = agg.max() + 1
    cnt = np.bincount(agg, minlength=nc)
    return np.stack([np.bincount(agg, weights=coords[:, d], minlength=nc) / cnt for d in range(coords.shape[1])], axis=1)


def prolong(agg, bs):
    n = agg.shape[0]
    nc = agg.max() + 1
    rows = np.arange(n * bs)
    cols = (agg[:, None] * bs + np.arange(bs)[None, :]).reshape(-1)
    return sp.csr_matrix((np.ones(n * bs), (rows, cols)), shape=(n * bs, nc * bs))


class MG:
    """aggregation multigrid, block-Jacobi smoothing (nu pre + nu post, damping w), V- or W(K)-cycle."""

    def __init__(self, A, coords, h, bs, nu=1, w=0.7, min_size=None, factor=2.0, cycle="V", over=1.0, smooth_p=False, gs=False):
        import os
        min_size = int(os.environ.get("MG_MIN", "400")) if min_size is None else min_size
        self.levels = []
        self.bs, self.nu, self.w, self.cycle, self.over = bs, nu, w, cycle, over
        self.gs = gs   # node-block Gauss-Seidel (forward before, backward after the coarse correction) instead of damped Jacobi
        while True:
            Dinv = block_diag_inv(A, bs)
            lev = dict(A=A, Dinv=Dinv)
            if gs:
                B = A.tobsr(blocksize=(bs, bs))
                B.sort_indices()
                n = B.shape[0] // bs
                rows = np.repeat(np.arange(n), np.diff(B.indptr))
                for name, keep in (("fwd", B.indices <= rows), ("bwd", B.indices >= rows)):
                    ptr = np.concatenate([[0], np.cumsum(np.bincount(rows[keep], minlength=n))])
                    T = sp.bsr_matrix((B.data[keep], B.indices[keep], ptr), shape=B.shape)
                    lev[name] = spl.splu(T.tocsc(), permc_spec="NATURAL")      # block triangle (D + L) / (D + U): no fill
            self.levels.append(lev)
            if A.shape[0] // bs <= min_size or len(self.levels) > 8:
                lev["lu"] = spl.splu(A.tocsc())
                break
            h = h * factor
            agg = aggregate_grid(coords, h)
            P = prolong(agg, bs)
            if smooth_p:   # smoothed aggregation: one damped (block-)Jacobi step on the tentative prolongator
                DinvA = sp.bsr_matrix((Dinv, np.arange(Dinv.shape[0]), np.arange(Dinv.shape[0] + 1)), shape=A.shape).tocsr() @ A
                P = (P - (2.0 / 3.0) * (DinvA @ P)).tocsr()
            lev["P"] = P
            A = (P.T @ A @ P).tocsr()
            coords = coarse_coords(coords, agg)
        print("   MG levels:", [l["A"].shape[0] // bs for l in self.levels], "nnz", [l["A"].nnz for l in self.levels])

    def smooth(self, lev, x, b, n, backward=False):
        for _ in range(n):
            if self.gs:
                x = x + lev["bwd" if backward else "fwd"].solve(b - lev["A"] @ x)
            else:
                x = x + self.w * apply_bdinv(lev["Dinv"], b - lev["A"] @ x)
        return x

    def cyc(self, k, b):
        lev = self.levels[k]
        if "lu" in lev:
            return lev["lu"].solve(b)
        if self.gs:
            x = self.smooth(lev, np.zeros_like(b), b, self.nu)
        else:
            x = self.w * apply_bdinv(lev["Dinv"], b)
            x = self.smooth(lev, x, b, self.nu - 1)
        r = b - lev["A"] @ x
        rc = lev["P"].T @ r
        ec = self.cyc(k + 1, rc)
        if self.cycle == "W":
            ec = ec + self.cyc(k + 1, rc - self.levels[k + 1]["A"] @ ec)
        x = x + self.over * (lev["P"] @ ec)
        x = self.smooth(lev, x, b, self.nu, backward=True)
        return x

    def __call__(self, r):
        return self.cyc(0, r)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    variants = [v for v in sys.argv[2:] if v != "cloud"] or ["jac", "mono", "schur"]
    t0 = time.time()
    mesh, A, b = build(n, cloud="cloud" in sys.argv)
    bs = mesh.dim + 1
    A_phys = A
    A, b, s = equilibrate(A, b)
    coords = mesh.coords()
    vol = np.abs(mg.det_j(mesh)).mean() / (2 if mesh.dim == 2 else 6)
    h = ((2 if mesh.dim == 2 else 6) * vol) ** (1.0 / mesh.dim)     # edge of the cube whose Kuhn split has the mean volume
    print(f"h estimate {h:.4f}  (1/n = {1.0 / n:.4f})")
    print(f"n={n} nodes={mesh.n_nodes} dof={A.shape[0]} nnz={A.nnz}  build {time.time() - t0:.1f}s")
    for v in variants:
        t0 = time.time()
        if v == "jac":
            Dinv = block_diag_inv(A, bs)
            M = lambda r: apply_bdinv(Dinv, r)     # noqa: E731
        elif v.startswith("mono"):
            # mono[:nu[:w[:factor[:cycle[:over]]]]]
            a = v.split(":")
            nu = int(a[1]) if len(a) > 1 else 1
            w = float(a[2]) if len(a) > 2 else 0.7
            fac = float(a[3]) if len(a) > 3 else 2.0
            cyc = a[4] if len(a) > 4 else "V"
            over = float(a[5]) if len(a) > 5 else 1.0
            sa = len(a) > 6 and a[6] == "sa"
            if v.startswith("monou") and sa:
                mgp = MG(A_phys, coords, h, bs, nu=nu, w=w, factor=fac, cycle=cyc, over=over, smooth_p=True)
                M = lambda r, mgp=mgp: mgp(r / s) / s     # noqa: E731
            elif v.startswith("monou"):      # hierarchy on the physical (unscaled) matrix: P is constant in (v, p)
                mgp = MG(A_phys, coords, h, bs, nu=nu, w=w, factor=fac, cycle=cyc, over=over)
                M = lambda r, mgp=mgp: mgp(r / s) / s     # noqa: E731
            else:
                M = MG(A, coords, h, bs, nu=nu, w=w, factor=fac, cycle=cyc, over=over)
        elif v.startswith("schur"):
            a = v.split(":")
            nu = int(a[1]) if len(a) > 1 else 1
            exact = len(a) > 2 and a[2] == "exact"
            nn = mesh.n_nodes
            iv = (np.arange(nn)[:, None] * bs + np.arange(bs - 1)[None, :]).reshape(-1)
            ip = np.arange(nn) * bs + bs - 1
            Avv, Avp, Apv, App = A[iv][:, iv], A[iv][:, ip], A[ip][:, iv], A[ip][:, ip]
            dvv = Avv.diagonal()
            Shat = (App - Apv @ sp.diags(1.0 / dvv) @ Avp).tocsr()
            print("   Shat nnz", Shat.nnz, "App nnz", App.nnz)
            if exact:
                lu = spl.splu(Shat.tocsc())
                Sinv = lu.solve
            else:
                # schur:nu:ncyc:cycle:over:sa -- ncyc stationary MG iterations per application, optional smoothed aggregation
                ncyc = int(a[2]) if len(a) > 2 else 1
                cyc = a[3] if len(a) > 3 else "V"
                over = float(a[4]) if len(a) > 4 else 1.0
                mgS = MG(Shat, coords, h, 1, nu=nu, w=0.7, cycle=cyc, over=over, smooth_p=(len(a) > 5 and a[5] == "sa"))

                def Sinv(r, mgS=mgS, ncyc=ncyc):
                    z = mgS(r)
                    for _ in range(ncyc - 1):
                        z = z + mgS(r - Shat @ z)
                    return z

            def M(r, Sinv=Sinv):
                z = np.zeros_like(r)
                zv = r[iv] / dvv
                for _ in range(2):
                    zv = zv + (r[iv] - Avv @ zv) / dvv
                zp = Sinv(r[ip] - Apv @ zv)
                z[iv] = zv - (Avp @ zp) / dvv      # block LDU: upper factor as well
                z[ip] = zp
                return z
        else:
            raise SystemExit(f"unknown variant {v}")
        t1 = time.time()
        x, it, res = bicgstab(A, b, M)
        print(f"{v:24s} iters={it:6d}  true rel res={res:.2e}  setup {t1 - t0:.1f}s solve {time.time() - t1:.1f}s")


if __name__ == "__main__":
    main()
