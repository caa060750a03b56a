"""Synthetic meshes and fields for parity tests and benchmarks.

Host-side input generation only (numpy); nothing here is on the hot path.  The reference
gets its meshes from gmsh + CGAL alpha shapes (srcs/mesh/Mesh.cpp:762-917, Mesh2D.cpp,
Mesh3D.cpp), neither of which is available here, so the stand-in is a Kuhn/Freudenthal
split of the unit box: 2 triangles per square / 6 tetrahedra per cube, all with detJ > 0
(the orientation CGAL alpha-shape cells have, Mesh3D.cpp:166-170).

Node flag bits follow the C ABI (include/pfem_b200.h): bit0 isBound, bit1 isFree,
bit2 isFixed, bit3 isOnFreeSurface (Node.hpp:93-105, Node.inl:48-66).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field

import numpy as np

F_BOUND, F_FREE, F_FIXED, F_FS = 1, 2, 4, 8


@dataclass
class Mesh:
    dim: int
    x: np.ndarray          # (dim*nNodes,) float64, layout x[n + d*nNodes]
    conn: np.ndarray       # (nElm, dim+1) int64, row-major, positive orientation
    flags: np.ndarray      # (nNodes,) uint8
    dir_mask: np.ndarray   # (nNodes,) uint8: bound node whose tag carries a velocity BC
    dir_val: np.ndarray    # (dim*nNodes,) float64 Dirichlet velocity, layout n + d*nNodes
    n_cells: int = 0
    meta: dict = field(default_factory=dict)

    @property
    def n_nodes(self) -> int:
        return self.flags.shape[0]

    @property
    def n_elems(self) -> int:
        return self.conn.shape[0]

    def coords(self) -> np.ndarray:
        """(nNodes, dim) view-copy of the positions."""
        return self.x.reshape(self.dim, self.n_nodes).T.copy()


def _perm_sign(p) -> int:
    s = 1
    p = list(p)
    for i in range(len(p)):
        while p[i] != i:
            j = p[i]
            p[i], p[j] = p[j], p[i]
            s = -s
    return s


def kuhn_box(dim: int, n: int, *, jitter: float = 0.1, seed: int = 1234, permute: bool = False,
             perm_seed: int = 4321, free_fraction: float = 0.0, free_seed: int = 7,
             extent: float = 1.0) -> Mesh:
    """Unit box split into n^dim cells, each cut into dim! simplices along the main diagonal.

    jitter: interior nodes are displaced by U(-jitter*h, jitter*h) per coordinate (seed).
    permute: apply a random permutation to node and element numbering (models PFEM node
             insert/delete renumbering, Mesh.cpp:139-144, 181-204).
    free_fraction: append that fraction of extra isolated nodes (isFree: in no element).
    Flags: the faces x=0, x=1, (y=0, y=1,) and the bottom (last coordinate = 0) are
    isBound|isFixed with a zero-velocity Dirichlet BC; the top face is isOnFreeSurface.
    """
    assert dim in (2, 3)
    m = n + 1
    h = extent / n
    axes = [np.arange(m)] * dim
    grid = np.meshgrid(*axes, indexing="ij")           # grid[d][i0,i1,(i2)]
    idx = [g.ravel() for g in grid]
    strides = [m ** (dim - 1 - d) for d in range(dim)]  # lexicographic, last coord fastest
    n_nodes = m ** dim
    pts = np.stack([i.astype(np.float64) * h for i in idx], axis=1)  # (nNodes, dim)

    on_lo = [i == 0 for i in idx]
    on_hi = [i == n for i in idx]
    boundary = np.zeros(n_nodes, dtype=bool)
    for d in range(dim):
        boundary |= on_lo[d] | on_hi[d]
    top = on_hi[dim - 1]
    wall = np.zeros(n_nodes, dtype=bool)
    for d in range(dim - 1):
        wall |= on_lo[d] | on_hi[d]
    wall |= on_lo[dim - 1]

    if jitter > 0:
        rng = np.random.default_rng(seed)
        disp = rng.uniform(-jitter * h, jitter * h, size=(n_nodes, dim))
        disp[boundary] = 0.0
        pts = pts + disp

    # cells and their simplices
    caxes = [np.arange(n)] * dim
    cgrid = np.meshgrid(*caxes, indexing="ij")
    base = sum(c.ravel().astype(np.int64) * s for c, s in zip(cgrid, strides))  # node id of cell origin
    simplices = []
    for perm in itertools.permutations(range(dim)):
        verts = [base]
        cur = base
        for a in perm:
            cur = cur + strides[a]
            verts.append(cur)
        if _perm_sign(perm) < 0:
            verts[-1], verts[-2] = verts[-2], verts[-1]
        simplices.append(np.stack(verts, axis=1))
    # interleave so that the dim! simplices of one cell are consecutive
    conn = np.stack(simplices, axis=1).reshape(-1, dim + 1).astype(np.int64)

    flags = np.zeros(n_nodes, dtype=np.uint8)
    flags[wall] |= F_BOUND | F_FIXED
    flags[top & ~wall] |= F_FS
    dir_mask = wall.astype(np.uint8)

    if free_fraction > 0:
        k = max(1, int(round(free_fraction * n_nodes)))
        rng = np.random.default_rng(free_seed)
        extra = rng.uniform(0.05 * extent, 0.95 * extent, size=(k, dim))
        extra[:, dim - 1] += extent  # detached drops above the box
        pts = np.concatenate([pts, extra], axis=0)
        flags = np.concatenate([flags, np.full(k, F_FREE, dtype=np.uint8)])
        dir_mask = np.concatenate([dir_mask, np.zeros(k, dtype=np.uint8)])
        n_nodes += k

    if permute:
        rng = np.random.default_rng(perm_seed)
        new_of_old = rng.permutation(n_nodes)
        old_of_new = np.empty_like(new_of_old)
        old_of_new[new_of_old] = np.arange(n_nodes)
        pts = pts[old_of_new]
        flags = flags[old_of_new]
        dir_mask = dir_mask[old_of_new]
        conn = new_of_old[conn]
        conn = conn[rng.permutation(conn.shape[0])]

    x = np.ascontiguousarray(pts.T).reshape(-1)
    dir_val = np.zeros(dim * n_nodes, dtype=np.float64)
    return Mesh(dim=dim, x=x, conn=np.ascontiguousarray(conn), flags=flags, dir_mask=dir_mask, dir_val=dir_val,
                n_cells=n, meta=dict(jitter=jitter, seed=seed, permute=permute, free_fraction=free_fraction))


def dam_break(dim: int, ncol: int, *, L: float = 0.146, jitter: float = 0.1, seed: int = 1234, permute: bool = False,
              perm_seed: int = 4321) -> Mesh:
    """Dam-break start configuration in the proportions of the reference's first three configs
    (examples/2D/damBreakKoshizuka/geometry.geo:2-44, examples/3D/damBreakKoshizuka/geometry.geo): a water column of
    width L (x L in 3-D) and height 2L in the left corner of a tank 4L wide and 3L high, node spacing d = L/ncol.
    The column is a Kuhn grid (stand-in for gmsh + CGAL, see the module docstring); the tank walls carry nodes at the
    same spacing.  Wall nodes touching the fluid are isBound|isFixed; the DRY wall nodes (floor right of the column,
    left wall above it, right wall) belong to no element: isBound|isFixed|isFree, the combination every PFEM run has
    and the unit-box meshes lack.  The column's right and top faces are the free surface.  All walls: velocity BC 0."""
    assert dim in (2, 3)
    d = L / ncol
    counts = [ncol] * (dim - 1) + [2 * ncol]                 # cells per axis of the column (last axis = height)
    m = [c + 1 for c in counts]
    grid = np.meshgrid(*[np.arange(k) for k in m], indexing="ij")
    idx = [g.ravel() for g in grid]
    strides = [int(np.prod(m[a + 1:])) for a in range(dim)]
    n_col_nodes = int(np.prod(m))
    pts = np.stack([i.astype(np.float64) * d for i in idx], axis=1)
    wall = (idx[0] == 0) | (idx[dim - 1] == 0)                 # left wall, floor
    if dim == 3:
        wall |= (idx[1] == 0) | (idx[1] == counts[1])          # front and back walls of a tank exactly L deep
    fs = ((idx[0] == counts[0]) | (idx[dim - 1] == counts[dim - 1])) & ~wall
    boundary = wall | fs
    if jitter > 0:
        disp = np.random.default_rng(seed).uniform(-jitter * d, jitter * d, size=pts.shape)
        disp[boundary] = 0.0
        pts = pts + disp
    cgrid = np.meshgrid(*[np.arange(c) for c in counts], indexing="ij")
    base = sum(c.ravel().astype(np.int64) * st for c, st in zip(cgrid, strides))
    simplices = []
    for perm in itertools.permutations(range(dim)):
        verts, cur = [base], base
        for a in perm:
            cur = cur + strides[a]
            verts.append(cur)
        if _perm_sign(perm) < 0:
            verts[-1], verts[-2] = verts[-2], verts[-1]
        simplices.append(np.stack(verts, axis=1))
    conn = np.stack(simplices, axis=1).reshape(-1, dim + 1).astype(np.int64)
    # dry wall nodes of the tank (spacing d): floor x in (L, 4L], left wall z in (2L, 3L], right wall z in (0, 3L]
    lat = [np.arange(m[1]) * d] if dim == 3 else []            # the y lattice of the 3-D tank
    def line(xs, zs):
        cols = [np.asarray(xs, dtype=np.float64)] + [None] * (dim - 2) + [np.asarray(zs, dtype=np.float64)]
        if dim == 2:
            return np.stack([cols[0], cols[1]], axis=1)
        yy = lat[0]
        return np.stack([np.repeat(cols[0], yy.size), np.tile(yy, cols[0].size), np.repeat(cols[2], yy.size)], axis=1)
    kx = np.arange(ncol + 1, 4 * ncol + 1) * d
    kz_left = np.arange(2 * ncol + 1, 3 * ncol + 1) * d
    kz_right = np.arange(1, 3 * ncol + 1) * d
    dry = np.concatenate([line(kx, np.zeros_like(kx)), line(np.zeros_like(kz_left), kz_left),
                          line(np.full_like(kz_right, 4 * L), kz_right)], axis=0)
    pts = np.concatenate([pts, dry], axis=0)
    n_nodes = pts.shape[0]
    flags = np.zeros(n_nodes, dtype=np.uint8)
    flags[:n_col_nodes][wall] |= F_BOUND | F_FIXED
    flags[:n_col_nodes][fs] |= F_FS
    flags[n_col_nodes:] |= F_BOUND | F_FIXED | F_FREE
    dir_mask = ((flags & F_BOUND) != 0).astype(np.uint8)
    if permute:
        rng = np.random.default_rng(perm_seed)
        new_of_old = rng.permutation(n_nodes)
        old_of_new = np.empty_like(new_of_old)
        old_of_new[new_of_old] = np.arange(n_nodes)
        pts, flags, dir_mask = pts[old_of_new], flags[old_of_new], dir_mask[old_of_new]
        conn = new_of_old[conn]
        conn = conn[rng.permutation(conn.shape[0])]
    return Mesh(dim=dim, x=np.ascontiguousarray(pts.T).reshape(-1), conn=np.ascontiguousarray(conn), flags=flags,
                dir_mask=dir_mask, dir_val=np.zeros(dim * n_nodes), n_cells=ncol, meta=dict(kind="dam_break", L=L, d=d))


def boundary_facets(mesh: Mesh) -> np.ndarray:
    """Boundary facets as the reference stores them after the alpha shape (Mesh3D.cpp:218-262, Mesh2D.cpp): one row
    per element face that belongs to exactly one element = [dim facet nodes, the opposite element node
    (Facet::m_outNodeIndex), the element index (Facet::m_elementIndex)], int64, ordered by element then local face."""
    dim, npe = mesh.dim, mesh.dim + 1
    conn = mesh.conn
    ne = conn.shape[0]
    faces = np.empty((ne, npe, dim), dtype=np.int64)
    for k in range(npe):                       # local face k = all element nodes but node k
        faces[:, k, :] = conn[:, [j for j in range(npe) if j != k]]
    key = np.sort(faces.reshape(-1, dim), axis=1)
    _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    once = (cnt[inv.reshape(-1)] == 1).reshape(ne, npe)
    e_idx, k_idx = np.nonzero(once)
    out = np.empty((e_idx.size, dim + 2), dtype=np.int64)
    out[:, :dim] = faces[e_idx, k_idx]
    out[:, dim] = conn[e_idx, k_idx]
    out[:, dim + 1] = e_idx
    return out


def det_j(mesh: Mesh) -> np.ndarray:
    """detJ per element (Element.cpp:71-86) -- used to check orientation of generated meshes."""
    c = mesh.coords()
    p0 = c[mesh.conn[:, 0]]
    J = np.stack([c[mesh.conn[:, k + 1]] - p0 for k in range(mesh.dim)], axis=2)  # (nElm, dim, dim) columns = edges
    return np.linalg.det(J)


# --------------------------------------------------------------------------------------
# Fields
# --------------------------------------------------------------------------------------
PSPG_PARAMS = dict(rho=1000.0, mu=1e-3, dt=1e-3)       # damBreakKoshizuka3DIncomp.lua:30-31,44,50
WC_PARAMS = dict(mu=1e-3, K0=2.2e5, K0p=7.6, rhoStar=1000.0, securityCoeff=0.1)  # ...3DComp.lua:32-34,46


def gravity(dim: int) -> np.ndarray:
    g = np.zeros(3)
    g[dim - 1] = -9.81
    return g


def velocity_field(mesh: Mesh, *, noise: float = 0.01, seed: int = 99, zero_on_walls: bool = True) -> np.ndarray:
    """v = (sin2pi x cos2pi y, -cos2pi x sin2pi y, 0.1 sin2pi z) + noise*N(0,1); layout n + d*nNodes."""
    c = mesh.coords()
    nn = mesh.n_nodes
    v = np.zeros((mesh.dim, nn))
    tp = 2 * np.pi
    v[0] = np.sin(tp * c[:, 0]) * np.cos(tp * c[:, 1])
    v[1] = -np.cos(tp * c[:, 0]) * np.sin(tp * c[:, 1])
    if mesh.dim == 3:
        v[2] = 0.1 * np.sin(tp * c[:, 2])
    rng = np.random.default_rng(seed)
    v += noise * rng.standard_normal(v.shape)
    if zero_on_walls:
        v[:, (mesh.flags & F_BOUND) != 0] = 0.0
    return v.reshape(-1)


def hydrostatic_pressure(mesh: Mesh, rho: float = 1000.0, g: float = 9.81, height: float = 1.0) -> np.ndarray:
    c = mesh.coords()
    return rho * g * np.maximum(height - c[:, mesh.dim - 1], 0.0)


def tait_density(p: np.ndarray, K0: float, K0p: float, rho_star: float) -> np.ndarray:
    """rho = rho*.((K0'/K0) p + 1)^(1/K0')  (WCompNewton/ContEquation.inl:319-331)."""
    return rho_star * np.power((K0p / K0) * p + 1.0, 1.0 / K0p)


def pspg_state(mesh: Mesh, **kw):
    """(q_cur, q_prev): each (dim+1)*nNodes in Q.hpp layout [u,v,(w),p]; v_prev = v."""
    v = velocity_field(mesh, **kw)
    p = hydrostatic_pressure(mesh)
    q = np.concatenate([v, p])
    return q.copy(), q.copy()


def wc_state(mesh: Mesh, **kw):
    """dict of SoA arrays for the weakly-compressible problem: v, p, rho, acc (WC/Problem.cpp:17-18)."""
    v = velocity_field(mesh, **kw)
    p = hydrostatic_pressure(mesh)
    rho = tait_density(p, WC_PARAMS["K0"], WC_PARAMS["K0p"], WC_PARAMS["rhoStar"])
    acc = np.zeros_like(v)
    return dict(v=v, p=p, rho=rho, acc=acc)


def delaunay_cloud(dim: int, n_points: int, *, seed: int = 11, alpha: float = 1.3, free_fraction: float = 0.0) -> Mesh:
    """Unstructured stand-in for the gmsh + CGAL alpha-shape meshes of the examples (no gmsh/CGAL here): Delaunay
    triangulation of a jittered point cloud in the unit box, cells kept when their circumradius < alpha*h
    (the criterion of Mesh2D.cpp:54 / Mesh3D.cpp:60), positive orientation enforced.  Node valence varies widely
    (3-D: up to ~40 incident tets), which exercises the multi-chunk paths of the gather kernels."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    m = max(2, int(round(n_points ** (1.0 / dim))))
    h = 1.0 / (m - 1)
    axes = np.meshgrid(*([np.linspace(0, 1, m)] * dim), indexing="ij")
    pts = np.stack([a.ravel() for a in axes], axis=1)
    on_b = ((pts == 0) | (pts == 1)).any(axis=1)
    pts = pts + rng.uniform(-0.35 * h, 0.35 * h, pts.shape) * (~on_b)[:, None]
    tri = Delaunay(pts)
    conn = tri.simplices.astype(np.int64)
    p0 = pts[conn[:, 0]]
    J = np.stack([pts[conn[:, k + 1]] - p0 for k in range(dim)], axis=2)
    det = np.linalg.det(J)
    neg = det < 0
    conn[neg, -1], conn[neg, -2] = conn[neg, -2].copy(), conn[neg, -1].copy()
    det = np.abs(det)
    # circumradius filter + drop slivers
    A = 2 * np.transpose(J, (0, 2, 1))
    rhs = (np.transpose(J, (0, 2, 1)) ** 2).sum(axis=2)
    ok = det > 1e-9 * h ** dim
    cc = np.zeros((conn.shape[0], dim))
    cc[ok] = np.linalg.solve(A[ok], rhs[ok][..., None])[..., 0]
    rad = np.linalg.norm(cc, axis=1)
    conn = conn[ok & (rad < alpha * h)]
    n_nodes = pts.shape[0]
    wall = np.zeros(n_nodes, dtype=bool)
    for d in range(dim - 1):
        wall |= (pts[:, d] == 0) | (pts[:, d] == 1)
    wall |= pts[:, dim - 1] == 0
    flags = np.zeros(n_nodes, dtype=np.uint8)
    flags[wall] |= F_BOUND | F_FIXED
    flags[(pts[:, dim - 1] == 1) & ~wall] |= F_FS
    used = np.zeros(n_nodes, dtype=bool)
    used[conn.ravel()] = True
    flags[~used] |= F_FREE                                       # nodes left without elements by the alpha criterion
    dir_mask = wall.astype(np.uint8)
    if free_fraction > 0:
        k = max(1, int(round(free_fraction * n_nodes)))
        extra = rng.uniform(0.05, 0.95, size=(k, dim))
        extra[:, dim - 1] += 1.0
        pts = np.concatenate([pts, extra])
        flags = np.concatenate([flags, np.full(k, F_FREE, dtype=np.uint8)])
        dir_mask = np.concatenate([dir_mask, np.zeros(k, dtype=np.uint8)])
        n_nodes += k
    x = np.ascontiguousarray(pts.T).reshape(-1)
    return Mesh(dim=dim, x=x, conn=np.ascontiguousarray(conn), flags=flags, dir_mask=dir_mask,
                dir_val=np.zeros(dim * n_nodes), n_cells=m, meta=dict(kind="delaunay", seed=seed))
