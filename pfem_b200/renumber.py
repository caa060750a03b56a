"""Spatial renumbering of a mesh before it is uploaded (host side, numpy; SURVEY.md section 8d: "headline = (L) after the
library's own renumbering; (R) reported beside it").

PFEM node/element numbering drifts towards random as nodes are added and removed (Mesh.cpp:27-147, 928-1017): the
gather kernels then lose the L1/L2 locality of neighbouring records (explicit step 2x slower on a randomly permuted
mesh).  Nothing in the C ABI depends on the numbering, so the caller (the shim's uploadMesh, tools/bench_wc.py) may
renumber per remesh: nodes along a Morton (Z-order) curve of their coordinates, elements by their smallest new node
index.  `Renumbering.to_new / to_old` map nodal arrays in the ABI layout q[n + s*nNodes] in both directions; results
are those of the original mesh up to the summation order of the element loops (1e-13, not bit-identical).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .meshgen import Mesh


def _part1by2(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint64) & np.uint64(0x1FFFFF)
    v = (v | (v << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
    v = (v | (v << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    v = (v | (v << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
    return v


def morton_codes(coords: np.ndarray) -> np.ndarray:
    """coords: (nNodes, dim) -> 63-bit Z-order keys on a 2^21 grid over the bounding box."""
    lo, hi = coords.min(0), coords.max(0)
    span = np.where(hi > lo, hi - lo, 1.0)
    g = np.minimum(((coords - lo) / span * (2 ** 21 - 1)).astype(np.uint64), np.uint64(2 ** 21 - 1))
    key = np.zeros(coords.shape[0], dtype=np.uint64)
    for d in range(coords.shape[1]):
        key |= _part1by2(g[:, d]) << np.uint64(d)
    return key


@dataclass
class Renumbering:
    mesh: Mesh              # the renumbered mesh
    new_of_old: np.ndarray  # node permutation: new index of old node
    old_of_new: np.ndarray
    elem_old_of_new: np.ndarray

    def to_new(self, q: np.ndarray) -> np.ndarray:
        """nodal array in ABI layout q[n + s*nNodes], old numbering -> new numbering"""
        nn = self.old_of_new.size
        return np.ascontiguousarray(q.reshape(-1, nn)[:, self.old_of_new].reshape(-1))

    def to_old(self, q: np.ndarray) -> np.ndarray:
        nn = self.old_of_new.size
        return np.ascontiguousarray(q.reshape(-1, nn)[:, self.new_of_old].reshape(-1))


def spatial_renumber(mesh: Mesh) -> Renumbering:
    nn = mesh.n_nodes
    old_of_new = np.argsort(morton_codes(mesh.coords()), kind="stable")
    new_of_old = np.empty(nn, dtype=np.int64)
    new_of_old[old_of_new] = np.arange(nn)
    conn = new_of_old[mesh.conn]
    e_order = np.argsort(conn.min(axis=1), kind="stable")
    out = Mesh(dim=mesh.dim, x=np.ascontiguousarray(mesh.x.reshape(mesh.dim, nn)[:, old_of_new].reshape(-1)),
               conn=np.ascontiguousarray(conn[e_order]), flags=np.ascontiguousarray(mesh.flags[old_of_new]),
               dir_mask=np.ascontiguousarray(mesh.dir_mask[old_of_new]),
               dir_val=np.ascontiguousarray(mesh.dir_val.reshape(mesh.dim, nn)[:, old_of_new].reshape(-1)),
               n_cells=mesh.n_cells, meta=dict(mesh.meta))
    return Renumbering(out, new_of_old, old_of_new, e_order)
