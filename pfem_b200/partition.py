"""Spatial (RCB) partition of a mesh over the GPUs of one box, with one ghost-element layer.

Host-side logic, once per remesh; the reference has no counterpart (it is single-process OpenMP).

Scheme -- "owner computes":
  * nodes are split by recursive coordinate bisection of their coordinates into `n_ranks` parts (n_ranks need not be a
    power of two: cuts are proportional);
  * a rank keeps every element incident to at least one of its nodes, sorted by GLOBAL element index, so the
    node-gather kernels sum element contributions in exactly the order a single GPU would: sharded results are
    bit-identical to the 1-GPU results;
  * local node numbering = owned nodes (ascending global id) followed by ghost nodes grouped by owner rank
    (ascending global id inside a group), so every halo receive lands in one contiguous range of the nodal arrays;
  * per peer: `send` = local ids of owned nodes that are ghosts on that peer, in the peer's ghost order.
After each element pass the owners send the updated nodal records of those nodes; there is no reduction of partial
sums, and dot products / the CFL minimum are all-reduced.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .meshgen import Mesh


def rcb_owner(coords: np.ndarray, n_ranks: int) -> np.ndarray:
    """Recursive coordinate bisection: rank id per point.  Deterministic: points are ordered along the widest axis by
    (coordinate, point index), the same total order csrc/partition.cu selects with."""
    n = coords.shape[0]
    owner = np.zeros(n, dtype=np.int32)

    def split(idx, r0, nr):
        if nr == 1:
            owner[idx] = r0
            return
        pts = coords[idx]
        axis = int(np.argmax(pts.max(axis=0) - pts.min(axis=0))) if idx.size else 0
        order = np.lexsort((idx, pts[:, axis]))         # total order (coordinate, node index): no dependence on tie handling
        nl = nr // 2
        cut = (idx.size * nl) // nr
        split(idx[order[:cut]], r0, nl)
        split(idx[order[cut:]], r0 + nl, nr - nl)

    split(np.arange(n, dtype=np.int64), 0, n_ranks)
    return owner


@dataclass
class LocalPart:
    rank: int
    n_ranks: int
    mesh: Mesh                      # local mesh: owned nodes first, then ghosts; elements in global order
    n_owned: int
    l2g_nodes: np.ndarray           # local -> global node id
    l2g_elems: np.ndarray           # local -> global element id
    elem_primary: np.ndarray        # bool per local element: this rank is the element's primary owner (owner of node 0)
    peers: list = field(default_factory=list)          # peer ranks, ascending
    send_idx: list = field(default_factory=list)       # per peer: local ids (owned) to send, in the peer's ghost order
    recv_start: list = field(default_factory=list)     # per peer: first local id of the contiguous ghost range
    recv_count: list = field(default_factory=list)

    def scatter_nodal(self, q_global: np.ndarray, n_comp: int, n_nodes_global: int) -> np.ndarray:
        """Global SoA vector q[n + s*N] -> local SoA vector (owned + ghost nodes)."""
        qg = q_global.reshape(n_comp, n_nodes_global)
        return np.ascontiguousarray(qg[:, self.l2g_nodes]).reshape(-1)


def partition_mesh(mesh: Mesh, n_ranks: int, rank: int, owner: np.ndarray | None = None) -> LocalPart:
    dim, nn = mesh.dim, mesh.n_nodes
    if owner is None:
        owner = rcb_owner(mesh.coords(), n_ranks)
    conn = mesh.conn
    eo = owner[conn]                                     # (nE, npe) owner of each element node
    mine = (eo == rank).any(axis=1)
    l2g_elems = np.flatnonzero(mine)                     # ascending global element id
    lconn_g = conn[l2g_elems]
    owned = np.flatnonzero(owner == rank)
    touched = np.unique(lconn_g)
    ghosts = touched[owner[touched] != rank]
    gorder = np.lexsort((ghosts, owner[ghosts]))         # by owner rank, then global id
    ghosts = ghosts[gorder]
    l2g_nodes = np.concatenate([owned, ghosts]).astype(np.int64)
    g2l = np.full(nn, -1, dtype=np.int64)
    g2l[l2g_nodes] = np.arange(l2g_nodes.size)
    lconn = g2l[lconn_g]
    assert (lconn >= 0).all()

    xg = mesh.x.reshape(dim, nn)
    dv = mesh.dir_val.reshape(dim, nn)
    lmesh = Mesh(dim=dim, x=np.ascontiguousarray(xg[:, l2g_nodes]).reshape(-1), conn=np.ascontiguousarray(lconn),
                 flags=np.ascontiguousarray(mesh.flags[l2g_nodes]), dir_mask=np.ascontiguousarray(mesh.dir_mask[l2g_nodes]),
                 dir_val=np.ascontiguousarray(dv[:, l2g_nodes]).reshape(-1), n_cells=mesh.n_cells, meta=dict(mesh.meta))

    part = LocalPart(rank=rank, n_ranks=n_ranks, mesh=lmesh, n_owned=int(owned.size), l2g_nodes=l2g_nodes,
                     l2g_elems=l2g_elems, elem_primary=(eo[l2g_elems, 0] == rank))

    # receive side: contiguous ghost ranges per owner rank
    gowner = owner[ghosts]
    # send side: my owned nodes that are ghosts on peer r  <=>  they share an element with a node owned by r
    pairs_node, pairs_rank = [], []
    ranks_in_elem = eo[l2g_elems]                        # only my local elements can contain my owned nodes
    for a in range(conn.shape[1]):
        na = lconn_g[:, a]
        own_a = ranks_in_elem[:, a] == rank
        for b in range(conn.shape[1]):
            if a == b:
                continue
            sel = own_a & (ranks_in_elem[:, b] != rank)
            pairs_node.append(na[sel])
            pairs_rank.append(ranks_in_elem[sel, b])
    if pairs_node:
        pn = np.concatenate(pairs_node)
        pr = np.concatenate(pairs_rank)
        key = np.unique(pr.astype(np.int64) * nn + pn)
        s_rank, s_node = key // nn, key % nn             # sorted by rank then global id == the peer's ghost order
    else:
        s_rank = s_node = np.zeros(0, dtype=np.int64)
    peers = sorted(set(np.unique(gowner).tolist()) | set(np.unique(s_rank).tolist()))
    for p in peers:
        part.peers.append(int(p))
        part.send_idx.append(g2l[s_node[s_rank == p]].astype(np.int32))
        sel = np.flatnonzero(gowner == p)
        part.recv_start.append(int(owned.size + (sel[0] if sel.size else 0)))
        part.recv_count.append(int(sel.size))
    return part


def partition_mesh_native(mesh: Mesh, n_ranks: int, rank: int, handle=None) -> LocalPart:
    """The same partition computed by the library (csrc/partition.cu: pfem_partition_*, C++/OpenMP) -- what the shim and the
    bench use; `partition_mesh` above is the numpy statement the tests compare it with."""
    from .capi import NativePartition

    own = handle is None
    h = handle or NativePartition(mesh.dim, mesh.conn, mesh.x, n_ranks)
    try:
        d = h.local(rank)
    finally:
        if own:
            h.close()
    dim, nn = mesh.dim, mesh.n_nodes
    l2g = d["l2g_nodes"]
    xg = mesh.x.reshape(dim, nn)
    dv = mesh.dir_val.reshape(dim, nn)
    lmesh = Mesh(dim=dim, x=np.ascontiguousarray(xg[:, l2g]).reshape(-1), conn=d["conn"].astype(np.int64),
                 flags=np.ascontiguousarray(mesh.flags[l2g]), dir_mask=np.ascontiguousarray(mesh.dir_mask[l2g]),
                 dir_val=np.ascontiguousarray(dv[:, l2g]).reshape(-1), n_cells=mesh.n_cells, meta=dict(mesh.meta))
    part = LocalPart(rank=rank, n_ranks=n_ranks, mesh=lmesh, n_owned=d["n_owned"], l2g_nodes=l2g, l2g_elems=d["l2g_elems"],
                     elem_primary=None)
    so = d["send_offsets"]
    for k, p in enumerate(d["peers"]):
        part.peers.append(int(p))
        part.send_idx.append(d["send_idx"][so[k]:so[k + 1]].astype(np.int32))
        part.recv_start.append(int(d["recv_start"][k]))
        part.recv_count.append(int(d["recv_count"][k]))
    return part


def gather_owned(parts_values, parts, n_comp: int, n_nodes_global: int) -> np.ndarray:
    """Inverse of scatter for tests: owned entries of each local SoA vector -> global SoA vector."""
    out = np.zeros((n_comp, n_nodes_global))
    for v, p in zip(parts_values, parts):
        nl = p.l2g_nodes.size
        out[:, p.l2g_nodes[: p.n_owned]] = v.reshape(n_comp, nl)[:, : p.n_owned]
    return out.reshape(-1)
