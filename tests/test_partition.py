"""Host-side partition logic of the multi-GPU path (CPU only; world_size-2 gloo run included)."""
import os
import sys

import numpy as np
import pytest

from pfem_b200 import meshgen as mg
from pfem_b200.partition import gather_owned, partition_mesh, rcb_owner


@pytest.mark.parametrize("dim,n,n_ranks", [(2, 9, 2), (2, 8, 3), (3, 5, 2), (3, 6, 4), (3, 6, 8)])
def test_partition_is_consistent(dim, n, n_ranks):
    mesh = mg.kuhn_box(dim, n, free_fraction=0.02, permute=True)
    owner = rcb_owner(mesh.coords(), n_ranks)
    counts = np.bincount(owner, minlength=n_ranks)
    assert counts.sum() == mesh.n_nodes and counts.max() - counts.min() <= n_ranks      # balanced bisection
    parts = [partition_mesh(mesh, n_ranks, r, owner) for r in range(n_ranks)]
    seen = np.zeros(mesh.n_nodes, dtype=int)
    for p in parts:
        seen[p.l2g_nodes[: p.n_owned]] += 1
        lm = p.mesh
        assert (np.diff(p.l2g_elems) > 0).all()                                          # global element order kept
        assert (mesh.conn[p.l2g_elems] == p.l2g_nodes[lm.conn]).all()
        # every element incident to an owned node is local
        inc = np.isin(mesh.conn, p.l2g_nodes[: p.n_owned]).any(axis=1)
        assert set(np.flatnonzero(inc)) == set(p.l2g_elems)
        # ghosts grouped by owner, contiguous receive ranges covering all ghosts
        gown = owner[p.l2g_nodes[p.n_owned:]]
        assert (np.diff(gown) >= 0).all()
        assert sum(p.recv_count) == lm.n_nodes - p.n_owned
    assert (seen == 1).all()
    # send list r->q names exactly q's ghost range owned by r, in the same order
    for r, pr in enumerate(parts):
        for k, q in enumerate(pr.peers):
            pq = parts[q]
            kq = pq.peers.index(r)
            sent = pr.l2g_nodes[pr.send_idx[k]]
            expected = pq.l2g_nodes[pq.recv_start[kq]: pq.recv_start[kq] + pq.recv_count[kq]]
            assert (sent == expected).all()
            assert (pr.send_idx[k] < pr.n_owned).all()
    # exactly one primary owner per element
    prim = np.zeros(mesh.n_elems, dtype=int)
    for p in parts:
        prim[p.l2g_elems[p.elem_primary]] += 1
    assert (prim == 1).all()


def test_owner_computes_gather_equals_global():
    """A node-gather quantity (lumped element size per node) computed shard by shard equals the global one bitwise."""
    mesh = mg.kuhn_box(3, 5)
    vol = mg.det_j(mesh) / 6.0
    ref = np.zeros(mesh.n_nodes)
    for e in range(mesh.n_elems):                      # ascending element order, like the reference's serial scatter
        ref[mesh.conn[e]] += vol[e] / 4
    parts = [partition_mesh(mesh, 4, r) for r in range(4)]
    vals = []
    for p in parts:
        lv = mg.det_j(p.mesh) / 6.0
        acc = np.zeros(p.mesh.n_nodes)
        for e in range(p.mesh.n_elems):
            acc[p.mesh.conn[e]] += lv[e] / 4
        vals.append(acc)
    got = gather_owned(vals, parts, 1, mesh.n_nodes)
    assert (got == ref).all()


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mesh = mg.kuhn_box(3, 4, permute=True)
    part = partition_mesh(mesh, world, rank)
    truth = np.sin(np.arange(mesh.n_nodes, dtype=np.float64))            # a global nodal field
    local = np.full(part.mesh.n_nodes, np.nan)
    local[: part.n_owned] = truth[part.l2g_nodes[: part.n_owned]]       # owners know their values, ghosts do not
    reqs = []
    recv_bufs = []
    for k, p in enumerate(part.peers):
        send = torch.from_numpy(np.ascontiguousarray(local[part.send_idx[k]]))
        reqs.append(dist.isend(send, dst=p))
        buf = torch.empty(part.recv_count[k], dtype=torch.float64)
        recv_bufs.append((k, buf))
        reqs.append(dist.irecv(buf, src=p))
    for r in reqs:
        r.wait()
    for k, buf in recv_bufs:
        local[part.recv_start[k]: part.recv_start[k] + part.recv_count[k]] = buf.numpy()
    ok = bool((local == truth[part.l2g_nodes]).all())
    # dot product / min all-reduce over owned entries
    t = torch.tensor([float((local[: part.n_owned] ** 2).sum())], dtype=torch.float64)
    dist.all_reduce(t)
    ok = ok and abs(t.item() - float((truth ** 2).sum())) < 1e-9
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok))


def test_halo_exchange_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


@pytest.mark.parametrize("dim,n,n_ranks", [(3, 10, 2), (3, 12, 4), (3, 9, 8), (2, 40, 3), (3, 11, 5)])
def test_native_partitioner_equals_the_numpy_statement(dim, n, n_ranks):
    """csrc/partition.cu (pfem_partition_*, what the shim and the bench call) against partition.py: same owners, same
    local meshes, same halo plans -- both order points along the widest axis by (coordinate, node index)."""
    from pfem_b200.capi import NativePartition
    from pfem_b200.partition import partition_mesh_native

    mesh = mg.kuhn_box(dim, n, free_fraction=0.002, permute=True)
    h = NativePartition(dim, mesh.conn, mesh.x, n_ranks)
    try:
        assert (h.owner() == rcb_owner(mesh.coords(), n_ranks)).all()
        for r in range(n_ranks):
            a = partition_mesh(mesh, n_ranks, r)
            b = partition_mesh_native(mesh, n_ranks, r, handle=h)
            assert a.n_owned == b.n_owned and (a.l2g_nodes == b.l2g_nodes).all() and (a.l2g_elems == b.l2g_elems).all()
            assert (a.mesh.conn == b.mesh.conn).all() and a.peers == b.peers
            assert a.recv_start == b.recv_start and a.recv_count == b.recv_count
            assert all((x == y).all() for x, y in zip(a.send_idx, b.send_idx))
    finally:
        h.close()
