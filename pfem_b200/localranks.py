"""Run a partitioned mesh on the ranks of ONE process (pfem_comm_local_*): one host thread and one context per rank.

The ranks may share a device (parity tests on a single-GPU box) or use one device each (the reference's single host
process driving several GPUs).  ctypes releases the GIL during every library call, so the rank threads really run
concurrently and meet at the library's internal barriers.
"""
from __future__ import annotations

import threading

from .capi import LocalGroup, NativePartition, PfemContext
from .partition import partition_mesh_native


def run_ranks(mesh, n_ranks, fn, devices=None, native_partition=True):
    """fn(rank, ctx, part) -> result, called concurrently for every rank with the local mesh already uploaded
    (topology, positions, Dirichlet data, halo plan).  Returns the list of results; re-raises the first failure."""
    devices = devices or [0] * n_ranks
    group = LocalGroup(n_ranks)
    handle = NativePartition(mesh.dim, mesh.conn, mesh.x, n_ranks)
    results, errors = [None] * n_ranks, [None] * n_ranks
    # contexts join the group before any thread starts, so that a rank failing early can still release the others
    ctxs = [PfemContext(mesh.dim, devices[r]) for r in range(n_ranks)]
    for r, cx in enumerate(ctxs):
        cx.comm_init_local(group, r)

    def work(r):
        ctx = ctxs[r]
        try:
            part = partition_mesh_native(mesh, n_ranks, r, handle=handle)
            ctx.set_mesh(part.mesh)
            ctx.set_partition(part)
            results[r] = fn(r, ctx, part)
        except BaseException as e:  # noqa: BLE001 -- reported to the caller below
            errors[r] = e
            if ctx is not None:
                ctx.comm_abort()
        finally:
            if ctx is not None:
                ctx.close()

    threads = [threading.Thread(target=work, args=(r,), name=f"pfem-rank-{r}") for r in range(n_ranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    handle.close()
    group.close()
    # report the root cause, not the "aborted by another rank" it caused on the other ranks
    real = [e for e in errors if e is not None and "aborted by another rank" not in str(e)]
    for e in real + [e for e in errors if e is not None]:
        raise e
    return results
