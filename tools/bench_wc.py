#!/usr/bin/env python
"""Explicit weakly-compressible step on the C5 mesh (BASELINE.json configs[4]): synthetic Kuhn box n=150 -> 20.25 M tets,
RCB-sharded over the ranks of one box (STRONG scaling: the mesh is fixed, as north_star names it).

    python tools/bench_wc.py [--cells 150] [--steps 20]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_wc.py ...

One step = pfem_wc_step (half kick + move, continuity, momentum, 2 halo exchanges) + pfem_wc_next_dt (CFL min + all-reduce).
Prints one JSON line: Melem/s, ms/step and the fraction of the HBM roofline with the ALGORITHMIC bytes of SURVEY.md 8(d):
B_wc_step = 2*nElm*npe*4 + nNodes*8*(11 + 11 + 12).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pfem_b200 import meshgen as mg  # noqa: E402
from pfem_b200.capi import PfemContext  # noqa: E402
from pfem_b200.partition import partition_mesh  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=150)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--permute", action="store_true", help="(R) ordering: random node/element numbering (seed 4321)")
    ap.add_argument("--renumber", action="store_true", help="Morton renumbering on the host before the upload (pfem_b200/renumber.py)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lrank = int(os.environ.get("LOCAL_RANK", "0"))
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(lrank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    dim = 3
    gmesh = mg.kuhn_box(dim, args.cells, permute=args.permute)
    t_renum = 0.0
    if args.renumber:
        from pfem_b200.renumber import spatial_renumber
        t0 = time.perf_counter()
        gmesh = spatial_renumber(gmesh).mesh
        t_renum = time.perf_counter() - t0
    n_elems, n_nodes = gmesh.n_elems, gmesh.n_nodes
    st = mg.wc_state(gmesh)
    packed = np.concatenate([st["v"], st["p"], st["rho"], st["acc"]])
    W = mg.WC_PARAMS
    ctx = PfemContext(dim, lrank)
    t_part = 0.0
    if world > 1:
        uid = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])
        t0 = time.perf_counter()
        part = partition_mesh(gmesh, world, rank)
        t_part = time.perf_counter() - t0
        mesh = part.mesh
        packed = part.scatter_nodal(packed, 2 * dim + 2, n_nodes)
        del gmesh
    else:
        part, mesh = None, gmesh
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    ctx.set_topology(mesh.conn, mesh.flags)
    if part is not None:
        ctx.set_partition(part)
    t_topo = time.perf_counter() - t0
    ctx.set_positions(mesh.x)
    ctx.set_dirichlet(mesh.dir_mask, mesh.dir_val)
    ctx.set_states(0, packed)
    wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], mg.gravity(dim), True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        dt = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
        for _ in range(args.warmup):
            ctx.wc_step(wp, dt)
            dt = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
        ctx.profile_enable(True)
        ctx.profile_reset()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            ctx.wc_step(wp, dt)
            dt = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1) / args.steps
        phases = {}
        for ph in ("Update solutions", "Solving continuity eq", "Solving momentum eq", "Compute next dt", "Halo exchange"):
            t, n = ctx.profile_get(ph)
            phases[ph] = t / max(n, 1)
    tmax = torch.tensor([ms] + [phases[k] for k in sorted(phases)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax[0])
    phases = {k: float(v) for k, v in zip(sorted(phases), tmax[1:])}
    peak = 6550.7
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    b_alg = 2 * n_elems * 4 * 4 + n_nodes * 8 * (11 + 11 + 12)
    kern_ms = phases["Update solutions"] + phases["Solving continuity eq"] + phases["Solving momentum eq"]
    line = {"metric": "explicit weakly-compressible step Melem/s", "value": n_elems / (ms * 1e-3) / 1e6, "unit": "Melem/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "scaling": "strong", "dtype": "f64",
            "config": {"workload": f"C5 synthetic 3D Kuhn box n={args.cells}: {n_elems} tets, {n_nodes} nodes, CDS_dpdt + Meduri, "
                                   "step + CFL dt per step", "partition": "RCB nodes + ghost-element layer" if world > 1 else "none",
                       "ordering": ("random (R)" if args.permute else "lexicographic (L)") + (" + host Morton renumbering" if args.renumber else ""),
                       "renumber_host_s": t_renum},
            "phases_ms": phases, "dt": dt, "pattern_build_s": t_topo, "partition_host_s": t_part,
            "roofline": {"bound": "hbm", "algorithmic_bytes": b_alg, "achieved": b_alg / (kern_ms * 1e-3) / 1e9,
                         "peak": peak * world, "unit": "GB/s", "frac": b_alg / (kern_ms * 1e-3) / 1e9 / (peak * world),
                         "kernels": "kick/move + continuity + momentum kernels (device time, max over ranks)"}}
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
