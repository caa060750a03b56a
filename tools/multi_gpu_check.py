"""Launched by torch.distributed.run on N GPUs: sharded explicit WC steps and a sharded PSPG assemble + BiCGSTAB solve,
each compared with the same computation on ONE GPU (rank 0 runs the single-GPU reference on its own device)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfem_b200 import meshgen as mg
from pfem_b200.capi import PfemContext
from pfem_b200.partition import partition_mesh


def main():
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    dim = 3
    mesh = mg.kuhn_box(dim, cells, free_fraction=0.002, permute=True)
    g = mg.gravity(dim)
    part = partition_mesh(mesh, world, rank)
    nn, nl = mesh.n_nodes, part.mesh.n_nodes

    ctx = PfemContext(dim, lrank)
    uid = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])
    ctx.set_mesh(part.mesh)
    ctx.set_partition(part)

    ok = True
    # ------------------------------------------------------------------ explicit weakly-compressible steps
    W = mg.WC_PARAMS
    for variant in (6, 11):   # pfem_wc_set_variant: gather kernels, two-pass element records
        ctx.set_positions(part.mesh.x)
        ctx.wc_set_variant(variant)
        st = mg.wc_state(mesh)
        st["acc"] = 0.3 * np.random.default_rng(4).standard_normal(st["acc"].shape)
        packed = np.concatenate([st["v"], st["p"], st["rho"], st["acc"]])
        ctx.set_states(0, part.scatter_nodal(packed, 2 * dim + 2, nn))
        wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, True)
        dts = []
        for _ in range(3):
            dt = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
            dts.append(dt)
            ctx.wc_step(wp, dt)
        loc = ctx.get_states(0, 2 * dim + 2).reshape(2 * dim + 2, nl)
        xloc = ctx.get_positions().reshape(dim, nl)
        # gather owned values on rank 0
        glob = torch.zeros((2 * dim + 2 + dim, nn), dtype=torch.float64, device="cuda")
        idx = torch.from_numpy(part.l2g_nodes[: part.n_owned]).cuda()
        glob[: 2 * dim + 2, idx] = torch.from_numpy(loc[:, : part.n_owned]).cuda()
        glob[2 * dim + 2:, idx] = torch.from_numpy(xloc[:, : part.n_owned]).cuda()
        dist.all_reduce(glob)
        # ghost copies must equal the owners' values after the last exchange
        gl = glob.cpu().numpy()
        ghost_ok = np.array_equal(loc[: dim + 2, part.n_owned:], gl[: dim + 2, part.l2g_nodes[part.n_owned:]])
        if rank == 0:
            with PfemContext(dim, lrank) as one:
                one.set_mesh(mesh)
                one.set_states(0, packed)
                one.wc_set_variant(variant)
                wp1 = one.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], g, True)
                dts1 = []
                for _ in range(3):
                    dt = one.wc_next_dt(wp1, W["securityCoeff"], 1e-3)
                    dts1.append(dt)
                    one.wc_step(wp1, dt)
                ref = one.get_states(0, 2 * dim + 2).reshape(2 * dim + 2, nn)
                xref = one.get_positions().reshape(dim, nn)
            same = np.array_equal(gl[: 2 * dim + 2], ref) and np.array_equal(gl[2 * dim + 2:], xref) and dts == dts1
            err = np.abs(gl[: 2 * dim + 2] - ref).max()
            print(f"[wc variant {variant}] {world} GPUs vs 1 GPU: bit-identical={same} max|diff|={err:.3e} dts={dts}")
            ok &= same
        ok &= ghost_ok

    # ------------------------------------------------------------------ PSPG assemble + distributed BiCGSTAB + Picard body
    P = mg.PSPG_PARAMS
    q, q_prev = mg.pspg_state(mesh)
    ctx.set_positions(part.mesh.x)
    ctx.set_states(0, part.scatter_nodal(q, dim + 1, nn))
    par = ctx.pspg_params(P["rho"], P["mu"], P["dt"], g)
    ctx.snapshot_positions()
    ctx.pspg_assemble(par, part.scatter_nodal(q_prev, dim + 1, nn))
    sol = ctx.pspg_solve(1e-12, 20000)
    out = ctx.pspg_picard_iter(par, None, 1e-12, 20000)
    gq = torch.zeros((2, dim + 1, nn), dtype=torch.float64, device="cuda")
    gq[0][:, idx] = torch.from_numpy(sol["q"].reshape(dim + 1, nl)[:, : part.n_owned]).cuda()
    gq[1][:, idx] = torch.from_numpy(out["q"].reshape(dim + 1, nl)[:, : part.n_owned]).cuda()
    dist.all_reduce(gq)
    if rank == 0:
        with PfemContext(dim, lrank) as one:
            one.set_mesh(mesh)
            one.set_states(0, q)
            par1 = one.pspg_params(P["rho"], P["mu"], P["dt"], g)
            one.snapshot_positions()
            one.pspg_assemble(par1, q_prev)
            s1 = one.pspg_solve(1e-12, 20000)
            o1 = one.pspg_picard_iter(par1, None, 1e-12, 20000)
        for name, a, b in (("solve", gq[0].cpu().numpy().reshape(-1), s1["q"]), ("picard", gq[1].cpu().numpy().reshape(-1), o1["q"])):
            ev = np.abs(a[: dim * nn] - b[: dim * nn]).max() / np.abs(b[: dim * nn]).max()
            ep = np.abs(a[dim * nn:] - b[dim * nn:]).max() / np.abs(b[dim * nn:]).max()
            print(f"[pspg {name}] {world} GPUs vs 1 GPU: rel|dv|={ev:.2e} rel|dp|={ep:.2e}")
            ok &= ev < 1e-8 and ep < 1e-8
        print(f"[pspg] iters {sol['iters']} vs {s1['iters']}, status {sol['status']}/{s1['status']}, picard res {out['res']:.3e} vs {o1['res']:.3e}")
        ok &= sol["status"] == 0 and out["status"] == 0 and abs(out["res"] - o1["res"]) <= 1e-6 * max(o1["res"], 1e-30) + 1e-9
    ctx.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "OK" if flag.item() == 1 else "FAILED")
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
