#!/usr/bin/env python
"""bench.py -- PFEM3D finite-element hot path on B200: FE assembly Melem/s + Krylov solve ms/step, % of HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--cells n]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

Workload at N=1 (BASELINE.json configs[3], "C4"): synthetic Kuhn box n=69 -> 1 971 054 tets, 343 000 nodes,
1 372 000 dof.  One STEP = one body of the PSPG Picard loop: m_buildAbPSPG + m_applyBCPSPG (assembly) followed by the
linear solve (multigrid-preconditioned BiCGSTAB to ||r||/||b|| <= 1e-10), inputs resident in HBM.
At N>1 the box grows with N (n = round(69 N^(1/3)): ~2 M tets per GPU, weak scaling); nodes are split by RCB, every
rank assembles the rows of its nodes from its elements + one ghost-element layer (no collective in the assembly), the
Krylov solve exchanges interface values of x before each SpMV and all-reduces its dot products over NCCL/NVLink.

`value`       = assembly throughput: elements of the whole mesh / max-over-ranks device time of the assembly inside the
                timed steps (CUDA events on the launching stream).
`krylov`      = the linear solve of the same steps (ms, iterations, SpMV time).  `ms_per_step` = assembly + solve.
`e2e`         = the same assembly metric through the host-buffer C-ABI calls (H2D of positions, states, qPrev and D2H of
                the nodal states inside the timed region); `e2e.picard_body_*` = assemble + solve + solution to the host.
`roofline`    = the fine-level smoothing sweep of the multigrid cycle (SpMV + block-Jacobi epilogue on the fp32 copy of A),
                the kernel with the largest share of the step; `roofline_spmv` = the fp64 BiCGSTAB SpMV and
                `roofline_assembly` = the assembly kernel with the ALGORITHMIC bytes of SURVEY.md section 8(d).  All:
                bytes per launch / average launch duration (CUDA events) / measured HBM copy peak.
`cpu_baseline`= the reference's CPU structure (oracle/pfem_oracle.cpp) on a bounded sample, rank 0 only.

--impl reference: the CPU arm alone (OpenMP element loop -> triplets -> serial duplicate-summing CSC compression ->
serial RHS -> serial BC), all host threads, bounded sample of the same workload.  Timed code: the reference's OWN
m_buildAbPSPG + m_applyBCPSPG from oracle/_ref/libpfem_ref.so (its sources compiled in place against stand-in
Eigen/sol2/gmsh headers, oracle/refbuild; `cpu_baseline.kind` = "reference"), with the oracle port's number beside it
(`port_value`); only the port when that library is absent.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from pfem_b200 import meshgen as mg  # noqa: E402

# dram__bytes_read.sum + dram__bytes_write.sum of one fine-level smoothing sweep at C4 (ncu --set full, profiles/r1_ncu_mg.md)
TRAFFIC_SMOOTH = 422.7e6
METRIC = "FE assembly Melem/s (+ Krylov solve ms/step under 'krylov'; % of HBM roofline under 'roofline*')"
UNIT = "Melem/s"
REL_TOL = 1e-10
MAX_ITER = 40000
C4_ELEMS = 1971054


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(n_nodes, n_elems, nnz, dim):
    """index = 4 B, value = 8 B."""
    npe, n_dof = dim + 1, (dim + 1) * n_nodes
    b_asm = n_elems * npe * 4 + n_nodes * (dim * 8 * 3 + 1) + nnz * 8 + n_dof * 8
    b_spmv = nnz * 12 + (n_dof + 1) * 4 + 2 * n_dof * 8
    return b_asm, b_spmv


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = [float(r[1]) for r in rows]
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].strip().lower().startswith("active"):
                    reasons.add(name)
        out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(rows[0][2]), reasons=sorted(reasons), samples=len(rows),
                   power_w_max=max(float(r[3]) for r in rows))
        return out


def cpu_baseline_sample(cells, want_solve_iters=10):
    """Oracle (the reference's CPU structure) on a bounded sample of the workload: Kuhn box n=cells."""
    import scipy.sparse as sp

    from oracle import oracle as orc

    mesh = mg.kuhn_box(3, cells)
    q, q_prev = mg.pspg_state(mesh)
    P = mg.PSPG_PARAMS
    par = orc.pspg_param_array(P["rho"], P["mu"], P["dt"], mg.gravity(3))
    vcur = q[: 3 * mesh.n_nodes].copy()
    ph = np.zeros(6)
    t0 = time.perf_counter()
    A, b = orc.pspg_build(mesh, vcur, q_prev, par, True, phase_sec=ph)
    t_asm = time.perf_counter() - t0
    A_csr = sp.csr_matrix(A)
    t1 = time.perf_counter()
    _, it, _ = orc.bicgstab(A_csr, b, 1e-30, want_solve_iters)
    t_it = (time.perf_counter() - t1) / max(it, 1)
    return dict(t_asm=t_asm, phases={k: float(v) for k, v in zip(orc.PHASES, ph)}, ms_per_iter=1e3 * t_it,
                cores=orc.num_threads(), n_elems=mesh.n_elems)


def reference_build_sample(cells):
    """The REFERENCE'S OWN m_buildAbPSPG + m_applyBCPSPG (MomContEquationPSPG.inl:7-235), compiled in place from the
    reference sources into oracle/_ref/libpfem_ref.so (oracle/refbuild, stand-in Eigen/sol2/gmsh headers), on all host
    threads, same bounded sample as the port.  None when the library did not travel / was not built."""
    from oracle import ref

    if not ref.available():
        return None
    from oracle import oracle as orc

    threads = ref.set_threads(0)
    mesh = mg.kuhn_box(3, cells)
    q, q_prev = mg.pspg_state(mesh)
    P = mg.PSPG_PARAMS
    par = orc.pspg_param_array(P["rho"], P["mu"], P["dt"], mg.gravity(3))
    with ref.RefCase(mesh, "pspg", par) as rc:
        rc.set_states(q)
        t0 = time.perf_counter()
        rc.pspg_build(q_prev, True)
        t = time.perf_counter() - t0
    ref.set_threads(1)
    return dict(t_asm=t, cores=threads, n_elems=mesh.n_elems)


REF_NOTE = ("the reference's own m_buildAbPSPG + m_applyBCPSPG, compiled from its sources (oracle/refbuild) against a "
            "stand-in for Eigen: omp element loop and triplet logic are the reference's, setFromTriplets and the dense "
            "products underneath are the stand-in's, not Eigen's")


def run_reference(args):
    """`--impl reference`: CPU arm.  Rank 0 only under torchrun; the other ranks exit without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import oracle as orc
    orc.build()
    cells = args.ref_cells
    times, iters_ms, last = [], [], None
    ref_times, ref_cores = [], 0
    for s in range(args.warmup + args.steps):
        last = cpu_baseline_sample(cells)
        rb = reference_build_sample(cells)
        if s >= args.warmup:
            times.append(last["t_asm"])
            iters_ms.append(last["ms_per_iter"])
            if rb:
                ref_times.append(rb["t_asm"])
                ref_cores = rb["cores"]
    t_port = float(np.mean(times))
    port_val = last["n_elems"] / t_port / 1e6
    kind = "reference" if ref_times else "port"
    t = float(np.mean(ref_times)) if ref_times else t_port
    val = last["n_elems"] / t / 1e6
    cores = ref_cores if ref_times else last["cores"]
    sample = (f"Kuhn box n={cells}: {last['n_elems']} tets ({100.0 * last['n_elems'] / C4_ELEMS:.1f}% of C4); "
              + (REF_NOTE if ref_times else
                 "omp element loop -> triplets -> serial CSC compression -> serial RHS -> serial BC (oracle/pfem_oracle.cpp; "
                 "oracle/_ref was not built)"))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C4 synthetic 3D Kuhn box, incompressible PSPG assembly + Krylov (bounded CPU sample)",
                   "cells": cells, "n_elems": last["n_elems"]},
        "krylov": {"ms_per_iter": float(np.mean(iters_ms)), "solver": "Jacobi-BiCGSTAB, omp CSR SpMV (the reference uses SparseLU)"},
        "phases_s": last["phases"],
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "port_value": port_val, "port_cores": last["cores"]},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from pfem_b200.capi import PfemContext
    from pfem_b200.partition import partition_mesh

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # the contract is ONE JSON line on stdout: park the real stdout and send everything libraries print (NCCL banner ...)
    # to stderr until the line is ready
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cells = args.cells if world == 1 else int(round(args.cells * world ** (1.0 / 3.0)))
    gmesh = mg.kuhn_box(3, cells)
    gq, gq_prev = mg.pspg_state(gmesh)
    n_elems_global, n_nodes_global = gmesh.n_elems, gmesh.n_nodes
    P = mg.PSPG_PARAMS
    g = mg.gravity(3)
    ctx = PfemContext(3, local_rank)
    t_part = 0.0
    if world > 1:
        uid = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])
        t0 = time.perf_counter()
        part = partition_mesh(gmesh, world, rank)
        t_part = time.perf_counter() - t0
        mesh = part.mesh
        q = part.scatter_nodal(gq, 4, n_nodes_global)
        q_prev = part.scatter_nodal(gq_prev, 4, n_nodes_global)
        del gmesh, gq, gq_prev
    else:
        part, mesh, q, q_prev = None, gmesh, gq, gq_prev
    def pinned(a):
        """Host inputs of the end-to-end leg live in pinned memory (bench contract), still plain numpy views."""
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        v = t.numpy()
        v[...] = a
        keep.append(t)
        return v

    keep = []
    x_host, q, q_prev = pinned(mesh.x), pinned(q), pinned(q_prev)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    t_topo0 = time.perf_counter()
    ctx.set_topology(mesh.conn, mesh.flags)
    if part is not None:
        ctx.set_partition(part)
    t_topo = time.perf_counter() - t_topo0
    ctx.set_positions(x_host)
    ctx.set_dirichlet(mesh.dir_mask, mesh.dir_val)
    ctx.set_states(0, q)
    ctx.pspg_set_qprev(q_prev)
    par = ctx.pspg_params(P["rho"], P["mu"], P["dt"], g)

    def step():
        ctx.pspg_assemble_resident(par)
        return ctx.pspg_solve(REL_TOL, MAX_ITER, fetch=False)

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            sol = step()
        ctx.profile_enable(True)
        ctx.profile_reset()
        launches0 = ctx.launch_count()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            sol = step()
        ev1.record(stream)
        barrier()
        clocks = sampler.stop() if sampler else {}
        step_ms = ev0.elapsed_time(ev1) / args.steps
        launches = (ctx.launch_count() - launches0) // args.steps
        asm_ms, asm_calls = ctx.profile_get("Assemble system")
        prep_ms, _ = ctx.profile_get("Prepare matrix assembly")
        spmv_ms, spmv_calls = ctx.profile_get("SpMV")
        solve_ms, solve_calls = ctx.profile_get("Solve system")
        halo_ms, halo_calls = ctx.profile_get("Halo exchange")
        pre_setup_ms, pre_setup_calls = ctx.profile_get("Preconditioner setup")
        pre_apply_ms, pre_apply_calls = ctx.profile_get("Preconditioner apply")
        precond_used, precond_levels = ctx.pspg_get_preconditioner()
        ctx.profile_enable(False)
        # one more (untimed) step with per-kernel phases: the multigrid cycle runs un-graphed so that its fine-level smoothing
        # sweep -- the kernel with the largest share of the step -- can be timed with events on the launching stream
        ctx.profile_reset()
        ctx.profile_enable(2)
        step()
        smooth_ms, smooth_calls = ctx.profile_get("MG smooth L0")
        ctx.profile_enable(False)

        # ---- end to end through the host-buffer ABI: H2D(x, states, qPrev) + assemble + D2H(nodal states) ---------
        e2e_t = []
        for s in range(2 + args.steps):
            barrier()
            t0 = time.perf_counter()
            ctx.set_positions(x_host)
            ctx.set_states(0, q)
            ctx.pspg_assemble(par, q_prev)
            _ = ctx.get_states(0, 4)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            if s >= 2:
                e2e_t.append(t1 - t0)
        e2e_s = float(np.mean(e2e_t))
        barrier()
        t0 = time.perf_counter()
        ctx.set_positions(x_host)
        ctx.set_states(0, q)
        ctx.pspg_assemble(par, q_prev)
        sol_e2e = ctx.pspg_solve(REL_TOL, MAX_ITER, fetch=True)
        e2e_picard_s = time.perf_counter() - t0

    if world == 1:
        nnz = ctx.pspg_reference_nnz()
    else:  # reference pattern of the global matrix: 16 (nNodes + 2 nEdges) minus masked rows; estimate from the block count
        nb = torch.tensor([float(ctx.info().nnzBlocks)], device="cuda", dtype=torch.float64)
        nnz = None
    info = ctx.info()
    peak, peak_src = measured_peaks()

    # max over ranks of the device times; sums of per-rank sizes
    asm_per = (asm_ms + prep_ms) / max(asm_calls, 1)
    t_max = torch.tensor([step_ms, asm_per, spmv_ms / max(spmv_calls, 1), solve_ms / max(solve_calls, 1), e2e_s, e2e_picard_s,
                          halo_ms / max(halo_calls, 1) if halo_calls else 0.0,
                          smooth_ms / max(smooth_calls, 1), pre_setup_ms / max(pre_setup_calls, 1),
                          pre_apply_ms / max(pre_apply_calls, 1)], device="cuda", dtype=torch.float64)
    sizes = torch.tensor([float(info.nnzBlocks), float(mesh.n_nodes), float(mesh.n_elems)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(sizes, op=dist.ReduceOp.SUM)
    (step_ms, asm_ms_per, spmv_ms_per, solve_ms_per, e2e_s, e2e_picard_s, halo_ms_per, smooth_ms_per, pre_setup_ms_per,
     pre_apply_ms_per) = [float(v) for v in t_max]
    if nnz is None:
        nnz = int(sizes[0]) * 16  # block storage (no masked-row savings): upper bound of the reference nnz
    value = n_elems_global / (asm_ms_per * 1e-3) / 1e6
    b_asm, b_spmv = algorithmic_bytes(n_nodes_global, n_elems_global, nnz, 3)
    spmv_us = 1e3 * spmv_ms_per
    # fine-level smoothing sweep of the multigrid cycle: bytes it has to move on its storage:
    # fp32 4x4 blocks + block column index, per dof the row of Dw (4 doubles), b, x (own) read and y written
    n_blocks_global = int(sizes[0])
    b_smooth = n_blocks_global * (16 * 4 + 4) + 4 * n_nodes_global * (4 * 8 + 3 * 8)
    smooth_us = 1e3 * smooth_ms_per

    line = None
    if rank == 0:
        cpu = cpu_baseline_sample(args.cpu_cells)
        cpu_ref = reference_build_sample(args.cpu_cells) if world == 1 else None
        agg_peak = peak * world
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"synthetic 3D Kuhn box n={cells} ({n_elems_global} tets, {n_elems_global / world / 1e6:.2f} M per GPU), "
                                   f"incompressible PSPG: assembly+BC then multigrid-preconditioned BiCGSTAB (rel tol {REL_TOL:g}) per step"
                                   + ("" if world == 1 else "; RCB node partition + ghost-element layer, NCCL halo/all-reduce in the solve"),
                       "n_elems": n_elems_global, "n_nodes": n_nodes_global, "n_dof": 4 * n_nodes_global, "nnz": nnz,
                       "l2_policy": "inputs larger than L2 (A = %.0f MB per GPU vs 126 MB L2)" % (nnz * 8 / 1e6 / world),
                       "value_definition": "n_elems / max-over-ranks mean device time of (assembly prologue + assembly kernel) in the timed steps"},
            "assembly_ms": asm_ms_per,
            "step_melem_s": n_elems_global / (step_ms * 1e-3) / 1e6,
            "pattern_build_ms": 1e3 * t_topo, "partition_host_s": t_part,
            "krylov": {"solve_ms": solve_ms_per, "iters": sol["iters"], "rel_res": sol["rel_res"], "status": sol["status"],
                       "ms_per_iter": solve_ms_per / max(sol["iters"], 1), "spmv_us": spmv_us,
                       "spmv_launches_per_step": spmv_calls // max(args.steps, 1), "halo_us": 1e3 * halo_ms_per,
                       "preconditioner": precond_used, "mg_levels": precond_levels,
                       "precond_setup_ms": pre_setup_ms_per, "precond_apply_us": 1e3 * pre_apply_ms_per,
                       "precond_applies_per_step": pre_apply_calls // max(args.steps, 1),
                       "mg_smooth_l0_us": smooth_us, "mg_smooth_l0_launches_per_step": smooth_calls},
            "roofline": ({"kernel": "k_spmv<4,float,EPI_SMOOTH> (fine-level multigrid smoothing sweep)", "bound": "hbm",
                          "achieved": b_smooth / (smooth_us * 1e-6) / 1e9, "peak": agg_peak, "unit": "GB/s",
                          "frac": b_smooth / (smooth_us * 1e-6) / 1e9 / agg_peak, "traffic": TRAFFIC_SMOOTH if world == 1 else None,
                          "algorithmic_bytes": b_smooth, "peak_source": peak_src + (" x n_gpus" if world > 1 else ""),
                          "frac_of_8TBs_nominal": b_smooth / (smooth_us * 1e-6) / 1e9 / (8000.0 * world)}
                         if smooth_calls else
                         {"kernel": "k_spmv<4>", "bound": "hbm", "achieved": b_spmv / (spmv_us * 1e-6) / 1e9, "peak": agg_peak,
                          "unit": "GB/s", "frac": b_spmv / (spmv_us * 1e-6) / 1e9 / agg_peak, "traffic": None,
                          "algorithmic_bytes": b_spmv, "peak_source": peak_src + (" x n_gpus" if world > 1 else ""),
                          "frac_of_8TBs_nominal": b_spmv / (spmv_us * 1e-6) / 1e9 / (8000.0 * world)}),
            "roofline_spmv": {"kernel": "k_spmv<4> (fp64 BiCGSTAB SpMV)", "bound": "hbm", "achieved": b_spmv / (spmv_us * 1e-6) / 1e9,
                              "peak": agg_peak, "unit": "GB/s", "frac": b_spmv / (spmv_us * 1e-6) / 1e9 / agg_peak, "traffic": None,
                              "algorithmic_bytes": b_spmv},
            "roofline_assembly": {"kernel": "k_pspg_assemble<3>", "bound": "hbm",
                                  "achieved": b_asm / (asm_ms_per * 1e-3) / 1e9, "peak": agg_peak, "unit": "GB/s",
                                  "frac": b_asm / (asm_ms_per * 1e-3) / 1e9 / agg_peak, "traffic": None, "algorithmic_bytes": b_asm,
                                  "frac_of_8TBs_nominal": b_asm / (asm_ms_per * 1e-3) / 1e9 / (8000.0 * world)},
            "cpu_baseline": ({"value": cpu_ref["n_elems"] / cpu_ref["t_asm"] / 1e6, "unit": UNIT, "cores": cpu_ref["cores"],
                              "kind": "reference",
                              "sample": f"Kuhn box n={args.cpu_cells} ({cpu_ref['n_elems']} tets) assembled once: {REF_NOTE}: "
                                        f"{cpu_ref['t_asm']:.2f} s",
                              "port_value": cpu["n_elems"] / cpu["t_asm"] / 1e6, "port_cores": cpu["cores"],
                              "phases_s_port": cpu["phases"], "bicgstab_ms_per_iter_port": cpu["ms_per_iter"]}
                             if cpu_ref else
                             {"value": cpu["n_elems"] / cpu["t_asm"] / 1e6, "unit": UNIT, "cores": cpu["cores"], "kind": "port",
                              "sample": f"Kuhn box n={args.cpu_cells} ({cpu['n_elems']} tets) assembled once by oracle/pfem_oracle.cpp "
                                        f"(omp element loop + serial CSC compression + serial BC): {cpu['t_asm']:.2f} s; "
                                        f"BiCGSTAB {cpu['ms_per_iter']:.1f} ms/iter",
                              "phases_s": cpu["phases"], "bicgstab_ms_per_iter": cpu["ms_per_iter"]}),
            "clocks": clocks,
            "e2e": {"value": n_elems_global / e2e_s / 1e6, "unit": UNIT,
                    "h2d_bytes_per_step": int(8 * 10 * sizes[1]), "d2h_bytes_per_step": int(8 * 4 * sizes[1]),
                    "picard_body_ms": 1e3 * e2e_picard_s, "picard_body_melem_s": n_elems_global / e2e_picard_s / 1e6,
                    "picard_body_iters": sol_e2e["iters"]},
            "gpu_launches": int(launches),
        }
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if line is not None:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpu", choices=["gpu", "reference"])
    ap.add_argument("--cells", type=int, default=69, help="Kuhn box cells per side at N=1 (69 -> C4); scaled by N^(1/3)")
    ap.add_argument("--cpu-cells", type=int, default=30, help="bounded CPU-baseline sample inside the GPU run")
    ap.add_argument("--ref-cells", type=int, default=34, help="bounded sample of the --impl reference arm")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
