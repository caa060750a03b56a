#!/usr/bin/env python
"""bench.py -- PFEM3D finite-element hot path on B200: FE assembly Melem/s + Krylov solve ms/step, % of HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--cells n]

Workload at N=1 (BASELINE.json configs[3], "C4"): synthetic Kuhn box n=69 -> 1 971 054 tets, 343 000 nodes,
1 372 000 dof; one STEP = one body of the PSPG Picard loop = m_buildAbPSPG + m_applyBCPSPG (assembly) followed by
the linear solve (Jacobi-BiCGSTAB to ||r||/||b|| <= 1e-10), inputs resident in HBM.
`value` = assembly throughput (elements assembled per second, CUDA events around the assembly inside the timed
steps); the Krylov solve of the same steps is reported under "krylov"; `ms_per_step` is the whole step.
`e2e` = the same assembly metric through the host-buffer C-ABI calls (H2D of positions, states, qPrev and D2H of the
assembled RHS inside the timed region).  `roofline` = SpMV (dominant kernel of the step), `roofline_assembly` = the
assembly kernel; both use the ALGORITHMIC bytes of SURVEY.md section 8(d) over the measured HBM copy peak.

--impl reference: the reference's own CPU structure (OpenMP element loop -> triplets -> serial duplicate-summing CSC
compression -> serial RHS -> serial BC; oracle/pfem_oracle.cpp, the reference itself cannot be compiled here) on the
box's host cores, on a bounded sample of the same workload.
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

METRIC = "FE assembly Melem/s (+ Krylov solve ms/step under 'krylov'; % of HBM roofline under 'roofline*')"
UNIT = "Melem/s"
REL_TOL = 1e-10
MAX_ITER = 20000


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
    from oracle import oracle as orc
    import scipy.sparse as sp

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
    return dict(mesh=mesh, t_asm=t_asm, phases={k: float(v) for k, v in zip(orc.PHASES, ph)}, ms_per_iter=1e3 * t_it,
                cores=orc.num_threads(), n_elems=mesh.n_elems)


def run_reference(args):
    """`--impl reference`: CPU arm.  Rank 0 only under torchrun."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    cells = args.ref_cells
    times, iters_ms, last = [], [], None
    for s in range(args.warmup + args.steps):
        last = cpu_baseline_sample(cells)
        if s >= args.warmup:
            times.append(last["t_asm"])
            iters_ms.append(last["ms_per_iter"])
    t = float(np.mean(times))
    val = last["n_elems"] / t / 1e6
    sample = (f"Kuhn box n={cells}: {last['n_elems']} tets ({100.0 * last['n_elems'] / 1971054:.1f}% of C4); omp element loop -> "
              "triplets -> serial CSC compression -> serial RHS -> serial BC (oracle/pfem_oracle.cpp; Eigen reference not buildable here)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C4 synthetic 3D Kuhn box PSPG assembly + Krylov (bounded CPU sample)", "cells": cells,
                   "n_elems": last["n_elems"]},
        "krylov": {"ms_per_iter": float(np.mean(iters_ms)), "solver": "Jacobi-BiCGSTAB, omp CSR SpMV (the reference uses SparseLU)"},
        "phases_s": last["phases"],
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from pfem_b200.capi import PfemContext

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
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

    cells = args.cells
    mesh = mg.kuhn_box(3, cells)
    q, q_prev = mg.pspg_state(mesh)
    P = mg.PSPG_PARAMS
    g = mg.gravity(3)
    ctx = PfemContext(3, local_rank)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    t_topo0 = time.perf_counter()
    ctx.set_topology(mesh.conn, mesh.flags)
    t_topo = time.perf_counter() - t_topo0
    ctx.set_positions(mesh.x)
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
        ctx.profile_enable(False)

        # ---- end-to-end through the host-buffer ABI: H2D(x, states, qPrev) + assemble + D2H(b) -----------------
        e2e_t = []
        for s in range(2 + args.steps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.set_positions(mesh.x)
            ctx.set_states(0, q)
            ctx.pspg_assemble(par, q_prev)
            _ = ctx.get_states(0, 4)  # device->host read of nodal results of the step
            t1 = time.perf_counter()
            if s >= 2:
                e2e_t.append(t1 - t0)
        e2e_s = float(np.mean(e2e_t))
        # e2e Picard body: assemble (host qPrev) + solve + q back
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.set_positions(mesh.x)
        ctx.set_states(0, q)
        ctx.pspg_assemble(par, q_prev)
        sol_e2e = ctx.pspg_solve(REL_TOL, MAX_ITER, fetch=True)
        e2e_picard_s = time.perf_counter() - t0

    info = ctx.info()
    n_dof = info.nDof
    nnz = int(info.nnzReference) if info.nnzReference >= 0 else ctx.pspg_reference_nnz()
    peak, peak_src = measured_peaks()
    b_asm, b_spmv = algorithmic_bytes(mesh.n_nodes, mesh.n_elems, nnz, 3)

    # max over ranks of the device time
    t_step = torch.tensor([step_ms, (asm_ms + prep_ms) / max(asm_calls, 1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_step, op=dist.ReduceOp.MAX)
    step_ms, asm_ms_per = float(t_step[0]), float(t_step[1])
    spmv_us = 1e3 * spmv_ms / max(spmv_calls, 1)
    value = world * mesh.n_elems / (asm_ms_per * 1e-3) / 1e6

    line = None
    if rank == 0:
        cpu = cpu_baseline_sample(args.cpu_cells)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"C4 synthetic 3D Kuhn box n={cells}, incompressible PSPG: assembly+BC then Jacobi-BiCGSTAB "
                                   f"(rel tol {REL_TOL:g}) per step; per-GPU replica at N>1",
                       "n_elems": mesh.n_elems, "n_nodes": mesh.n_nodes, "n_dof": int(n_dof), "nnz": nnz,
                       "l2_policy": "inputs larger than L2 (A = %.0f MB vs 126 MB L2)" % (nnz * 8 / 1e6),
                       "value_definition": "n_elems / mean device time of (assembly prologue + assembly kernel) inside the timed steps"},
            "assembly_ms": asm_ms_per,
            "step_melem_s": world * mesh.n_elems / (step_ms * 1e-3) / 1e6,
            "pattern_build_ms": 1e3 * t_topo,
            "krylov": {"solve_ms": solve_ms / max(solve_calls, 1), "iters": sol["iters"], "rel_res": sol["rel_res"],
                       "status": sol["status"], "ms_per_iter": solve_ms / max(solve_calls, 1) / max(sol["iters"], 1),
                       "spmv_us": spmv_us, "spmv_launches_per_step": spmv_calls // max(args.steps, 1)},
            "roofline": {"kernel": "k_spmv<4>", "bound": "hbm", "achieved": b_spmv / (spmv_us * 1e-6) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": b_spmv / (spmv_us * 1e-6) / 1e9 / peak, "traffic": None,
                         "algorithmic_bytes": b_spmv, "peak_source": peak_src,
                         "frac_of_8TBs_nominal": b_spmv / (spmv_us * 1e-6) / 1e9 / 8000.0},
            "roofline_assembly": {"kernel": "k_pspg_assemble<3>", "bound": "hbm",
                                  "achieved": b_asm / (asm_ms_per * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": b_asm / (asm_ms_per * 1e-3) / 1e9 / peak, "traffic": None, "algorithmic_bytes": b_asm,
                                  "frac_of_8TBs_nominal": b_asm / (asm_ms_per * 1e-3) / 1e9 / 8000.0},
            "cpu_baseline": {"value": cpu["n_elems"] / cpu["t_asm"] / 1e6, "unit": UNIT, "cores": cpu["cores"], "kind": "port",
                             "sample": f"Kuhn box n={args.cpu_cells} ({cpu['n_elems']} tets) assembled once by oracle/pfem_oracle.cpp "
                                       f"(omp element loop + serial CSC compression + serial BC): {cpu['t_asm']:.2f} s; "
                                       f"BiCGSTAB {cpu['ms_per_iter']:.1f} ms/iter",
                             "phases_s": cpu["phases"], "bicgstab_ms_per_iter": cpu["ms_per_iter"]},
            "clocks": clocks,
            "e2e": {"value": world * mesh.n_elems / e2e_s / 1e6, "unit": UNIT,
                    "h2d_bytes_per_step": int(8 * (3 * mesh.n_nodes + 4 * mesh.n_nodes + 3 * mesh.n_nodes)),
                    "d2h_bytes_per_step": int(8 * 4 * mesh.n_nodes),
                    "picard_body_ms": 1e3 * e2e_picard_s, "picard_body_melem_s": mesh.n_elems / e2e_picard_s / 1e6,
                    "picard_body_iters": sol_e2e["iters"]},
            "gpu_launches": int(launches),
        }
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpu", choices=["gpu", "reference"])
    ap.add_argument("--cells", type=int, default=69, help="Kuhn box cells per side (69 -> C4)")
    ap.add_argument("--cpu-cells", type=int, default=30, help="bounded CPU-baseline sample inside the GPU run")
    ap.add_argument("--ref-cells", type=int, default=34, help="bounded sample of the --impl reference arm")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
