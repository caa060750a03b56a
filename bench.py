#!/usr/bin/env python
"""bench.py -- PFEM3D finite-element hot path on B200: FE assembly Melem/s + Krylov solve ms/step, % of HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--cells n] [--wc-cells n]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

N = 1 -- BASELINE.json configs[3] ("C4"): synthetic Kuhn box n=69 -> 1 971 054 tets, 343 000 nodes, 1 372 000 dof.
  One STEP = one body of the PSPG Picard loop: m_buildAbPSPG + m_applyBCPSPG (assembly) followed by the linear solve
  (multigrid-preconditioned flexible GMRES to ||r||/||b|| <= 1e-12, the tolerance the parity tests solve at), inputs resident.
  `value`    = assembly Melem/s (device events); `krylov` = the solve of the same steps; `ms_per_step` = both.
  `e2e`      = the WHOLE step through the host-buffer C ABI, per step: pfem_set_topology (pattern build of the remeshed
               connectivity: the incompressible solver remeshes every step, IncompNewton/Solver.cpp:242) + positions,
               Dirichlet data, states, qPrev H2D + assembly + solve + solution D2H.  `e2e.assembly_only` = the round-1
               definition (H2D + assembly + D2H) for continuity.
  `roofline` = the kernel with the largest share of the step (fine-level multigrid smoothing sweep); `roofline_spmv`,
               `roofline_assembly` (+ fp64 GFLOP/s) and `roofline_wc` (explicit weakly-compressible step at C5, one GPU)
               beside it.  `traffic` is read from profiles/r2_traffic.json (written from ncu dram__bytes captures).
N > 1 -- BASELINE.json configs[4] ("C5"), the north star's multi-GPU case: synthetic Kuhn box n=150 -> 20 250 000 tets,
  explicit weakly-compressible step, RCB-sharded, STRONG scaling.  One STEP = kick/move + continuity + momentum with
  their two halo exchanges + the CFL time step with its min-all-reduce (pfem_wc_run: dt chained on the device, no host
  round trip).  `value` = whole-mesh Melem/s of that step (max over ranks).  In the same job rank 0 also runs the
  identical chain on ONE GPU: `wc_1gpu` (its ms/step -> `strong_efficiency_vs_1gpu`) and `parity_vs_1gpu.wc_*` (the
  sharded fields must be bit-identical).  The PSPG numbers ride along under `pspg_weak` (n = 69 N^(1/3): ~2 M tets per
  GPU, assembly + solve) and `parity_vs_1gpu.pspg_*` (C4 sharded over the N ranks vs the rank-0 single-GPU solve).
The run FAILS (non-zero exit, "failed" in the line) when a solve does not converge, when a sharded solve needs more
than 5x the single-GPU iterations, or when a parity bar is missed.

--impl reference: the CPU arm alone, all host threads, bounded sample of the same workload: at N=1 the reference's own
m_buildAbPSPG + m_applyBCPSPG, at N>1 its own explicit step (m_solveWCompNewtonNoT + computeNextDT), both from
oracle/_ref/libpfem_ref.so (the reference's sources compiled in place against stand-in Eigen/sol2/gmsh headers,
oracle/refbuild; `cpu_baseline.kind` = "reference"); the oracle port when that library is absent.
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
METRIC_WC = "explicit weakly-compressible FE step Melem/s (element loop + halo exchanges + CFL dt; % of HBM roofline under 'roofline')"
UNIT = "Melem/s"
REL_TOL = 1e-12   # the tolerance the parity tests demonstrate 1e-8 fields at (tests/test_gpu_pspg.py, test_gpu_mg.py)
MAX_ITER = 40000
C4_ELEMS = 1971054
# fp64 operations of one element's share of the PSPG system with the closed forms of SURVEY.md appendix A, computed once:
# geometry + tau + RHS ~ 130, 16 node-pair blocks x 24 fused multiply-adds (48 flop)
FLOPS_PER_ELEM_ASM = 130 + 16 * 48


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch (or per step) from the committed ncu captures."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(p):
        return None
    v = json.load(open(p)).get(key)
    return None if v is None else float(v["bytes"])


def algorithmic_bytes(n_nodes, n_elems, nnz, dim):
    """index = 4 B, value = 8 B.  SpMV: the survey's figure is for scalar CSR (12 B per non-zero);
    the node-block CSR this library stores needs one 4-byte column index per (dim+1)^2 block, and THAT is the algorithmic
    traffic of the kernel measured here (the scalar-CSR figure would put the kernel above the HBM peak)."""
    npe, n_dof = dim + 1, (dim + 1) * n_nodes
    b_asm = n_elems * npe * 4 + n_nodes * (dim * 8 * 3 + 1) + nnz * 8 + n_dof * 8
    b_spmv = nnz * 8 + (nnz // (npe * npe)) * 4 + (n_nodes + 1) * 4 + 2 * n_dof * 8
    return b_asm, b_spmv


def wc_step_bytes(n_nodes, n_elems, dim=3):
    """B_wc_step = 2 nElm npe 4 + nNodes 8 (R + W + accumulators)."""
    return 2 * n_elems * (dim + 1) * 4 + n_nodes * 8 * (11 + 11 + 12)


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


# ---------------------------------------------------------------------------------------------------------------------
# CPU arms (the only places bench.py executes anything under oracle/)
# ---------------------------------------------------------------------------------------------------------------------
def host_threads():
    """All host threads, also under torchrun (which exports OMP_NUM_THREADS=1 to its workers)."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    from oracle import oracle as orc
    orc.lib().oracle_set_num_threads(int(n))
    return n


def cpu_baseline_sample(cells, want_solve_iters=10):
    """Oracle (the reference's CPU structure) on a bounded sample of the workload: Kuhn box n=cells."""
    import scipy.sparse as sp

    from oracle import oracle as orc

    host_threads()
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


def wc_cpu_sample(cells, steps=3):
    """The explicit weakly-compressible step on the host: the reference's own m_solveWCompNewtonNoT + computeNextDT
    (oracle/_ref) when that library is there, else the oracle port; seconds per step on a Kuhn box n=cells."""
    from oracle import oracle as orc
    from oracle import ref

    threads = host_threads()
    mesh = mg.kuhn_box(3, cells)
    st = mg.wc_state(mesh)
    W = mg.WC_PARAMS
    wpar = orc.wc_param_array(W["mu"], W["K0"], W["K0p"], W["rhoStar"], mg.gravity(3), True, "CDS_dpdt")
    if ref.available():
        threads = ref.set_threads(0)
        with ref.RefCase(mesh, "wc", np.concatenate([wpar, [1e-6, 1e-3, W["securityCoeff"]]])) as rc:
            rc.set_states(np.concatenate([st["v"], st["p"], st["rho"], st["acc"]]))
            dt = rc.wc_next_dt()
            rc.wc_step(dt)
            t0 = time.perf_counter()
            for _ in range(steps):
                dt = rc.wc_next_dt()
                rc.wc_step(dt)
            t = (time.perf_counter() - t0) / steps
        ref.set_threads(1)
        return dict(t_step=t, cores=threads, n_elems=mesh.n_elems, kind="reference")
    x = mesh.x
    dt = orc.wc_next_dt(mesh, x, st, wpar, W["securityCoeff"], 1e-3)
    x, st = orc.wc_step(mesh, x, st, wpar, dt)
    t0 = time.perf_counter()
    for _ in range(steps):
        dt = orc.wc_next_dt(mesh, x, st, wpar, W["securityCoeff"], 1e-3)
        x, st = orc.wc_step(mesh, x, st, wpar, dt)
    t = (time.perf_counter() - t0) / steps
    return dict(t_step=t, cores=orc.num_threads(), n_elems=mesh.n_elems, kind="port")


REF_NOTE = ("the reference's own m_buildAbPSPG + m_applyBCPSPG, compiled from its sources (oracle/refbuild) against a "
            "stand-in for Eigen: omp element loop and triplet logic are the reference's, setFromTriplets and the dense "
            "products underneath are the stand-in's, not Eigen's")
REF_NOTE_WC = ("the reference's own m_solveWCompNewtonNoT + computeNextDT (omp element loops, serial nodal scatter), "
               "compiled from its sources (oracle/refbuild) against a stand-in for Eigen")


def run_reference(args):
    """`--impl reference`: CPU arm.  Rank 0 only under torchrun; the other ranks exit without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import oracle as orc
    orc.build()
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if world > 1 or args.gpus > 1:
        # the multi-GPU arm's workload: explicit weakly-compressible step (C5), bounded sample
        cells = args.ref_wc_cells
        ts, last = [], None
        for s in range(args.warmup + args.steps):
            last = wc_cpu_sample(cells, steps=2)
            if s >= args.warmup:
                ts.append(last["t_step"])
        t = float(np.mean(ts))
        val = last["n_elems"] / t / 1e6
        sample = (f"Kuhn box n={cells}: {last['n_elems']} tets ({100.0 * last['n_elems'] / 20250000:.2f}% of C5), step + CFL dt; "
                  + (REF_NOTE_WC if last["kind"] == "reference" else "oracle/pfem_oracle.cpp port (oracle/_ref was not built)"))
        line = {
            "impl": "reference", "metric": METRIC_WC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C5 synthetic 3D Kuhn box, explicit weakly-compressible step (bounded CPU sample)",
                       "cells": cells, "n_elems": last["n_elems"]},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line), flush=True)
        return
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


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def log(msg):
    """Progress to stderr (stdout carries the one JSON line): a hung leg is then visible in the captured log."""
    print(f"[bench {time.strftime('%H:%M:%S')} rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


class Env:
    """torch / distributed plumbing of one rank."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.keep = []

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def pinned(self, a, dtype=None):
        """Host inputs of the end-to-end leg live in pinned memory (bench contract), still plain numpy views."""
        a = np.ascontiguousarray(a, dtype=dtype)
        t = self.torch.empty(max(a.nbytes, 1), dtype=self.torch.uint8, pin_memory=True)
        v = t.numpy()[: a.nbytes].view(a.dtype).reshape(a.shape)
        v[...] = a
        self.keep.append(t)
        return v

    def max_over_ranks(self, values):
        t = self.torch.tensor([float(v) for v in values], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def sum_over_ranks(self, values):
        t = self.torch.tensor([float(v) for v in values], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t]

    def comm_init(self, ctx):
        uid = [ctx.comm_unique_id() if self.rank == 0 else None]
        self.dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(self.world, self.rank, uid[0])

    def gather_owned_to_all(self, local, part, n_comp, n_nodes_global):
        """Owned entries of a local SoA vector -> the global SoA vector on every rank (sum of disjoint contributions)."""
        torch = self.torch
        nl = part.l2g_nodes.size
        g = torch.zeros((n_comp, n_nodes_global), dtype=torch.float64, device="cuda")
        idx = torch.from_numpy(part.l2g_nodes[: part.n_owned]).cuda()
        g[:, idx] = torch.from_numpy(np.ascontiguousarray(local.reshape(n_comp, nl)[:, : part.n_owned])).cuda()
        self.dist.all_reduce(g)
        return g.cpu().numpy().reshape(-1)


def make_part(env, mesh):
    from pfem_b200.partition import partition_mesh_native
    t0 = time.perf_counter()
    part = partition_mesh_native(mesh, env.world, env.rank)
    return part, time.perf_counter() - t0


def cpu_baseline_block(args, world):
    """`cpu_baseline` of the PSPG assembly (rank 0): the reference's own code when oracle/_ref travelled, else the port."""
    cpu = cpu_baseline_sample(args.cpu_cells)
    cpu_ref = reference_build_sample(args.cpu_cells)
    if cpu_ref:
        return {"value": cpu_ref["n_elems"] / cpu_ref["t_asm"] / 1e6, "unit": UNIT, "cores": cpu_ref["cores"], "kind": "reference",
                "sample": f"Kuhn box n={args.cpu_cells} ({cpu_ref['n_elems']} tets) assembled once: {REF_NOTE}: {cpu_ref['t_asm']:.2f} s",
                "port_value": cpu["n_elems"] / cpu["t_asm"] / 1e6, "port_cores": cpu["cores"], "phases_s_port": cpu["phases"],
                "bicgstab_ms_per_iter_port": cpu["ms_per_iter"]}
    return {"value": cpu["n_elems"] / cpu["t_asm"] / 1e6, "unit": UNIT, "cores": cpu["cores"], "kind": "port",
            "sample": f"Kuhn box n={args.cpu_cells} ({cpu['n_elems']} tets) assembled once by oracle/pfem_oracle.cpp (omp element loop + "
                      f"serial CSC compression + serial BC): {cpu['t_asm']:.2f} s; BiCGSTAB {cpu['ms_per_iter']:.1f} ms/iter",
            "phases_s": cpu["phases"], "bicgstab_ms_per_iter": cpu["ms_per_iter"]}


def pspg_leg(env, args, cells, steps, warmup, with_e2e, failures, tag):
    """PSPG assemble + solve on a Kuhn box n=cells, sharded over the ranks of env.  Returns a dict of numbers (max over
    ranks where that applies) and, for parity checks, the gathered solution."""
    from pfem_b200.capi import PfemContext
    torch = env.torch
    gmesh = mg.kuhn_box(3, cells)
    gq, gq_prev = mg.pspg_state(gmesh)
    n_elems_g, n_nodes_g = gmesh.n_elems, gmesh.n_nodes
    P = mg.PSPG_PARAMS
    g = mg.gravity(3)
    ctx = PfemContext(3, env.local_rank)
    t_part = 0.0
    if env.world > 1:
        env.comm_init(ctx)
        part, t_part = make_part(env, gmesh)
        mesh = part.mesh
        q = part.scatter_nodal(gq, 4, n_nodes_g)
        q_prev = part.scatter_nodal(gq_prev, 4, n_nodes_g)
    else:
        part, mesh, q, q_prev = None, gmesh, gq, gq_prev
    conn_h = env.pinned(mesh.conn, np.uint64)
    flags_h = env.pinned(mesh.flags, np.uint8)
    x_h, q_h, qp_h = env.pinned(mesh.x), env.pinned(q), env.pinned(q_prev)
    dmask_h, dval_h = env.pinned(mesh.dir_mask, np.uint8), env.pinned(mesh.dir_val)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)

    def upload_mesh():
        ctx.set_topology(conn_h, flags_h)
        if part is not None:
            ctx.set_partition(part)
        ctx.set_positions(x_h)
        ctx.set_dirichlet(dmask_h, dval_h)

    t0 = time.perf_counter()
    upload_mesh()
    torch.cuda.synchronize()
    t_topo = time.perf_counter() - t0
    ctx.set_states(0, q_h)
    ctx.pspg_set_qprev(qp_h)
    par = ctx.pspg_params(P["rho"], P["mu"], P["dt"], g)

    def step():
        ctx.pspg_assemble_resident(par)
        return ctx.pspg_solve(REL_TOL, MAX_ITER, fetch=False)

    out = {}
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            sol = step()
        ctx.profile_enable(True)
        ctx.profile_reset()
        launches0 = ctx.launch_count()
        sampler = ClockSampler(env.local_rank) if env.rank == 0 else None
        env.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(steps):
            sol = step()
        ev1.record(stream)
        env.barrier()
        out["clocks"] = sampler.stop() if sampler else {}
        step_ms = ev0.elapsed_time(ev1) / steps
        out["launches"] = (ctx.launch_count() - launches0) // steps
        ph = {}
        for name in ("Assemble system", "Prepare matrix assembly", "SpMV", "Solve system", "Halo exchange", "Preconditioner setup",
                     "Preconditioner apply", "All-gather"):
            ms, n = ctx.profile_get(name)
            ph[name] = (ms, n)
        precond_used, precond_levels = ctx.pspg_get_preconditioner()
        ctx.profile_enable(False)
        # one more (untimed) step with per-kernel phases: the multigrid cycle runs un-graphed so that its fine-level smoothing
        # sweep -- the kernel with the largest share of the step -- can be timed with events on the launching stream
        ctx.profile_reset()
        ctx.profile_enable(2)
        step()
        smooth_ms, smooth_calls = ctx.profile_get("MG smooth L0")
        ctx.profile_enable(False)
        sol_full = ctx.pspg_solve(REL_TOL, MAX_ITER, fetch=True)

        e2e = None
        if with_e2e:
            # ---- end to end through the host-buffer ABI, the whole step: remeshed connectivity -> pattern, fields H2D,
            #      assembly, solve at the parity tolerance, solution D2H ------------------------------------------------
            e2e_t, e2e_asm_t, topo_t = [], [], []
            for s in range(2 + steps):
                env.barrier()
                t0 = time.perf_counter()
                upload_mesh()
                t1 = time.perf_counter()
                ctx.set_states(0, q_h)
                ctx.pspg_assemble(par, qp_h)
                sol_e = ctx.pspg_solve(REL_TOL, MAX_ITER, fetch=True)
                torch.cuda.synchronize()
                t2 = time.perf_counter()
                if s >= 2:
                    e2e_t.append(t2 - t0)
                    topo_t.append(t1 - t0)
            for s in range(2 + steps):  # round-1 definition: H2D(x, states, qPrev) + assembly + D2H(nodal states)
                env.barrier()
                t0 = time.perf_counter()
                ctx.set_positions(x_h)
                ctx.set_states(0, q_h)
                ctx.pspg_assemble(par, qp_h)
                _ = ctx.get_states(0, 4)
                torch.cuda.synchronize()
                if s >= 2:
                    e2e_asm_t.append(time.perf_counter() - t0)
            nn, ne = mesh.n_nodes, mesh.n_elems
            e2e = dict(step_s=float(np.mean(e2e_t)), topo_s=float(np.mean(topo_t)), asm_only_s=float(np.mean(e2e_asm_t)),
                       iters=sol_e["iters"], status=sol_e["status"],
                       h2d=int(ne * 4 * 8 + nn + 3 * 8 * nn + nn + 3 * 8 * nn + 4 * 8 * nn + 4 * 8 * nn), d2h=int(4 * 8 * nn),
                       h2d_asm_only=int(8 * 10 * nn), d2h_asm_only=int(8 * 4 * nn))
    nnz = ctx.pspg_reference_nnz() if env.world == 1 else None
    info = ctx.info()
    asm_ms = (ph["Assemble system"][0] + ph["Prepare matrix assembly"][0]) / max(ph["Assemble system"][1], 1)
    per = lambda k: ph[k][0] / max(ph[k][1], 1)  # noqa: E731
    vals = [step_ms, asm_ms, per("SpMV"), per("Solve system"), per("Halo exchange"), smooth_ms / max(smooth_calls, 1),
            per("Preconditioner setup"), per("Preconditioner apply"), t_topo, t_part]
    if e2e:
        vals += [e2e["step_s"], e2e["topo_s"], e2e["asm_only_s"]]
    mx = env.max_over_ranks(vals)
    sizes = env.sum_over_ranks([info.nnzBlocks, mesh.n_nodes, mesh.n_elems])
    out.update(cells=cells, n_elems=n_elems_g, n_nodes=n_nodes_g, step_ms=mx[0], asm_ms=mx[1], spmv_ms=mx[2], solve_ms=mx[3],
               halo_ms=mx[4], smooth_ms=mx[5], pre_setup_ms=mx[6], pre_apply_ms=mx[7], topo_s=mx[8], part_s=mx[9],
               iters=sol["iters"], rel_res=sol["rel_res"], status=sol["status"], precond=precond_used, levels=precond_levels,
               spmv_calls=ph["SpMV"][1] // max(steps, 1), pre_apply_calls=ph["Preconditioner apply"][1] // max(steps, 1),
               smooth_calls=smooth_calls, n_blocks=int(sizes[0]), nnz=nnz, local_nodes_sum=int(sizes[1]))
    if e2e:
        out["e2e"] = dict(e2e, step_s=mx[10], topo_s=mx[11], asm_only_s=mx[12])
    if sol["status"] != 0 or sol_full["status"] != 0:
        failures.append(f"{tag}: solve status {sol['status']}/{sol_full['status']} after {sol['iters']} iterations")
    # the solution on every rank's owned nodes -> global vector (for the parity legs)
    if env.world > 1:
        out["q_global"] = env.gather_owned_to_all(sol_full["q"], part, 4, n_nodes_g)
    else:
        out["q_global"] = sol_full["q"]
    ctx.close()
    return out


def wc_leg(env, args, cells, steps, warmup, profile_phases):
    """Explicit weakly-compressible chain on a Kuhn box n=cells, sharded over env's ranks: dt0 = CFL, then pfem_wc_run
    (warmup) untimed and pfem_wc_run (steps) timed.  Returns timings (max over ranks) and the final local states."""
    from pfem_b200.capi import PfemContext
    torch = env.torch
    gmesh = mg.kuhn_box(3, cells)
    n_elems_g, n_nodes_g = gmesh.n_elems, gmesh.n_nodes
    st = mg.wc_state(gmesh)
    packed = np.concatenate([st["v"], st["p"], st["rho"], st["acc"]])
    del st
    W = mg.WC_PARAMS
    ctx = PfemContext(3, env.local_rank)
    t_part = 0.0
    if env.world > 1:
        env.comm_init(ctx)
        part, t_part = make_part(env, gmesh)
        mesh = part.mesh
        packed_l = part.scatter_nodal(packed, 8, n_nodes_g)
    else:
        part, mesh, packed_l = None, gmesh, packed
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    ctx.set_topology(mesh.conn, mesh.flags)
    if part is not None:
        ctx.set_partition(part)
    t_topo = time.perf_counter() - t0
    ctx.set_positions(mesh.x)
    ctx.set_dirichlet(mesh.dir_mask, mesh.dir_val)
    ctx.set_states(0, packed_l)
    wp = ctx.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], mg.gravity(3), True)
    phases = {}
    with torch.cuda.stream(stream):
        dt = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
        dt, _ = ctx.wc_run(wp, warmup, W["securityCoeff"], 1e-3, dt)
        launches0 = ctx.launch_count()
        sampler = ClockSampler(env.local_rank) if env.rank == 0 else None
        env.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        dt, elapsed = ctx.wc_run(wp, steps, W["securityCoeff"], 1e-3, dt)
        ev1.record(stream)
        env.barrier()
        clocks = sampler.stop() if sampler else {}
        ms = ev0.elapsed_time(ev1) / steps
        launches = (ctx.launch_count() - launches0) // steps
        states = ctx.get_states(0, 8)
        xs = ctx.get_positions()
        if profile_phases:  # kernel phases of the same step, un-chained (device events per phase; extra steps, untimed)
            ctx.profile_enable(True)
            d2 = dt
            for _ in range(2):  # the serialised (un-overlapped) path has its own first-use costs: keep them out of the phases
                ctx.wc_step(wp, d2)
                d2 = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
            ctx.profile_reset()
            for _ in range(3):
                ctx.wc_step(wp, d2)
                d2 = ctx.wc_next_dt(wp, W["securityCoeff"], 1e-3)
            for name in ("Update solutions", "Solving continuity eq", "Solving momentum eq", "CFL nodal pass", "Halo exchange"):
                t, n = ctx.profile_get(name)
                phases[name] = t / max(n, 1)
            ctx.profile_enable(False)
    names = sorted(phases)
    mx = env.max_over_ranks([ms, t_topo, t_part] + [phases[k] for k in names])
    out = dict(cells=cells, n_elems=n_elems_g, n_nodes=n_nodes_g, ms=mx[0], topo_s=mx[1], part_s=mx[2], launches=int(launches),
               dt=dt, clocks=clocks, phases={k: v for k, v in zip(names, mx[3:])}, part=part, states=states, xs=xs,
               packed=packed, gmesh=gmesh)
    ctx.close()
    return out


def run_single(env, args):
    """N = 1: C4 PSPG step (headline) + C5 explicit step (roofline_wc) on one GPU."""
    failures = []
    peak, peak_src = measured_peaks()
    log("PSPG C4 step")
    r = pspg_leg(env, args, args.cells, args.steps, args.warmup, True, failures, "pspg C4")
    log(f"  -> assembly {r['asm_ms']:.3f} ms, solve {r['solve_ms']:.1f} ms ({r['iters']} iterations); C5 explicit step")
    n_elems, n_nodes, nnz = r["n_elems"], r["n_nodes"], r["nnz"]
    b_asm, b_spmv = algorithmic_bytes(n_nodes, n_elems, nnz, 3)
    b_smooth = r["n_blocks"] * (16 * 4 + 4) + 4 * n_nodes * (4 * 4 + 3 * 4)   # fp32 A blocks + index; per dof: fp32 Dw row, x, b, y
    smooth_us, spmv_us = 1e3 * r["smooth_ms"], 1e3 * r["spmv_ms"]
    value = n_elems / (r["asm_ms"] * 1e-3) / 1e6

    def roof(kernel, nbytes, seconds, traffic_key, extra=None):
        ach = nbytes / seconds / 1e9
        d = {"kernel": kernel, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
             "traffic": profiled_traffic(traffic_key), "algorithmic_bytes": nbytes, "peak_source": peak_src,
             "frac_of_8TBs_nominal": ach / 8000.0}
        if extra:
            d.update(extra)
        return d

    # ---- explicit weakly-compressible step at C5 on this GPU -------------------------------------------------------------
    w = wc_leg(env, args, args.wc_cells, max(args.steps, 5), 3, True)
    b_wc = wc_step_bytes(w["n_nodes"], w["n_elems"])
    wc_kern_ms = sum(w["phases"].get(k, 0.0) for k in ("Update solutions", "Solving continuity eq", "Solving momentum eq", "CFL nodal pass"))
    wc_cpu = wc_cpu_sample(args.ref_wc_cells, steps=2)
    e = r["e2e"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["step_ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"C4 synthetic 3D Kuhn box n={r['cells']} ({n_elems} tets), incompressible PSPG: assembly+BC then "
                               f"multigrid-preconditioned FGMRES(40) (rel tol {REL_TOL:g}) per step",
                   "n_elems": n_elems, "n_nodes": n_nodes, "n_dof": 4 * n_nodes, "nnz": nnz,
                   "l2_policy": "inputs larger than L2 (A = %.0f MB vs 126 MB L2)" % (nnz * 8 / 1e6),
                   "value_definition": "n_elems / mean device time of (assembly prologue + assembly kernel) in the timed steps"},
        "assembly_ms": r["asm_ms"], "step_melem_s": n_elems / (r["step_ms"] * 1e-3) / 1e6,
        "pattern_build_ms": 1e3 * r["topo_s"],
        "krylov": {"solve_ms": r["solve_ms"], "iters": r["iters"], "rel_res": r["rel_res"], "status": r["status"], "rel_tol": REL_TOL,
                   "solver": "FGMRES(40), one multigrid V(3,3) cycle + one fp64 SpMV per iteration (r1/early r2: BiCGSTAB, two of each)",
                   "ms_per_iter": r["solve_ms"] / max(r["iters"], 1), "spmv_us": spmv_us, "spmv_launches_per_step": r["spmv_calls"],
                   "preconditioner": r["precond"], "mg_levels": r["levels"], "precond_setup_ms": r["pre_setup_ms"],
                   "precond_apply_us": 1e3 * r["pre_apply_ms"], "precond_applies_per_step": r["pre_apply_calls"],
                   "mg_smooth_l0_us": smooth_us, "mg_smooth_l0_launches_per_step": r["smooth_calls"],
                   "mg_smooth_share_of_step": (smooth_us * 1e-3 * r["smooth_calls"]) / r["step_ms"] if r["step_ms"] else None},
        "roofline": roof("k_spmv<4,float,EPI_SMOOTH,float> (fine-level multigrid smoothing sweep, fp32 matrix copy and vectors)", b_smooth, smooth_us * 1e-6, "mg_smooth_c4"),
        "roofline_spmv": roof("k_spmv<4> (fp64 Krylov SpMV, node-block CSR)", b_spmv, spmv_us * 1e-6, "spmv_c4",
                              {"scalar_csr_bytes_survey_8d": nnz * 12 + (4 * n_nodes + 1) * 4 + 2 * 4 * n_nodes * 8}),
        "roofline_assembly": roof("k_pspg_assemble2<3>", b_asm, r["asm_ms"] * 1e-3, "pspg_assemble_c4",
                                  {"fp64_gflops": FLOPS_PER_ELEM_ASM * n_elems / (r["asm_ms"] * 1e-3) / 1e9,
                                   "fp64_flops_per_element": FLOPS_PER_ELEM_ASM,
                                   "assembly_plus_spmv_frac": (b_asm + b_spmv) / ((r["asm_ms"] + r["spmv_ms"]) * 1e-3) / 1e9 / peak}),
        "roofline_wc": roof("explicit step: kick/move + continuity + momentum + CFL kernels (two-pass element records)", b_wc,
                            w["ms"] * 1e-3, "wc_step_c5",
                            {"workload": f"C5 synthetic 3D Kuhn box n={w['cells']} ({w['n_elems']} tets), CDS_dpdt + Meduri, step + CFL dt "
                                         "chained on the device (pfem_wc_run)",
                             "ms_per_step": w["ms"], "melem_s": w["n_elems"] / (w["ms"] * 1e-3) / 1e6, "kernel_ms": wc_kern_ms,
                             "phases_ms": w["phases"], "launches_per_step": w["launches"], "pattern_build_ms": 1e3 * w["topo_s"],
                             "cpu_baseline": {"value": wc_cpu["n_elems"] / wc_cpu["t_step"] / 1e6, "unit": UNIT, "cores": wc_cpu["cores"],
                                              "kind": wc_cpu["kind"],
                                              "sample": f"Kuhn box n={args.ref_wc_cells} ({wc_cpu['n_elems']} tets), step + CFL dt"}}),
        "cpu_baseline": cpu_baseline_block(args, 1),
        "clocks": r["clocks"],
        "e2e": {"value": n_elems / e["step_s"] / 1e6, "unit": UNIT, "h2d_bytes_per_step": e["h2d"], "d2h_bytes_per_step": e["d2h"],
                "ms_per_step": 1e3 * e["step_s"], "set_topology_ms": 1e3 * e["topo_s"], "iters": e["iters"], "rel_tol": REL_TOL,
                "definition": "per step through host buffers: pfem_set_topology + positions/Dirichlet/states H2D + pfem_pspg_assemble "
                              "(qPrev H2D) + pfem_pspg_solve + solution D2H",
                "assembly_only": {"value": n_elems / e["asm_only_s"] / 1e6, "unit": UNIT, "h2d_bytes_per_step": e["h2d_asm_only"],
                                  "d2h_bytes_per_step": e["d2h_asm_only"]}},
        "gpu_launches": int(r["launches"]),
    }
    return line, failures


def run_multi(env, args):
    """N > 1: C5 explicit weakly-compressible step, strong scaling (headline) + PSPG weak leg + parity vs one GPU."""
    from pfem_b200.capi import PfemContext
    torch = env.torch
    failures = []
    peak, peak_src = measured_peaks()
    world = env.world
    W = mg.WC_PARAMS
    # ---- C5 explicit step, sharded -------------------------------------------------------------------------------------
    log("C5 explicit step, sharded")
    w = wc_leg(env, args, args.wc_cells, args.steps, args.warmup, True)
    log(f"  -> {w['ms']:.3f} ms/step")
    n_elems, n_nodes = w["n_elems"], w["n_nodes"]
    sharded_states = env.gather_owned_to_all(w["states"], w["part"], 8, n_nodes)
    sharded_x = env.gather_owned_to_all(w["xs"], w["part"], 3, n_nodes)
    # ---- the same chain on ONE GPU (rank 0): strong-scaling base and parity -----------------------------------------------
    one = None
    log("C5 explicit step on one GPU (rank 0) + parity")
    if env.rank == 0:
        gmesh = w["gmesh"]
        with PfemContext(3, env.local_rank) as c1:
            stream = torch.cuda.Stream()
            c1.set_stream(stream.cuda_stream)
            c1.set_mesh(gmesh)
            c1.set_states(0, w["packed"])
            wp = c1.wc_params(W["mu"], W["K0"], W["K0p"], W["rhoStar"], mg.gravity(3), True)
            with torch.cuda.stream(stream):
                dt = c1.wc_next_dt(wp, W["securityCoeff"], 1e-3)
                dt, _ = c1.wc_run(wp, args.warmup, W["securityCoeff"], 1e-3, dt)
                torch.cuda.synchronize()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record(stream)
                dt, _ = c1.wc_run(wp, args.steps, W["securityCoeff"], 1e-3, dt)
                ev1.record(stream)
                torch.cuda.synchronize()
                ms1 = ev0.elapsed_time(ev1) / args.steps
                s1, x1 = c1.get_states(0, 8), c1.get_positions()
        same = bool(np.array_equal(s1, sharded_states) and np.array_equal(x1, sharded_x) and dt == w["dt"])
        one = dict(ms=ms1, bit_identical=same, max_abs_diff=float(np.abs(s1 - sharded_states).max()),
                   max_rel_diff=float(np.abs(s1 - sharded_states).max() / max(np.abs(s1).max(), 1e-300)))
        if one["max_rel_diff"] > 1e-12:
            failures.append(f"wc parity vs 1 GPU: max rel diff {one['max_rel_diff']:.3e} > 1e-12")
    del sharded_states, sharded_x
    w.pop("gmesh"), w.pop("packed"), w.pop("states"), w.pop("xs")
    env.barrier()
    # ---- PSPG: weak leg (~2 M tets per GPU) and parity at C4 ----------------------------------------------------------------
    cells_weak = int(round(args.cells * world ** (1.0 / 3.0)))
    log(f"PSPG weak leg n={cells_weak}")
    pw = pspg_leg(env, args, cells_weak, max(2, min(args.steps, 5)), 2, False, failures, f"pspg weak n={cells_weak}")
    weak_1gpu = None
    if env.rank == 0:  # the same (large) mesh on ONE GPU: iteration count and fields to compare the sharded solve with
        gmesh = mg.kuhn_box(3, cells_weak)
        gq, gq_prev = mg.pspg_state(gmesh)
        P = mg.PSPG_PARAMS
        with PfemContext(3, env.local_rank) as c1:
            c1.set_mesh(gmesh)
            c1.set_states(0, gq)
            par = c1.pspg_params(P["rho"], P["mu"], P["dt"], mg.gravity(3))
            c1.pspg_assemble(par, gq_prev)
            s1 = c1.pspg_solve(REL_TOL, MAX_ITER)
        nn = gmesh.n_nodes
        qa, qb = pw["q_global"], s1["q"]
        weak_1gpu = dict(iters=s1["iters"], status=s1["status"],
                         rel_dv=float(np.abs(qa[: 3 * nn] - qb[: 3 * nn]).max() / np.abs(qb[: 3 * nn]).max()),
                         rel_dp=float(np.abs(qa[3 * nn:] - qb[3 * nn:]).max() / np.abs(qb[3 * nn:]).max()))
        if weak_1gpu["rel_dv"] > 1e-8 or weak_1gpu["rel_dp"] > 1e-8:
            failures.append(f"pspg weak parity vs 1 GPU: rel|dv|={weak_1gpu['rel_dv']:.2e} rel|dp|={weak_1gpu['rel_dp']:.2e} > 1e-8")
        if pw["iters"] > 5 * max(s1["iters"], 1):
            failures.append(f"pspg weak: {pw['iters']} iterations > 5x the single-GPU count {s1['iters']} on the same mesh")
        del gmesh, gq, gq_prev, qa, qb
    pw.pop("q_global")
    log(f"  -> {pw['iters']} iterations, solve {pw['solve_ms']:.1f} ms; PSPG C4 sharded + parity")
    pc = pspg_leg(env, args, args.cells, 2, 1, False, failures, f"pspg C4 sharded x{world}")
    parity_pspg = None
    if env.rank == 0:
        gmesh = mg.kuhn_box(3, args.cells)
        gq, gq_prev = mg.pspg_state(gmesh)
        P = mg.PSPG_PARAMS
        with PfemContext(3, env.local_rank) as c1:
            c1.set_mesh(gmesh)
            c1.set_states(0, gq)
            par = c1.pspg_params(P["rho"], P["mu"], P["dt"], mg.gravity(3))
            c1.pspg_assemble(par, gq_prev)
            s1 = c1.pspg_solve(REL_TOL, MAX_ITER)
        nn = gmesh.n_nodes
        qa, qb = pc["q_global"], s1["q"]
        ev = float(np.abs(qa[: 3 * nn] - qb[: 3 * nn]).max() / np.abs(qb[: 3 * nn]).max())
        ep = float(np.abs(qa[3 * nn:] - qb[3 * nn:]).max() / np.abs(qb[3 * nn:]).max())
        parity_pspg = dict(rel_dv=ev, rel_dp=ep, iters_sharded=pc["iters"], iters_1gpu=s1["iters"], status_1gpu=s1["status"])
        if ev > 1e-8 or ep > 1e-8:
            failures.append(f"pspg parity vs 1 GPU: rel|dv|={ev:.2e} rel|dp|={ep:.2e} > 1e-8")
        if pc["iters"] > 5 * max(s1["iters"], 1):
            failures.append(f"pspg C4 sharded: {pc['iters']} iterations > 5x the single-GPU count {s1['iters']}")
    pc.pop("q_global")
    b_wc = wc_step_bytes(n_nodes, n_elems)
    agg_peak = peak * world
    ach = b_wc / (w["ms"] * 1e-3) / 1e9
    line = None
    if env.rank == 0:
        wc_cpu = wc_cpu_sample(args.ref_wc_cells, steps=2)
        value = n_elems / (w["ms"] * 1e-3) / 1e6
        line = {
            "metric": METRIC_WC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": w["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"C5 synthetic 3D Kuhn box n={w['cells']} ({n_elems} tets, {n_nodes} nodes), explicit weakly-compressible "
                                   "step (CDS_dpdt + Meduri): kick/move + continuity + momentum + CFL dt per step, dt chained on the device",
                       "partition": f"RCB nodes + ghost-element layer over {world} GPUs (csrc/partition.cu), 2 halo exchanges + 1 min-all-reduce "
                                    "per step over NCCL (serialised with the node passes; PFEM_WC_OVERLAP=1 overlaps them: measured slower at 4 GPUs)",
                       "n_elems": n_elems, "n_nodes": n_nodes,
                       "l2_policy": "inputs larger than L2 (element records %.0f MB per GPU vs 126 MB L2)" % (n_elems * 160 / 1e6 / world),
                       "value_definition": "n_elems / max-over-ranks device time per step of pfem_wc_run"},
            "phases_ms": w["phases"], "partition_host_s": w["part_s"], "pattern_build_ms": 1e3 * w["topo_s"],
            "wc_1gpu": {"ms_per_step": one["ms"], "melem_s": n_elems / (one["ms"] * 1e-3) / 1e6,
                        "note": "the same chain on one GPU, run by rank 0 in this job"},
            "strong_efficiency_vs_1gpu": one["ms"] / (w["ms"] * world),
            "parity_vs_1gpu": {"wc_bit_identical": one["bit_identical"], "wc_max_abs_diff": one["max_abs_diff"],
                               "wc_max_rel_diff": one["max_rel_diff"], "wc_steps_compared": args.warmup + args.steps,
                               "pspg_rel_dv": parity_pspg["rel_dv"], "pspg_rel_dp": parity_pspg["rel_dp"],
                               "pspg_iters_sharded": parity_pspg["iters_sharded"], "pspg_iters_1gpu": parity_pspg["iters_1gpu"],
                               "pspg_workload": f"C4 n={args.cells} sharded over {world} GPUs vs one GPU, rel tol {REL_TOL:g}"},
            "pspg_weak": {"workload": f"Kuhn box n={cells_weak} ({pw['n_elems']} tets, {pw['n_elems'] / world / 1e6:.2f} M per GPU), assembly + "
                                      f"multigrid-FGMRES (rel tol {REL_TOL:g})",
                          "assembly_melem_s": pw["n_elems"] / (pw["asm_ms"] * 1e-3) / 1e6, "assembly_ms": pw["asm_ms"],
                          "step_ms": pw["step_ms"], "solve_ms": pw["solve_ms"], "iters": pw["iters"], "status": pw["status"],
                          "rel_res": pw["rel_res"], "mg_levels": pw["levels"], "halo_us": 1e3 * pw["halo_ms"],
                          "precond_setup_ms": pw["pre_setup_ms"], "iters_vs_1gpu_c4": pw["iters"] / max(parity_pspg["iters_1gpu"], 1),
                          "same_mesh_on_1gpu": {"iters": weak_1gpu["iters"], "status": weak_1gpu["status"],
                                                "iters_ratio": pw["iters"] / max(weak_1gpu["iters"], 1),
                                                "rel_dv": weak_1gpu["rel_dv"], "rel_dp": weak_1gpu["rel_dp"]},
                          "partition_host_s": pw["part_s"], "pattern_build_ms": 1e3 * pw["topo_s"]},
            "roofline": {"kernel": "explicit step (all kernels + exchanges of one step)", "bound": "hbm", "achieved": ach, "peak": agg_peak,
                         "unit": "GB/s", "frac": ach / agg_peak, "traffic": None, "algorithmic_bytes": b_wc,
                         "peak_source": peak_src + " x n_gpus", "frac_of_8TBs_nominal": ach / (8000.0 * world)},
            "cpu_baseline": {"value": wc_cpu["n_elems"] / wc_cpu["t_step"] / 1e6, "unit": UNIT, "cores": wc_cpu["cores"],
                             "kind": wc_cpu["kind"],
                             "sample": f"Kuhn box n={args.ref_wc_cells} ({wc_cpu['n_elems']} tets), step + CFL dt: "
                                       + (REF_NOTE_WC if wc_cpu["kind"] == "reference" else "oracle port")},
            "clocks": w["clocks"],
            # the step operates on device-resident state between remeshes (WCompNewton/Solver.cpp:236-276 remeshes every maxDT of
            # simulated time, ~100 steps): its end-to-end form is the same call -- no per-step host buffers exist on this path
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 24,
                    "note": "pfem_wc_run keeps the states on the device between remeshes; per step only dt/elapsed/NaN flag return"},
            "gpu_launches": w["launches"],
        }
    return line, failures


def run_gpu(args):
    # the contract is ONE JSON line on stdout: park the real stdout and send everything libraries print (NCCL banner ...)
    # to stderr until the line is ready
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    env = Env()
    if env.world == 1:
        line, failures = run_single(env, args)
    else:
        line, failures = run_multi(env, args)
    nfail = env.sum_over_ranks([len(failures)])[0]
    if env.world > 1:
        env.dist.barrier()
        env.dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if line is not None:
        if failures:
            line["failed"] = failures
        print(json.dumps(line), flush=True)
    if nfail:
        for f in failures:
            print("bench.py: FAILED: " + f, file=sys.stderr)
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpu", choices=["gpu", "reference"])
    ap.add_argument("--cells", type=int, default=69, help="Kuhn box cells per side of the PSPG workload (69 -> C4)")
    ap.add_argument("--wc-cells", type=int, default=150, help="Kuhn box cells per side of the explicit-step workload (150 -> C5)")
    ap.add_argument("--cpu-cells", type=int, default=30, help="bounded CPU-baseline sample inside the GPU run")
    ap.add_argument("--ref-cells", type=int, default=34, help="bounded sample of the --impl reference arm (PSPG)")
    ap.add_argument("--ref-wc-cells", type=int, default=40, help="bounded sample of the CPU explicit step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
