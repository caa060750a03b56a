"""Fixtures of the fractional-step solver, made by the reference's own code built in place
(oracle/refbuild -> oracle/_ref): the three linear systems of one Picard body of MomContEquationFracStep.inl, each built by
the reference's m_buildMat* + m_applyBC* from the inputs stored next to it, and the stand-in ConjugateGradient's solution and
iteration count for the two velocity systems.  Run from the repo root: python tests/golden/make_fracstep.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import oracle as orc      # noqa: E402
from oracle import ref                # noqa: E402
from pfem_b200 import meshgen as mg   # noqa: E402
from make_golden import mesh_arrays, save   # noqa: E402


def csc(prefix, A):
    return {prefix + "_indptr": A.indptr.astype(np.int64), prefix + "_indices": A.indices.astype(np.int32), prefix + "_data": A.data}


def main():
    if not ref.available():
        raise SystemExit("oracle/_ref not built and /root/reference absent")
    P = mg.PSPG_PARAMS
    gamma_fs = 1.0
    for dim, n in ((2, 8), (3, 4)):
        mesh = mg.kuhn_box(dim, n, free_fraction=0.02, permute=True)
        nn = mesh.n_nodes
        # non-zero Dirichlet data on part of the walls: the column elimination of m_applyBCVAppStep is live
        vals = mesh.dir_val.reshape(dim, nn)
        sel = (mesh.dir_mask != 0) & (np.arange(nn) % 2 == 0)
        vals[:, sel] = 0.05 * np.random.default_rng(8).standard_normal((dim, int(sel.sum())))
        mesh.dir_val = np.ascontiguousarray(vals.reshape(-1))
        _, q_prev = mg.pspg_state(mesh)
        q_prev = q_prev + 0.02 * np.random.default_rng(3).standard_normal(q_prev.shape)
        par = orc.pspg_param_array(P["rho"], P["mu"], P["dt"], mg.gravity(dim))
        delta_p = 50.0 * np.random.default_rng(4).standard_normal(nn) + 200.0 * mesh.coords()[:, 0]
        with ref.RefCase(mesh, "pspg", np.concatenate([par, [10, 1e-6, gamma_fs, 2.0]]), solver_id="FracStep") as rc:
            rc.set_states(q_prev)
            A0, b0 = rc.fs_build(0, q_prev[: dim * nn], q_prev[dim * nn:])
            v_tilde, it0, err0, info0 = rc.fs_solve(0)
            A1, b1 = rc.fs_build(1, v_tilde, q_prev[dim * nn:])
            _, it1, err1, info1 = rc.fs_solve(1)
            A2, b2 = rc.fs_build(2, delta_p)
            dv, it2, err2, info2 = rc.fs_solve(2)
        assert info0 == 0 and info2 == 0 and info1 == 2, (info0, info1, info2)
        save(f"fs_{dim}d_kuhn", **mesh_arrays(mesh), q_prev=q_prev, par=par, gamma_fs=np.float64(gamma_fs), delta_p=delta_p,
             **csc("A0", A0), b0=b0, v_tilde=v_tilde, cg0=np.array([it0, err0, info0]),
             **csc("A1", A1), b1=b1, cg1=np.array([it1, err1, info1]),
             **csc("A2", A2), b2=b2, dv=dv, cg2=np.array([it2, err2, info2]))
        print(f"  dim {dim}: {nn} nodes, CG iterations {it0} / {it1} (info {info1}) / {it2}")


if __name__ == "__main__":
    main()
