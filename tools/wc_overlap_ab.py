"""A/B of the explicit step on the sharded C5 mesh under torchrun: prints ms/step of bench.py's wc_leg for the environment it
is started with (PFEM_WC_OVERLAP=0|1, PFEM_WC_CFG=...).  Usage: torchrun ... tools/wc_overlap_ab.py [cells] [steps]"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 150
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
env = bench.Env()
args = types.SimpleNamespace(steps=steps, warmup=3)
w = bench.wc_leg(env, args, cells, steps, 3, True)
if env.rank == 0:
    print(f"N={env.world} cells={cells} env={ {k: v for k, v in os.environ.items() if k.startswith('PFEM_')} }: {w['ms']:.3f} ms/step, "
          f"phases {{ {', '.join(f'{k}: {v:.3f}' for k, v in sorted(w['phases'].items()))} }}", flush=True)
