#!/usr/bin/env bash
# The gpurun calls behind the round-2 numbers in DESIGN.md / profiles/ (run from the repo root in the dev container).
# Each line is one call; they were not run as one script (GPU budget), the order does not matter.  Multi-GPU calls carry an
# inner `timeout` shorter than gpurun's: a hung collective must not take the box down with it.
set -euo pipefail
G=/usr/local/graft/bin/gpurun

# parity + smoke (216 GPU tests, ~75 s)
$G --timeout 1200 -- 'timeout 1100 python -m pytest tests -m gpu -x -q; python -c "import __graft_entry__ as g; g.smoke()"'
# headline bench at N=1 -> profiles/r2_bench_n1.json
$G --timeout 900 -- 'timeout 800 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json'
# N>1: C5 explicit step, strong scaling, with the 1-GPU chain, parity and the PSPG weak leg in the same job -> profiles/r2_bench_n{2,4,8}.json
for N in 2 4 8; do
  $G --gpus $N --timeout 420 -- "timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json"
done
# the Krylov solve alone, its variants: PFEM_GMRES_M=0 (BiCGSTAB), PFEM_MG_FP32V=0, PFEM_MG_NU, PFEM_MG_PRE/POST/PREC/POSTC, PFEM_MG_OVER
$G --timeout 600 -- 'for sz in "69 1e-12 3" "133000 1e-12 3 cloud" "700 1e-12 2"; do python tools/solve_bench.py $sz; done'
# whole remeshed step through host buffers, phase by phase
$G --timeout 300 -- 'python tools/e2e_breakdown.py 69'
# ncu: launch list of one solve -> profiles/r2_gmres_launches.md ; --set full of the sweep and the Krylov SpMV -> profiles/r2_ncu_mg.md
$G --timeout 600 -- 'SOLVE_REPS=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_gmres_launches.csv python tools/solve_bench.py 69; python tools/ncu_traffic.py gpurun_out/r2_gmres_launches.csv'
$G --timeout 600 -- 'SOLVE_REPS=1 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_spmv<\(int\)4, \(int\)4, \(int\)2, float, \(int\)2, float>|k_spmv<\(int\)4, \(int\)3, \(int\)0, double" -s 10 -c 3 -o gpurun_out/r2_smooth_f32b -f python tools/solve_bench.py 69; python tools/ncu_summary.py gpurun_out/r2_smooth_f32b.ncu-rep gpurun_out/r2_smooth.md'
# ncu: assembly kernel -> profiles/r2_ncu_pspg.md ; explicit step traffic (two-pass and tiles) -> profiles/r2_ncu_wc.md
$G --timeout 600 -- 'ncu --set full --clock-control none --import-source on -k regex:k_pspg_assemble2 -s 2 -c 1 -o gpurun_out/r2_asm5 -f python tools/quick_asm_bench.py'
$G --timeout 900 -- 'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_wc --csv --log-file gpurun_out/r2_wc_c5_launches.csv python tools/wc_steps.py 150 3; python tools/ncu_traffic.py gpurun_out/r2_wc_c5_launches.csv 3'
