#!/usr/bin/env bash
# The gpurun calls behind the round-1 numbers in DESIGN.md / profiles/ (run from the repo root in the dev container).
# Each line is one call; they were not run as one script (GPU budget), the order does not matter.
set -euo pipefail
G=/usr/local/graft/bin/gpurun

# parity + smoke (118 GPU tests, ~50 s)
$G --timeout 900 -- 'python -m pytest tests -m gpu -x -q; python -c "import __graft_entry__ as g; g.smoke()"'
# headline bench, both arms -> profiles/r1_bench_n1.json, r1_bench_reference.json
$G --timeout 900 -- 'python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json; python bench.py > gpurun_out/bench_n1.json'
# explicit step on C5, one GPU -> profiles/r1_wc_c5_n1.json ; variants: PFEM_WC_CFG=6 (gather) | 11 (two-pass) | 12 (mixed), PFEM_WC_EB=2|3|4
$G --timeout 900 -- 'python tools/bench_wc.py --cells 150 --steps 20 | tail -1 > gpurun_out/r1_wc_c5_n1.json'
# random numbering and host Morton renumbering (8.69 / 2.93 ms per step)
$G --timeout 900 -- 'python tools/bench_wc.py --cells 150 --steps 10 --permute | tail -1; python tools/bench_wc.py --cells 150 --steps 10 --permute --renumber | tail -1'
# strong scaling of the explicit step -> profiles/r1_wc_c5_n{2,4,8}.json
for N in 2 4 8; do
  $G --gpus $N --timeout 600 -- "python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N tools/bench_wc.py --cells 150 --steps 20 | tail -1 > gpurun_out/r1_wc_c5_n$N.json"
done
# multi-GPU parity (bit-identical to one GPU, both kernel variants)
$G --gpus 2 --timeout 600 -- 'python -m pytest tests/test_gpu_multi.py -x -q'
# ncu: per-launch time + DRAM bytes, and --set full of the two-pass kernels -> profiles/r1_ncu_wc*.md
$G --timeout 900 -- 'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_wc --launch-skip 14 --launch-count 7 --csv --log-file gpurun_out/wc.csv python tools/bench_wc.py --cells 150 --steps 2 --warmup 2'
$G --timeout 500 -- 'ncu --set full --clock-control none --import-source on -k regex:"k_wc_(cont|mom)_(elem|node)|k_wc_dt_fast" --launch-skip 10 --launch-count 5 -o gpurun_out/r1_wc_twopass -f python tools/bench_wc.py --cells 150 --steps 2 --warmup 2; python tools/ncu_summary.py gpurun_out/r1_wc_twopass.ncu-rep gpurun_out/r1_ncu_wc_twopass.md'
