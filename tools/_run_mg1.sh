python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/r1_bench_n1.json
cat gpurun_out/r1_bench_n1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','assembly_ms','krylov','roofline','roofline_spmv','e2e','gpu_launches')})"
tail -3 gpurun_out/bench_err.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r1_launches_bench.csv gpurun_out/r1_launches_bench_c4.md | head -30
