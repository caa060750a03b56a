N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 2>gpurun_out/b${N}err.log | tail -1 > gpurun_out/r1_bench_n${N}.json
python -c "import json; d=json.load(open('gpurun_out/r1_bench_n${N}.json')); print({k:d[k] for k in ('value','ms_per_step','krylov')})"
