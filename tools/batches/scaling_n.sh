N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "rc=$?"
tail -c 300 gpurun_out/r02_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 tools/per_rank_probe.py 32 2>&1 | grep "world="
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29563 bench.py --config 4 --gpus $N --steps 1 --warmup 1 > gpurun_out/r02_bench_cfg4_n$N.json 2> gpurun_out/r02_bench_cfg4_n$N.err; echo "rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29564 tools/mgpu_soak.py 32 1000 2>&1 | grep "SOAK\|rror\|differs" | head -5
grep -o '"value": [0-9.]*, "unit": "GDoF/s", "n_gpus": [0-9]*\|"ms_per_step": [0-9.]*\|"iterations": [0-9]*\|"inner_iterations": [0-9]*' gpurun_out/r02_bench_n$N.json gpurun_out/r02_bench_cfg4_n$N.json | head -12
