# Weak scaling of BASELINE configs[3] (profiles/r02_scaling.md):  gpurun --gpus N --timeout 600 -- 'bash tools/batches/cfg4_scaling.sh N'
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29563 bench.py --config 4 --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_cfg4_n$N.json 2> gpurun_out/bench_cfg4_n$N.err
tail -1 gpurun_out/bench_cfg4_n$N.json | cut -c1-400
