python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_multi_b.log; tail -6 gpurun_out/r02_pytest_multi_b.log
for e in "" "DPP_NO_IPC_BOX=1"; do
echo "== env: $e"
env $e python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29554 tools/mgpu_check.py 32 1 2>&1 | grep "rank 0"
env $e python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tools/mgpu_check.py 16 2 2>&1 | grep "rank 0"
env $e python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --config 4 --gpus 2 --steps 1 --warmup 1 2>/dev/null | grep -o '"value": [0-9.]*, "unit": "GDoF/s", "n_gpus": [0-9]*\|"ms_per_step": [0-9.]*\|"inner_iterations": [0-9]*' | head -3 | paste - - -
done
