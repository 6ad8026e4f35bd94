set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_pytest_gpu_a.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
python tools/csr_timing.py > gpurun_out/r02_csr_timing_a.log 2>&1
python tools/sanitize_small.py 10 > gpurun_out/r02_sanitize_plain.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck.log python tools/sanitize_small.py 8 > gpurun_out/r02_memcheck.out 2>&1
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r02_racecheck.log python tools/sanitize_small.py 8 > gpurun_out/r02_racecheck.out 2>&1
timeout 600 compute-sanitizer --tool synccheck --log-file gpurun_out/r02_synccheck.log python tools/sanitize_small.py 8 cg > gpurun_out/r02_synccheck.out 2>&1
tail -5 gpurun_out/r02_pytest_gpu_a.log; tail -c 600 gpurun_out/r02_bench_a.json; tail -3 gpurun_out/r02_memcheck.log gpurun_out/r02_racecheck.log gpurun_out/r02_synccheck.log
