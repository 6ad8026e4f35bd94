# Evidence for DESIGN 4.8 (the iteration without the w vector), one B200
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err
# the two kernels of one iteration (launches 61, 62 of the first solve), full counters
ncu --set full --clock-control none --import-source on -k regex:"k_cg_fused_apply" -s 60 -c 2 -f -o gpurun_out/cg_kernels_b python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_cg_b.log 2>&1
# launch list of a window of the timed solve (shares of the two modes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/launches_b.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/launches_b.log 2>&1
# python tools/profile_summary.py kernels gpurun_out/cg_kernels_b.ncu-rep profiles/<name>.md profiles/<name>.json "<description>"
# python tools/profile_summary.py launches gpurun_out/launches_b.csv profiles/<name>.md
