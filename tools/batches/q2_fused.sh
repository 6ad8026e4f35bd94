# Evidence for DESIGN 4.5 (fused degree-2 iteration), one B200:  gpurun --timeout 1700 -- 'bash tools/batches/q2_fused.sh'
mkdir -p gpurun_out
python tools/prof_q2.py 192 solve                                     # kernel timings (plain apply, fused apply, r-update, one-field block)
CFG4_N=192 python tools/run_configs.py 4 > gpurun_out/cfg4_192.jsonl  # config 4 at full size on one GPU
python bench.py --config 4 --steps 2 --warmup 1 > gpurun_out/bench_cfg4.json
# full ncu captures (launch 15 = fused apply <2,...>, launch 16 onwards = the one-field kernel)
ncu --set full --clock-control none --import-source on -k regex:"k_cg_fused_apply_q2|k_apply_q2u" -s 14 -c 4 -f -o gpurun_out/q2f_b python tools/prof_q2.py 192 > gpurun_out/ncu_q2f_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_cg_fused_apply_q2" -s 15 -c 1 -f -o gpurun_out/q2f_c python tools/prof_q2.py 192 > gpurun_out/ncu_q2f_c.log 2>&1
# python tools/profile_summary.py kernels gpurun_out/q2f_b.ncu-rep profiles/<name>.md profiles/<name>.json "<description>"
