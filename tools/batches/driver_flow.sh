set -x
( time python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" ) 2>&1 | tail -6
( time python bench.py --steps 20 --warmup 20 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err ) 2>&1 | tail -4
tail -c 500 gpurun_out/r02_bench_default.err
( time python bench.py --impl reference --steps 20 --warmup 20 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err ) 2>&1 | tail -4
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_default.json") if l.startswith("{")][-1])
for k in ("value", "ms_per_step", "iterations", "e2e", "roofline", "cpu_baseline", "roofline_general", "assembly", "first_call_ms", "clocks", "gpu_launches", "solve_roofline"):
    print(k, json.dumps(d.get(k))[:900])
r = json.loads([l for l in open("gpurun_out/r02_bench_ref.json") if l.startswith("{")][-1])
print("ref", r["value"], r["ms_per_step"], r["cpu_baseline"])
PY
