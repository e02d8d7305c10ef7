# quick perf iteration: rebuild, two parity tests, bench line, per-kernel event profile
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 400 gpurun_out/r2_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench.json"))
print("RESULT it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["phases_ms_per_iteration"], "roofline %.3f" % d["roofline"]["frac"], "e2e %.1f it/s %.1f ms" % (d["e2e"]["value"], d["e2e"]["ms_per_optimize_call"]), "cost", d["config"]["final_cost"])
PY
CB2_PROFILE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_kprof.json 2> gpurun_out/r2_kprof.txt
grep "cb2 profile" gpurun_out/r2_kprof.txt | tail -26
