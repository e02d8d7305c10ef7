mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_q.json"))
print("RESULT it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["phases_ms_per_iteration"], "roofline %.3f" % d["roofline"]["frac"], "e2e %.1f" % d["e2e"]["value"], "cost", d["config"]["final_cost"])
PY
