mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 300 gpurun_out/bench_full.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_full.json"))
print("RESULT it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["phases_ms_per_iteration"], "roofline %.3f sweep %.3f" % (d["roofline"]["frac"], d["roofline"]["whole_sweep"]["frac"]), "e2e %.1f" % d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], d["clocks"])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
