mkdir -p gpurun_out
python -c "from calico_b200 import build; build.build()" || exit 1
for c in C2 C3 C5; do
timeout 900 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; tail -c 400 gpurun_out/bench_$c.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$c.json").read().strip().splitlines()[-1])
    print("RESULT $c it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["phases_ms_per_iteration"], "roofline %.3f" % d["roofline"]["frac"], "e2e %.1f" % d["e2e"]["value"], "cost %.4e -> %.4e" % (d["config"]["initial_cost"], d["config"]["final_cost"]), d["config"]["residual_blocks"])
except Exception as e:
    print("RESULT $c FAILED", e, open("gpurun_out/bench_$c.json").read()[-300:])
PY
done
