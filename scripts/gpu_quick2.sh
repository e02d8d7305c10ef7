mkdir -p gpurun_out
python -c "from calico_b200 import build; build.build()" || exit 1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_q.json"))
print("RESULT it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["phases_ms_per_iteration"], "roofline %.3f" % d["roofline"]["frac"], "e2e %.1f" % d["e2e"]["value"], "cost", d["config"]["final_cost"])
PY
CB2_PROFILE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep "cb2 profile" | tail -24 | head -12
