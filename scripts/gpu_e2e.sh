python -c "from calico_b200 import build; build.build()" || exit 1
python scripts/e2e_breakdown.py C4 10 2>&1 | tail -3
for v in 0 1; do
if [ $v = 1 ]; then export CB2_IMU_STREAM=1; fi
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q$v.json 2> gpurun_out/bench_q.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_q$v.json"))
print("RESULT imu_stream=$v it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["phases_ms_per_iteration"], "roofline %.3f sweep %.3f" % (d["roofline"]["frac"], d["roofline"]["whole_sweep"]["frac"]), "e2e %.1f" % d["e2e"]["value"], "cost", d["config"]["final_cost"])
PY
done
