set -x
mkdir -p gpurun_out && rm -f gpurun_out/*.ncu-rep
python -c "from calico_b200 import build; build.build(True)"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
for cps in 32 48 72 110 160; do
  CB2_CHUNK_CPS=$cps timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C4_cps$cps.json 2>> gpurun_out/bench_sweep.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_C4_cps$cps.json"))
print("cps", $cps, "it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["phases_ms_per_iteration"], "roofline %.3f" % d["roofline"]["frac"])
PY
done
CB2_NVCC_EXTRA="-DCB2_EVAL_MINBLOCKS=3" python -c "from calico_b200 import build; build.build(True)"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C4_mb3.json 2>> gpurun_out/bench_sweep.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_C4_mb3.json"))
print("minblocks 3", "it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["phases_ms_per_iteration"], "roofline %.3f" % d["roofline"]["frac"])
PY
python -c "from calico_b200 import build; build.build(True)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_C4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
