set -x
mkdir -p gpurun_out && rm -f gpurun_out/*.ncu-rep
python -c "from calico_b200 import build; build.build()"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g.json 2>> gpurun_out/bench_g.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_g.json"))
print("RESULT it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["phases_ms_per_iteration"], "roofline %.3f" % d["roofline"]["frac"], "e2e %.1f" % d["e2e"]["value"])
PY
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"eval_kernel|camera_frame_kernel" -s 8 --launch-count 4 -f -o gpurun_out/prof_eval2 $B > gpurun_out/ncu_g.log 2>&1; tail -1 gpurun_out/ncu_g.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_C4_g.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
du -sh gpurun_out
