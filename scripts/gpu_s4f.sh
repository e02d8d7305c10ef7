set -x
mkdir -p gpurun_out && rm -f gpurun_out/*.ncu-rep
python -c "from calico_b200 import build; build.build()"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s4f.json 2>> gpurun_out/bench_s4f.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_s4f.json"))
print("RESULT it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["phases_ms_per_iteration"], "roofline %.3f" % d["roofline"]["frac"], "e2e %.1f" % d["e2e"]["value"], "cost", d["config"]["final_cost"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_C4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'border_gram_dmma|reduced_solve_smem|level3_build|apply_step|cr_level_kernel<\(bool\)1>' --launch-count 5 -f -o gpurun_out/prof_s4f $B > gpurun_out/ncu_s4f.log 2>&1; tail -1 gpurun_out/ncu_s4f.log
