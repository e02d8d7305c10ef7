set -x
mkdir -p gpurun_out && rm -f gpurun_out/*.ncu-rep
python -c "from calico_b200 import build; build.build()"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_h.json 2>> gpurun_out/bench_h.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_h.json"))
print("RESULT it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["phases_ms_per_iteration"], "roofline %.3f" % d["roofline"]["frac"], "e2e %.1f" % d["e2e"]["value"])
PY
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'eval_kernel<\(int\)., \(int\)2>|accumulate_kernel' --launch-count 4 -f -o gpurun_out/prof_eval3 $B > gpurun_out/ncu_g.log 2>&1; tail -1 gpurun_out/ncu_g.log
du -sh gpurun_out
