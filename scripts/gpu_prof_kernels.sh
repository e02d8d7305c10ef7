mkdir -p gpurun_out
python -c "from calico_b200 import build; build.build()"
CB2_PROFILE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_kprof.json 2> gpurun_out/kprof.txt
python - <<PY
import json
d=json.load(open("gpurun_out/bench_kprof.json"))
print("RESULT it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["phases_ms_per_iteration"])
PY
grep "cb2 profile" gpurun_out/kprof.txt | head -45
