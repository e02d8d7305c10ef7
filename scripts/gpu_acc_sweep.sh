set -x
mkdir -p gpurun_out && rm -f gpurun_out/*.ncu-rep
run() {
  CB2_NVCC_EXTRA="$1" python -c "from calico_b200 import build; build.build(True)"
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_acc.json 2>> gpurun_out/bench_acc.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_acc.json"))
print("RESULT [$1] it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["phases_ms_per_iteration"], "roofline %.3f" % d["roofline"]["frac"], "e2e %.1f" % d["e2e"]["value"])
PY
}
python -c "from calico_b200 import build; build.build(True)"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
run ""
run "-DCB2_ACC_ROWS=64"
run "-DCB2_ACC_ROWS=16"
run "-DCB2_ACC_MINBLOCKS=8 -DCB2_ACC_ROWS=24"
run "-DCB2_ACC_MINBLOCKS=7"
run "-DCB2_EVAL_MINBLOCKS=4"
python -c "from calico_b200 import build; build.build(True)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_C4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
