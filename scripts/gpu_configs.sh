# the other BASELINE configs on one B200 (C2, C3 with the measured CPU baseline beside them; C5 GPU only)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
for c in C2 C3; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err; tail -c 200 gpurun_out/r2_bench_$c.err
done
timeout 900 python bench.py --config C5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_C5.json 2> gpurun_out/r2_bench_C5.err; tail -c 200 gpurun_out/r2_bench_C5.err
python - <<PY
import json
for c in ("C2", "C3", "C5"):
    try:
        d = json.load(open("gpurun_out/r2_bench_%s.json" % c))
        cb = d.get("cpu_baseline", {})
        print("RESULT", c, "it/s %.1f" % d["value"], "jac evals/s %.3g" % (d["jacobian_evals_per_sec"] or 0), "roofline %.3f" % d["roofline"]["frac"], "e2e %.1f" % d["e2e"]["value"],
              "cpu", cb.get("value"), cb.get("cores"), d["phases_ms_per_iteration"])
    except Exception as e:
        print("RESULT", c, "FAILED", e)
PY
