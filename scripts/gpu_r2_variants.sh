# A/B of the late round-2 changes on one B200: border Gram product of the early levels' rows beside the late reduction levels.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
run() {
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_var_$tag.json 2> gpurun_out/r2_var_$tag.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_var_$tag.json"))
print("RESULT $tag it/s %.1f" % d["value"], "ms/step %.4f" % d["ms_per_step"], {k: round(v, 4) for k, v in d["phases_ms_per_iteration"].items()}, "e2e %.1f" % d["e2e"]["value"], "cost %.9g" % d["config"]["final_cost"])
PY
}
run a4r64 CB2_EARLY_GRAM_AFTER=4 CB2_EARLY_GRAM_RESERVE=64
run a5r64 CB2_EARLY_GRAM_AFTER=5 CB2_EARLY_GRAM_RESERVE=64
run a4r96 CB2_EARLY_GRAM_AFTER=4 CB2_EARLY_GRAM_RESERVE=96
run l4a4r64 CB2_EARLY_GRAM_LEVELS=4 CB2_EARLY_GRAM_AFTER=4 CB2_EARLY_GRAM_RESERVE=64
run l5a5r64 CB2_EARLY_GRAM_LEVELS=5 CB2_EARLY_GRAM_AFTER=5 CB2_EARLY_GRAM_RESERVE=64
run l5a4r48 CB2_EARLY_GRAM_LEVELS=5 CB2_EARLY_GRAM_AFTER=4 CB2_EARLY_GRAM_RESERVE=48
