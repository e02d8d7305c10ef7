# compute-sanitizer (memcheck + racecheck + synccheck) over one small LM solve through the C ABI (tiny: 2 cameras + IMU; freed chart pose too).
# Logs land in gpurun_out/; the summaries are copied into profiles/ by hand.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
cat > /tmp/san_run.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
from calico_b200 import _capi, synthetic
from oracle import oracle_py
cfg, free_pose = sys.argv[1], len(sys.argv) > 2
truth, prob = synthetic.generate(cfg, oracle_py.oracle_api, noise=True)
if free_pose:
    prob.bodies[0].pose_const = False
a = _capi.CApi(); a.set_device(0)
ids = prob.clone().push(a)
s, log = a.optimize(_capi.Options(minimizer_progress_to_stdout=0, max_num_iterations=4))
for sid in ids:
    a.get_residuals(sid)
print("solve ok:", len(log) - 1, "iterations, final cost", s.final_cost)
a.close()
PY
for tool in memcheck racecheck synccheck; do
  for cfg in "tiny" "tiny free_pose"; do
    tag=$(echo $cfg | tr ' ' '_')
    timeout 120 compute-sanitizer --tool $tool --error-exitcode 0 --print-limit 20 python /tmp/san_run.py $cfg > gpurun_out/sanitizer_${tool}_${tag}.log 2>&1
    echo "== $tool $cfg: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_${tag}.log | tail -1) $(grep 'solve ok' gpurun_out/sanitizer_${tool}_${tag}.log)"
  done
done
