set -x
mkdir -p gpurun_out && rm -f gpurun_out/*.ncu-rep
nvidia-smi -L
python -c "from calico_b200 import build; build.build(True)"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C4_e1.json 2> gpurun_out/bench_e1.err; tail -3 gpurun_out/bench_e1.err
NG=$(nvidia-smi -L | wc -l)
for n in 2 4 8; do
  if [ $n -le $NG ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C4_e$n.json 2> gpurun_out/bench_e$n.err; tail -5 gpurun_out/bench_e$n.err
  fi
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_C4_e*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("RESULT", f, "gpus", d["n_gpus"], "it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["phases_ms_per_iteration"], "e2e %.1f" % d["e2e"]["value"])
    except Exception as e:
        print("RESULT", f, "unreadable", e)
PY
