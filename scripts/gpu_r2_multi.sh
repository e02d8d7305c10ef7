# usage: bash scripts/gpu_r2_multi.sh N   (run under gpurun --gpus N): multi-GPU parity tests that fit N GPUs + the C4 bench at N ranks
N=${1:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r2_pytest_multi_n$N.log
for n in $(seq 2 $N); do
  case $n in 2|4|8) ;; *) continue;; esac
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err
  tail -c 300 gpurun_out/r2_bench_n$n.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_n$n.json").read().strip().splitlines()[-1])
    print("RESULT n=$n it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["phases_ms_per_iteration"], "e2e %.1f" % d["e2e"]["value"], "cost", d["config"]["final_cost"])
except Exception as e:
    print("RESULT n=$n FAILED", e)
PY
done
