mkdir -p gpurun_out
python -c "from calico_b200 import build; build.build()" || exit 1
for n in 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
tail -c 300 gpurun_out/bench_n$n.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$n.json").read().strip().splitlines()[-1])
print("RESULT n=$n it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["phases_ms_per_iteration"], "roofline %.3f" % d["roofline"]["frac"], "e2e %.1f" % d["e2e"]["value"], "cost", d["config"]["final_cost"])
PY
done
