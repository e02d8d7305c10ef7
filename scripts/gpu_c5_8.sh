mkdir -p gpurun_out
python -c "from calico_b200 import build; build.build()" || exit 1
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 --config C5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n8.json 2> gpurun_out/bench_c5_n8.err
tail -c 300 gpurun_out/bench_c5_n8.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c5_n8.json").read().strip().splitlines()[-1])
print("RESULT C5 n=8 it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["phases_ms_per_iteration"], "roofline %.3f" % d["roofline"]["frac"], "e2e %.1f" % d["e2e"]["value"], "cost %.4e -> %.4e" % (d["config"]["initial_cost"], d["config"]["final_cost"]))
PY
