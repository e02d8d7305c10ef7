mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
CB2_PROFILE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n8_kprof.json 2> gpurun_out/r2_kprof_n8.txt
grep "cb2 profile" gpurun_out/r2_kprof_n8.txt | head -34
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29657 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('RESULT n=8 it/s %.1f' % d['value'], d['phases_ms_per_iteration'], 'e2e', d['e2e']['value'])"
