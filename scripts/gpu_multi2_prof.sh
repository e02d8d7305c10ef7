mkdir -p gpurun_out
python -c "from calico_b200 import build; build.build()" || exit 1
CB2_PROFILE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2p.json 2> gpurun_out/bench_n2p.err
grep "cb2 profile" gpurun_out/bench_n2p.err | head -34
