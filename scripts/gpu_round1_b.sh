set -x
mkdir -p gpurun_out
python -c "from calico_b200 import build; build.build()"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C4_b.json 2> gpurun_out/bench_C4_b.err; tail -3 gpurun_out/bench_C4_b.err; cat gpurun_out/bench_C4_b.json
timeout 1200 ncu --set full --clock-control none --import-source on --launch-skip 82 --launch-count 42 -f -o gpurun_out/prof_r1_iter python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
