set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -5 gpurun_out/sanitizer.log
timeout 600 python bench.py --config C2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C2.json 2> gpurun_out/bench_C2.err; tail -3 gpurun_out/bench_C2.err; cat gpurun_out/bench_C2.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_C4.json 2> gpurun_out/bench_C4.err; tail -3 gpurun_out/bench_C4.err; cat gpurun_out/bench_C4.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_C4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log
