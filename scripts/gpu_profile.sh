# ncu captures of the hot kernels on the C4 workload (one GPU). Reports go to gpurun_out/ (scratch, <= 64 MiB).
set -x
mkdir -p gpurun_out && rm -f gpurun_out/*.ncu-rep
python -c "from calico_b200 import build; build.build()"
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_kernel --launch-skip 82 --launch-count 1 -f -o gpurun_out/prof_eval_cam $B > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:accumulate_kernel --launch-count 1 -f -o gpurun_out/prof_accumulate $B > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:band_factor_kernel --launch-count 2 -f -o gpurun_out/prof_factor $B > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log
timeout 600 ncu --set full --clock-control none -k regex:"band_backsolve_kernel|reduced_solve_kernel|border_gram_kernel" --launch-count 5 -f -o gpurun_out/prof_solve_rest $B > gpurun_out/ncu4.log 2>&1; tail -2 gpurun_out/ncu4.log
ls -la gpurun_out; du -sh gpurun_out
