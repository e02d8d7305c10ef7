set -x
mkdir -p gpurun_out && rm -f gpurun_out/*.ncu-rep
python -c "from calico_b200 import build; build.build()"
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"band_factor_kernel|band_backsolve_kernel|reduced_solve_kernel" --launch-count 5 -f -o gpurun_out/prof_solve $B > gpurun_out/ncu3.log 2>&1; tail -1 gpurun_out/ncu3.log
du -sh gpurun_out
