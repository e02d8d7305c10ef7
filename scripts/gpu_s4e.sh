set -x
mkdir -p gpurun_out && rm -f gpurun_out/*.ncu-rep
python -c "from calico_b200 import build; build.build()"
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'cr_level_kernel|border_gram_dmma|reduced_solve_smem' --launch-skip 1 --launch-count 4 -f -o gpurun_out/prof_s4e $B > gpurun_out/ncu_s4e.log 2>&1; tail -1 gpurun_out/ncu_s4e.log
