mkdir -p gpurun_out && rm -f gpurun_out/*.ncu-rep
python -c "from calico_b200 import build; build.build()" || exit 1
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'eval_kernel<\(int\)., \(int\)2>|accumulate_kernel' --launch-skip 0 --launch-count 4 -f -o gpurun_out/prof_s4i $B > gpurun_out/ncu_s4i.log 2>&1; tail -1 gpurun_out/ncu_s4i.log
