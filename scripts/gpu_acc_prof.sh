mkdir -p gpurun_out && rm -f gpurun_out/*.ncu-rep
python -c "from calico_b200 import build; build.build()" || exit 1
CB2_ACC_GENERIC=1 CB2_PROFILE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep "accumulate" | cut -c1-150 | tail -1
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'accumulate_kernel<' --launch-count 1 -f -o gpurun_out/prof_acc $B > gpurun_out/ncu_acc.log 2>&1; tail -1 gpurun_out/ncu_acc.log
