mkdir -p gpurun_out && rm -f gpurun_out/*.ncu-rep
python -c "from calico_b200 import build; build.build()" || exit 1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_C4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'eval_kernel<\(int\)., \(int\)2>|accumulate_kernel|cr_level_kernel<\(bool\)1>|border_gram_dmma|reduced_solve_smem|cr_back_kernel' --launch-count 7 -f -o gpurun_out/prof_final $B > gpurun_out/ncu_final.log 2>&1; tail -1 gpurun_out/ncu_final.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'cr_level_kernel<\(bool\)0>' --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_final_cr $B > gpurun_out/ncu_final2.log 2>&1; tail -1 gpurun_out/ncu_final2.log
