# End-of-round evidence run on ONE B200: clean-tree rebuild ON the box, the whole GPU suite, smoke, the bench line, the reference arm,
# the ncu launch list and --set full captures of the dominant kernels. Everything lands in gpurun_out/ (summaries are then copied to profiles/).
mkdir -p gpurun_out && rm -f gpurun_out/*.ncu-rep
rm -f calico_b200/libcalico_b200.so oracle/liboracle.so
python -c "import __graft_entry__ as g; g.build()" || exit 1
ls -la calico_b200/libcalico_b200.so
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 300 gpurun_out/r2_bench_final.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_C4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'eval_kernel<\(int\)., \(int\)2>|accumulate_kernel<|expand_gram|cr_level_kernel<\(bool\)1>|border_gram_dmma|reduced_solve_smem|cr_back_kernel' --launch-count 8 -f -o gpurun_out/prof_r2_final $B > gpurun_out/ncu_r2_final.log 2>&1; tail -1 gpurun_out/ncu_r2_final.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'cr_level_kernel<\(bool\)0>' --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_r2_final_cr $B > gpurun_out/ncu_r2_final2.log 2>&1; tail -1 gpurun_out/ncu_r2_final2.log
