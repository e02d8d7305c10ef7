mkdir -p gpurun_out && rm -f gpurun_out/*.ncu-rep
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_C4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'eval_kernel<\(int\)., \(int\)2>|accumulate_kernel<|cr_level_kernel<\(bool\)1>|border_gram_dmma|reduced_solve_smem|cr_back_kernel' --launch-count 7 -f -o gpurun_out/prof_final $B > gpurun_out/ncu_final.log 2>&1; tail -1 gpurun_out/ncu_final.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'cr_level_kernel<\(bool\)0>' --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_final_cr $B > gpurun_out/ncu_final2.log 2>&1; tail -1 gpurun_out/ncu_final2.log
