# ncu --set full of the kernels named in $1 (regex), one bench iteration; reports into gpurun_out/
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$1" --launch-count ${2:-4} -f -o gpurun_out/prof_r2 $B > gpurun_out/ncu_r2.log 2>&1; tail -2 gpurun_out/ncu_r2.log
ncu -i gpurun_out/prof_r2.ncu-rep --page raw --csv > gpurun_out/prof_r2_raw.csv 2>/dev/null; wc -l gpurun_out/prof_r2_raw.csv
