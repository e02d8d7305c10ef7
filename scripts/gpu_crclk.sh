mkdir -p gpurun_out
CB2_NVCC_EXTRA="-DCB2_CR_CLOCKS" python -c "from calico_b200 import build; build.build(force=True)" || exit 1
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | grep "cr clocks" | tail -6
rm -f calico_b200/libcalico_b200.so
