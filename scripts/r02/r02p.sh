export DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_w5t.so
python bench.py --chains 512 --steps 1 --warmup 3 --sweeps-per-step 1 --no-cpu-baseline --no-e2e --no-uncached --no-self-check --sweep-mode 3 > gpurun_out/r02p_trace.txt 2>&1
grep -c "^TR" gpurun_out/r02p_trace.txt
