export DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_lzt.so
for m in 3 4; do
timeout 120 python bench.py --chains 512 --steps 1 --warmup 3 --sweeps-per-step 1 --no-cpu-baseline --no-e2e --no-uncached --no-self-check --sweep-mode $m > gpurun_out/r02z_trace_m$m.txt 2>&1
grep -c "^TR" gpurun_out/r02z_trace_m$m.txt
done
