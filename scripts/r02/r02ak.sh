# ncu --set full of the regrouped cache_apply_kernel (Lorenz-only build), steady state (skip the first 6 launches: cache build + first applies)
export DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_lz.so
B="python bench.py --steps 3 --warmup 3 --sweeps-per-step 2 --no-cpu-baseline --no-e2e --no-uncached --no-self-check"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:cache_apply_kernel -s 8 -c 2 -o gpurun_out/prof_r02ak -f $B > gpurun_out/r02ak_ncu.log 2>&1
tail -3 gpurun_out/r02ak_ncu.log
