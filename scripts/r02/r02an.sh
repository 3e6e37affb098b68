# ncu --set full of the step-parallel sweep kernel at 512 chains (one warp per scheduler): what is a tile's 4,000 cycles made of?
export DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_lz.so
B="python bench.py --steps 2 --warmup 2 --sweeps-per-step 2 --no-cpu-baseline --no-e2e --no-uncached --no-self-check --chains 512 --sweep-mode 5"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:sweep_sp -s 4 -c 1 -o gpurun_out/prof_r02an -f $B > gpurun_out/r02an_ncu.log 2>&1
tail -2 gpurun_out/r02an_ncu.log | cut -c1-200
