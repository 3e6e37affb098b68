export DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_w5.so
B="python bench.py --steps 3 --warmup 1 --sweeps-per-step 2 --no-cpu-baseline --no-e2e --no-uncached --no-self-check --sweep-mode 3"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sweep_ws -s 2 -c 2 -o gpurun_out/prof_r02m_ws512 -f $B --chains 512 > gpurun_out/r02m_ncu512.log 2>&1
ls -la gpurun_out/*.ncu-rep
