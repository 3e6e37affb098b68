export DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_lzws.so
timeout 300 python bench.py --chains 512 --steps 3 --warmup 1 --sweeps-per-step 2 --no-cpu-baseline --no-e2e --no-uncached --no-self-check --sweep-mode 3 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_sweep_pipeline.py -m gpu -x -q -k "lorenz and ws_large" 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_sweep_pipeline.py -m gpu -x -q -k "lorenz and 64 and ws_small" 2>&1 | tail -15
