for v in v1 v2 v3; do
export DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_$v.so
echo "== $v"
timeout 600 python -m pytest tests/test_gpu_sweep_pipeline.py -m gpu -x -q -k "lorenz and 64 and ws_small" 2>&1 | tail -3
done
