export DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_ca.so
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "(guiding_cache and lorenz) or bench_configuration or full_size_properties or slice_of_the_full" 2>&1 | tail -2
B="python bench.py --steps 10 --warmup 3 --sweeps-per-step 4 --no-cpu-baseline --no-e2e --no-uncached --no-self-check"
for ch in 4096 2048 1024 512; do timeout 200 $B --chains $ch > gpurun_out/r02ad_b${ch}.json 2>gpurun_out/r02ad.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02ad_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_sweep'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()}, '%.4g'%d['value'])
    except Exception as e: print(f,'ERR',e)
PY
