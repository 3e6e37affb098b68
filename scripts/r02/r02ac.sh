B="python bench.py --steps 10 --warmup 3 --sweeps-per-step 4 --no-cpu-baseline --no-e2e --no-uncached --no-self-check"
for v in g1 g2 g4; do
export DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_$v.so
for ch in 4096 512; do timeout 200 $B --chains $ch > gpurun_out/r02ac_${v}_b${ch}.json 2>gpurun_out/r02ac.err; done
done
export DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_g2.so
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "guiding_cache and lorenz" 2>&1 | tail -2
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02ac_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_sweep'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()})
    except Exception as e: print(f,'ERR',e)
PY
