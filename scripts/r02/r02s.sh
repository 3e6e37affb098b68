export DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_lz.so
timeout 200 python -m pytest tests/test_gpu_sweep_pipeline.py -m gpu -x -q -k "lorenz" 2>&1 | tail -2
B="python bench.py --steps 10 --warmup 3 --sweeps-per-step 4 --no-cpu-baseline --no-e2e --no-uncached --no-self-check"
for ch in 256 512 768 1024 1536 2048 4096; do
  for m in 1 2 3; do timeout 120 $B --chains $ch --sweep-mode $m > gpurun_out/r02s_b${ch}_m$m.json 2>gpurun_out/r02s.err; done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02s_b*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_sweep'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()}, d['roofline']['kernel'], round(d['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
PY
