export DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_p2.so
B="python bench.py --steps 10 --warmup 3 --sweeps-per-step 4 --no-cpu-baseline --no-e2e --no-uncached --no-self-check"
for ch in 512 1024 1536 2048; do timeout 200 $B --chains $ch --sweep-mode 2 --fwd-lanes 2 > gpurun_out/r02ah_b${ch}.json 2>gpurun_out/r02ah.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02ah_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_sweep'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()}, d['roofline']['kernel'])
    except Exception as e: print(f,'ERR',e)
PY
