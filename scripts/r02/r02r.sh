timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02r_tests.log 2>&1; tail -4 gpurun_out/r02r_tests.log
B="python bench.py --steps 10 --warmup 3 --sweeps-per-step 4 --no-cpu-baseline --no-e2e --no-uncached --no-self-check"
for ch in 512 1024 1536 2048 4096; do
  for m in 0 2 3; do timeout 120 $B --chains $ch --sweep-mode $m > gpurun_out/r02r_b${ch}_m$m.json 2>gpurun_out/r02r.err; done
  for m in 1 3; do timeout 120 $B --chains $ch --sweep-mode $m --eager-noise > gpurun_out/r02r_b${ch}_m${m}e.json 2>gpurun_out/r02r.err; done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02r_b*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_sweep'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()}, round(d['roofline']['frac'],3), '%.3g'%d['value'])
    except Exception as e: print(f,'ERR',e)
PY
