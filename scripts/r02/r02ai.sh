# grouped F0/Psi layout (FPG = 4 tiles per lane line): cache tests, then the default workload's kernel times
python -m pytest tests -m gpu -x -q -k "cache or full_size" > gpurun_out/r02ai_tests.log 2>&1; tail -3 gpurun_out/r02ai_tests.log
B="python bench.py --steps 10 --warmup 3 --sweeps-per-step 4 --no-cpu-baseline --no-e2e --no-uncached --no-self-check"
timeout 300 $B > gpurun_out/r02ai_b4096.json 2> gpurun_out/r02ai.err
timeout 300 $B --chains 512 > gpurun_out/r02ai_b512.json 2>> gpurun_out/r02ai.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02ai_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_sweep'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()}, d['roofline']['kernel'], '%.4g'%d['value'])
    except Exception as e: print(f,'ERR',e)
PY
