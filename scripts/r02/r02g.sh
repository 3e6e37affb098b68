python -m pytest tests/test_gpu_sweep_pipeline.py tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or lazy or tsit5" > gpurun_out/r02g_tests.log 2>&1; tail -3 gpurun_out/r02g_tests.log
for ch in 512 1024 2048 4096; do for m in 2 3; do
python bench.py --chains $ch --steps 10 --warmup 3 --sweeps-per-step 4 --no-cpu-baseline --no-e2e --no-uncached --no-self-check --sweep-mode $m > gpurun_out/r02g_b${ch}_m$m.json 2>gpurun_out/r02g.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02g_b*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_sweep'], {k:round(v,3) for k,v in d['kernel_ms'].items()}, round(d['roofline']['frac'],3), '%.3g'%d['value'])
    except Exception as e: print(f,'ERR',e)
PY
