export DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_w6.so
timeout 150 python -m pytest tests/test_gpu_sweep_pipeline.py -m gpu -x -q -k "lorenz and (ws or lazy)" 2>&1 | tail -2
for ch in 512 1024 2048 4096; do
timeout 120 python bench.py --chains $ch --steps 10 --warmup 3 --sweeps-per-step 4 --no-cpu-baseline --no-e2e --no-uncached --no-self-check --sweep-mode 3 > gpurun_out/r02q_b${ch}_m3.json 2>gpurun_out/r02q.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02q_b*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_sweep'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()}, round(d['roofline']['frac'],3), '%.3g'%d['value'])
    except Exception as e: print(f,'ERR',e)
PY
export DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_w6t.so
timeout 120 python bench.py --chains 512 --steps 1 --warmup 3 --sweeps-per-step 1 --no-cpu-baseline --no-e2e --no-uncached --no-self-check --sweep-mode 3 > gpurun_out/r02p_trace.txt 2>&1
grep -c "^TR" gpurun_out/r02p_trace.txt
