timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "jr or jansen" 2>&1 | tail -3
timeout 600 python bench.py --config c5 --steps 4 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r02ae_bench_c5.json 2> gpurun_out/r02ae_bench_c5.err
DMT_K1_DENSE=1 timeout 600 python bench.py --config c5 --steps 4 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r02ae_bench_c5_dense.json 2> gpurun_out/r02ae_bench_c5_dense.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02ae_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms'].items()})
    except Exception as e: print(f,'ERR',e)
PY
