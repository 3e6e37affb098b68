timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02x_tests.log 2>&1; tail -4 gpurun_out/r02x_tests.log
timeout 900 python bench.py > gpurun_out/r02x_bench_default.json 2> gpurun_out/r02x_bench_default.err; tail -c 300 gpurun_out/r02x_bench_default.json
timeout 600 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02x_bench_c4.json 2> gpurun_out/r02x_bench_c4.err
for ch in 2048 1024 512; do timeout 300 python bench.py --chains $ch --steps 10 --warmup 3 --no-cpu-baseline --no-uncached --no-self-check > gpurun_out/r02x_bench_c3_${ch}.json 2>gpurun_out/r02x.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02x_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], round(d.get('ms_per_sweep',0),3), {k:round(v,3) for k,v in d.get('kernel_ms',{}).items()}, d.get('roofline',{}).get('kernel'), round(d.get('roofline',{}).get('frac') or 0,3), 'e2e %.4g'%d.get('e2e',{}).get('value',0))
    except Exception as e: print(f,'ERR',e)
PY
python __graft_entry__.py --smoke 2>&1 | tail -2
