timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02v_tests.log 2>&1; tail -4 gpurun_out/r02v_tests.log
for cfg in c1 c2 c4 c5; do
  timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02v_bench_$cfg.json 2> gpurun_out/r02v_bench_$cfg.err
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02v_bench_reference.json 2> gpurun_out/r02v_bench_reference.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02v_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.3g'%d['value'], d.get('ms_per_step'), {k:round(v,3) for k,v in d.get('kernel_ms',{}).items()}, d.get('roofline',{}).get('kernel'), d.get('roofline',{}).get('frac'))
    except Exception as e: print(f,'ERR',e)
PY
