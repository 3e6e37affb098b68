# final evidence of the round: full GPU test suite, default bench line, the 2 / 4 / 8-GPU shard sizes on one GPU, reference arm
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02ar_tests.log 2>&1; tail -3 gpurun_out/r02ar_tests.log
timeout 900 python bench.py > gpurun_out/r02ar_bench_default.json 2> gpurun_out/r02ar_bench_default.err; tail -c 300 gpurun_out/r02ar_bench_default.json
B="python bench.py --no-cpu-baseline --no-uncached"
for ch in 2048 1024 512; do timeout 300 $B --chains $ch > gpurun_out/r02ar_bench_c3_${ch}.json 2>gpurun_out/r02ar.err || tail -5 gpurun_out/r02ar.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02ar_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_sweep'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()}, d['roofline']['kernel'], '%.4g'%d['value'], '%.4g'%d['e2e']['value'], round(d['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
PY
