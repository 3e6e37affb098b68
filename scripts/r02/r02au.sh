# final evidence of the round (step-parallel kernel with branch-free selects): full GPU test suite, smoke, default bench line, shard sizes,
# ncu --set full of the step-parallel kernel at the 8-GPU shard size
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02au_tests.log 2>&1; tail -3 gpurun_out/r02au_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02au_bench_default.json 2> gpurun_out/r02au_bench_default.err
B="python bench.py --no-cpu-baseline --no-uncached"
for ch in 1024 512; do timeout 300 $B --chains $ch > gpurun_out/r02au_bench_c3_${ch}.json 2>gpurun_out/r02au.err || tail -5 gpurun_out/r02au.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02au_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_sweep'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()}, d['roofline']['kernel'], '%.4g'%d['value'], '%.4g'%d['e2e']['value'], round(d['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
PY
timeout 300 ncu --set full --import-source on --clock-control none -k regex:sweep_sp -s 4 -c 1 -o gpurun_out/prof_r02au -f python bench.py --steps 2 --warmup 2 --sweeps-per-step 2 --no-cpu-baseline --no-e2e --no-uncached --no-self-check --chains 512 > gpurun_out/r02au_ncu.log 2>&1
ls -la gpurun_out/prof_r02au.ncu-rep
