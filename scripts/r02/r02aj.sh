# grouped F0/Psi layout, FPG lanes per (parameter set, block): FPG = 4 (lz) and 2 (lz2), Lorenz-only builds
B="python bench.py --steps 10 --warmup 3 --sweeps-per-step 4 --no-cpu-baseline --no-e2e --no-uncached --no-self-check"
for v in lz; do
  export DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_$v.so
  timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_guiding_cache_matches_uncached_sweeps and lorenz" > gpurun_out/r02aj_tests_$v.log 2>&1; tail -1 gpurun_out/r02aj_tests_$v.log
  timeout 200 $B > gpurun_out/r02aj_${v}_b4096.json 2> gpurun_out/r02aj.err
  timeout 200 $B --chains 512 > gpurun_out/r02aj_${v}_b512.json 2>> gpurun_out/r02aj.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02aj_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_sweep'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()}, d['roofline']['kernel'], '%.4g'%d['value'])
    except Exception as e: print(f,'ERR',e)
PY
