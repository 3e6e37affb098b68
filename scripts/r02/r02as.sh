# step-parallel sweep kernel, branch-free selects (v2 + sel4): parity, times
export DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_lz.so
timeout 300 python -m pytest tests/test_gpu_sweep_pipeline.py -m gpu -x -q -k "lorenz and (sp or 5)" > gpurun_out/r02as_tests.log 2>&1; tail -12 gpurun_out/r02as_tests.log
B="python bench.py --steps 10 --warmup 3 --sweeps-per-step 4 --no-cpu-baseline --no-e2e --no-uncached --no-self-check"
for ch in 512 1024 1280; do timeout 200 $B --chains $ch --sweep-mode 5 > gpurun_out/r02as_sp_b${ch}.json 2>gpurun_out/r02as.err || tail -5 gpurun_out/r02as.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02as_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_sweep'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()}, d['roofline']['kernel'], '%.4g'%d['value'])
    except Exception as e: print(f,'ERR',e)
PY
