B="python bench.py --steps 10 --warmup 3 --sweeps-per-step 4 --no-cpu-baseline --no-e2e --no-uncached --no-self-check"
for v in wca wcb wcc; do
export DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_$v.so
timeout 100 python -m pytest tests/test_gpu_sweep_pipeline.py -m gpu -x -q -k "lorenz and ws_compact" 2>&1 | tail -1
for ch in 512 896; do timeout 120 $B --chains $ch --sweep-mode 4 > gpurun_out/r02aa_${v}_b${ch}.json 2>gpurun_out/r02aa.err; done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02aa_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_sweep'],3), round(d['kernel_ms']['sweep_fused'],3), d['roofline']['kernel'])
    except Exception as e: print(f,'ERR',e)
PY
