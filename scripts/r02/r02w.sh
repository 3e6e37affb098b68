export DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_lz.so
timeout 300 python -m pytest tests/test_gpu_sweep_pipeline.py tests/test_gpu_abi_c.py tests/test_gpu_host_api.py -m gpu -x -q -k "lorenz or abi or c_program or history or views" 2>&1 | tail -3
B="python bench.py --steps 10 --warmup 3 --sweeps-per-step 4 --no-cpu-baseline --no-e2e --no-uncached --no-self-check"
for ch in 256 512 768 896 1024 1536; do
  for m in 3 4; do timeout 120 $B --chains $ch --sweep-mode $m > gpurun_out/r02w_b${ch}_m$m.json 2>gpurun_out/r02w.err; done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02w_b*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_sweep'],3), round(d['kernel_ms']['sweep_fused'],3), d['roofline']['kernel'])
    except Exception as e: print(f,'ERR',e)
PY
