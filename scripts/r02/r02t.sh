export DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_lz.so
B="python bench.py --steps 10 --warmup 3 --sweeps-per-step 4 --no-cpu-baseline --no-e2e --no-uncached --no-self-check"
for ch in 512 768 1024 1280 1536; do
  timeout 120 $B --chains $ch --sweep-mode 2 --fwd-lanes 1 > gpurun_out/r02t_b${ch}_pipe1.json 2>gpurun_out/r02t.err
  for l in 1 2 4; do timeout 120 $B --chains $ch --sweep-mode 1 --fwd-lanes $l > gpurun_out/r02t_b${ch}_cl$l.json 2>gpurun_out/r02t.err; done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02t_b*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_sweep'],3), round(d['kernel_ms']['sweep_fused'],3), d['roofline']['kernel'])
    except Exception as e: print(f,'ERR',e)
PY
