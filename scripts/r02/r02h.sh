# X-sector over-fetch experiment: evict-first off / 256-bit load split in two, Lorenz-only builds
B="python bench.py --chains 4096 --steps 10 --warmup 3 --sweeps-per-step 4 --no-cpu-baseline --no-e2e --no-uncached --no-self-check"
for v in lz lzef0 lzsplit; do
  DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_$v.so $B > gpurun_out/r02h_$v.json 2> gpurun_out/r02h_$v.err
  DMT_LIB=$PWD/diffusionmcmctools.jl_b200/libdmt_$v.so ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:sweep_pipe -c 2 --csv --log-file gpurun_out/r02h_ncu_$v.csv $B > /dev/null 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02h_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_sweep'], {k:round(v,3) for k,v in d['kernel_ms'].items()}, round(d['roofline']['frac'],3), '%.3g'%d['value'])
    except Exception as e: print(f,'ERR',e)
PY
grep -h "sweep_pipe" gpurun_out/r02h_ncu_*.csv | cut -d, -f1,5,12- | head -30
