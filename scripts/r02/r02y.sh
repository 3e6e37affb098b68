# 8-GPU box: multi-rank correctness under pytest, then the strong-scaling bench at N = 8 and N = 4
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02y_multi.log 2>&1; tail -3 gpurun_out/r02y_multi.log
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-uncached > gpurun_out/r02y_bench_n$n.json 2> gpurun_out/r02y_bench_n$n.err
tail -c 1500 gpurun_out/r02y_bench_n$n.json | head -c 400; echo
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02y_bench_n*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['n_gpus'], '%.4g'%d['value'], d['scaling'], round(d.get('ms_per_sweep',0),3), {k:round(v,3) for k,v in d.get('kernel_ms',{}).items()}, d.get('roofline',{}).get('kernel'), 'e2e %.4g'%d.get('e2e',{}).get('value',0), d.get('self_check',{}).get('allreduce_ok'), d.get('weak',{}))
    except Exception as e: print(f,'ERR',e)
PY
