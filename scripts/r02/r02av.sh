# two ranks x 512 chains (the step-parallel kernel under torch.distributed.run: peer-memory all-reduce, self-check with the oracle replay)
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --chains 512 --steps 10 --warmup 3 --sweeps-per-step 4 --no-cpu-baseline --no-uncached --no-weak > gpurun_out/r02av_bench_2x512.json 2> gpurun_out/r02av.err || tail -5 gpurun_out/r02av.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02av_bench_2x512.json').read().strip().splitlines()[-1])
print(d['n_gpus'], '%.4g'%d['value'], d['scaling'], round(d['ms_per_sweep'],3), d['roofline']['kernel'], 'e2e %.4g'%d['e2e']['value'], d.get('self_check'))
PY
