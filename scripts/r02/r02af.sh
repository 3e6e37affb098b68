timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02af_tests.log 2>&1; tail -3 gpurun_out/r02af_tests.log
timeout 900 python bench.py > gpurun_out/r02af_bench_default.json 2> gpurun_out/r02af_bench_default.err; python -c "
import json; d=json.loads(open('gpurun_out/r02af_bench_default.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['ms_per_sweep'], d['roofline']['kernel'], d['roofline']['frac'], d['clocks'])"
python __graft_entry__.py --smoke 2>&1 | tail -1
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r02af_bench_20_5.json 2>gpurun_out/r02af_bench_20_5.err ) 2>&1 | grep real
