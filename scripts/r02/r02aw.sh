# sweep_pipe_kernel with the TMA issue behind elect.sync (no per-copy elect loops): parity (all models), oracle replay at 4096 chains, bench
timeout 120 python -m pytest tests/test_gpu_sweep_pipeline.py -m gpu -x -q > gpurun_out/r02aw_tests.log 2>&1; tail -2 gpurun_out/r02aw_tests.log
timeout 120 python -m pytest tests/test_gpu_full_size.py -m gpu -x -q -k "bench_configuration and 4096" > gpurun_out/r02aw_tests2.log 2>&1; tail -2 gpurun_out/r02aw_tests2.log
timeout 100 python bench.py --no-cpu-baseline --no-uncached > gpurun_out/r02aw_bench_default.json 2> gpurun_out/r02aw.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02aw_bench_default.json').read().strip().splitlines()[-1]); print(round(d['ms_per_sweep'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()}, d['roofline']['kernel'], '%.4g'%d['value'], '%.4g'%d['e2e']['value'], round(d['roofline']['frac'],3), d['clocks'])
PY
