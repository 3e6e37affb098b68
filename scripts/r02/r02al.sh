# evidence run after the regrouped guiding cache: full GPU test suite, default bench line, launch list + full ncu capture
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02al_tests.log 2>&1; tail -3 gpurun_out/r02al_tests.log
timeout 900 python bench.py > gpurun_out/r02al_bench_default.json 2> gpurun_out/r02al_bench_default.err; tail -c 900 gpurun_out/r02al_bench_default.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02al.csv python bench.py --steps 2 --warmup 1 --sweeps-per-step 2 --no-cpu-baseline --no-e2e --no-uncached --no-self-check > gpurun_out/r02al_ncu_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'sweep_pipe|cache_apply_kernel' -s 12 -c 4 -o gpurun_out/prof_r02al -f python bench.py --steps 2 --warmup 1 --sweeps-per-step 4 --no-cpu-baseline --no-e2e --no-uncached --no-self-check > gpurun_out/r02al_ncu_full.log 2>&1
ls -la gpurun_out/prof_r02al.ncu-rep
