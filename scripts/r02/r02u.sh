# round-2 evidence run: full GPU test suite, default bench line, launch list + full ncu capture of the default command
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02u_tests.log 2>&1; tail -3 gpurun_out/r02u_tests.log
timeout 900 python bench.py > gpurun_out/r02u_bench_default.json 2> gpurun_out/r02u_bench_default.err; tail -c 600 gpurun_out/r02u_bench_default.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02u_launches.csv python bench.py --steps 2 --warmup 1 --sweeps-per-step 2 --no-cpu-baseline --no-e2e --no-uncached --no-self-check > gpurun_out/r02u_ncu_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sweep_pipe -s 6 -c 2 -o gpurun_out/prof_r02u_pipe -f python bench.py --steps 2 --warmup 1 --sweeps-per-step 4 --no-cpu-baseline --no-e2e --no-uncached --no-self-check > gpurun_out/r02u_ncu_full.log 2>&1
ls -la gpurun_out/prof_r02u_pipe.ncu-rep
