# step-parallel kernel for every model (full library): the sweep-pipeline tests and the bench-configuration replays
timeout 600 python -m pytest tests/test_gpu_sweep_pipeline.py -m gpu -x -q > gpurun_out/r02ap_tests.log 2>&1; tail -3 gpurun_out/r02ap_tests.log
timeout 600 python -m pytest tests/test_gpu_full_size.py -m gpu -x -q -k "bench_configuration" > gpurun_out/r02ap_tests2.log 2>&1; tail -3 gpurun_out/r02ap_tests2.log
