mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_energy.py -m gpu -x -q -k "generic") > gpurun_out/s22_tests.log 2>&1; tail -3 gpurun_out/s22_tests.log
