TAG=${1:-r02c17}
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_energy.py tests/test_gpu_widening.py tests/test_gpu_baseline_configs.py -q -m gpu -k "not 1e4" --durations=4) > gpurun_out/${TAG}_tests.log 2>&1
tail -12 gpurun_out/${TAG}_tests.log
