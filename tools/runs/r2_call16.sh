TAG=${1:-r02c16}
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_npt.py -q -m gpu --durations=3) > gpurun_out/${TAG}_tests.log 2>&1
tail -12 gpurun_out/${TAG}_tests.log
