TAG=${1:-r02c14}
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_npt.py tests/test_gpu_periodic_general.py tests/test_gpu_triclinic.py tests/test_gpu_widening.py tests/test_gpu_energy.py tests/test_gpu_fullsize.py tests/test_gpu_baseline_configs.py -q -m gpu -k "not 1e4" --durations=6) > gpurun_out/${TAG}_tests.log 2>&1
tail -16 gpurun_out/${TAG}_tests.log
