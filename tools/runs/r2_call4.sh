# round 2, call 4: new parity tests (all, no -x), launch list of config 4 (SW NVT)
TAG=${1:-r02c4}
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_periodic_general.py tests/test_gpu_widening.py tests/test_gpu_npt.py -q -m gpu --durations=5) > gpurun_out/${TAG}_newtests.log 2>&1
tail -5 gpurun_out/${TAG}_newtests.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_c4_launches.csv \
  python -c "import sys; sys.path.insert(0,'benchmarks'); import configs; configs.c4(20, 5)" > gpurun_out/${TAG}_c4_launches.log 2>&1
python profiles/launch_summary.py gpurun_out/${TAG}_c4_launches.csv > gpurun_out/${TAG}_c4_launches.txt 2>&1
head -30 gpurun_out/${TAG}_c4_launches.txt
