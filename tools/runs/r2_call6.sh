# round 2, call 6: triclinic boxes + reciprocal SW triplets: parity tests, headline bench (unchanged?), config 4
TAG=${1:-r02c6}
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_triclinic.py tests/test_gpu_periodic_general.py tests/test_gpu_npt.py tests/test_gpu_energy.py tests/test_gpu_fullsize.py -q -m gpu --durations=5) > gpurun_out/${TAG}_tests.log 2>&1
tail -15 gpurun_out/${TAG}_tests.log
(time python -m pytest tests/test_gpu_baseline_configs.py -q -m gpu -k "sw or neighbor_sets") > gpurun_out/${TAG}_tests2.log 2>&1
tail -3 gpurun_out/${TAG}_tests2.log
python bench.py --steps 20 --warmup 5 --no-cpu --no-extra --no-variants > gpurun_out/${TAG}_bench20.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/${TAG}_bench20.log'):
  try: d = json.loads(l)
  except Exception: continue
  print('bench20', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['neighbor_rebuild_ms'], d['roofline']['traffic'])
PY
python -c "
import sys; sys.path.insert(0,'benchmarks'); import configs
configs.c4(100, 50)" > gpurun_out/${TAG}_c4.log 2>&1
tail -1 gpurun_out/${TAG}_c4.log | cut -c1-200
