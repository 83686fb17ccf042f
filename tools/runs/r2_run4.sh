TAG=${1:-r02d}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_neighbor.py tests/test_gpu_energy.py tests/test_gpu_baseline_configs.py tests/test_gpu_fullsize.py tests/test_oracle_c.py -m gpu -q -k "not 1e4") > gpurun_out/${TAG}_tests.log 2>&1
tail -5 gpurun_out/${TAG}_tests.log
(time python bench.py --steps 20 --warmup 5 --no-cpu --no-variants) > gpurun_out/${TAG}_bench20.log 2>&1
(time python bench.py --steps 1000 --warmup 500 --no-cpu --no-variants) > gpurun_out/${TAG}_bench1000.log 2>&1
python - "$TAG" <<'PY'
import json, sys
for name in ('bench20', 'bench1000'):
  for l in open('gpurun_out/%s_%s.log' % (sys.argv[1], name)):
    try: d = json.loads(l)
    except Exception: continue
    print(name, 'value %.4g  ms/step %.4f  rebuild %.3f ms  force %.4f ms  step_frac %.3f  e2e %.4g rebuilds %d' % (
        d['value'], d['ms_per_step'], d['neighbor_rebuild_ms'], d['roofline']['kernel_ms'],
        d['roofline']['step_frac'], d['e2e']['value'], d['config']['rebuilds_in_timed_region']))
PY
CMD="python bench.py --steps 40 --warmup 10 --no-cpu --kernel-reps 2 --no-variants"
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_nbr_cell_test -s 1 -c 1 \
    -f -o gpurun_out/${TAG}_celltest $CMD > gpurun_out/${TAG}_celltest.log 2>&1
ls -la gpurun_out/${TAG}_*
