mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/s18_tests.log 2>&1; tail -3 gpurun_out/s18_tests.log
python bench.py --no-cpu > gpurun_out/s18_bench.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/s18_bench.log'):
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['ms_per_step'], d['neighbor_rebuild_ms'], d['roofline']['kernel_ms'], d['config']['rebuilds_in_timed_region'], d['e2e']['value'])
PY
