mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/s21_tests.log 2>&1; tail -3 gpurun_out/s21_tests.log
(time python bench.py --no-cpu) > gpurun_out/s21_bench.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/s21_bench.log'):
    try: d=json.loads(l)
    except Exception: print(l[:300].rstrip()); continue
    print(d['ms_per_step'], d['neighbor_rebuild_ms'], d['roofline']['kernel_ms'], d['config']['rebuilds_in_timed_region'], d['e2e']['value'], d['variants'])
PY
