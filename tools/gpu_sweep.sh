mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/s29_tests.log 2>&1; tail -3 gpurun_out/s29_tests.log
(time python bench.py) > gpurun_out/s29_bench.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/s29_bench.log'):
    try: d=json.loads(l)
    except Exception: print(l[:200].rstrip()); continue
    print(d['value'], d['ms_per_step'], d['neighbor_rebuild_ms'], d['roofline']['kernel_ms'], d['roofline']['step_frac'], d['e2e']['value'], d['variants']['lazy_idx']['ms_per_step'])
PY
bash tools/profile.sh r01s5 > /dev/null 2>&1
ls gpurun_out/r01s5* | head
