mkdir -p gpurun_out
rm -f gpurun_out/s16_var.log
for l in eager graph; do
  echo "== loop $l" >> gpurun_out/s16_var.log
  python bench.py --no-cpu --loop $l >> gpurun_out/s16_var.log 2>&1
done
python - <<'PY'
import json
for l in open('gpurun_out/s16_var.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['ms_per_step'], d['neighbor_rebuild_ms'], d['roofline']['kernel_ms'], d['config']['rebuilds_in_timed_region'], d['e2e']['value'])
PY
