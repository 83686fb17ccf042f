mkdir -p gpurun_out
(time python bench.py) > gpurun_out/s20_bench.log 2>&1
(time python bench.py --impl reference --steps 20 --warmup 3) > gpurun_out/s20_ref.log 2>&1
(time python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/s20_smoke.log 2>&1
tail -4 gpurun_out/s20_smoke.log
python - <<'PY'
import json
for f in ['gpurun_out/s20_bench.log','gpurun_out/s20_ref.log']:
  for l in open(f):
    try: d=json.loads(l)
    except Exception: print(l[:200].rstrip()); continue
    print(d.get('impl','b200'), d['value'], d['ms_per_step'], d.get('neighbor_rebuild_ms'), (d.get('roofline') or {}).get('kernel_ms'), d['e2e']['value'], d.get('cpu_baseline'))
PY
