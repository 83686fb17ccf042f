TAG=${1:-r02c8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/probe_dd_steps.py 63 80 > gpurun_out/${TAG}_probe.log 2>&1
grep -E "^step|overflow" gpurun_out/${TAG}_probe.log | awk '{print}' | head -90
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 --unroll 1000 --no-cpu --no-extra > gpurun_out/${TAG}_bench2_old.log 2>&1
python - "$TAG" <<'PY'
import json, sys
for name in ('bench2_old',):
  for l in open('gpurun_out/%s_%s.log' % (sys.argv[1], name)):
    try: d = json.loads(l)
    except Exception: continue
    if isinstance(d, dict) and 'value' in d:
      print(name, d['value'], d['ms_per_step'], d['config'].get('rebuilds_in_timed_region'), d['config'].get('untimed_steps_before_window'))
PY
