TAG=${1:-r02c7}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-extra --no-variants > gpurun_out/${TAG}_bench1.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/${TAG}_bench2.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/${TAG}_bench2b.log 2>&1
python - "$TAG" <<'PY'
import json, sys
for name in ('bench1', 'bench2', 'bench2b'):
  lines = open('gpurun_out/%s_%s.log' % (sys.argv[1], name)).read().splitlines()
  ok = False
  for l in lines:
    try: d = json.loads(l)
    except Exception: continue
    if not isinstance(d, dict) or 'value' not in d: continue
    ok = True
    print(name, 'value %.4g ms/step %.4f e2e %.4g rebuilds %s untimed %s' % (d['value'], d['ms_per_step'], d['e2e']['value'],
          d['config'].get('rebuilds_in_timed_region'), d['config'].get('untimed_steps_before_window')))
  if not ok: print(name, 'NO JSON', lines[-10:])
PY
