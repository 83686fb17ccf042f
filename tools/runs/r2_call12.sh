TAG=${1:-r02c12}
mkdir -p gpurun_out
python tools/probe_build.py > gpurun_out/${TAG}_build.log 2>&1
tail -8 gpurun_out/${TAG}_build.log
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-extra --no-variants > gpurun_out/${TAG}_bench1.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/${TAG}_bench2.log 2>&1
python - "$TAG" <<'PY'
import json, sys
for name in ('bench1', 'bench2'):
  lines = open('gpurun_out/%s_%s.log' % (sys.argv[1], name)).read().splitlines()
  ok = False
  for l in lines:
    try: d = json.loads(l)
    except Exception: continue
    if not isinstance(d, dict) or 'value' not in d: continue
    ok = True
    print(name, 'value %.4g ms/step %.4f e2e %.4g rebuilds %s untimed %s rebuild_ms %s' % (d['value'], d['ms_per_step'], d['e2e']['value'],
          d['config'].get('rebuilds_in_timed_region'), d['config'].get('untimed_steps_before_window'), d.get('neighbor_rebuild_ms')))
  if not ok: print(name, 'NO JSON', lines[-10:])
PY
(time python -m pytest tests/test_gpu_triclinic.py tests/test_gpu_neighbor.py tests/test_gpu_periodic_general.py -q -m gpu) > gpurun_out/${TAG}_tests.log 2>&1
tail -4 gpurun_out/${TAG}_tests.log
