TAG=${1:-r02h}
N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.log
(time timeout 600 python -m pytest tests/test_domain.py -m gpu -q -x) > gpurun_out/${TAG}_tests.log 2>&1
tail -5 gpurun_out/${TAG}_tests.log
(time python bench.py --steps 20 --warmup 5 --no-cpu --no-variants) > gpurun_out/${TAG}_bench1.log 2>&1
for n in $N; do
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5) > gpurun_out/${TAG}_bench$n.log 2>&1
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 200 --warmup 50) > gpurun_out/${TAG}_bench${n}_long.log 2>&1
done
python - "$TAG" "$N" <<'PY'
import json, sys
names = ['bench1'] + ['bench%s%s' % (n, s) for n in sys.argv[2].split() for s in ('', '_long')]
for name in names:
  try: lines = open('gpurun_out/%s_%s.log' % (sys.argv[1], name)).read().splitlines()
  except Exception as e: print(name, e); continue
  ok = False
  for l in lines:
    try: d = json.loads(l)
    except Exception: continue
    ok = True
    print(name, 'value %.4g  ms/step %.4f  force %.4f ms  e2e %.4g rebuilds %s' % (
        d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['config'].get('rebuilds_in_timed_region')))
  if not ok: print(name, 'NO JSON', lines[-8:])
PY
