# round 2: 8-GPU validation.  DD parity tests on real ranks, the scaling series 1/2/4/8 in the driver's
# shape (20 steps, 5 warm-up), the long 8-GPU run, the 32M-atom and strong-scaling extras at N=8.
TAG=${1:-r02dd8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.log
if [ -z "$SKIP_TESTS" ]; then (time timeout 900 python -m pytest tests/test_domain.py -m gpu -q) > gpurun_out/${TAG}_tests.log 2>&1; fi
[ -z "$SKIP_TESTS" ] && tail -3 gpurun_out/${TAG}_tests.log
run() {  # name nproc flags...
  local name=$1 n=$2; shift 2
  if [ "$n" = 1 ]; then
    (time timeout 900 python bench.py --gpus 1 "$@") > gpurun_out/${TAG}_$name.log 2>&1
  else
    (time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$n bench.py --gpus $n "$@") > gpurun_out/${TAG}_$name.log 2>&1
  fi
}
run bench8 8 --steps 20 --warmup 5 --no-cpu
run bench8_long 8 --steps 200 --warmup 50 --no-cpu --no-extra
run bench4 4 --steps 20 --warmup 5 --no-cpu --no-extra
run bench2 2 --steps 20 --warmup 5 --no-cpu --no-extra
run bench1 1 --steps 20 --warmup 5 --no-cpu --no-extra --no-variants
python - "$TAG" <<'PY'
import json, sys
for name in ('bench1', 'bench2', 'bench4', 'bench8', 'bench8_long'):
  try: lines = open('gpurun_out/%s_%s.log' % (sys.argv[1], name)).read().splitlines()
  except Exception as e: print(name, e); continue
  ok = False
  for l in lines:
    try: d = json.loads(l)
    except Exception: continue
    ok = True
    print(name, 'value %.4g  ms/step %.4f  force %.4f ms  e2e %.4g rebuilds %s' % (
        d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['config'].get('rebuilds_in_timed_region')),
        {k: (round(v['value'] / 1e9, 3), round(v['ms_per_step'], 4)) for k, v in d.items() if isinstance(v, dict) and 'atoms' in v})
  if not ok: print(name, 'NO JSON', lines[-12:])
PY
