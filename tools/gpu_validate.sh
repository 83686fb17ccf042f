# One-call GPU validation (run under gpurun): parity tests, smoke, default bench.
#   gpurun --timeout 900 -- 'bash tools/gpu_validate.sh <tag>'
TAG=${1:-validate}
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time python bench.py) > gpurun_out/${TAG}_bench.log 2>&1
python - "$TAG" <<'PY'
import json, sys
for l in open('gpurun_out/%s_bench.log' % sys.argv[1]):
    try: d = json.loads(l)
    except Exception: continue
    print('value %.4g  ms/step %.4f  rebuild %.3f ms  force %.4f ms  step_frac %.3f  e2e %.4g  lazy %.4f ms' % (
        d['value'], d['ms_per_step'], d['neighbor_rebuild_ms'], d['roofline']['kernel_ms'],
        d['roofline']['step_frac'], d['e2e']['value'], d['variants']['lazy_idx']['ms_per_step']))
PY
