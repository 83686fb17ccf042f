mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/s12_tests.log 2>&1; tail -3 gpurun_out/s12_tests.log
rm -f gpurun_out/s12_var.log
for v in "" _m5 _m6 _u2 _u2m5; do
  echo "== lib '$v'" >> gpurun_out/s12_var.log
  JMD_B200_LIB=$PWD/jax_md_b200/libjmd_b200$v.so python bench.py --no-cpu --steps 300 --warmup 200 >> gpurun_out/s12_var.log 2>&1
done
python - <<'PY'
import json
for l in open('gpurun_out/s12_var.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['ms_per_step'], d['neighbor_rebuild_ms'], d['roofline']['kernel_ms'], d['config']['rebuilds_in_timed_region'])
PY
