mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/s5_tests.log 2>&1; tail -3 gpurun_out/s5_tests.log
rm -f gpurun_out/s5_var.log
for b in 0 1 2; do for v in "" _wrap; do
  echo "== brick $b variant '$v'" >> gpurun_out/s5_var.log
  JMD_BRICK_SHIFT=$b JMD_B200_LIB=$PWD/jax_md_b200/libjmd_b200$v.so python bench.py --no-cpu --steps 400 --warmup 300 >> gpurun_out/s5_var.log 2>&1
done; done
python - <<'PY'
import json
for l in open('gpurun_out/s5_var.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['ms_per_step'], d['neighbor_rebuild_ms'], d['roofline']['kernel_ms'], d['config']['rebuilds_in_timed_region'])
PY
