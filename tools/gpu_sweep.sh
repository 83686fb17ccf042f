mkdir -p gpurun_out
rm -f gpurun_out/s13_var.log
for v in "" _u2m6 _u3m5 _u4m5 _u4m4; do
  echo "== lib '$v'" >> gpurun_out/s13_var.log
  JMD_B200_LIB=$PWD/jax_md_b200/libjmd_b200$v.so python bench.py --no-cpu --steps 300 --warmup 200 >> gpurun_out/s13_var.log 2>&1
done
python - <<'PY'
import json
for l in open('gpurun_out/s13_var.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['ms_per_step'], d['neighbor_rebuild_ms'], d['roofline']['kernel_ms'], d['config']['rebuilds_in_timed_region'])
PY
