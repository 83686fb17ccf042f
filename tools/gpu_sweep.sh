mkdir -p gpurun_out
rm -f gpurun_out/s28_var.log
for v in "" _m4; do
  echo "== lib '$v'" >> gpurun_out/s28_var.log
  JMD_B200_LIB=$PWD/jax_md_b200/libjmd_b200$v.so python bench.py --no-cpu --no-variants --steps 400 --warmup 300 >> gpurun_out/s28_var.log 2>&1
done
python - <<'PY'
import json
for l in open('gpurun_out/s28_var.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l[:300].rstrip()); continue
    print(d['ms_per_step'], d['neighbor_rebuild_ms'], d['roofline']['kernel_ms'], d['config']['rebuilds_in_timed_region'])
PY
