mkdir -p gpurun_out
(JMD_B200_LIB=$PWD/jax_md_b200/libjmd_b200_tma.so timeout 600 python -m pytest tests/test_gpu_energy.py tests/test_gpu_fullsize.py -m gpu -x -q) > gpurun_out/s26_tests.log 2>&1; tail -3 gpurun_out/s26_tests.log
rm -f gpurun_out/s26_var.log
for v in "" _tma _tma8; do
  echo "== lib '$v'" >> gpurun_out/s26_var.log
  JMD_B200_LIB=$PWD/jax_md_b200/libjmd_b200$v.so timeout 300 python bench.py --no-cpu --no-variants --steps 400 --warmup 300 >> gpurun_out/s26_var.log 2>&1
done
python - <<'PY'
import json
for l in open('gpurun_out/s26_var.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l[:300].rstrip()); continue
    print(d['ms_per_step'], d['neighbor_rebuild_ms'], d['roofline']['kernel_ms'], d['config']['rebuilds_in_timed_region'])
PY
