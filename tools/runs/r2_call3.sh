# round 2, call 3: new parity tests, bench repeatability (driver-shaped), the configs that failed, missing ncu captures
TAG=${1:-r02c3}
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_periodic_general.py tests/test_gpu_widening.py tests/test_gpu_npt.py -q -m gpu -x --durations=5) > gpurun_out/${TAG}_newtests.log 2>&1
tail -5 gpurun_out/${TAG}_newtests.log
for i in 1 2 3; do
  python bench.py --steps 20 --warmup 5 --no-cpu --no-extra --no-variants > gpurun_out/${TAG}_bench20_$i.log 2>&1
  python - <<PY
import json
for l in open('gpurun_out/${TAG}_bench20_$i.log'):
  try: d = json.loads(l)
  except Exception: continue
  print('bench20 #$i', d['value'], d['ms_per_step'], d['config']['rebuilds_in_timed_region'], d['clocks'])
PY
done
(time python bench.py --steps 20 --warmup 5) > gpurun_out/${TAG}_bench20.log 2>&1
(time python benchmarks/configs.py --quick) > gpurun_out/${TAG}_configs.log 2>&1
grep -c '"config"' gpurun_out/${TAG}_configs.log
prof() {   # name regex skip cmd...
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 \
      -f -o gpurun_out/${TAG}_$name "$@" > gpurun_out/${TAG}_$name.log 2>&1
  python profiles/ncu_summary.py gpurun_out/${TAG}_$name.ncu-rep > gpurun_out/${TAG}_ncu_$name.txt 2>&1
  python tools/ncu_lines.py gpurun_out/${TAG}_$name.ncu-rep 1.5 >> gpurun_out/${TAG}_ncu_$name.txt 2>&1
  rm -f gpurun_out/${TAG}_$name.ncu-rep
}
CMD="python bench.py --steps 40 --warmup 10 --no-cpu --kernel-reps 2 --no-variants --no-extra"
prof update '^k_update$' 3 $CMD
prof sw '^k_sw$' 2 python -c "import sys; sys.path.insert(0,'benchmarks'); import configs; configs.c4(20, 5)"
ls -la gpurun_out/${TAG}_* | awk '{print $5, $9}'
