# round 2: final single-GPU validation of HEAD: build check, full GPU suite, smoke, the driver's two bench commands,
# the long bench, all BASELINE configs, launch list + ncu of the kernels that changed late (SW, Dense scan).
TAG=${1:-r02f}
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --durations=8) > gpurun_out/${TAG}_tests.log 2>&1
tail -14 gpurun_out/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/${TAG}_bench20.log 2>&1
(time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5) > gpurun_out/${TAG}_ref.log 2>&1
(time python bench.py --no-cpu --no-extra) > gpurun_out/${TAG}_bench1000.log 2>&1
python - "$TAG" <<'PY'
import json, sys
for name in ('bench20', 'bench1000', 'ref'):
  for l in open('gpurun_out/%s_%s.log' % (sys.argv[1], name)):
    try: d = json.loads(l)
    except Exception: continue
    if not isinstance(d, dict): continue
    if 'roofline' in d:
      print(name, 'value %.4g  ms/step %.4f  rebuild %.3f ms  force %.4f ms  step_frac %.3f  e2e %.4g rebuilds %d traffic %s' % (
          d['value'], d['ms_per_step'], d['neighbor_rebuild_ms'], d['roofline']['kernel_ms'],
          d['roofline']['step_frac'], d['e2e']['value'], d['config']['rebuilds_in_timed_region'], d['roofline']['traffic']),
          {k: (round(v['value'] / 1e9, 3), round(v['ms_per_step'], 4)) for k, v in d.items() if isinstance(v, dict) and 'atoms' in v})
    else:
      print(name, d.get('value'), (d.get('cpu_baseline') or {}).get('sample'))
PY
(time python benchmarks/configs.py --quick) > gpurun_out/${TAG}_configs.log 2>&1
grep -c '"config"' gpurun_out/${TAG}_configs.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_c4_launches.csv \
  python -c "import sys; sys.path.insert(0,'benchmarks'); import configs; configs.c4(20, 5)" > gpurun_out/${TAG}_c4_launches.log 2>&1
python profiles/launch_summary.py gpurun_out/${TAG}_c4_launches.csv > gpurun_out/${TAG}_c4_launches.txt 2>&1
head -8 gpurun_out/${TAG}_c4_launches.txt
prof() {   # name regex skip cmd...
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 \
      -f -o gpurun_out/${TAG}_$name "$@" > gpurun_out/${TAG}_$name.log 2>&1
  python profiles/ncu_summary.py gpurun_out/${TAG}_$name.ncu-rep > gpurun_out/${TAG}_ncu_$name.txt 2>&1
  python tools/ncu_lines.py gpurun_out/${TAG}_$name.ncu-rep 1.5 >> gpurun_out/${TAG}_ncu_$name.txt 2>&1
  rm -f gpurun_out/${TAG}_$name.ncu-rep
}
prof sw '^k_sw$' 40 python -c "import sys; sys.path.insert(0,'benchmarks'); import configs; configs.c4(20, 5)"
prof swcompact '^k_sw_compact$' 40 python -c "import sys; sys.path.insert(0,'benchmarks'); import configs; configs.c4(20, 5)"
ls -la gpurun_out/${TAG}_* | awk '{print $5, $9}'
