# round 2: full single-GPU validation + evidence.  usage: bash tools/r2_full.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --durations=8) > gpurun_out/${TAG}_tests.log 2>&1
tail -14 gpurun_out/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time python bench.py --steps 20 --warmup 5) > gpurun_out/${TAG}_bench20.log 2>&1
(time python bench.py --impl reference --steps 20 --warmup 5) > gpurun_out/${TAG}_ref.log 2>&1
(time python bench.py --no-cpu --no-extra) > gpurun_out/${TAG}_bench1000.log 2>&1
python - "$TAG" <<'PY'
import json, sys
for name in ('bench20', 'bench1000', 'ref'):
  for l in open('gpurun_out/%s_%s.log' % (sys.argv[1], name)):
    try: d = json.loads(l)
    except Exception: continue
    if 'roofline' in d:
      print(name, 'value %.4g  ms/step %.4f  rebuild %.3f ms  force %.4f ms  step_frac %.3f  e2e %.4g rebuilds %d' % (
          d['value'], d['ms_per_step'], d['neighbor_rebuild_ms'], d['roofline']['kernel_ms'],
          d['roofline']['step_frac'], d['e2e']['value'], d['config']['rebuilds_in_timed_region']),
          {k: (round(v['value'] / 1e9, 3), round(v['ms_per_step'], 4)) for k, v in d.items() if isinstance(v, dict) and 'atoms' in v})
    else:
      print(name, d.get('value'), d.get('cpu_baseline', {}).get('sample'))
PY
(time python benchmarks/configs.py --quick) > gpurun_out/${TAG}_configs.log 2>&1
grep -c config gpurun_out/${TAG}_configs.log
CMD="python bench.py --steps 40 --warmup 10 --no-cpu --kernel-reps 2 --no-variants --no-extra"
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
prof() {   # name regex skip cmd...
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 \
      -f -o gpurun_out/${TAG}_$name "$@" > gpurun_out/${TAG}_$name.log 2>&1
}
prof force k_pair_force 12 $CMD
prof scan k_nbr_stencil_scan 1 $CMD
prof export k_nbr_export_fin 0 $CMD
prof drift k_kick_drift 5 $CMD
prof update 'k_update<' 3 $CMD
prof offsets k_nbr_offsets 0 $CMD
C4="python -c \"import sys; sys.argv=['x','--quick']; sys.path.insert(0,'benchmarks'); import configs; configs.c4(20, 5)\""
ncu --set full --clock-control none --import-source on -k regex:k_sw -s 2 -c 2 -f -o gpurun_out/${TAG}_sw \
    python -c "import sys; sys.path.insert(0,'benchmarks'); import configs; configs.c4(20, 5)" > gpurun_out/${TAG}_sw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_nhc_half_step -s 2 -c 1 -f -o gpurun_out/${TAG}_nhc \
    python -c "import sys; sys.path.insert(0,'benchmarks'); import configs; configs.c4(20, 5)" > gpurun_out/${TAG}_nhc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fire_mix -s 2 -c 1 -f -o gpurun_out/${TAG}_fire \
    python -c "import sys; sys.path.insert(0,'benchmarks'); import configs; configs.c3(20, 5)" > gpurun_out/${TAG}_fire.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_dd_comm -s 8 -c 2 -f -o gpurun_out/${TAG}_ddcomm \
    python tools/probe_dd1.py 63 > gpurun_out/${TAG}_ddcomm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_dd_select_ordered -s 0 -c 1 -f -o gpurun_out/${TAG}_ddselect \
    python tools/probe_dd1.py 63 > gpurun_out/${TAG}_ddselect.log 2>&1
ls -la gpurun_out/${TAG}_* | awk '{print $5, $9}'
