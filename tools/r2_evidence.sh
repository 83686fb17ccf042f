# round 2 evidence (one GPU): driver-shaped bench, reference arm, configs, ncu launch list and
# --set full captures of every kernel.  Reports are summarised ON the box (the merge back is
# capped at 64 MiB) into gpurun_out/<tag>_ncu_<kernel>.txt and deleted.
TAG=${1:-r02}
mkdir -p gpurun_out
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
      print(name, d.get('value'), (d.get('cpu_baseline') or {}).get('sample'))
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
  python profiles/ncu_summary.py gpurun_out/${TAG}_$name.ncu-rep > gpurun_out/${TAG}_ncu_$name.txt 2>&1
  python tools/ncu_lines.py gpurun_out/${TAG}_$name.ncu-rep 1.5 >> gpurun_out/${TAG}_ncu_$name.txt 2>&1
  if [ "$name" = force ]; then python tools/ncu_traffic.py gpurun_out/${TAG}_$name.ncu-rep k_pair_force 1000188 OrderedSparse > gpurun_out/${TAG}_traffic.log 2>&1; cp profiles/traffic.json gpurun_out/${TAG}_traffic.json; fi
  rm -f gpurun_out/${TAG}_$name.ncu-rep
}
prof force k_pair_force 12 $CMD
prof scan k_nbr_stencil_scan 1 $CMD
prof export k_nbr_export_fin 0 $CMD
prof drift k_kick_drift 5 $CMD
prof update 'k_update<' 3 $CMD
prof offsets k_nbr_offsets 0 $CMD
prof sw 'k_sw<' 2 python -c "import sys; sys.path.insert(0,'benchmarks'); import configs; configs.c4(20, 5)"
prof swcompact k_sw_compact 2 python -c "import sys; sys.path.insert(0,'benchmarks'); import configs; configs.c4(20, 5)"
prof nhc k_nhc_half_step 2 python -c "import sys; sys.path.insert(0,'benchmarks'); import configs; configs.c4(20, 5)"
prof fire k_fire_mix 2 python -c "import sys; sys.path.insert(0,'benchmarks'); import configs; configs.c3(20, 5)"
prof ddpush k_dd_comm_push 8 python tools/probe_dd1.py 63
prof ddwait k_dd_comm_wait 8 python tools/probe_dd1.py 63
prof ddselect k_dd_select_ordered 0 python tools/probe_dd1.py 63
ls -la gpurun_out/${TAG}_* | awk '{print $5, $9}'
