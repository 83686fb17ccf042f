# after r2_final.sh: refresh the launch list and the `ncu --set full` summaries of the headline kernels with HEAD
TAG=${1:-r02h}
mkdir -p gpurun_out
CMD="python bench.py --steps 40 --warmup 10 --no-cpu --kernel-reps 2 --no-variants --no-extra"
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
python profiles/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.txt 2>&1
head -9 gpurun_out/${TAG}_launches.txt
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
prof update '^k_update$' 3 $CMD
ls -la gpurun_out/${TAG}_* | awk '{print $5, $9}'
