# ncu evidence for one round (run under gpurun, one GPU).  usage: bash tools/profile.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 40 --warmup 10 --no-cpu --kernel-reps 2"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_force -s 12 -c 1 \
    -f -o gpurun_out/${TAG}_force $CMD > gpurun_out/${TAG}_force.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_nbr_stencil_scan -s 1 -c 1 \
    -f -o gpurun_out/${TAG}_scan $CMD > gpurun_out/${TAG}_scan.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_nbr_export -c 1 \
    -f -o gpurun_out/${TAG}_export $CMD > gpurun_out/${TAG}_export.log 2>&1
ls -la gpurun_out/${TAG}_*
