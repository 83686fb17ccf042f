TAG=${1:-r01s3}
mkdir -p gpurun_out
CMD="python bench.py --steps 40 --warmup 10 --no-cpu --kernel-reps 2"
JMD_STAGE=1 ncu --set full --clock-control none --import-source on -k regex:k_pair_force -s 12 -c 1 \
    -f -o gpurun_out/${TAG}_force_staged $CMD > gpurun_out/${TAG}_force_staged.log 2>&1
JMD_STAGE=0 ncu --set full --clock-control none --import-source on -k regex:k_pair_force -s 12 -c 1 \
    -f -o gpurun_out/${TAG}_force_direct $CMD > gpurun_out/${TAG}_force_direct.log 2>&1
ls -la gpurun_out/${TAG}_*
