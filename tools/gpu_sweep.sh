mkdir -p gpurun_out
CMD="python bench.py --steps 40 --warmup 10 --no-cpu --no-variants --kernel-reps 2"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01s6_launches.csv $CMD > gpurun_out/r01s6_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_force -s 12 -c 1 -f -o gpurun_out/r01s6_force $CMD > gpurun_out/r01s6_force.log 2>&1
ls gpurun_out/r01s6*
