mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_energy.py -m gpu -x -q -k "lammps or jammed") > gpurun_out/s25_tests.log 2>&1; tail -30 gpurun_out/s25_tests.log
