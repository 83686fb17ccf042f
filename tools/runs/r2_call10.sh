TAG=${1:-r02c10}
mkdir -p gpurun_out
python tools/probe_build.py > gpurun_out/${TAG}_build.log 2>&1
cat gpurun_out/${TAG}_build.log | tail -12
