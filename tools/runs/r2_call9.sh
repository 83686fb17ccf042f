TAG=${1:-r02c9}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/probe_dd_steps.py 63 40 > gpurun_out/${TAG}_probe.log 2>&1
grep -E "^phase|overflow|Error|error|dirty|R min" gpurun_out/${TAG}_probe.log | head -40
grep -E "^step" gpurun_out/${TAG}_probe.log | grep " R " | head -5
