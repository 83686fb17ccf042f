# round 2, call 5: full GPU suite on the rebuilt library (SW records kernel), config 4 timing + launch list + ncu
TAG=${1:-r02c5}
mkdir -p gpurun_out
(time python -m pytest tests -q -m gpu -x --durations=8) > gpurun_out/${TAG}_tests.log 2>&1
tail -4 gpurun_out/${TAG}_tests.log
python -c "
import sys; sys.path.insert(0,'benchmarks'); import configs
configs.c4(100, 50)" > gpurun_out/${TAG}_c4.log 2>&1
tail -2 gpurun_out/${TAG}_c4.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_c4_launches.csv \
  python -c "import sys; sys.path.insert(0,'benchmarks'); import configs; configs.c4(20, 5)" > gpurun_out/${TAG}_c4_launches.log 2>&1
python profiles/launch_summary.py gpurun_out/${TAG}_c4_launches.csv > gpurun_out/${TAG}_c4_launches.txt 2>&1
head -12 gpurun_out/${TAG}_c4_launches.txt
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
