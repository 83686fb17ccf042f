"""DRAM bytes per launch of a kernel from an `ncu --set full` report -> profiles/traffic.json
(read by bench.py for roofline.traffic).
usage: python tools/ncu_traffic.py report.ncu-rep key atoms format"""
import csv, io, json, os, subprocess, sys
rep, key, atoms, fmt = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
def b(name):
  v = float(d[name].replace(',', ''))
  return int(v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u[name]])
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'traffic.json')
db = json.load(open(path)) if os.path.exists(path) else {}
db[key] = {'kernel': d.get('Kernel Name', '')[:80], 'atoms': atoms, 'format': fmt,
           'dram_bytes_read': b('dram__bytes_read.sum'), 'dram_bytes_write': b('dram__bytes_write.sum'),
           'duration_us': float(d['gpu__time_duration.sum'].replace(',', '')) * ({'us': 1, 'ms': 1e3, 'ns': 1e-3}[u['gpu__time_duration.sum']]),
           'source': 'ncu --set full, ' + os.path.basename(rep)}
json.dump(db, open(path, 'w'), indent=1)
print(db[key])
