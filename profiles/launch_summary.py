"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python profiles/launch_summary.py gpurun_out/launches.csv [min_us]
Empty (gated no-op) launches are separated with min_us (default 3.5)."""
import collections
import csv
import re
import sys


def main(path, min_us=3.5):
  rows = list(csv.reader(open(path)))
  hdr = None
  d = collections.OrderedDict()
  tot = 0.0
  n = 0
  for r in rows:
    if hdr is None:
      if 'Kernel Name' in r:
        hdr = r
        ki, vi, ui = r.index('Kernel Name'), r.index('Metric Value'), r.index('Metric Unit')
      continue
    if len(r) <= vi:
      continue
    name = re.sub(r'\(.*', '', r[ki]).replace('(anonymous namespace)::', '')
    name = name.replace('void ', '').replace('<unnamed>::', '')
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    d.setdefault(name, []).append(v)
    tot += v
    n += 1
  print('total %.1f us over %d launches' % (tot, n))
  print('%-58s %5s %5s %11s %10s %10s %7s' % ('kernel', 'n', 'real', 'sum_us', 'mean_real', 'max', 'share'))
  for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    real = [x for x in v if x > min_us]
    mr = sum(real) / len(real) if real else 0.0
    print('%-58s %5d %5d %11.1f %10.1f %10.1f %6.1f%%' % (
        k[:58], len(v), len(real), sum(v), mr, max(v), 100 * sum(v) / tot))


if __name__ == '__main__':
  main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 3.5)
