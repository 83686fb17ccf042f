"""Per-source-line instruction counts of an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [min_pct]"""
import csv, io, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(io.StringIO(out)))
hdr = None
lines = []
fname = ''
for r in rows:
  if r and r[0] == 'File Path':
    fname = r[1].split('/')[-1]
  if r and r[0] == 'Line No' and 'Instructions Executed' in r:
    hdr = r
    ie = hdr.index('Instructions Executed'); ism = hdr.index('# Samples')
    continue
  if hdr and r and r[0].isdigit() and len(r) > ie:
    try:
      lines.append((fname, int(r[0]), int(r[ie]), int(r[ism] or 0), r[1]))
    except ValueError:
      pass
tot = sum(l[2] for l in lines)
print('total warp instructions', tot)
for f, l, e, sm, src in lines:
  if e >= minp / 100 * tot:
    print('%-24s %4d %6.2f%% %6d  %s' % (f, l, 100 * e / tot, sm, src.strip()[:110]))
