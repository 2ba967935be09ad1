"""Per-source-line instruction / stall-sample shares from an ncu report.
usage: python profiles/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv',
                      '--print-source', 'cuda,sass'], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur, hdr, agg = None, None, {}
for r in rows:
  if len(r) >= 2 and r[0] == 'File Path':
    cur = r[1].split('/')[-1]
  elif len(r) > 4 and r[0] == 'Line No':
    hdr = r
  elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == '-':
    ih, isamp = hdr.index('Instructions Executed'), hdr.index('# Samples')
    key = (cur, int(r[0]))
    a = agg.setdefault(key, [0, 0, r[1].strip()[:80]])
    a[0] += int(r[ih] or 0)
    a[1] += int(r[isamp] or 0)
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[1] for a in agg.values()) or 1
print('total warp-instructions', tot, 'samples', tots)
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
  print('%5.1f%% inst %5.1f%% samp  %s:%d  %s' %
        (100 * a[0] / tot, 100 * a[1] / tots, f, ln, a[2]))

if len(sys.argv) > 3:  # ranges: file:lo-hi,... -> share per range
  for spec in sys.argv[3].split(','):
    f, rng = spec.split(':')
    lo, hi = (int(v) for v in rng.split('-'))
    ins = sum(a[0] for (ff, ln), a in agg.items() if ff == f and lo <= ln <= hi)
    smp = sum(a[1] for (ff, ln), a in agg.items() if ff == f and lo <= ln <= hi)
    print('%-36s %5.1f%% inst %5.1f%% samp' % (spec, 100 * ins / tot,
                                              100 * smp / tots))
