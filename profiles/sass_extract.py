"""Opcode counts per kernel from the built library's SASS (evidence that the
Blackwell paths are what runs: tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM,
tcgen05.commit = UTCBAR, cp.async.bulk = UBLKCP, mbarrier = SYNCS).
usage: python profiles/sass_extract.py > profiles/r02_sass_extract.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'putting-dune_b200', 'lib', 'libpdune_b200.so')
WANT = ('UTCHMMA', 'LDTM', 'UTCBAR', 'UBLKCP', 'SYNCS', 'MUFU.EX2', 'MUFU.LG2',
        'MUFU.RCP', 'IMAD.WIDE.U32', 'DFMA', 'DMUL', 'DADD', 'FFMA', 'REDG')
sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True,
                      text=True).stdout
names = subprocess.run(['c++filt'], input='\n'.join(
    re.findall(r'Function : (\S+)', sass)), capture_output=True,
    text=True).stdout.split('\n')
print('# SASS extract of libpdune_b200.so (sm_100a), end of round 2')
print('# cuobjdump -sass putting-dune_b200/lib/libpdune_b200.so, opcode counts '
      'per kernel')
print('# (tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM, tcgen05.commit = UTCBAR, '
      'cp.async.bulk = UBLKCP,')
print('#  mbarrier = SYNCS; fast kernels: MUFU.EX2/LG2/RCP float32 iteration, '
      'IMAD.WIDE.U32 Philox)')
print()
blocks = re.split(r'\n\s*Function : ', sass)[1:]
rows = []
for name, block in zip(names, blocks):
  cnt = collections.Counter()
  for op in re.findall(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)',
                       block, flags=re.M):
    for w in WANT:
      if op == w or op.startswith(w + '.'):
        cnt[w] += 1
  if cnt:
    rows.append((name, cnt))
for name, cnt in sorted(rows):
  print(name[:100])
  print('    ' + ', '.join(f'{k}: {v}' for k, v in sorted(cnt.items())))
