"""Cycles per renderer phase (library built with PD_NVCC_EXTRA=-DPD_RENDER_PHASE_CLOCKS).
python profiles/prof_render_phases.py [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'putting-dune_b200'))

import torch

import putting_dune_b200 as pd
from putting_dune_b200 import imaging

m = int(sys.argv[1]) if len(sys.argv) > 1 else 150
b = pd.EnvBatch(m, seed=3)
b.reset()
out = torch.empty((m, 512, 512), dtype=torch.float32, device=b.device)
imaging.render_batch(b, out=out)
torch.cuda.synchronize()
ws = imaging._workspace(b.device, 512)
off = 256 + (1 << 18)
ws[off:off + 8 * 16 * 512].zero_()
imaging.render_batch(b, out=out)
torch.cuda.synchronize()
t = ws[off:off + 8 * 16 * 512].view(torch.int64).view(-1, 16).cpu().numpy()
t = t[t.sum(axis=1) > 0]
names = ['P0 setup', 'P1 clean', 'P2 blur', 'P3 poisson', 'P4 jitter..uniform',
         'P5 exp', 'P6 gauss', 'P7 hist+maps', 'P8 blend', 'P9 out']
tot = t.sum(axis=1).mean()
print('CTAs', t.shape[0], 'cycles per CTA', tot)
names += ['(P0a params loaded)', '(P0b scan done)', '(P0c tables done)', '(P7a hist done)', '(P7b maps done)', '-']
for i, n in enumerate(names):
  print('%-20s %6.1f%%  (min %.0f max %.0f per CTA)' % (
      n, 100 * t[:, i].mean() / tot, t[:, i].min(), t[:, i].max()))
