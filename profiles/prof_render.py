"""Driver for ncu: a few 512x512 frames.  python profiles/prof_render.py [m]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'putting-dune_b200'))

import torch

import putting_dune_b200 as pd
from putting_dune_b200 import imaging

m = int(sys.argv[1]) if len(sys.argv) > 1 else 148
b = pd.EnvBatch(m, seed=3)
b.reset()
out = torch.empty((m, 512, 512), dtype=torch.float32, device=b.device)
for _ in range(3):
  s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(
      enable_timing=True)
  s.record()
  imaging.render_batch(b, out=out)
  e.record()
  torch.cuda.synchronize()
  print(f'{m} frames: {s.elapsed_time(e):.3f} ms')
