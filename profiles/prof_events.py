"""Small driver for ncu: a few large-batch event steps (1Mi envs).

  python profiles/prof_events.py [prior|simple] [n_envs] [dwell_us]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'putting-dune_b200'))

import numpy as np
import torch

import putting_dune_b200 as pd

rate = sys.argv[1] if len(sys.argv) > 1 else 'prior'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
dwell = int(sys.argv[3]) if len(sys.argv) > 3 else 1500000
spec = pd.RateSpec.prior() if rate == 'prior' else pd.RateSpec.simple()
b = pd.EnvBatch(n, seed=0)
b.reset()
rng = np.random.default_rng(0)
ctl = torch.as_tensor(0.5 + rng.uniform(-1, 1, (n, 1, 2)) * 1.42 / 22.5,
                      device=b.device)
for _ in range(6):
  out = b.step_and_image(ctl, dwell, spec)
torch.cuda.synchronize()
print('events/step', out.events.double().mean().item(), 'transitions/step',
      out.transitions.double().mean().item())
