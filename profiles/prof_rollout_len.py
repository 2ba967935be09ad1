"""ms per launch of pd_rollout_actions over 4096 envs vs the number of steps
per launch (fixed per-launch costs of the small-batch stepping kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'putting-dune_b200'))

import numpy as np
import torch

import putting_dune_b200 as pd
from putting_dune_b200 import _native as nat

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = pd.EnvBatch(n, seed=1)
b.reset()
rate = pd.RateSpec(1)
rng = np.random.default_rng(0)
for t in (1, 8, 32, 64, 128, 256, 1024):
  acts = torch.as_tensor(rng.uniform(-1, 1, (t, n, 2)), device=b.device)
  for _ in range(3):
    b.rollout(acts, 1500000, rate, action_mode=nat.ACTION_RELATIVE_TO_SILICON)
  torch.cuda.synchronize()
  evs = [(torch.cuda.Event(enable_timing=True),
          torch.cuda.Event(enable_timing=True)) for _ in range(5)]
  for s, e in evs:
    s.record()
    b.rollout(acts, 1500000, rate, action_mode=nat.ACTION_RELATIVE_TO_SILICON)
    e.record()
  torch.cuda.synchronize()
  ms = sorted(s.elapsed_time(e) for s, e in evs)[2]
  print('%5d steps: %.4f ms  %.3f us/step  %.3e env-steps/s' % (
      t, ms, 1e3 * ms / t, n * t / (ms / 1e3)))
