"""Timing / ncu driver for the small-batch rollout (BASELINE configs[1]):
4096 envs x 256 relative_random steps per launch, prior rates.

  python profiles/prof_rollout_small.py [n_envs] [steps] [reps]
"""
import ctypes as C
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
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
# L2 policy between launches: 'memset' (256 MB fill, leaves L2 full of dirty
# lines), 'pool' (rotate through 10 action buffers, 168 MB > L2), 'read' (read
# a 256 MB buffer: clean lines), 'none'
mode = sys.argv[4] if len(sys.argv) > 4 else 'memset'
rate = pd.RateSpec.prior()
b = pd.EnvBatch(n, seed=0)
b.reset()
dev = b.device
gen = torch.Generator(device=dev)
gen.manual_seed(1)
acts = [(torch.rand((steps, n, 2), generator=gen, device=dev,
                    dtype=torch.float64) * 2 - 1)
        for _ in range(10 if mode == 'pool' else 3)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
P = lambda t: C.c_void_p(t.data_ptr())


def launch(i):
  nat.check(nat.lib.pd_rollout_actions(
      C.byref(b.lattice_tables.c), C.byref(b.c), C.byref(rate.c),
      P(acts[i % len(acts)]), nat.ACTION_RELATIVE_TO_SILICON, 1.42, 1500000, steps,
      2000000, None, None, stream))


for i in range(3):
  launch(i)
torch.cuda.synchronize()
ms = []
for i in range(reps):
  if mode == 'memset':
    flush.zero_()
  elif mode == 'read':
    flush.view(torch.int64).sum()
  s, e = (torch.cuda.Event(enable_timing=True),
          torch.cuda.Event(enable_timing=True))
  s.record()
  launch(i)
  e.record()
  torch.cuda.synchronize()
  ms.append(s.elapsed_time(e))
m = float(np.median(ms))
print('n=%d steps=%d l2=%s spec=%s prepass=%s stride=%s: %.4f ms  %.3e env-steps/s' %
      (n, steps, mode, os.environ.get('PD_ROLLOUT_SPEC', '1'),
       os.environ.get('PD_PREPASS', '1'),
       os.environ.get('PD_LANE_STRIDE', '-'), m, n * steps / (m / 1e3)))
