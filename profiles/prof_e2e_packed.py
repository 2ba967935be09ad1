"""Host-buffer rollout in the packed format (pd_rollout_actions_host_packed)
at configs[1]: wall time per call.

  PD_PACKED_SCHEDULE=5,5,4,2 python profiles/prof_e2e_packed.py
"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..',
                                'putting-dune_b200'))
import putting_dune_b200 as pd  # noqa: E402
from putting_dune_b200 import _native as nat  # noqa: E402

n = int(os.environ.get('N_ENVS', 4096))
t_steps = int(os.environ.get('N_STEPS', 256))
reps = int(os.environ.get('REPS', 60))
dev = torch.device('cuda:0')
batch = pd.EnvBatch(n, seed=0, device=dev)
batch.reset()
rate = pd.RateSpec(nat.RATE_PRIOR)
rng = np.random.default_rng(5)
pool = 4
h_a = [torch.as_tensor(rng.uniform(-1, 1, size=(t_steps, n, 2))
                       .astype(np.float32)).pin_memory() for _ in range(pool)]
h_p = torch.empty((t_steps, n), dtype=torch.uint16).pin_memory()
P = lambda t: C.c_void_p(t.data_ptr())
stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def call(i):
  nat.check(nat.lib.pd_rollout_actions_host_packed(
      C.byref(batch.lattice_tables.c), C.byref(batch.c), C.byref(rate.c),
      P(h_a[i % pool]), nat.ACTION_RELATIVE_TO_SILICON, 1.42, 1500000,
      t_steps, 2000000, P(h_p), stream))


for i in range(5):
  call(i)
times = []
for i in range(reps):
  t0 = time.perf_counter()
  call(i)
  times.append(time.perf_counter() - t0)
times = np.asarray(times) * 1e3
print('schedule=%s n=%d steps=%d: median %.4f ms  min %.4f ms  %.3e '
      'env-steps/s' % (os.environ.get('PD_PACKED_SCHEDULE', 'default'), n,
                       t_steps, np.median(times), times.min(),
                       n * t_steps / np.median(times) * 1e3))
