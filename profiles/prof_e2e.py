"""Host-buffer rollout (pd_rollout_actions_host_f32) at configs[1]: wall time
per call and a digest of the results.

  python profiles/prof_e2e.py                        # chunked copy-engine pipeline
  PD_HOST_STREAMED=1 python profiles/prof_e2e.py     # streamed launch (opt-in)
  PD_HOST_TRACE=1 REPS=4 python profiles/prof_e2e.py # timeline of each call
  OWNED=1 python profiles/prof_e2e.py                # library-owned stagings
  PD_HOST_COPY_SMS=8 python profiles/prof_e2e.py     # SMs given to the writers
"""
import ctypes as C
import hashlib
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
reps = int(os.environ.get('REPS', 50))
dev = torch.device('cuda:0')
batch = pd.EnvBatch(n, seed=0, device=dev)
batch.reset()
rate = pd.RateSpec(nat.RATE_PRIOR)
rng = np.random.default_rng(5)
pool = 4
h_a = [torch.as_tensor(rng.uniform(-1, 1, size=(t_steps, n, 2))
                       .astype(np.float32)).pin_memory() for _ in range(pool)]
d_a32 = torch.empty((t_steps, n, 2), dtype=torch.float32, device=dev)
d_ctl = torch.empty((t_steps, n, 2), dtype=torch.float64, device=dev)
d_si = torch.empty((t_steps, n), dtype=torch.int32, device=dev)
d_el = torch.empty((t_steps, n), dtype=torch.int64, device=dev)
d_el32 = torch.empty((t_steps, n), dtype=torch.int32, device=dev)
h_si = torch.empty((t_steps, n), dtype=torch.int32).pin_memory()
h_el = torch.empty((t_steps, n), dtype=torch.int32).pin_memory()
P = lambda t: C.c_void_p(t.data_ptr())
stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


# OWNED=1: stagings kept by the library (re-filled behind each call)
STAGE = ((None,) * 5 if os.environ.get('OWNED') == '1' else
         (P(d_a32), P(d_ctl), P(d_si), P(d_el), P(d_el32)))


def call(i):
  nat.check(nat.lib.pd_rollout_actions_host_f32(
      C.byref(batch.lattice_tables.c), C.byref(batch.c), C.byref(rate.c),
      P(h_a[i % pool]), nat.ACTION_RELATIVE_TO_SILICON, 1.42, 1500000,
      t_steps, 2000000, *STAGE, P(h_si), P(h_el), stream))


digest = hashlib.sha256()
for i in range(pool):
  h_si.zero_()
  h_el.zero_()
  call(i)
  digest.update(h_si.numpy().tobytes())
  digest.update(h_el.numpy().tobytes())
torch.cuda.synchronize()
digest.update(batch.si_idx.cpu().numpy().tobytes())
digest.update(batch.sim_time_us.cpu().numpy().tobytes())
times = []
checksum = 0  # CHECKSUM=1: over the results of every timed call (stress test)
for i in range(reps):
  t0 = time.perf_counter()
  call(i)
  times.append(time.perf_counter() - t0)
  if os.environ.get('CHECKSUM') == '1':
    checksum = (checksum * 1000003 + int(h_si.numpy().sum(dtype=np.int64)) +
                3 * int(h_el.numpy().sum(dtype=np.int64))) % (1 << 61)
if os.environ.get('CHECKSUM') == '1':
  digest.update(str(checksum).encode())
times = np.asarray(times) * 1e3
print('streamed=%s owned=%s n=%d steps=%d: median %.4f ms  min %.4f ms  '
      '%.3e env-steps/s  digest %s' % (
          os.environ.get('PD_HOST_STREAMED', '0'),
          os.environ.get('OWNED', '0'), n, t_steps,
          np.median(times), times.min(), n * t_steps / np.median(times) * 1e3,
          digest.hexdigest()[:16]))
