"""Timing driver for the large-batch stepping kernel (k_walk) on the bench's
workload: relative_random actions through the on-device
RelativeToSiliconActionAdapter, prior rates, dwell 1.5 s.

  python profiles/prof_walk.py [n_envs] [steps_per_launch] [rate] [reps]

Prints env-steps/s (median of `reps` launches, L2 flushed between them).
Knobs are read by the library from the environment: PD_FAST (0: float64
kernels), PD_PREPASS, PD_LANE_STRIDE; OUTPUTS=1 writes the per-step results.
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rate_name = sys.argv[3] if len(sys.argv) > 3 else 'prior'
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 8
rate = pd.RateSpec.prior() if rate_name == 'prior' else pd.RateSpec.simple()
b = pd.EnvBatch(n, seed=0)
b.reset()
dev = b.device
gen = torch.Generator(device=dev)
gen.manual_seed(1)
acts = [(torch.rand((steps, n, 2), generator=gen, device=dev,
                    dtype=torch.float64) * 2 - 1) for _ in range(3)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
P = lambda t: C.c_void_p(t.data_ptr())


# OUTPUTS=1: per-step Si site (int32) and elapsed microseconds (int64) written
# for every env-step, as bench.py times it
d_si = d_el = None
if os.environ.get('OUTPUTS') == '1':
  d_si = torch.empty((steps, n), dtype=torch.int32, device=dev)
  d_el = torch.empty((steps, n), dtype=torch.int64, device=dev)


def launch(i):
  nat.check(nat.lib.pd_rollout_actions(
      C.byref(b.lattice_tables.c), C.byref(b.c), C.byref(rate.c),
      P(acts[i % 3]), nat.ACTION_RELATIVE_TO_SILICON, 1.42, 1500000, steps,
      2000000, P(d_si) if d_si is not None else None,
      P(d_el) if d_el is not None else None, stream))


# (WARM=0 EACH=1: every launch since the reset, timed and printed)
for i in range(int(os.environ.get('WARM', '3'))):
  launch(i)
torch.cuda.synchronize()
ms = []
for i in range(reps):
  flush.zero_()
  s, e = (torch.cuda.Event(enable_timing=True),
          torch.cuda.Event(enable_timing=True))
  s.record()
  launch(i)
  e.record()
  torch.cuda.synchronize()
  ms.append(s.elapsed_time(e))
if os.environ.get('EACH') == '1':
  print('ms per launch:', ' '.join('%.4f' % v for v in ms))
m = float(np.median(ms))
ev = float(b.n_events.sum().item()) / float(b.ctrl_count.sum().item())
print('n=%d steps=%d rate=%s fast=%s plan=%s outputs=%s: %.4f ms  '
      '%.3e env-steps/s  (%.2f iterations per control)' %
      (n, steps, rate_name, os.environ.get('PD_FAST', '1'),
       os.environ.get('PD_PLAN', '0') + '/' + os.environ.get('PD_WALK_PLAN', '1'), os.environ.get('OUTPUTS', '0'), m, n * steps / (m / 1e3), ev))

# PLAN_CLOCKS=1 (library built with -DPD_PLAN_CLOCKS): phase times of the last
# k_rollout_plan launch, per CTA (globaltimer, ns)
if os.environ.get('PLAN_CLOCKS') == '1':
  buf = (C.c_ulonglong * (1024 * 8))()
  nat.lib.pd_debug_plan_clocks(buf)
  a = np.frombuffer(buf, dtype=np.uint64).reshape(1024, 8).astype(np.int64)
  a = a[a[:, 0] > 0]
  t00 = a[:, 0].min()
  names = ['prologue', 'dense', 'queue', 'commit', 'fill', 'store']
  cnt = a[:, 4]
  a = a[:, [0, 1, 2, 3, 5, 6, 7]]
  d = np.diff(a, axis=1) / 1e3
  print('CTAs %d; start spread %.1f us; end (max) %.1f us' %
        (len(a), (a[:, 0].max() - t00) / 1e3, (a[:, -1].max() - t00) / 1e3))
  for i, nm in enumerate(names):
    print('  %-10s mean %6.1f  p50 %6.1f  max %6.1f us' %
          (nm, d[:, i].mean(), np.median(d[:, i]), d[:, i].max()))
  ev = np.stack([(cnt >> sh) & 0xFFFF for sh in (0, 16, 32, 48)], axis=1)
  print('  warp-walker events per launch: replays %d, area checks %d, serial '
        'controls %d, busy masks %d' % tuple(ev.sum(axis=0)))
  tail = (a[:, 5] - a[:, 3]) / 1e3
  order = np.argsort(-tail)[:6]
  for i in order:
    print('    CTA %4d commit+fill %6.1f us: replays %d, area %d, serial %d, '
          'masks %d' % ((i, tail[i]) + tuple(ev[i])))
  quiet = ev.sum(axis=1) == 0
  if quiet.any():
    print('  CTAs without events: %d, commit+fill mean %.1f max %.1f us' %
          (quiet.sum(), tail[quiet].mean(), tail[quiet].max()))

# WALK_CLOCKS=1 (library built with -DPD_PLAN_CLOCKS): phase times of the
# first 1024 CTAs of the last k_walk_plan launch (globaltimer, ns)
if os.environ.get('WALK_CLOCKS') == '1':
  buf = (C.c_ulonglong * (1024 * 8))()
  nat.lib.pd_debug_plan_clocks(buf)
  a = np.frombuffer(buf, dtype=np.uint64).reshape(1024, 8).astype(np.int64)
  a = a[a[:, 0] > 0][:, :7]
  names = ['prologue', 'dense', 'queue', 'commit', 'store (+ later chunks)',
           'epilogue']
  d = np.diff(a, axis=1) / 1e3
  print('CTAs %d; first CTA start to last of these ending %.1f us; CTA '
        'lifetime mean %.1f us' % (len(a), (a[:, 6].max() - a[:, 0].min()) / 1e3,
                                   (a[:, 6] - a[:, 0]).mean() / 1e3))
  for i, nm in enumerate(names):
    print('  %-24s mean %6.2f  p50 %6.2f  max %6.2f us' %
          (nm, d[:, i].mean(), np.median(d[:, i]), d[:, i].max()))
