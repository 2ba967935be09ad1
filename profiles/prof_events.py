"""Small driver for ncu / timing: large-batch event steps.

  python profiles/prof_events.py [prior|simple] [n_envs] [dwell_us] [mode]

mode = step (one step_and_image per launch) or rollout8 (8 fused steps).
Beam = Si position + U(-1,1)^2 bond lengths (the relative_random workload,
registry.py:263-266), refreshed from the device state before each launch.
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
mode = sys.argv[4] if len(sys.argv) > 4 else 'step'
spec = pd.RateSpec.prior() if rate == 'prior' else pd.RateSpec.simple()
b = pd.EnvBatch(n, seed=0)
b.reset()
gen = torch.Generator(device=b.device)
gen.manual_seed(0)


def near_si_controls(steps):
  """Microscope-frame beam within one bond of the Si (the FOV re-centres, so
  the Si sits in [0.25, 0.75]^2 of the frame)."""
  p = b.silicon_position()
  q = (p - b.fov[:, :2]) / (b.fov[:, 2:] - b.fov[:, :2])
  a = torch.rand((steps, n, 2), generator=gen, device=b.device,
                 dtype=torch.float64) * 2 - 1
  return q[None] + a * (1.42 / b.fov_scale)[None, :, None]


times = []
for i in range(8):
  if mode == 'step':
    ctl = near_si_controls(1)[0][:, None, :].contiguous()
  else:
    ctl = near_si_controls(8).contiguous()
  torch.cuda.synchronize()
  s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(
      enable_timing=True)
  s.record()
  if mode == 'step':
    out = b.step_and_image(ctl, dwell, spec)
  else:
    b.rollout(ctl, dwell, spec)
  e.record()
  torch.cuda.synchronize()
  times.append(s.elapsed_time(e))
steps = n * (1 if mode == 'step' else 8)
ms = float(np.median(times[2:]))
print(f'{rate} dwell={dwell/1e6}s {mode} kernel={os.environ.get("PD_STEP_KERNEL","auto")}: '
      f'{ms:.4f} ms/launch  {steps/ms*1e3:.3e} env-steps/s  '
      f'events/ctl={b.n_events.double().sum().item()/b.ctrl_count.double().sum().item():.3f} '
      f'transitions/ctl={b.n_transitions.double().sum().item()/b.ctrl_count.double().sum().item():.3f}')
