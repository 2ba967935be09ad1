"""Cycles per phase of k_step_learned (library built with
PD_NVCC_EXTRA=-DPD_MLP_PHASE_CLOCKS).  python profiles/prof_mlp_phases.py [H] [tc]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'putting-dune_b200'))

import numpy as np
import torch

import putting_dune_b200 as pd
from oracle import pdune_oracle as po
from putting_dune_b200 import _native as nat

h = int(sys.argv[1]) if len(sys.argv) > 1 else 256
tc = {'tc': 1, 'split': 2}.get(sys.argv[2] if len(sys.argv) > 2 else '', 0)
n = 65536
mlp = po.MlpParams.synthetic(7, hidden=(h, h))
w = pd.MlpWeights(**{k: getattr(mlp, k) for k in pd.MlpWeights.NAMES})
rate = pd.RateSpec(2, mlp=w, tensor_core=tc)
b = pd.EnvBatch(n, seed=11)
b.reset()
rng = np.random.default_rng(0)
ctl = torch.as_tensor(0.5 + rng.uniform(-1, 1, (n, 1, 2)) * 1.42 / 22.5,
                      device=b.device)
buf = (C.c_ulonglong * 16)()
b.step_and_image(ctl, 1500000, rate)
nat.lib.pd_debug_mlp_phases(buf)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
b.step_and_image(ctl, 1500000, rate)
e.record()
torch.cuda.synchronize()
nat.lib.pd_debug_mlp_phases(buf)
t = np.array(list(buf), dtype=np.float64)
print('H', h, {0: 'fp32', 1: 'tc bf16', 2: 'tc split'}[tc], 'ms', s.elapsed_time(e))
for i, name in enumerate(['build items (f64 canonicalise)', 'network wave',
                          'events + finalise', 'compaction']):
  print('%-32s %5.1f%%  %.0f cycles/CTA' % (name, 100 * t[i] / t[:4].sum(),
                                            t[i] / 148))
if tc:
  w_tot = t[8:12].sum()
  for i, name in zip(range(8, 12), ['operands (h1 split, W1 copy)',
                                    'MMA issue + wait',
                                    'epilogue (tcgen05.ld, bias, swish, W2)',
                                    'head reduction + softplus']):
    print('  wave: %-38s %5.1f%%  %.0f cycles/CTA' %
          (name, 100 * t[i] / max(w_tot, 1), t[i] / 148))
e_tot = t[12:16].sum()
for i, name in zip(range(12, 16), ['rates + ctrl_count + Philox', 'kmc_event',
                                   'hop bookkeeping', 'finalise']):
  print('  events (warp 0): %-28s %5.1f%%  %.0f cycles/CTA' %
        (name, 100 * t[i] / max(e_tot, 1), t[i] / 148))
b_tot = t[4:8].sum()
for i, name in zip(range(4, 8), ['state loads + dwell loop', 'key barrier',
                                 'site / neighbour positions', 'canonicalise']):
  print('  build (warp 0): %-28s %5.1f%%  %.0f cycles/CTA' %
        (name, 100 * t[i] / max(b_tot, 1), t[i] / 148))
