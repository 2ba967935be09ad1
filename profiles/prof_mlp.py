"""Driver for ncu: learned-rate step, 65536 envs.
python profiles/prof_mlp.py [H] [fp32 | tc (bf16) | split (fp16 hi + lo)]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'putting-dune_b200'))

import numpy as np
import torch

import putting_dune_b200 as pd
from oracle import pdune_oracle as po

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
for _ in range(4):
  b.step_and_image(ctl, 1500000, rate)
torch.cuda.synchronize()
print('ok')
