"""Times the learned-rate step (bench.py's measure_mlp) with the library named
by PDUNE_B200_LIB: one line per shape / path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'putting-dune_b200'))

import torch

import bench
import putting_dune_b200 as pd

dev = torch.device('cuda:0')
batch = pd.EnvBatch(1024, seed=1, device=dev)
out = bench.measure_mlp(pd, batch, dev, None)
for k, v in out.items():
  print(os.environ.get('PDUNE_B200_LIB', 'default').split('/')[-1], k,
        '%.4f ms' % v['launch_ms'], '%.3e' % v['env_steps_per_s'])
