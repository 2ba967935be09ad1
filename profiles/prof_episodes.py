"""Driver for timing / ncu: greedy goal-reaching episodes (BASELINE configs[4])
on one GPU.  python profiles/prof_episodes.py [n_envs] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'putting-dune_b200'))

import torch

import putting_dune_b200 as pd
from putting_dune_b200 import episodes as ep

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
b = pd.EnvBatch(n, seed=5)
rate = pd.RateSpec.simple()
ep.run_greedy_episodes(b, rate)
torch.cuda.synchronize()
for _ in range(reps):
  s, e = (torch.cuda.Event(enable_timing=True),
          torch.cuda.Event(enable_timing=True))
  s.record()
  stats, _, _ = ep.run_greedy_episodes(b, rate)
  e.record()
  torch.cuda.synchronize()
  agg = ep.aggregate_results(ep.gather_episode_stats(stats))
  print('%d episodes: %.3f ms, %d actions, %.3e env-steps/s' %
        (n, s.elapsed_time(e), agg['total_actions'],
         agg['total_actions'] / (s.elapsed_time(e) / 1e3)))
