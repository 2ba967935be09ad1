"""CPU (gloo, world_size 2) test of the multi-rank path: env sharding and the
all-gather of packed episode statistics (the only collective on the path)."""

import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
  for p in (ROOT, os.path.join(ROOT, 'putting-dune_b200')):
    if p not in sys.path:
      sys.path.insert(0, p)
  from putting_dune_b200 import episodes
  dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}',
                          rank=rank, world_size=world)
  lo, n = episodes.shard_bounds(total, rank, world)
  rec = np.zeros(n, dtype=episodes.STATS_DTYPE)
  ids = np.arange(lo, lo + n)
  rec['num_actions'] = ids % 11 + 1
  rec['reached_goal'] = ids % 3 != 0
  rec['env_seconds'] = np.where(rec['reached_goal'], ids * 0.5, np.nan)
  rec['total_reward'] = np.where(rec['reached_goal'], 0.9, 0.0)
  local = torch.from_numpy(rec.view(np.uint8).reshape(n, 16).copy())
  full = episodes.gather_episode_stats(local)
  assert full.shape == (total, 16)
  agg = episodes.aggregate_results(full)
  np.save(os.path.join(out_dir, f'agg_{rank}.npy'),
          np.array([agg['average_num_times_reached_goal'],
                    agg['average_num_actions_taken'],
                    agg['average_environment_seconds_to_goal'],
                    agg['average_total_reward'], agg['total_actions']]))
  dist.destroy_process_group()


def test_gloo_allgather_of_episode_stats(tmp_path):
  world, total = 2, 64
  mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)),
           nprocs=world, join=True)
  a0 = np.load(tmp_path / 'agg_0.npy')
  a1 = np.load(tmp_path / 'agg_1.npy')
  np.testing.assert_array_equal(a0, a1)  # every rank sees the whole job
  ids = np.arange(total)
  reached = ids % 3 != 0
  acts = ids % 11 + 1
  np.testing.assert_allclose(a0[0], reached.mean())
  np.testing.assert_allclose(a0[1], acts[reached].sum() / reached.sum())
  np.testing.assert_allclose(a0[2], (ids * 0.5)[reached].sum() / reached.sum())
  np.testing.assert_allclose(a0[3], 0.9, rtol=1e-6)
  assert a0[4] == acts.sum()


def test_shard_bounds():
  import pytest
  sys.path.insert(0, os.path.join(ROOT, 'putting-dune_b200'))
  from putting_dune_b200 import episodes
  assert episodes.shard_bounds(1 << 20, 3, 8) == (3 * 131072, 131072)
  with pytest.raises(ValueError):
    episodes.shard_bounds(10, 0, 3)
