"""Whole goal-reaching episodes on the device and their statistics.

Reference: putting_dune/eval_lib.py (``evaluate`` :77-184, ``EvalResult``
:47-59, ``aggregate_results`` :187-214) for the ``greedy_on_neighbor``
experiment (experiments/registry.py:287-298).  Environments shard across GPUs
by global env id with no inter-step communication; the only collective is one
all-gather of the packed 16-byte per-env records.
"""

from __future__ import annotations

import ctypes as C
import dataclasses
import datetime as dt
from typing import Optional

import numpy as np
import torch

from putting_dune_b200 import _native as nat
from putting_dune_b200 import engine

STATS_DTYPE = np.dtype([('num_actions', '<i4'), ('env_seconds', '<f4'),
                        ('total_reward', '<f4'), ('reached_goal', 'u1'),
                        ('pad', 'u1', (3,))])
assert STATS_DTYPE.itemsize == 16


@dataclasses.dataclass(frozen=True)
class EpisodeConfig:
  dwell_time: dt.timedelta = dt.timedelta(seconds=5.0)  # registry.py:291-294
  image_duration: dt.timedelta = dt.timedelta(seconds=2.0)
  timeout: dt.timedelta = dt.timedelta(minutes=10)  # eval_lib.py:82
  step_limit: int = 600  # run_helpers.py:34
  argmax: tuple = (1.42, 0.0)  # registry.py:289

  def to_c(self) -> nat.PdEpisodeConfig:
    us = lambda t: t // dt.timedelta(microseconds=1)
    return nat.PdEpisodeConfig(us(self.dwell_time), us(self.image_duration),
                               us(self.timeout), self.step_limit, 0,
                               float(self.argmax[0]), float(self.argmax[1]))


def run_greedy_episodes(batch: engine.EnvBatch, rate: engine.RateSpec,
                        cfg: Optional[EpisodeConfig] = None):
  """Resets every env of ``batch``, draws goals and runs the greedy controller
  to the end of each episode, in one kernel launch.  Returns (stats uint8
  [n, 16] device tensor of ``pd_episode_stats`` records, goal_site int32 [n],
  goal_xy float64 [n, 2])."""
  cfg = cfg or EpisodeConfig()
  n, dev = batch.num_envs, batch.device
  stats = torch.zeros((n, 16), dtype=torch.uint8, device=dev)
  goal_xy = torch.zeros((n, 2), dtype=torch.float64, device=dev)
  goal_site = torch.zeros((n,), dtype=torch.int32, device=dev)
  c_cfg = cfg.to_c()
  P = lambda t: C.c_void_p(t.data_ptr())
  with torch.cuda.device(dev):
    nat.check(nat.lib.pd_run_episodes(
        C.byref(batch.lattice_tables.c), C.byref(batch.c), C.byref(rate.c),
        C.byref(c_cfg), P(goal_xy), P(goal_site), P(stats),
        C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
  return stats, goal_site, goal_xy


def stats_to_numpy(stats: torch.Tensor) -> np.ndarray:
  """uint8 [n, 16] -> structured array with EvalResult's fields."""
  return stats.detach().cpu().numpy().reshape(-1).view(STATS_DTYPE)


def gather_episode_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
  """All-gathers the per-rank records (NCCL over NVLink on GPUs; any backend
  torch.distributed offers).  Every rank must hold the same number of envs."""
  import torch.distributed as dist
  if not (dist.is_available() and dist.is_initialized()):
    return stats
  world = dist.get_world_size(group)
  out = torch.empty((world * stats.shape[0], stats.shape[1]),
                    dtype=stats.dtype, device=stats.device)
  dist.all_gather_into_tensor(out, stats.contiguous(), group=group)
  return out


def shard_bounds(total_envs: int, rank: int, world: int):
  """Contiguous block of global env ids owned by ``rank``."""
  per = total_envs // world
  if per * world != total_envs:
    raise ValueError('total_envs must divide evenly across ranks')
  return rank * per, per


def aggregate_results(stats) -> dict:
  """eval_lib.py:187-214: averages over the episodes that reached the goal."""
  s = stats_to_numpy(stats) if torch.is_tensor(stats) else stats
  reached = s['reached_goal'].astype(bool)
  den = max(int(reached.sum()), 1)
  return {
      'average_num_times_reached_goal': float(reached.mean()) if s.size else 0.,
      'average_num_actions_taken':
          float(s['num_actions'][reached].astype(np.int64).sum()) / den,
      'average_environment_seconds_to_goal':
          float(s['env_seconds'][reached].astype(np.float64).sum()) / den,
      'average_total_reward':
          float(s['total_reward'][reached].astype(np.float64).sum()) / den,
      'episodes': int(s.size),
      'total_actions': int(s['num_actions'].astype(np.int64).sum()),
  }
