"""Batched RL environment: ``num_envs`` PuttingDuneEnvironments stepped by one
call, observations / rewards / discounts as device tensors.

Reference: putting_dune/putting_dune_environment.py:36-158 (dm_env semantics:
a step on a fresh or finished episode resets and returns FIRST) wrapped in
run_helpers.StepLimitWrapper (:120-153), with the action adapters of
action_adapters.py:53-274, the two 10-float feature constructors of
feature_constructors.py:79-228 and goals.SingleSiliconGoalReaching.
"""

from __future__ import annotations

import collections
import ctypes as C
import datetime as dt
from typing import Optional

import torch

from putting_dune_b200 import _native as nat
from putting_dune_b200 import constants
from putting_dune_b200 import engine

TimeStep = collections.namedtuple(
    'TimeStep', ['step_type', 'reward', 'discount', 'observation'])

ADAPTERS = {'direct': nat.ADAPTER_DIRECT, 'delta_position': nat.ADAPTER_DELTA,
            'relative_to_silicon': nat.ADAPTER_RELATIVE,
            'relative_to_silicon_material_frame':
                nat.ADAPTER_RELATIVE_MATERIAL}
FEATURES = {'microscope_frame': nat.FEATURES_MICROSCOPE,
            'material_frame': nat.FEATURES_MATERIAL}


class BatchedPuttingDuneEnvironment:
  """``step(actions)`` advances every env; finished ones start a new episode
  on their next step, exactly like the reference's single env."""

  def __init__(self, num_envs: int, *, rate: Optional[engine.RateSpec] = None,
               action_adapter: str = 'relative_to_silicon',
               feature_constructor: str = 'microscope_frame',
               dwell_time_range=(dt.timedelta(seconds=1.5),
                                 dt.timedelta(seconds=1.5)),
               max_distance_angstroms: float =
               constants.CARBON_BOND_DISTANCE_ANGSTROMS,
               image_duration: dt.timedelta = dt.timedelta(seconds=2.0),
               step_limit: int = 600, seed: int = 0, device=None,
               env_offset: int = 0):
    self.rate = rate or engine.RateSpec.simple()
    self.batch = engine.EnvBatch(num_envs, seed=seed, device=device,
                                 env_offset=env_offset)
    adapter = ADAPTERS[action_adapter]
    d0, d1 = (t.total_seconds() for t in dwell_time_range)
    relative = adapter >= nat.ADAPTER_RELATIVE
    self.action_dim = 3 if relative and d0 != d1 else 2
    self.cfg = nat.PdEnvConfig(
        adapter, FEATURES[feature_constructor], self.action_dim, step_limit,
        d0, d1, float(max_distance_angstroms),
        image_duration // dt.timedelta(microseconds=1))
    e, dev = num_envs, self.batch.device
    z = lambda shape, t: torch.zeros(shape, dtype=t, device=dev)
    self._bufs = dict(
        goal_xy=z((e, 2), torch.float64), beam_pos=z((e, 2), torch.float64),
        elapsed_steps=z((e,), torch.int32),
        needs_reset=torch.ones((e,), dtype=torch.uint8, device=dev),
        controls_xy=z((e, 2), torch.float64), dwell_us=z((e,), torch.int64),
        elapsed_us=z((e,), torch.int64), resetting=z((e,), torch.uint8))
    self._buf_c = nat.PdEnvBuffers(
        *[self._bufs[k].data_ptr() for k in (
            'goal_xy', 'beam_pos', 'elapsed_steps', 'needs_reset',
            'controls_xy', 'dwell_us', 'elapsed_us', 'resetting')])
    self._obs = z((e, 10), torch.float32)
    self._reward = z((e,), torch.float32)
    self._discount = z((e,), torch.float32)
    self._step_type = z((e,), torch.int32)

  @property
  def num_envs(self) -> int:
    return self.batch.num_envs

  @property
  def goal_position_material_frame(self) -> torch.Tensor:
    return self._bufs['goal_xy']

  def reset(self) -> TimeStep:
    """Starts a new episode in every env (returns FIRST everywhere)."""
    self._bufs['needs_reset'].fill_(1)
    return self.step(torch.zeros((self.num_envs, self.action_dim),
                                 dtype=torch.float64,
                                 device=self.batch.device))

  def step(self, actions) -> TimeStep:
    b = self.batch
    a = torch.as_tensor(actions, dtype=torch.float64,
                        device=b.device).reshape(self.num_envs,
                                                 self.action_dim).contiguous()
    P = lambda t: C.c_void_p(t.data_ptr())
    with torch.cuda.device(b.device):
      nat.check(nat.lib.pd_env_step(
          C.byref(b.lattice_tables.c), C.byref(b.c), C.byref(self.rate.c),
          C.byref(self.cfg), C.byref(self._buf_c), P(a), P(self._obs),
          P(self._reward), P(self._discount), P(self._step_type),
          C.c_void_p(torch.cuda.current_stream(b.device).cuda_stream)))
    return TimeStep(self._step_type, self._reward, self._discount, self._obs)
