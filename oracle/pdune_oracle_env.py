"""CPU oracle for the batched RL environment layer.

TEST INFRASTRUCTURE ONLY (same rules as pdune_oracle.py).

Restates, for many envs at once, `PuttingDuneEnvironment` wrapped in
`StepLimitWrapper`:

  putting_dune_environment.py:72-158  seed / reset / step (dm_env semantics)
  run_helpers.py:120-153              StepLimitWrapper
  action_adapters.py:53-274           Direct, DeltaPosition, RelativeToSilicon,
                                      RelativeToSiliconMaterialFrame adapters
  feature_constructors.py:79-228      the two 10-float feature constructors
  goals.py:70-185                     SingleSiliconGoalReaching

Parity status: PINNED -- tests/golden/env_reference.npz holds TimeStep
sequences of the unmodified reference environment (oracle/refrun.py
`run_reference_env_stack`, InjectedRng, canonical neighbour order); this file
reproduces step types, rewards, discounts bit-exactly and observations to
float32 rounding.

Draw order on the RESET stream per episode: 13 simulator draws (pdune_oracle
.reset), then DeltaPositionActionAdapter.reset's two uniforms (if that adapter
is used), then the goal draw.
"""

from __future__ import annotations

import dataclasses

import numpy as np

from oracle import pdune_oracle as po
from oracle import pdune_oracle_episode as oe

ADAPTER_DIRECT = 0  # action_adapters.py:53-84
ADAPTER_DELTA = 1  # :87-128
ADAPTER_RELATIVE = 2  # :131-216
ADAPTER_RELATIVE_MATERIAL = 3  # :219-274
FEATURES_MICROSCOPE = 0  # feature_constructors.py:79-154
FEATURES_MATERIAL = 1  # :157-228
STEP_FIRST, STEP_MID, STEP_LAST = 0, 1, 2


@dataclasses.dataclass
class EnvConfig:
  adapter: int = ADAPTER_RELATIVE
  features: int = FEATURES_MICROSCOPE
  min_dwell_s: float = 1.5
  max_dwell_s: float = 1.5
  max_distance: float = po.BOND
  image_duration_us: int = 2000000
  step_limit: int = 600
  rate_fn: int = po.RATE_SIMPLE

  @property
  def action_dim(self) -> int:
    relative = self.adapter in (ADAPTER_RELATIVE, ADAPTER_RELATIVE_MATERIAL)
    return 3 if relative and self.min_dwell_s != self.max_dwell_s else 2


class OracleEnv:
  """Batched PuttingDuneEnvironment + StepLimitWrapper."""

  def __init__(self, num_envs: int, seed: int, cfg: EnvConfig, mlp=None,
               gmm=None, env_offset: int = 0):
    self.cfg, self.mlp, self.gmm = cfg, mlp, gmm
    self.state = po.make_state(num_envs, seed, env_offset=env_offset)
    e = num_envs
    self.goal_pos = np.zeros((e, 2))
    self.beam_pos = np.zeros((e, 2))
    self.elapsed_steps = np.zeros(e, dtype=np.int32)
    self.needs_reset = np.ones(e, dtype=bool)  # _requires_reset = True

  # -- helpers --------------------------------------------------------------
  def _reset_envs(self, idx: np.ndarray) -> None:
    st = self.state
    mask = np.zeros(st.num_envs, dtype=bool)
    mask[idx] = True
    po.reset(st, mask)
    k = 13
    if self.cfg.adapter == ADAPTER_DELTA:  # rng.uniform(0, 1, size=2)
      ep = st.episode[idx] - np.uint32(1)
      self.beam_pos[idx, 0] = po.draw_linear(st.seed, st.env_ids[idx], ep,
                                             po.STREAM_RESET, 13)
      self.beam_pos[idx, 1] = po.draw_linear(st.seed, st.env_ids[idx], ep,
                                             po.STREAM_RESET, 14)
      k = 15
    _, pos = oe.choose_goals(st, envs=idx, draw_index=k)
    self.goal_pos[idx] = pos[idx]
    self.elapsed_steps[idx] = 0
    self.needs_reset[idx] = False

  def features(self, idx: np.ndarray) -> np.ndarray:
    st, f = self.state, self.state.fov[idx]
    si = st.si_idx[idx]
    q_si = oe.observe_site(st, idx, si)
    q_n = oe.observe_site(st, idx, st.nbr[si])
    si_m = oe.observed_to_material(f, q_si)
    goal_delta = self.goal_pos[idx] - si_m  # feature_constructors.py:60-76
    if self.cfg.features == FEATURES_MICROSCOPE:
      d = q_n - q_si[:, None, :]
      dist = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])
      body = np.concatenate((q_si, (d / dist[..., None]).reshape(-1, 6)),
                            axis=1)
    else:
      nbr_m = oe.observed_to_material(f[:, None, :], q_n)
      body = np.concatenate(
          (si_m, (nbr_m - si_m[:, None, :]).reshape(-1, 6)), axis=1)
    return np.concatenate((body, goal_delta), axis=1).astype(np.float32)

  def _controls(self, idx: np.ndarray, actions: np.ndarray):
    """Adapter: actions -> (control position [n, 2], dwell_us [n])."""
    cfg, st = self.cfg, self.state
    a = np.asarray(actions, dtype=np.float64)
    fixed = np.full(idx.size, 1500000, dtype=np.int64)
    if cfg.adapter == ADAPTER_DIRECT:
      return np.clip(a[:, :2], 0.0, 1.0), fixed
    if cfg.adapter == ADAPTER_DELTA:
      self.beam_pos[idx] = np.clip(self.beam_pos[idx] + a[:, :2], 0.0, 1.0)
      return self.beam_pos[idx].copy(), fixed
    f = st.fov[idx]
    q_si = oe.observe_site(st, idx, st.si_idx[idx])
    if cfg.adapter == ADAPTER_RELATIVE:
      radius = np.stack((cfg.max_distance / (f[:, 2] - f[:, 0]),
                         cfg.max_distance / (f[:, 3] - f[:, 1])), axis=1)
      ctl = np.clip(q_si + np.clip(a[:, :2], -1.0, 1.0) * radius, 0.0, 1.0)
    else:
      si_m = oe.observed_to_material(f, q_si)
      t = si_m + a[:, :2]
      ctl = np.clip(np.stack(((t[:, 0] - f[:, 0]) / (f[:, 2] - f[:, 0]),
                              (t[:, 1] - f[:, 1]) / (f[:, 3] - f[:, 1])),
                             axis=1), 0.0, 1.0)
    if cfg.min_dwell_s == cfg.max_dwell_s:
      dwell = np.full(idx.size, po.seconds_to_us(np.float64(cfg.min_dwell_s)),
                      dtype=np.int64)
    else:
      frac = np.clip(a[:, 2], 0.0, 1.0)
      secs = frac * (cfg.max_dwell_s - cfg.min_dwell_s) + cfg.min_dwell_s
      dwell = po.seconds_to_us(secs)
    return ctl, dwell

  # -- dm_env API -----------------------------------------------------------
  def step(self, actions: np.ndarray) -> dict:
    """One `env.step(action)` per env; envs whose previous step was LAST (or
    that were never reset) reset instead and return FIRST."""
    cfg, st = self.cfg, self.state
    e = st.num_envs
    resetting = self.needs_reset | (self.elapsed_steps == -1)
    step_type = np.full(e, STEP_MID, dtype=np.int32)
    reward = np.zeros(e, dtype=np.float32)
    discount = np.zeros(e, dtype=np.float32)
    ridx = np.nonzero(resetting)[0]
    sidx = np.nonzero(~resetting)[0]
    controls = np.full((e, 1, 2), 0.5)
    dwell = np.zeros((e, 1), dtype=np.int64)
    if sidx.size:
      ctl, dw = self._controls(sidx, np.asarray(actions)[sidx])
      controls[sidx, 0] = ctl
      dwell[sidx, 0] = dw
    if ridx.size:
      self._reset_envs(ridx)
      step_type[ridx] = STEP_FIRST
      discount[ridx] = po.GAMMA_PER_SECOND ** (cfg.image_duration_us / 1e6)
    if sidx.size:
      out = po.step_and_image(st, controls, dwell, cfg.image_duration_us,
                              rate_fn=cfg.rate_fn, mlp=self.mlp, gmm=self.gmm,
                              skip=resetting)
      el = out['elapsed_us'][sidx]
      q_si = oe.observe_site(st, sidx, st.si_idx[sidx])
      si_m = oe.observed_to_material(st.fov[sidx], q_si)
      d = si_m - self.goal_pos[sidx]
      term = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) < oe.GOAL_RADIUS
      gamma = np.asarray([po.GAMMA_PER_SECOND ** (int(x) / 10**6)
                          for x in el])
      reward[sidx] = np.where(term, gamma, 0.0)
      discount[sidx] = np.where(term, 0.0, gamma)
      step_type[sidx[term]] = STEP_LAST
      self.needs_reset[sidx[term]] = True
      # StepLimitWrapper (run_helpers.py:133-153)
      self.elapsed_steps[sidx] += 1
      trunc = self.elapsed_steps[sidx] >= cfg.step_limit
      self.elapsed_steps[sidx[trunc]] = -1
      step_type[sidx[trunc]] = STEP_LAST
    obs = self.features(np.arange(e))
    return {'step_type': step_type, 'reward': reward, 'discount': discount,
            'observation': obs}
