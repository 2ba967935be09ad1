"""CPU oracle for whole goal-reaching episodes (BASELINE configs[4]).

TEST INFRASTRUCTURE ONLY (same rules as pdune_oracle.py).

Restates, for many envs at once, the loop `eval_lib.evaluate` runs for the
`greedy_on_neighbor` experiment (experiments/registry.py:287-298):

  putting_dune_environment.py:87-158   PuttingDuneEnvironment.reset / step
  goals.py:70-185                      SingleSiliconGoalReaching
  feature_constructors.py:157-228      SingleSiliconMaterialFrameFeatureConstructor
  agents/agent_lib.py:163-183          GreedyAgent.step (argmax = [1.42, 0])
  action_adapters.py:219-274           RelativeToSiliconMaterialFrameActionAdapter
  run_helpers.py:120-153               StepLimitWrapper(600)
  eval_lib.py:77-214                   evaluate / aggregate_results

Parity status: PINNED -- `tests/golden/episodes_reference.npz` holds the
EvalResults of the unmodified reference stack (run under oracle/refshim.py's
dm_env stand-in with InjectedRng, canonical neighbour order and agent wall time
excluded); this file reproduces reached_goal / num_actions / environment
seconds / reward for every episode of the fixture.

Conventions: the goal draw is draw 13 of the RESET stream (`rng.choice(n)` =
floor(u * n)); the agent's own generator is unused (position_noise_sigma = 0);
agent wall-clock time counts as zero.
"""

from __future__ import annotations

import dataclasses

import numpy as np

from oracle import pdune_oracle as po

GOAL_RANGE = (0.1, 50.0)  # goals.py:59
GOAL_RADIUS = po.BOND * 0.5  # goals.py:160
ARGMAX = np.asarray([1.42, 0.0])  # registry.py:289


@dataclasses.dataclass
class EpisodeConfig:
  rate_fn: int = po.RATE_SIMPLE
  dwell_us: int = 5000000  # registry.py:291-294
  image_duration_us: int = 2000000
  step_limit: int = 600  # run_helpers.py:34
  timeout_us: int = 600 * 1000000  # eval_lib.py:82


def observed_to_material(fov, q):
  """fov.microscope_frame_to_material_frame(ndarray): q * scale + ll."""
  return np.stack((q[..., 0] * (fov[..., 2] - fov[..., 0]) + fov[..., 0],
                   q[..., 1] * (fov[..., 3] - fov[..., 1]) + fov[..., 1]),
                  axis=-1)


def observe_site(state: po.OracleState, envs, sites):
  """Observed (normalised) coordinate of lattice `sites` [E] or [E, K]."""
  p = po.site_positions(state, sites, envs)
  f = state.fov[envs]
  if p.ndim == 3:
    f = f[:, None, :]
  return np.stack(((p[..., 0] - f[..., 0]) / (f[..., 2] - f[..., 0]),
                   (p[..., 1] - f[..., 1]) / (f[..., 3] - f[..., 1])), axis=-1)


def choose_goals(state: po.OracleState, envs=None, draw_index: int = 13):
  """goals.py:84-121 for every env (or `envs`), after reset.  Returns (goal
  site [E], goal position in the material frame [E, 2]); rows of envs not
  listed are zero.  `draw_index`: position of the goal draw on the RESET
  stream (13, or 15 after the two draws of DeltaPositionActionAdapter.reset)."""
  e = state.num_envs
  goal_site = np.zeros(e, dtype=np.int32)
  goal_pos = np.zeros((e, 2))
  u = po.draw_linear(state.seed, state.env_ids,
                     state.episode - np.uint32(1), po.STREAM_RESET,
                     draw_index)
  for i in (range(e) if envs is None else envs):
    q, _, sites = po.get_atoms_in_bounds(state, i)
    f = state.fov[i]
    q_si = q[sites == state.si_idx[i]].reshape(1, 2)
    scale = np.asarray([f[2] - f[0], f[3] - f[1]])
    d = scale * (q - q_si)
    dist = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
    valid = (dist < GOAL_RANGE[1]) & (dist > GOAL_RANGE[0])
    n = int(valid.sum())
    if n == 0:
      raise RuntimeError("Couldn't find any valid goals.")
    k = int(np.floor(u[i] * n))
    goal_site[i] = sites[valid][k]
    goal_pos[i] = observed_to_material(f, q[valid][k])
  return goal_site, goal_pos


def greedy_controls(state: po.OracleState, envs, goal_pos):
  """Features -> GreedyAgent.step -> RelativeToSiliconMaterialFrame adapter.
  Returns the control position in the microscope frame, float64 [n, 2]."""
  f = state.fov[envs]
  si = state.si_idx[envs]
  q_si = observe_site(state, envs, si)
  si_m = observed_to_material(f, q_si)
  nbr = state.nbr[si]  # canonical order (ascending site index when bonded)
  q_n = observe_site(state, envs, nbr)
  nbr_m = observed_to_material(f[:, None, :], q_n)
  deltas = (nbr_m - si_m[:, None, :]).astype(np.float32)
  goal_delta = (goal_pos[envs] - si_m).astype(np.float32)
  # agent_lib.py:163-183 in float32
  diff = deltas - goal_delta[:, None, :]
  scores = np.sqrt(diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1])
  best = np.argmin(scores, axis=1)
  bd = np.take_along_axis(deltas, best[:, None, None], axis=1)[:, 0, :]
  # float32 trigonometry, evaluated in float64 and rounded (the correctly
  # rounded float32 value up to ~1e-8 of the cases): libm / NumPy / CUDA
  # float32 routines differ from each other by an ulp now and then, their
  # float64 ones rounded to float32 do not -- the device does the same
  angle = np.arctan2(bd[:, 1].astype(np.float64),
                     bd[:, 0].astype(np.float64)).astype(np.float32)
  c = np.cos(angle.astype(np.float64)).astype(np.float32)
  s = np.sin(angle.astype(np.float64)).astype(np.float32)
  action = np.stack((ARGMAX[0] * c.astype(np.float64) +
                     ARGMAX[1] * (-s).astype(np.float64),
                     ARGMAX[0] * s.astype(np.float64) +
                     ARGMAX[1] * c.astype(np.float64)), axis=1)
  # action_adapters.py:231-256
  target = si_m + action
  ctl = np.stack(((target[:, 0] - f[:, 0]) / (f[:, 2] - f[:, 0]),
                  (target[:, 1] - f[:, 1]) / (f[:, 3] - f[:, 1])), axis=1)
  return np.clip(ctl, 0.0, 1.0)


def run_episodes(state: po.OracleState, cfg: EpisodeConfig, mlp=None) -> dict:
  """reset + goal + greedy control loop until goal, step limit or timeout."""
  e = state.num_envs
  po.reset(state)
  goal_site, goal_pos = choose_goals(state)
  env_time = np.full(e, cfg.image_duration_us, dtype=np.int64)  # eval_lib:121
  actions = np.zeros(e, dtype=np.int32)
  reached = np.zeros(e, dtype=bool)
  reward = np.zeros(e)
  active = env_time < cfg.timeout_us
  all_envs = np.arange(e)
  while active.any():
    idx = all_envs[active]
    ctl = np.full((e, 1, 2), 0.5)
    ctl[idx, 0] = greedy_controls(state, idx, goal_pos)
    # step only the active envs: inactive ones get a zero dwell, and their
    # state (FOV, counters) is restored afterwards.
    saved = (state.fov.copy(), state.ctrl_count.copy(),
             state.sim_time_us.copy())
    dwell = np.where(active, cfg.dwell_us, 0)[:, None]
    out = po.step_and_image(state, ctl, dwell, cfg.image_duration_us,
                            rate_fn=cfg.rate_fn, mlp=mlp)
    inact = ~active
    state.fov[inact] = saved[0][inact]
    state.ctrl_count[inact] = saved[1][inact]
    state.sim_time_us[inact] = saved[2][inact]
    el = out['elapsed_us']
    env_time[idx] += el[idx]
    actions[idx] += 1
    # goals.py:143-181 on the new observation
    q_si = observe_site(state, idx, state.si_idx[idx])
    si_m = observed_to_material(state.fov[idx], q_si)
    d = si_m - goal_pos[idx]
    dist = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
    term = dist < GOAL_RADIUS
    for j in np.nonzero(term)[0]:
      reward[idx[j]] = po.GAMMA_PER_SECOND ** (int(el[idx[j]]) / 10**6)
    reached[idx[term]] = True
    done = np.zeros(e, dtype=bool)
    done[idx[term]] = True
    done[idx] |= actions[idx] >= cfg.step_limit  # StepLimitWrapper truncation
    done[idx] |= ~(env_time[idx] < cfg.timeout_us)  # eval_lib.py:128
    active = active & ~done
  return {'reached': reached, 'num_actions': actions,
          'env_seconds': np.where(reached, env_time / 1e6, np.nan),
          'env_time_us': env_time, 'total_reward': reward,
          'goal_site': goal_site, 'goal_pos': goal_pos,
          'final_si': state.si_idx.copy()}


def aggregate(res: dict) -> dict:
  """eval_lib.py:187-214 aggregate_results."""
  r = res['reached']
  den = max(int(r.sum()), 1)
  return {
      'average_num_times_reached_goal': float(r.mean()),
      'average_num_actions_taken': float(res['num_actions'][r].sum()) / den,
      'average_environment_seconds_to_goal':
          float(np.nansum(res['env_seconds'][r])) / den,
      'average_total_reward': float(res['total_reward'][r].sum()) / den,
  }


def relative_to_silicon_controls(state: po.OracleState, actions: np.ndarray,
                                 max_distance: float = po.BOND) -> np.ndarray:
  """action_adapters.py:163-188 `RelativeToSiliconActionAdapter.get_action`
  for every env: clip(si_observed + clip(a, -1, 1) * max_distance / fov, 0, 1).
  actions: [E, 2] -> control positions [E, 2] in the microscope frame."""
  envs = np.arange(state.num_envs)
  a = np.clip(np.asarray(actions, dtype=np.float64), -1.0, 1.0)
  q_si = observe_site(state, envs, state.si_idx)
  f = state.fov
  radius = np.stack((max_distance / (f[:, 2] - f[:, 0]),
                     max_distance / (f[:, 3] - f[:, 1])), axis=1)
  return np.clip(q_si + a * radius, 0.0, 1.0)
