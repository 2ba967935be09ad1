"""GPU tests of the batched RL environment layer (pd_env_step) against the
oracle and the TimeSteps of the unmodified reference environment."""

import os

import numpy as np
import pytest

from oracle import pdune_oracle as po
from oracle import pdune_oracle_env as oenv
from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu

ENV_CASES = ((2, 0, 1.5, 1.5, 1.42, 600), (2, 1, 1.0, 5.0, 2.84, 600),
             (3, 1, 5.0, 5.0, 2.84, 7), (0, 0, 1.5, 1.5, 1.42, 600),
             (1, 0, 1.5, 1.5, 1.42, 600))
ADAPTER_NAMES = ['direct', 'delta_position', 'relative_to_silicon',
                 'relative_to_silicon_material_frame']
FEATURE_NAMES = ['microscope_frame', 'material_frame']


def _make(case, n, seed, rate_fn=po.RATE_SIMPLE):
  import datetime as dt
  import putting_dune_b200 as pd
  ad, ft, d0, d1, md, lim = case
  return pd.BatchedPuttingDuneEnvironment(
      n, rate=gh.rate_spec(rate_fn), action_adapter=ADAPTER_NAMES[ad],
      feature_constructor=FEATURE_NAMES[ft],
      dwell_time_range=(dt.timedelta(seconds=d0), dt.timedelta(seconds=d1)),
      max_distance_angstroms=md, step_limit=lim, seed=seed)


@pytest.mark.parametrize('case', range(5))
def test_env_matches_reference_timesteps(golden_dir, case):
  fix = np.load(os.path.join(golden_dir, 'env_reference.npz'))
  acts = fix[f'actions_{case}']
  env = _make(ENV_CASES[case], acts.shape[1], int(fix['seed']))
  assert env.action_dim == acts.shape[2]
  for t in range(acts.shape[0]):
    ts = env.step(acts[t])
    np.testing.assert_array_equal(gh.np_(ts.step_type),
                                  fix[f'step_type_{case}'][t])
    np.testing.assert_allclose(gh.np_(ts.reward), fix[f'reward_{case}'][t],
                               rtol=1e-6)
    np.testing.assert_allclose(gh.np_(ts.discount), fix[f'discount_{case}'][t],
                               rtol=1e-6)
    np.testing.assert_allclose(gh.np_(ts.observation),
                               fix[f'observation_{case}'][t], rtol=0,
                               atol=2e-6)


@pytest.mark.parametrize('case,rate_fn', [(0, po.RATE_SIMPLE),
                                          (2, po.RATE_SIMPLE),
                                          (1, po.RATE_PRIOR),
                                          (4, po.RATE_SIMPLE)])
def test_env_matches_oracle_with_episode_turnover(case, rate_fn):
  """2048 envs, greedy-ish actions so that goals are reached: termination,
  reward, auto-reset and the step limit all occur."""
  ad, ft, d0, d1, md, lim = ENV_CASES[case]
  n, seed, t_steps = 2048, 99, 40
  cfg = oenv.EnvConfig(adapter=ad, features=ft, min_dwell_s=d0, max_dwell_s=d1,
                       max_distance=md, step_limit=lim, rate_fn=rate_fn)
  ref = oenv.OracleEnv(n, seed, cfg)
  env = _make(ENV_CASES[case], n, seed, rate_fn)
  rng = np.random.default_rng(0)
  obs = None
  n_last = n_term = 0
  for t in range(t_steps):
    acts = rng.uniform(-1, 1, size=(n, cfg.action_dim))
    if obs is not None and ad in (2, 3):
      # steer towards the goal: unit vector of the goal delta (features 8, 9)
      g = obs[:, 8:10].astype(np.float64)
      g /= np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-9)
      acts[:, :2] = g * (1.0 if ad == 2 else 1.42)
    if cfg.action_dim == 3:
      acts[:, 2] = rng.uniform(-0.2, 1.2, size=n)
    want = ref.step(acts)
    ts = env.step(acts)
    np.testing.assert_array_equal(gh.np_(ts.step_type), want['step_type'])
    np.testing.assert_allclose(gh.np_(ts.reward), want['reward'], rtol=1e-6)
    np.testing.assert_allclose(gh.np_(ts.discount), want['discount'],
                               rtol=1e-6)
    np.testing.assert_allclose(gh.np_(ts.observation), want['observation'],
                               rtol=0, atol=2e-6)
    obs = want['observation']
    n_last += int((want['step_type'] == 2).sum())
    n_term += int((want['reward'] > 0).sum())
  np.testing.assert_array_equal(gh.np_(env.batch.si_idx), ref.state.si_idx)
  if ad in (2, 3) and rate_fn == po.RATE_SIMPLE:
    assert n_term > 100  # goals were reached and rewarded
  if lim < 600:
    assert n_last > n_term  # step-limit truncations too


def test_env_reset_semantics():
  # putting_dune_environment_test.py:99-134
  env = _make(ENV_CASES[0], 64, 5)
  first = env.step(np.zeros((64, 2)))
  assert (gh.np_(first.step_type) == 0).all()  # a fresh env resets first
  assert np.allclose(gh.np_(first.discount), 0.9967 ** 2.0, rtol=1e-6)
  assert (gh.np_(first.reward) == 0).all()
  mid = env.step(np.zeros((64, 2)))
  assert (gh.np_(mid.step_type) == 1).all()
  assert gh.np_(mid.observation).shape == (64, 10)
  again = env.reset()
  assert (gh.np_(again.step_type) == 0).all()
  assert (gh.np_(env.batch.episode) == 2).all()


def test_env_with_learned_rates_skips_resetting_envs():
  import putting_dune_b200 as pd
  mlp = po.MlpParams.synthetic(3, hidden=(32, 32))
  w = pd.MlpWeights(**{k: getattr(mlp, k) for k in pd.MlpWeights.NAMES})
  env = pd.BatchedPuttingDuneEnvironment(
      500, rate=pd.RateSpec(po.RATE_LEARNED, mlp=w), step_limit=3, seed=2)
  types = [gh.np_(env.step(np.zeros((500, 2))).step_type).copy()
           for _ in range(9)]
  # FIRST, MID, MID, LAST(limit), FIRST, ...
  assert [int(t[0]) for t in types[:5]] == [0, 1, 1, 2, 0]
  assert (gh.np_(env.batch.ctrl_count) == 6).all()
