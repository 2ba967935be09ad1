"""GPU tests of whole episodes (BASELINE configs[4]) through the C ABI."""

import os

import numpy as np
import pytest
import torch

from oracle import pdune_oracle as po
from oracle import pdune_oracle_episode as oe
from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu


def _run(n, seed, rate_fn, env_offset=0, **cfg):
  from putting_dune_b200 import engine, episodes
  b = engine.EnvBatch(n, seed=seed, env_offset=env_offset)
  stats, goal_site, goal_xy = episodes.run_greedy_episodes(
      b, gh.rate_spec(rate_fn), episodes.EpisodeConfig(**cfg))
  return b, episodes.stats_to_numpy(stats), gh.np_(goal_site), gh.np_(goal_xy)


@pytest.mark.parametrize('name,rate_fn', [('simple', po.RATE_SIMPLE),
                                          ('prior', po.RATE_PRIOR)])
def test_episodes_match_reference_eval(golden_dir, name, rate_fn):
  """EvalResults of the unmodified reference's eval_lib.evaluate."""
  fix = np.load(os.path.join(golden_dir, 'episodes_reference.npz'))
  n = fix[f'{name}_reached'].shape[0]
  _, s, _, _ = _run(n, int(fix[f'{name}_philox_seed']), rate_fn)
  np.testing.assert_array_equal(s['reached_goal'].astype(bool),
                                fix[f'{name}_reached'])
  np.testing.assert_array_equal(s['num_actions'], fix[f'{name}_num_actions'])
  np.testing.assert_allclose(s['env_seconds'], fix[f'{name}_env_seconds'],
                             rtol=1e-6, equal_nan=True)
  np.testing.assert_allclose(s['total_reward'], fix[f'{name}_total_reward'],
                             rtol=1e-6)


def test_episodes_match_oracle_4096():
  from putting_dune_b200 import episodes
  n, seed = 4096, 99
  st = po.make_state(n, seed)
  want = oe.run_episodes(st, oe.EpisodeConfig(rate_fn=po.RATE_SIMPLE))
  b, s, goal_site, goal_xy = _run(n, seed, po.RATE_SIMPLE)
  np.testing.assert_array_equal(goal_site, want['goal_site'])
  np.testing.assert_allclose(goal_xy, want['goal_pos'], rtol=0, atol=1e-12)
  # The controller's float32 trigonometry is evaluated in float64 and rounded
  # on both sides (pd_episode.cuh, pdune_oracle_episode.py), so the control
  # positions agree to the bit except when a float64 result sits within
  # ~2 ulp of a float32 rounding boundary (~1e-8 per evaluation, 1e-3 for the
  # ~1e5 evaluations of this test): every episode must agree.
  same = (s['num_actions'] == want['num_actions']) & (
      s['reached_goal'].astype(bool) == want['reached'])
  assert same.all(), (int((~same).sum()), 'episodes differ')
  ok = same & want['reached']
  np.testing.assert_allclose(s['env_seconds'][ok], want['env_seconds'][ok],
                             rtol=1e-6)
  np.testing.assert_allclose(s['total_reward'][ok], want['total_reward'][ok],
                             rtol=1e-6)
  np.testing.assert_array_equal(gh.np_(b.si_idx)[same], want['final_si'][same])
  agg = episodes.aggregate_results(s)
  ref = oe.aggregate(want)
  for k in ref:
    assert abs(agg[k] - ref[k]) <= 2e-3 * max(1.0, abs(ref[k])), k
  assert agg['average_num_times_reached_goal'] > 0.9


def test_episode_limits():
  # step limit (StepLimitWrapper truncation): not reached, actions == limit
  _, s, _, _ = _run(256, 3, po.RATE_PRIOR, step_limit=7)
  missed = ~s['reached_goal'].astype(bool)
  assert missed.mean() > 0.9 and (s['num_actions'][missed] == 7).all()
  assert (s['num_actions'] <= 7).all()
  assert np.isnan(s['env_seconds'][missed]).all()
  assert (s['total_reward'][missed] == 0).all()
  # simulated-time limit: 2 s + k * (5 + 2 [+2]) s must pass 30 s
  import datetime as dt
  _, s, _, _ = _run(256, 3, po.RATE_PRIOR, timeout=dt.timedelta(seconds=30))
  missed = ~s['reached_goal'].astype(bool)
  assert np.isin(s['num_actions'][missed], (4, 5)).all(), np.unique(
      s['num_actions'])


def test_sharded_episodes_equal_unsharded():
  n, seed = 2048, 17
  _, full, _, _ = _run(n, seed, po.RATE_SIMPLE)
  parts = [_run(n // 4, seed, po.RATE_SIMPLE, env_offset=r * (n // 4))[1]
           for r in range(4)]
  np.testing.assert_array_equal(np.concatenate(parts).tobytes(),
                                full.tobytes())


def test_goal_selection_row_path_equals_scan(monkeypatch):
  """k_choose_goal's row-analytic path (one lane per lattice row: the in-view
  run from the FOV bounds, exact test at the run's ends) picks the same goal
  site as the exhaustive scan for every env -- at full batch size, on the
  default lattice and on a small sheet that the FOV overhangs."""
  from putting_dune_b200 import engine, episodes
  for n, cols in ((300000, 50), (20000, 14)):
    out = []
    for flag in ('0', '1'):
      monkeypatch.setenv('PD_GOAL_SCAN', flag)
      b = engine.EnvBatch(n, seed=7, lattice=engine.Lattice(cols))
      stats, goal_site, goal_xy = episodes.run_greedy_episodes(
          b, gh.rate_spec(po.RATE_SIMPLE),
          episodes.EpisodeConfig(step_limit=3))
      out.append((gh.np_(goal_site), gh.np_(goal_xy),
                  episodes.stats_to_numpy(stats)))
    (site_a, xy_a, st_a), (site_b, xy_b, st_b) = out
    np.testing.assert_array_equal(site_a, site_b)
    np.testing.assert_array_equal(xy_a, xy_b)
    np.testing.assert_array_equal(st_a['num_actions'], st_b['num_actions'])
    assert (site_a >= 0).all()
    assert len(np.unique(site_a)) > 100
