"""The guarded float32 iteration (csrc/pd_fast.cuh, pd_step_fast.cu).

Rollouts on the prior / simple rates decide every KMC iteration of
graphene.py:658-694 in float32 when an error bound allows it and replay the
control with the float64 code otherwise.  These tests hold that claim to
'bit for bit': against the float64 kernels of the same library
(pd_set_fast_path(0)), against the oracle at the benchmarked configurations
(BASELINE configs[1]: 4096 envs x 256 steps; 1 Mi envs x 8 steps), and through
pd_fast_path_audit, which measures the float32 quantities against the bounds
the decisions assume.
"""

import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import pdune_oracle as po
from oracle import pdune_oracle_episode as oe
from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu

STATE_KEYS = ('si_idx', 'fov', 'ctrl_count', 'sim_time_us', 'n_events',
              'n_transitions', 'status')


@pytest.fixture(scope='module')
def eng():
  from putting_dune_b200 import engine
  return engine


@pytest.fixture
def nat():
  from putting_dune_b200 import _native
  yield _native
  _native.lib.pd_set_fast_path(1)
  _native.lib.pd_set_option(b'plan', 0)
  _native.lib.pd_set_option(b'walk_plan', 1)


def _select(nat, kernels):
  """'plan': k_rollout_plan (small batches) / k_walk_plan (large batches) where
  they apply; 'fast': k_rollout_fast / k_walk_fast only; 'exact': the float64
  kernels.  (The default is k_rollout_fast for small, k_walk_plan for large
  batches.)"""
  nat.check(nat.lib.pd_set_option(b'fast_path', 0 if kernels == 'exact' else 1))
  nat.check(nat.lib.pd_set_option(b'plan', 1 if kernels == 'plan' else 0))
  # (2: k_walk_plan whatever the batch size and the number of steps)
  nat.check(nat.lib.pd_set_option(b'walk_plan', 2 if kernels == 'plan' else 0))


def _rollout(eng, nat, kernels, n, seed, ctl, dwell, spec, mode, shift_fov,
             env_offset=0):
  _select(nat, kernels)
  b = eng.EnvBatch(n, seed=seed, env_offset=env_offset)
  b.reset()
  if shift_fov:
    # a FOV the Si is not centred in: the t = 0 safe-area check and the
    # relative adapter's clip to the frame matter
    b.fov[::3] += 3.9
    b.fov[1::7] -= 9.0
  si, el = b.rollout(ctl, dwell, spec, record=True, action_mode=mode,
                     max_distance_angstroms=1.42)
  sd = b.state_dict()
  return gh.np_(si), gh.np_(el), {k: gh.np_(sd[k]) for k in STATE_KEYS}


@pytest.mark.parametrize('n', [37, 700, 4096, 21000, 300000])
def test_fast_rollout_equals_float64_kernels(eng, nat, n):
  """k_rollout_fast (small batches) and k_walk_fast (large ones) return what
  the float64 kernels return: per-step Si site and elapsed time, FOV, clocks,
  event / transition counts, Philox control counter, status bits."""
  t_steps, seed = (45 if n < 100000 else 9), 77
  rng = np.random.default_rng(n)
  acts = rng.uniform(-1.2, 1.2, size=(t_steps, n, 2))
  direct = 0.5 + rng.uniform(-0.1, 0.1, size=(t_steps, n, 2))
  direct[0, :5] = np.nan           # non-finite controls: exact replay
  direct[1, 5:10] = 1e9
  direct[2, 10:15] = np.inf
  far = 0.5 + rng.uniform(-0.6, 0.6, size=(t_steps, n, 2))  # beam off the Si
  # non-finite actions under the relative adapter: np.clip keeps a NaN, so the
  # control is NaN and the rate function flags it (graphene.py:258); +-inf
  # clip to +-1
  acts_nan = acts.copy()
  acts_nan[0, :3] = np.nan
  acts_nan[min(3, t_steps - 1), 3:6, 0] = np.nan
  acts_nan[1, 6:9] = np.inf
  acts_nan[2, 9:12] = -np.inf
  cases = [
      (po.RATE_PRIOR, 1500000, acts, nat.ACTION_RELATIVE_TO_SILICON, True),
      (po.RATE_PRIOR, 1500000, acts, nat.ACTION_RELATIVE_TO_SILICON, False),
      (po.RATE_SIMPLE, 5000000, acts, nat.ACTION_RELATIVE_TO_SILICON, True),
      (po.RATE_SIMPLE, 700000, direct, nat.ACTION_DIRECT, True),
      (po.RATE_PRIOR, 30000000, direct, nat.ACTION_DIRECT, False),
      (po.RATE_PRIOR, 1500000, far, nat.ACTION_DIRECT, False),
      (po.RATE_SIMPLE, 3, acts, nat.ACTION_RELATIVE_TO_SILICON, False),
      (po.RATE_PRIOR, 1500000, acts_nan, nat.ACTION_RELATIVE_TO_SILICON,
       False),
  ]
  for rate_fn, dwell, ctl, mode, shift in cases:
    spec = gh.rate_spec(rate_fn)
    si_b, el_b, st_b = _rollout(eng, nat, 'exact', n, seed, ctl, dwell, spec,
                                mode, shift)
    for kernels in ('plan', 'fast'):
      si_a, el_a, st_a = _rollout(eng, nat, kernels, n, seed, ctl, dwell,
                                  spec, mode, shift)
      np.testing.assert_array_equal(si_a, si_b, err_msg=kernels)
      np.testing.assert_array_equal(el_a, el_b, err_msg=kernels)
      for k in STATE_KEYS:
        np.testing.assert_array_equal(st_a[k], st_b[k],
                                      err_msg=f'{kernels} {k}')
    if dwell > 1000:
      assert st_b['n_transitions'].sum() > 0
    if ctl is acts_nan:
      # PD_ENV_BAD_RATE (1) on the envs that saw a NaN action, nowhere else
      bad = (st_b['status'] & 1) != 0
      assert bad[:6].all() and not bad[6:].any()


def test_fast_rollout_single_step_and_edge_sites(eng, nat):
  """One-step rollouts (the fast walk kernel without look-ahead) and a Si that
  starts on the sheet's edge, where the three nearest sites are not the three
  bonded ones (geometry class 2: tables instead of the bulk constants)."""
  seed = 5
  rng = np.random.default_rng(3)
  cls = None
  # (600 steps: three 256-step chunks of k_rollout_plan, the last one ragged)
  for n, t_steps, rate_fn in ((5000, 1, po.RATE_SIMPLE),
                              (5000, 40, po.RATE_SIMPLE),
                              (5000, 60, po.RATE_PRIOR),
                              (301, 600, po.RATE_PRIOR),
                              # k_walk_plan with half of the envs handed over
                              # to k_walk_fast<LIST>; a ragged last CTA, two
                              # chunks
                              (200001, 21, po.RATE_PRIOR)):
    spec = gh.rate_spec(rate_fn)
    acts = rng.uniform(-1, 1, size=(t_steps, n, 2))
    res = []
    for kernels in ('plan', 'fast', 'exact'):
      _select(nat, kernels)
      b = eng.EnvBatch(n, seed=seed)
      b.reset()
      if cls is None:
        cls = gh.np_(b.lattice_tables.nbr)[:, 3] >> 24
      edge = np.nonzero(cls == 2)[0]
      assert edge.size > 50
      # park every second env's Si on an edge site, FOV centred on it
      sites = edge[np.arange((n + 1) // 2) % edge.size]
      b.si_idx[::2] = torch.as_tensor(sites.astype(np.int32), device=b.device)
      xy = gh.np_(b.silicon_position())
      half = gh.np_(b.fov_scale)[:, None] / 2
      b.fov.copy_(torch.as_tensor(np.concatenate([xy - half, xy + half], 1),
                                  device=b.device))
      si, el = b.rollout(acts, 5000000, spec, record=True,
                         action_mode=nat.ACTION_RELATIVE_TO_SILICON)
      sd = b.state_dict()
      res.append((gh.np_(si), gh.np_(el),
                  {k: gh.np_(sd[k]) for k in STATE_KEYS}))
    for other in res[:2]:
      np.testing.assert_array_equal(other[0], res[2][0])
      np.testing.assert_array_equal(other[1], res[2][1])
      for k in STATE_KEYS:
        np.testing.assert_array_equal(other[2][k], res[2][2][k], err_msg=k)
    assert res[2][2]['n_transitions'][::2].sum() > 100 * (t_steps > 1)


@pytest.mark.parametrize('kernels', ['fast', 'plan'])
@pytest.mark.parametrize('rate_fn', [po.RATE_PRIOR, po.RATE_SIMPLE])
def test_benchmarked_config_vs_oracle(eng, nat, rate_fn, kernels):
  """BASELINE configs[1] exactly as bench.py runs it -- 4096 envs x 256 beam
  steps, relative_random actions, dwell 1.5 s -- against the oracle's
  step_and_image for all 1,048,576 env-steps: Si site and elapsed
  microseconds of every step, final FOV, clocks and counters.  'fast' is
  the kernel bench.py times (k_rollout_fast), 'plan' the opt-in
  k_rollout_plan."""
  n, t_steps, seed = 4096, 256, 0
  rng = np.random.default_rng(100)
  acts = rng.uniform(-1.0, 1.0, size=(t_steps, n, 2))
  st = po.make_state(n, seed)
  po.reset(st)
  _select(nat, kernels)
  b = gh.batch_from_oracle(st)
  si, el = b.rollout(acts, 1500000, gh.rate_spec(rate_fn), record=True,
                     action_mode=nat.ACTION_RELATIVE_TO_SILICON,
                     max_distance_angstroms=1.42)
  si, el = gh.np_(si), gh.np_(el)
  for t in range(t_steps):
    ctl = oe.relative_to_silicon_controls(st, acts[t])[:, None, :]
    want = po.step_and_image(st, ctl, 1500000, rate_fn=rate_fn)
    np.testing.assert_array_equal(si[t], st.si_idx, err_msg=f'step {t}')
    np.testing.assert_array_equal(el[t], want['elapsed_us'],
                                  err_msg=f'step {t}')
  np.testing.assert_array_equal(gh.np_(b.n_transitions), st.n_transitions)
  np.testing.assert_array_equal(gh.np_(b.n_events), st.n_events)
  np.testing.assert_array_equal(gh.np_(b.sim_time_us), st.sim_time_us)
  np.testing.assert_allclose(gh.np_(b.fov), st.fov, rtol=0, atol=1e-13)
  assert st.n_transitions.sum() > 50000


def test_at_scale_config_vs_oracle_sample(eng, nat):
  """bench.py's at_scale case: 1 Mi envs x 8 steps through k_walk_fast; a
  block of 65,536 of those envs (Philox is keyed by the global env id)
  against the oracle, every step."""
  n, t_steps, seed, lo, m = 1 << 20, 8, 1, 413696, 65536
  rng = np.random.default_rng(17)
  acts = rng.uniform(-1.0, 1.0, size=(t_steps, n, 2))
  nat.lib.pd_set_fast_path(1)
  b = eng.EnvBatch(n, seed=seed)
  b.reset()
  st = po.make_state(m, seed, env_offset=lo)
  po.reset(st)
  np.testing.assert_array_equal(gh.np_(b.si_idx[lo:lo + m]), st.si_idx)
  # identical transform / FOV on both sides (reset's sincos differs by an ulp)
  b.lattice[lo:lo + m] = torch.as_tensor(st.lattice, device=b.device)
  b.fov[lo:lo + m] = torch.as_tensor(st.fov, device=b.device)
  si, el = b.rollout(acts, 1500000, gh.rate_spec(po.RATE_PRIOR), record=True,
                     action_mode=nat.ACTION_RELATIVE_TO_SILICON,
                     max_distance_angstroms=1.42)
  si, el = gh.np_(si[:, lo:lo + m]), gh.np_(el[:, lo:lo + m])
  for t in range(t_steps):
    ctl = oe.relative_to_silicon_controls(st, acts[t, lo:lo + m])[:, None, :]
    want = po.step_and_image(st, ctl, 1500000, rate_fn=po.RATE_PRIOR)
    np.testing.assert_array_equal(si[t], st.si_idx, err_msg=f'step {t}')
    np.testing.assert_array_equal(el[t], want['elapsed_us'])
  np.testing.assert_array_equal(gh.np_(b.n_events[lo:lo + m]), st.n_events)
  assert st.n_transitions.sum() > 30000


@pytest.mark.parametrize('rate_fn,dwell,dist', [
    (po.RATE_PRIOR, 1500000, 1.42), (po.RATE_SIMPLE, 1500000, 1.42),
    (po.RATE_PRIOR, 5000000, 1.42), (po.RATE_SIMPLE, 200000, 1.42),
    (po.RATE_PRIOR, 1500000, 4.0), (po.RATE_SIMPLE, 60000000, 6.0)])
def test_fast_path_audit(eng, nat, rate_fn, dwell, dist):
  """5e7 random iterations per case: no decided iteration differs from the
  float64 code, the exact waiting time always lies inside the float32
  interval, and the largest observed errors stay well below the bounds the
  decisions assume (the printed ratios are the safety factors DESIGN.md
  quotes)."""
  lat = eng.Lattice(50)
  out = nat.PdFastAudit()
  nat.check(nat.lib.pd_fast_path_audit(C.byref(lat.c), rate_fn, 12345,
                                       50_000_000, dwell, dist, C.byref(out),
                                       None))
  print(f'\naudit rate={rate_fn} dwell={dwell} dist={dist}: '
        f'no_hop={out.no_hop} hop={out.hop} unsure={out.unsure} '
        f'tot={out.total_rate_error_over_bound:.3f} '
        f't={out.waiting_time_error_over_bound:.3f} '
        f'choice={out.choice_error_over_bound:.3f} '
        f'draw={out.draw_error_over_bound:.3f}')
  assert out.samples == 50_000_000
  assert out.wrong_decision == 0 and out.wrong_slot == 0
  assert out.waiting_time_outside_bounds == 0
  assert out.total_rate_error_over_bound < 0.5
  assert out.waiting_time_error_over_bound < 0.6
  assert out.choice_error_over_bound < 0.5
  assert out.draw_error_over_bound < 0.5
  assert out.unsure < 0.01 * out.samples
  assert out.hop > 0.01 * out.samples


@pytest.mark.parametrize('n,t_steps', [(4096, 256), (37, 19), (300000, 5)])
def test_packed_host_rollout(eng, nat, n, t_steps):
  """pd_rollout_actions_host_packed (float32 actions in, uint16 Si site |
  re-centred << 15 out, chunked copy-engine pipeline) against the
  device-resident rollout of the same actions: sites, elapsed times and the
  whole env state."""
  rng = np.random.default_rng(n)
  acts = rng.uniform(-1.1, 1.1, size=(t_steps, n, 2)).astype(np.float32)
  for rate_fn, dwell in ((po.RATE_PRIOR, 1500000), (po.RATE_SIMPLE, 5000000)):
    spec = gh.rate_spec(rate_fn)
    a = eng.EnvBatch(n, seed=3)
    b = eng.EnvBatch(n, seed=3)
    a.reset()
    b.reset()
    a.fov[::5] += 4.1
    b.fov[::5] += 4.1
    si, el = a.rollout(acts.astype(np.float64), dwell, spec, record=True,
                       action_mode=nat.ACTION_RELATIVE_TO_SILICON)
    packed = b.rollout_host_packed(acts, dwell, spec,
                                   action_mode=nat.ACTION_RELATIVE_TO_SILICON)
    # twice through the same pinned result buffer (staging re-use)
    c = eng.EnvBatch(n, seed=3)
    c.reset()
    c.fov[::5] += 4.1
    packed2 = c.rollout_host_packed(acts, dwell, spec, out=packed.clone(),
                                    action_mode=nat.ACTION_RELATIVE_TO_SILICON)
    for pk in (packed, packed2):
      si_p, el_p = eng.EnvBatch.unpack_rollout(pk, dwell)
      np.testing.assert_array_equal(si_p.numpy(), gh.np_(si))
      np.testing.assert_array_equal(el_p.numpy(), gh.np_(el))
    sa, sb = a.state_dict(), b.state_dict()
    for k in STATE_KEYS:
      np.testing.assert_array_equal(gh.np_(sa[k]), gh.np_(sb[k]), err_msg=k)
    assert (gh.np_(el) > dwell + 2000000).any()  # some step re-centred


def test_fast_rollout_consecutive_calls(eng, nat):
  """Four consecutive 256-step rollouts of the same 4096 envs (1024 steps: a
  few envs walk all the way to the sheet's edge, FOVs re-centre many times):
  after every call the fast and the float64 kernels agree on every output and
  on the env state."""
  n, t_steps = 4096, 256
  rng = np.random.default_rng(5)
  acts = [rng.uniform(-1, 1, size=(t_steps, n, 2)).astype(np.float32).astype(
      np.float64) for _ in range(4)]
  spec = gh.rate_spec(po.RATE_PRIOR)
  runs = []
  for kernels in ('plan', 'fast', 'exact'):
    _select(nat, kernels)
    b = eng.EnvBatch(n, seed=0)
    b.reset()
    out = []
    for a in acts:
      si, el = b.rollout(a, 1500000, spec, record=True,
                         action_mode=nat.ACTION_RELATIVE_TO_SILICON)
      sd = b.state_dict()
      out.append((gh.np_(si), gh.np_(el),
                  {k: gh.np_(sd[k]).copy() for k in STATE_KEYS}))
    runs.append(out)
  for run in runs[:2]:
    for i, (x, y) in enumerate(zip(run, runs[2])):
      np.testing.assert_array_equal(x[0], y[0], err_msg=f'call {i} si')
      np.testing.assert_array_equal(x[1], y[1], err_msg=f'call {i} elapsed')
      for k in STATE_KEYS:
        np.testing.assert_array_equal(x[2][k], y[2][k],
                                      err_msg=f'call {i} {k}')


def test_rate_ops_audit(eng, nat):
  """the default build's float64 rate expressions are not
  the reference's operation sequence.  Over 3 x 10^8 evaluations per rate
  function: the simple rate's float32 value never differs from the operation
  sequence's (the cast guard catches every case; the simple rate is pure
  float64 NumPy upstream, so that is the reference's value), and the human
  prior's differs at the printed rate (it is JAX float32 upstream: covered by
  the stated rate tolerance, not defined to the bit)."""
  lat = eng.Lattice(50)
  out = nat.PdRateOpsStats()
  nat.check(nat.lib.pd_rate_ops_audit(C.byref(lat.c), 99, 100_000_000, 1.42,
                                      C.byref(out), None))
  print(f'\nrate ops audit: {out.evaluations} evaluations per rate fn; simple: '
        f'max {out.simple_max_ulps} ulp between the forms (guard '
        f'{out.guard_ulps}), casts differing before the guard '
        f'{out.simple_cast_differs_unguarded}, guard taken '
        f'{out.simple_guard_taken}, after the guard {out.simple_cast_differs}; '
        f'prior: max {out.prior_max_ulps} ulp, casts differing '
        f'{out.prior_cast_differs}')
  assert out.evaluations == 300_000_000
  assert out.simple_cast_differs == 0
  assert out.simple_max_ulps * 4 <= out.guard_ulps
  assert out.simple_guard_taken < 2e-4 * out.evaluations
  assert out.prior_cast_differs < 1e-5 * out.evaluations


@pytest.mark.parametrize('name', ['events_simple_large.npz',
                                  'events_prior_large.npz',
                                  'events_prior_single1000.npz'])
def test_fast_rollout_vs_reference_golden(eng, nat, golden_dir, name):
  """The fast rollout kernels against the unmodified reference itself: the
  reference's own trajectories (256 envs x 100 steps per rate function; one
  env x 1000 steps = BASELINE configs[0]) replayed as ONE pd_rollout_actions
  launch per dwell time (the fixtures hold two; a launch takes one, the envs
  are independent, so each launch is compared on the envs that used its dwell
  time).  Si site and elapsed microseconds of every step, final FOV."""
  from tests.test_oracle import expand_controls
  fix = np.load(os.path.join(golden_dir, name))
  controls, dwell = expand_controls(fix)
  n_steps, n = controls.shape[:2]
  spec = gh.rate_spec(int(fix['rate_fn']))
  nat.lib.pd_set_fast_path(1)
  checked = 0
  for d in np.unique(dwell):
    mask = (dwell[:, :, 0] == d).all(axis=0)
    b = eng.EnvBatch(n, seed=int(fix['seed']))
    b.reset()
    np.testing.assert_array_equal(gh.np_(b.si_idx), fix['si0'])
    si, el = b.rollout(controls[:, :, 0, :], int(d), spec, record=True,
                       action_mode=nat.ACTION_DIRECT)
    si, el = gh.np_(si), gh.np_(el)
    np.testing.assert_array_equal(si[:, mask], fix['si'][mask].T)
    np.testing.assert_array_equal(el[:, mask], fix['elapsed_us'][mask].T)
    if 'fov_last' in fix:
      np.testing.assert_allclose(gh.np_(b.fov)[mask], fix['fov_last'][mask],
                                 rtol=0, atol=1e-13)
    checked += int(mask.sum())
  assert checked == n


@pytest.mark.parametrize('n,t_steps', [(100000, 24), (4096, 96)])
def test_plan_kernels_host_formats(eng, nat, n, t_steps):
  """The plan kernels in the float32 / packed host formats (IO 1 and 2):
  pd_rollout_actions_host_packed and pd_rollout_actions_host_f32 with
  k_walk_plan (large batch) / k_rollout_plan (small batch) selected for every
  chunk, against the device-resident rollout through k_walk_fast /
  k_rollout_fast; every fifth FOV moved off-centre (first-step area test,
  clip)."""
  dwell = 1500000
  rng = np.random.default_rng(8)
  acts = rng.uniform(-1.1, 1.1, size=(t_steps, n, 2)).astype(np.float32)
  spec = gh.rate_spec(po.RATE_PRIOR)

  def batch():
    b = eng.EnvBatch(n, seed=21)
    b.reset()
    b.fov[::5] += 4.1
    return b

  _select(nat, 'fast')
  a = batch()
  si, el = a.rollout(acts.astype(np.float64), dwell, spec, record=True,
                     action_mode=nat.ACTION_RELATIVE_TO_SILICON)
  si, el = gh.np_(si), gh.np_(el)
  want = {k: gh.np_(v) for k, v in a.state_dict().items() if k in STATE_KEYS}
  _select(nat, 'plan')
  b = batch()
  packed = b.rollout_host_packed(acts, dwell, spec,
                                 action_mode=nat.ACTION_RELATIVE_TO_SILICON)
  si_p, el_p = eng.EnvBatch.unpack_rollout(packed, dwell)
  np.testing.assert_array_equal(si_p.numpy(), si)
  np.testing.assert_array_equal(el_p.numpy(), el)
  c = batch()
  si_h, el_h = c.rollout_host(acts, dwell, spec,
                              action_mode=nat.ACTION_RELATIVE_TO_SILICON)
  np.testing.assert_array_equal(si_h.numpy(), si)
  np.testing.assert_array_equal(el_h.numpy().astype(np.int64), el)
  for other in (b, c):
    sd = other.state_dict()
    for k in STATE_KEYS:
      np.testing.assert_array_equal(gh.np_(sd[k]), want[k], err_msg=k)


@pytest.mark.parametrize('kernels', ['fast', 'plan'])
@pytest.mark.parametrize('rate_fn', [po.RATE_PRIOR, po.RATE_SIMPLE])
def test_sheet_edge_sites_vs_oracle(eng, nat, rate_fn, kernels):
  """Every Si parked on a sheet-edge site (the three nearest sites are not
  the three bonded ones; where the Si atoms of a long rollout end up),
  FOV centred on it, 48 relative_random steps against the oracle: Si site and
  elapsed microseconds of every step, counters, final FOV."""
  n, t_steps, seed = 2048, 48, 9
  st = po.make_state(n, seed)
  po.reset(st)
  nbr = po.neighbor_table(50)
  base = po.base_lattice(50)
  d = np.linalg.norm(base[nbr] - base[:, None, :], axis=2)
  edge = np.nonzero((d > po.BOND * 1.01).any(axis=1))[0]
  assert 100 < edge.size < 200
  st.si_idx[:] = edge[np.arange(n) % edge.size]
  xy = po.site_positions(st, st.si_idx, np.arange(n))
  half = st.fov_scale[:, None] / 2
  st.fov[:] = np.concatenate([xy - half, xy + half], axis=1)
  rng = np.random.default_rng(4)
  acts = rng.uniform(-1.0, 1.0, size=(t_steps, n, 2))
  _select(nat, kernels)
  b = gh.batch_from_oracle(st)
  si, el = b.rollout(acts, 1500000, gh.rate_spec(rate_fn), record=True,
                     action_mode=nat.ACTION_RELATIVE_TO_SILICON,
                     max_distance_angstroms=1.42)
  si, el = gh.np_(si), gh.np_(el)
  on_edge = 0
  for t in range(t_steps):
    ctl = oe.relative_to_silicon_controls(st, acts[t])[:, None, :]
    want = po.step_and_image(st, ctl, 1500000, rate_fn=rate_fn)
    np.testing.assert_array_equal(si[t], st.si_idx, err_msg=f'step {t}')
    np.testing.assert_array_equal(el[t], want['elapsed_us'],
                                  err_msg=f'step {t}')
    on_edge += int(np.isin(st.si_idx, edge).sum())
  np.testing.assert_array_equal(gh.np_(b.n_transitions), st.n_transitions)
  np.testing.assert_array_equal(gh.np_(b.n_events), st.n_events)
  np.testing.assert_allclose(gh.np_(b.fov), st.fov, rtol=0, atol=1e-13)
  assert st.n_transitions.sum() > 5000
  # (a good part of the env-steps stays on edge sites)
  print(f'rate {rate_fn}: {on_edge / (n * t_steps):.2f} of the env-steps on '
        f'edge sites, {int(st.n_transitions.sum())} transitions')
  assert on_edge > 0.1 * n * t_steps


def test_large_batch_kernel_choice_across_regimes(eng, nat):
  """The default large-batch path picks its kernel per batch (k_walk_plan
  while few envs leave the plan, k_walk_fast once the Si atoms have gathered
  on the sheet's edge, the plan again after a reset): 5120 steps, a reset,
  480 more, launch by launch against k_walk_fast alone -- per-step results and
  final state."""
  n, t_steps, dwell = 460800, 8, 1500000
  spec = gh.rate_spec(po.RATE_PRIOR)
  gen = torch.Generator(device='cuda')
  gen.manual_seed(3)
  acts = [torch.rand((t_steps, n, 2), generator=gen, device='cuda',
                     dtype=torch.float64) * 2 - 1 for _ in range(5)]
  runs = []
  for walk_plan in (1, 0):
    nat.check(nat.lib.pd_set_option(b'fast_path', 1))
    nat.check(nat.lib.pd_set_option(b'plan', 0))
    nat.check(nat.lib.pd_set_option(b'walk_plan', walk_plan))
    b = eng.EnvBatch(n, seed=12)
    b.reset()
    digest = torch.zeros((), dtype=torch.int64, device='cuda')
    for i in range(700):
      if i == 640:
        b.reset()
      si, el = b.rollout(acts[i % 5], dwell, spec, record=True,
                         action_mode=nat.ACTION_RELATIVE_TO_SILICON)
      digest = digest * 1000003 + (si.to(torch.int64) * 31 + el).sum()
    sd = b.state_dict()
    runs.append((int(digest.item()),
                 {k: gh.np_(sd[k]) for k in STATE_KEYS}))
  assert runs[0][0] == runs[1][0]
  for k in STATE_KEYS:
    np.testing.assert_array_equal(runs[0][1][k], runs[1][1][k], err_msg=k)
