"""GPU parity tests of the event path, through the C ABI, against the oracle
and the committed reference golden vectors.

Bars (BASELINE.json north_star): bit-exact Si lattice index, chosen
transition, event counts and microsecond clocks; float64 geometry within
1e-13 A; float32 rates: simple bit-exact, human prior |d| <= 2e-6*max + 1e-12.
"""

import os

import numpy as np
import pytest
import torch

from oracle import pdune_oracle as po
from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
  from putting_dune_b200 import engine
  return engine


@pytest.fixture(autouse=True)
def _default_options():
  """Kernel-selection options back to their defaults after every test."""
  yield
  from putting_dune_b200 import _native as nat
  for name in (b'fast_path', b'prepass', b'rollout_spec'):
    nat.lib.pd_set_option(name, 1)
  nat.lib.pd_set_option(b'race_sampling', 0)


def site_classes(cols):
  """0 / 1: bulk site whose three nearest neighbours are the bonded ones at
  (-1/2, -r), (1, 0), (-1/2, r) bond lengths (r = sqrt(3)/2) / the negatives
  in reverse order; 2: anything else (sheet edge)."""
  base, nbr = po.base_lattice(cols), po.neighbor_table(cols)
  off = (base[nbr] - base[:, None, :]) / po.BOND
  r = np.sqrt(3.0) / 2
  a = np.array([[-0.5, -r], [1.0, 0.0], [-0.5, r]])
  cls = np.full(len(base), 2)
  cls[np.abs(off - a).max(axis=(1, 2)) < 1e-9] = 0
  cls[np.abs(off + a[::-1]).max(axis=(1, 2)) < 1e-9] = 1
  return cls


def test_lattice_tables_bit_exact(eng):
  for cols in (50, 10, 13):
    lat = eng.Lattice(cols)
    np.testing.assert_array_equal(gh.np_(lat.base_xy), po.base_lattice(cols))
    np.testing.assert_array_equal(gh.np_(lat.nbr)[:, :3],
                                  po.neighbor_table(cols))
    # column 3, low 24 bits: the sites within 2.6 A of the lattice centre,
    # then 0xFFFFFF; bits 24-25: the neighbour-geometry class of the site
    col3 = gh.np_(lat.nbr)[:, 3]
    cand = col3 & 0xFFFFFF
    want = np.nonzero((po.base_lattice(cols) ** 2).sum(axis=1) <= 2.6 ** 2)[0]
    np.testing.assert_array_equal(cand[:want.size], want)
    assert (cand[want.size:] == 0xFFFFFF).all()
    np.testing.assert_array_equal(col3 >> 24, site_classes(cols))


def test_reset_matches_oracle(eng):
  n, seed = 4096, 11
  st = po.make_state(n, seed, env_offset=1000)
  po.reset(st)
  b = eng.EnvBatch(n, seed=seed, env_offset=1000)
  b.reset()
  np.testing.assert_array_equal(gh.np_(b.si_idx), st.si_idx)
  # offsets are pure arithmetic (exact); cos/sin may differ in the last ulp
  np.testing.assert_array_equal(gh.np_(b.lattice)[:, :2], st.lattice[:, :2])
  np.testing.assert_allclose(gh.np_(b.lattice)[:, 2:], st.lattice[:, 2:],
                             rtol=0, atol=3e-16)
  np.testing.assert_array_equal(gh.np_(b.fov_scale), st.fov_scale)
  np.testing.assert_allclose(gh.np_(b.fov), st.fov, rtol=0, atol=1e-13)
  np.testing.assert_allclose(gh.np_(b.image_params), st.image_params,
                             rtol=1e-15, atol=0)
  assert (gh.np_(b.status) == 0).all()
  assert (gh.np_(b.episode) == 1).all()
  # masked second reset: only masked envs advance their episode
  mask = np.arange(n) % 3 == 0
  po.reset(st, mask)
  b.reset(mask)
  np.testing.assert_array_equal(gh.np_(b.episode), st.episode.astype(np.int32))
  np.testing.assert_array_equal(gh.np_(b.si_idx), st.si_idx)
  np.testing.assert_allclose(gh.np_(b.fov), st.fov, rtol=0, atol=1e-13)


def test_rates_match_oracle_and_reference_golden(eng, golden_dir):
  fix = np.load(os.path.join(golden_dir, 'rates_reference.npz'))
  n = fix['beam'].shape[0]
  st = po.make_state(n, int(fix['seed']))
  po.reset(st)
  b = gh.batch_from_oracle(st)
  for name, rate_fn in (('simple', po.RATE_SIMPLE), ('prior', po.RATE_PRIOR)):
    r, nb = b.rates(fix['beam'], gh.rate_spec(rate_fn))
    r, nb = gh.np_(r), gh.np_(nb)
    want, _ = po.rates_for(st, np.arange(n), fix['beam'], rate_fn)
    np.testing.assert_array_equal(nb, fix[f'succ_{name}'])
    if rate_fn == po.RATE_SIMPLE:
      np.testing.assert_array_equal(r, want)
      np.testing.assert_array_equal(r, fix['rates_simple'])
    else:
      tol = 2e-6 * want.max() + 1e-12
      assert np.abs(r - want).max() <= tol
      assert np.abs(r - fix['rates_prior']).max() <= tol


@pytest.mark.parametrize('name', [
    'events_simple.npz', 'events_prior.npz',
    # 256 envs x 100 steps per rate function and one env x 1000 steps
    # (BASELINE configs[0]) from the unmodified reference
    'events_simple_large.npz', 'events_prior_large.npz',
    'events_prior_single1000.npz'])
def test_reference_golden_trajectories(eng, golden_dir, name):
  """The reference's own trajectories, replayed through the C ABI with the
  device doing its own reset."""
  from tests.test_oracle import expand_controls
  fix = np.load(os.path.join(golden_dir, name))
  controls, dwell = expand_controls(fix)
  n_steps, n = controls.shape[:2]
  compact = 'fov_last' in fix
  b = eng.EnvBatch(n, seed=int(fix['seed']), log_capacity=64)
  b.reset()
  np.testing.assert_array_equal(gh.np_(b.si_idx), fix['si0'])
  np.testing.assert_allclose(gh.np_(b.fov), fix['fov0'], rtol=0, atol=1e-13)
  np.testing.assert_allclose(gh.np_(b.image_params), fix['image_params'],
                             rtol=1e-15)
  spec = gh.rate_spec(int(fix['rate_fn']))
  trans = []
  for t in range(n_steps):
    out = b.step_and_image(controls[t], dwell[t], spec)
    np.testing.assert_array_equal(gh.np_(b.si_idx), fix['si'][:, t])
    np.testing.assert_array_equal(gh.np_(out.elapsed_us),
                                  fix['elapsed_us'][:, t])
    np.testing.assert_allclose(
        gh.np_(b.fov),
        fix['fov_last'] if compact else fix['fov'][:, t], rtol=0,
        atol=1e-13) if (not compact or t == n_steps - 1) else None
    cnt = gh.np_(out.log_count)
    el, site = gh.np_(out.log_elapsed_us), gh.np_(out.log_site)
    for e in np.nonzero(cnt)[0]:
      for k in range(cnt[e]):
        trans.append((e, t, el[e, k], site[e, k]))
  want = fix['transitions']
  want = want[np.lexsort((want[:, 2], want[:, 1], want[:, 0]))]
  got = np.asarray(sorted(trans), dtype=np.int64)
  np.testing.assert_array_equal(got, want)
  xy, z, count = b.get_atoms_in_bounds()
  np.testing.assert_array_equal(
      gh.np_(count),
      fix['n_observed_last'] if compact else fix['n_observed'][:, -1])


PRIOR_CUSTOM = {  # as in make_golden: HumanPriorRatePredictor(mean, cov, max)
    'mean': (0.7, 0.15),
    'cov': ((0.12, 0.03), (0.03, 0.07)),
    'max_rate': 0.4,
}


def test_custom_human_prior(eng, golden_dir):
  """HumanPriorRatePredictor(mean, cov, max_rate) with non-default parameters
  (graphene.py:181-229): rates against the reference's own (golden) and
  4096 envs x 40 steps of trajectories against the oracle, through
  step_and_image and through the rollout kernels; parameters equal to the
  defaults take the specialised path and give identical results."""
  fix = np.load(os.path.join(golden_dir, 'rates_reference.npz'))
  n = fix['beam'].shape[0]
  st = po.make_state(n, int(fix['seed']))
  po.reset(st)
  b = gh.batch_from_oracle(st)
  spec = eng.RateSpec.prior(**PRIOR_CUSTOM)
  r, nb = b.rates(fix['beam'], spec)
  np.testing.assert_array_equal(gh.np_(nb), fix['succ_prior_custom'])
  want = fix['rates_prior_custom']
  assert np.abs(gh.np_(r) - want).max() <= 2e-6 * want.max() + 1e-12
  assert want.max() > 0.05
  # trajectories
  n, t_steps, seed = 4096, 40, 19
  st = po.make_state(n, seed)
  po.reset(st)
  b = gh.batch_from_oracle(st)
  c = gh.batch_from_oracle(st)
  rng = np.random.default_rng(2)
  ctl_all = np.zeros((t_steps, n, 2))
  for t in range(t_steps):
    ctl = gh.closed_loop_control(st, rng)
    ctl_all[t] = ctl
    wanted = po.step_and_image(st, ctl[:, None, :], 3000000,
                               rate_fn=po.RATE_PRIOR, prior=PRIOR_CUSTOM)
    out = b.step_and_image(ctl[:, None, :], 3000000, spec)
    np.testing.assert_array_equal(gh.np_(b.si_idx), st.si_idx)
    np.testing.assert_array_equal(gh.np_(out.elapsed_us), wanted['elapsed_us'])
    np.testing.assert_array_equal(gh.np_(out.transitions),
                                  wanted['transitions'])
  assert st.n_transitions.sum() > 5000
  si, _ = c.rollout(ctl_all, 3000000, spec, record=True)
  np.testing.assert_array_equal(gh.np_(si[-1]), st.si_idx)
  np.testing.assert_array_equal(gh.np_(c.n_events), st.n_events)
  # defaults passed explicitly == the specialised human prior
  d0, d1 = eng.EnvBatch(512, seed=3), eng.EnvBatch(512, seed=3)
  d0.reset()
  d1.reset()
  ctl = np.full((512, 1, 2), 0.52)
  o0 = d0.step_and_image(ctl, 5000000, eng.RateSpec.prior())
  o1 = d1.step_and_image(ctl, 5000000, eng.RateSpec.prior(
      (0.85, 0.0), ((0.1, 0.0), (0.0, 0.1)), np.log(2) / 3))
  np.testing.assert_array_equal(gh.np_(d0.si_idx), gh.np_(d1.si_idx))
  np.testing.assert_array_equal(gh.np_(o0.events), gh.np_(o1.events))


@pytest.mark.parametrize('rate_fn', [po.RATE_SIMPLE, po.RATE_PRIOR])
def test_config2_parity_4096_envs(eng, rate_fn):
  """BASELINE config 2: 4096 envs, injected-RNG parity over 100 steps."""
  n, n_steps, seed = 4096, 100, 42 + rate_fn
  st = po.make_state(n, seed)
  po.reset(st)
  b = gh.batch_from_oracle(st)
  spec = gh.rate_spec(rate_fn)
  rng = np.random.default_rng(1)
  dwell = np.where(np.arange(n) % 2 == 0, 1500000, 5000000)[:, None]
  for t in range(n_steps):
    ctl = gh.closed_loop_control(st, rng)[:, None, :]
    want = po.step_and_image(st, ctl, dwell, rate_fn=rate_fn)
    out = b.step_and_image(ctl, dwell, spec)
    np.testing.assert_array_equal(gh.np_(b.si_idx), st.si_idx)
    np.testing.assert_array_equal(gh.np_(out.elapsed_us), want['elapsed_us'])
    np.testing.assert_array_equal(gh.np_(out.transitions),
                                  want['transitions'])
    np.testing.assert_array_equal(gh.np_(out.events), want['events'])
    np.testing.assert_array_equal(gh.np_(out.recentred).astype(bool),
                                  want['recentred'])
  np.testing.assert_allclose(gh.np_(b.fov), st.fov, rtol=0, atol=1e-13)
  np.testing.assert_array_equal(gh.np_(b.n_events), st.n_events)
  np.testing.assert_array_equal(gh.np_(b.n_transitions), st.n_transitions)
  np.testing.assert_array_equal(gh.np_(b.sim_time_us), st.sim_time_us)
  np.testing.assert_array_equal(gh.np_(b.ctrl_count),
                                st.ctrl_count.astype(np.int32))
  assert st.n_transitions.sum() > 1000


def test_multiple_controls_and_ragged_dwell(eng):
  n, seed = 1000, 5  # not a multiple of the block size
  st = po.make_state(n, seed)
  po.reset(st)
  b = gh.batch_from_oracle(st, log_capacity=32)
  rng = np.random.default_rng(2)
  ctl = np.stack([gh.closed_loop_control(st, rng) for _ in range(3)], axis=1)
  dwell = rng.integers(0, 6000000, size=(n, 3))
  dwell[::7, 1] = 0  # zero dwell: no rate evaluation at all
  want = po.step_and_image(st, ctl, dwell)
  out = b.step_and_image(ctl, dwell, gh.rate_spec(po.RATE_SIMPLE))
  np.testing.assert_array_equal(gh.np_(b.si_idx), st.si_idx)
  np.testing.assert_array_equal(gh.np_(out.elapsed_us), want['elapsed_us'])
  np.testing.assert_array_equal(gh.np_(out.events), want['events'])
  np.testing.assert_array_equal(gh.np_(b.ctrl_count),
                                st.ctrl_count.astype(np.int32))
  # the transition log names the control each hop happened in
  cnt = gh.np_(out.log_count)
  np.testing.assert_array_equal(cnt, want['transitions'])
  assert gh.np_(out.log_ctrl)[cnt > 0].max() <= 2


def test_empty_and_no_control_edge_cases(eng):
  spec = gh.rate_spec(po.RATE_SIMPLE)
  b0 = eng.EnvBatch(0)
  b0.reset()
  b0.step_and_image(np.zeros((0, 1, 2)), 1500000, spec)
  b = eng.EnvBatch(8, seed=1)
  b.reset()
  si = gh.np_(b.si_idx).copy()
  out = b.step_and_image(np.zeros((8, 0, 2)), 0, spec, 2000000)
  np.testing.assert_array_equal(gh.np_(out.elapsed_us), 2000000)
  np.testing.assert_array_equal(gh.np_(b.si_idx), si)
  assert (gh.np_(b.ctrl_count) == 0).all()


def test_constant_rate_seam_reproduces_reference_tests(eng):
  # simulator_test.py:147-168: zero rates => elapsed = sum(dwell) + image.
  b = eng.EnvBatch(16, seed=3)
  b.reset()
  zero = gh.rate_spec(po.RATE_CONSTANT, constant=(0.0, 0.0, 0.0))
  ctl = np.full((16, 3, 2), 0.5)
  dwell = np.tile(np.array([[1500000, 3000000, 7230000]]), (16, 1))
  out = b.step_and_image(ctl, dwell, zero, 3500000)
  np.testing.assert_array_equal(gh.np_(out.elapsed_us), 15230000)
  assert (gh.np_(out.transitions) == 0).all()
  assert (gh.np_(out.events) == 3).all()
  # graphene_test.py:228-281: rate 5 => several hops within the dwell.
  st = po.make_state(256, seed=9)
  po.reset(st)
  b = gh.batch_from_oracle(st)
  five_o = lambda s, idx, beam, it: (np.full((idx.size, 3), 5.0, np.float32),
                                     s.nbr[s.si_idx[idx]])
  want = po.apply_control(st, np.zeros((256, 2)), np.full(256, 1500000),
                          rates_override=five_o)
  out = b.apply_control(np.zeros((256, 2)), 1500000,
                        gh.rate_spec(po.RATE_CONSTANT, constant=(5, 5, 5)))
  np.testing.assert_array_equal(gh.np_(out.transitions), want['transitions'])
  np.testing.assert_array_equal(gh.np_(b.si_idx), st.si_idx)
  assert want['transitions'].mean() > 10
  # negative rates: the reference asserts; the batch flags the env instead.
  bad = gh.rate_spec(po.RATE_CONSTANT, constant=(-1.0, 1.0, 1.0))
  b.apply_control(np.zeros((256, 2)), 1000, bad)
  assert (gh.np_(b.status) & 1).all()


def test_rollout_equals_repeated_steps(eng):
  n, t_steps, seed = 5000, 12, 8
  rng = np.random.default_rng(4)
  ctl = 0.5 + rng.uniform(-0.08, 0.08, size=(t_steps, n, 2))
  for rate_fn in (po.RATE_SIMPLE, po.RATE_PRIOR):
    spec = gh.rate_spec(rate_fn)
    a = eng.EnvBatch(n, seed=seed)
    b = eng.EnvBatch(n, seed=seed)
    a.reset()
    b.reset()
    si, el = a.rollout(ctl, 1500000, spec, record=True)
    for t in range(t_steps):
      out = b.step_and_image(ctl[t][:, None, :], 1500000, spec)
      np.testing.assert_array_equal(gh.np_(si[t]), gh.np_(b.si_idx))
      np.testing.assert_array_equal(gh.np_(el[t]), gh.np_(out.elapsed_us))
    for f in ('si_idx', 'fov', 'ctrl_count', 'sim_time_us', 'n_events',
              'n_transitions'):
      np.testing.assert_array_equal(gh.np_(getattr(a, f)),
                                    gh.np_(getattr(b, f)))


def test_staged_and_global_table_paths_agree(eng):
  """Large batches stage the lattice in shared memory; results must not
  depend on the path (size-independent property at full size)."""
  n, seed = 300000, 13
  big = eng.EnvBatch(n, seed=seed)
  big.reset()
  small = eng.EnvBatch(2048, seed=seed)
  small.reset()
  rng = np.random.default_rng(6)
  ctl = 0.5 + rng.uniform(-0.06, 0.06, size=(n, 1, 2))
  spec = gh.rate_spec(po.RATE_SIMPLE)
  o_big = big.step_and_image(ctl, 5000000, spec)
  o_small = small.step_and_image(ctl[:2048], 5000000, spec)
  np.testing.assert_array_equal(gh.np_(big.si_idx)[:2048],
                                gh.np_(small.si_idx))
  np.testing.assert_array_equal(gh.np_(o_big.elapsed_us)[:2048],
                                gh.np_(o_small.elapsed_us))
  # every hop lands on one of the three neighbours of the previous site
  tr = gh.np_(o_big.transitions)
  assert tr.sum() > n // 2 and tr.max() >= 4


def test_sharding_invariance(eng):
  """Philox is keyed by the global env id: a shard reproduces its slice."""
  seed = 21
  full = eng.EnvBatch(4096, seed=seed)
  shard = eng.EnvBatch(1024, seed=seed, env_offset=2048)
  full.reset()
  shard.reset()
  ctl = np.full((4096, 1, 2), 0.52)
  spec = gh.rate_spec(po.RATE_SIMPLE)
  for _ in range(5):
    full.step_and_image(ctl, 5000000, spec)
    shard.step_and_image(ctl[:1024], 5000000, spec)
  np.testing.assert_array_equal(gh.np_(full.si_idx)[2048:3072],
                                gh.np_(shard.si_idx))
  np.testing.assert_array_equal(gh.np_(full.sim_time_us)[2048:3072],
                                gh.np_(shard.sim_time_us))


def test_queries_match_oracle(eng):
  n, seed = 64, 17
  st = po.make_state(n, seed)
  po.reset(st)
  b = gh.batch_from_oracle(st)
  np.testing.assert_allclose(
      gh.np_(b.silicon_position()),
      po.site_positions(st, st.si_idx, np.arange(n)), rtol=0, atol=0)
  xy, z, count, site = b.get_atoms_in_bounds(with_sites=True)
  xy, z, count, site = gh.np_(xy), gh.np_(z), gh.np_(count), gh.np_(site)
  for e in range(n):
    q, zz, idx = po.get_atoms_in_bounds(st, e)
    assert count[e] == q.shape[0]
    np.testing.assert_array_equal(site[e, :count[e]], idx)
    np.testing.assert_array_equal(z[e, :count[e]], zz)
    np.testing.assert_array_equal(xy[e, :count[e]], q)
  # arbitrary (non-square, partly off-sheet) windows
  fov = np.tile(np.array([[-5.5, -6.3, 12.0, 9.1]]), (n, 1))
  fov[1] = [20.0, 20.0, 60.0, 60.0]
  fov[2] = [100.0, 100.0, 110.0, 110.0]  # empty
  xy, z, count = b.get_atoms_in_bounds(fov)
  count = gh.np_(count)
  for e in range(4):
    q, zz, _ = po.get_atoms_in_bounds(st, e, fov[e])
    assert count[e] == q.shape[0]
    np.testing.assert_array_equal(gh.np_(xy)[e, :count[e]], q)
  assert count[2] == 0
  g = gh.np_(b.grid_positions([0, 5]))
  np.testing.assert_array_equal(g[1], po.all_positions(st, 5))


def test_host_buffer_entry_point(eng):
  """pd_step_and_image_host: host controls in, host observation out."""
  import ctypes as C
  from putting_dune_b200 import _native as nat
  n, seed = 2048, 23
  a = eng.EnvBatch(n, seed=seed)
  b = eng.EnvBatch(n, seed=seed)
  a.reset()
  b.reset()
  spec = gh.rate_spec(po.RATE_PRIOR)
  ctl = torch.full((n, 1, 2), 0.5, dtype=torch.float64).pin_memory()
  d_ctl = torch.empty((n, 1, 2), dtype=torch.float64, device=b.device)
  h_el = torch.empty(n, dtype=torch.int64).pin_memory()
  h_xy = torch.empty((n, 2), dtype=torch.float64).pin_memory()
  h_fov = torch.empty((n, 4), dtype=torch.float64).pin_memory()
  P = lambda t: C.c_void_p(t.data_ptr())
  nat.check(nat.lib.pd_step_and_image_host(
      C.byref(b.lattice_tables.c), C.byref(b.c), C.byref(spec.c), P(ctl), None,
      5000000, 1, 2000000, P(d_ctl), None, C.byref(b._out_c), P(h_el),
      P(h_xy), P(h_fov), None))
  out = a.step_and_image(ctl, 5000000, spec)
  np.testing.assert_array_equal(h_el.numpy(), gh.np_(out.elapsed_us))
  np.testing.assert_array_equal(h_xy.numpy(), gh.np_(out.si_xy))
  np.testing.assert_array_equal(h_fov.numpy(), gh.np_(a.fov))


def test_checkpoint_resume(eng):
  spec = gh.rate_spec(po.RATE_SIMPLE)
  a = eng.EnvBatch(512, seed=31)
  a.reset()
  ctl = np.full((512, 1, 2), 0.51)
  a.step_and_image(ctl, 5000000, spec)
  saved = a.state_dict()
  b = eng.EnvBatch(512, seed=0)
  b.load_state_dict(saved)
  a.step_and_image(ctl, 5000000, spec)
  b.step_and_image(ctl, 5000000, spec)
  np.testing.assert_array_equal(gh.np_(a.si_idx), gh.np_(b.si_idx))
  np.testing.assert_array_equal(gh.np_(a.sim_time_us), gh.np_(b.sim_time_us))


def test_rollout_with_relative_to_silicon_adapter(eng):
  """On-device RelativeToSiliconActionAdapter (action_adapters.py:131-216):
  actions in [-1, 1]^2 around the Si, checked step by step against the
  oracle's adapter + step_and_image."""
  from oracle import pdune_oracle_episode as oe
  from putting_dune_b200 import _native as nat
  n, t_steps, seed = 3000, 10, 33
  rng = np.random.default_rng(8)
  actions = rng.uniform(-1.3, 1.3, size=(t_steps, n, 2))  # exercises the clip
  for rate_fn in (po.RATE_SIMPLE, po.RATE_PRIOR):
    st = po.make_state(n, seed)
    po.reset(st)
    b = gh.batch_from_oracle(st)
    si, el = b.rollout(actions, 5000000, gh.rate_spec(rate_fn), record=True,
                       action_mode=nat.ACTION_RELATIVE_TO_SILICON,
                       max_distance_angstroms=1.42)
    for t in range(t_steps):
      ctl = oe.relative_to_silicon_controls(st, actions[t])[:, None, :]
      want = po.step_and_image(st, ctl, 5000000, rate_fn=rate_fn)
      np.testing.assert_array_equal(gh.np_(si[t]), st.si_idx)
      np.testing.assert_array_equal(gh.np_(el[t]), want['elapsed_us'])
    np.testing.assert_allclose(gh.np_(b.fov), st.fov, rtol=0, atol=1e-13)
    np.testing.assert_array_equal(gh.np_(b.n_transitions), st.n_transitions)
    assert st.n_transitions.sum() > (2000 if rate_fn == po.RATE_SIMPLE else 300)
  # the same through the large-batch kernel
  big = eng.EnvBatch(160000, seed=seed)
  small = eng.EnvBatch(3000, seed=seed)
  big.reset()
  small.reset()
  acts = rng.uniform(-1, 1, size=(4, 160000, 2))
  spec = gh.rate_spec(po.RATE_SIMPLE)
  kw = dict(action_mode=nat.ACTION_RELATIVE_TO_SILICON, record=True)
  s_big, _ = big.rollout(acts, 5000000, spec, **kw)
  s_small, _ = small.rollout(acts[:, :3000].copy(), 5000000, spec, **kw)
  np.testing.assert_array_equal(gh.np_(s_big)[:, :3000], gh.np_(s_small))


@pytest.mark.parametrize('n', [37, 700, 4096, 11000, 21000])
def test_rollout_speculative_equals_serial(eng, n, monkeypatch):
  """k_rollout_spec (small batches: a group of 2..32 lanes evaluates the next
  iterations of one env speculatively) commits exactly what the serial
  k_rollout computes: per-step Si site and elapsed time, FOV, clocks, event
  and transition counts, Philox control counter."""
  from putting_dune_b200 import _native as nat
  t_steps, seed = 45, 77
  rng = np.random.default_rng(n)
  acts = rng.uniform(-1.2, 1.2, size=(t_steps, n, 2))
  direct = 0.5 + rng.uniform(-0.1, 0.1, size=(t_steps, n, 2))
  cases = [
      (po.RATE_PRIOR, None, 1500000, acts, nat.ACTION_RELATIVE_TO_SILICON),
      (po.RATE_SIMPLE, None, 5000000, acts, nat.ACTION_RELATIVE_TO_SILICON),
      (po.RATE_SIMPLE, None, 700000, direct, 0),
      # many hops per control, re-centres every few steps
      (po.RATE_CONSTANT, (2.0, 1.0, 3.0), 1500000, direct, 0),
      # one-microsecond clock ticks: controls that end exactly on the dwell
      (po.RATE_CONSTANT, (4e5, 3e5, 3e5), 3, direct, 0),
  ]
  for rate_fn, const, dwell, ctl, mode in cases:
    spec = gh.rate_spec(rate_fn, constant=const)
    res = []
    # k_rollout_pre (float32 look-ahead), k_rollout_spec (float64
    # speculation), k_rollout (serial)
    for flag, pre in ((1, 1), (1, 0), (0, 0)):
      # (the float64 kernels: the fast path would take the prior / simple
      # rollouts otherwise)
      nat.check(nat.lib.pd_set_option(b'fast_path', 0))
      nat.check(nat.lib.pd_set_option(b'rollout_spec', flag))
      nat.check(nat.lib.pd_set_option(b'prepass', pre))
      b = eng.EnvBatch(n, seed=seed)
      b.reset()
      # a FOV the Si is not centred in: the t = 0 safe-area check matters
      b.fov[::3] += 3.9
      si, el = b.rollout(ctl, dwell, spec, record=True, action_mode=mode,
                         max_distance_angstroms=1.42)
      res.append((gh.np_(si), gh.np_(el), b.state_dict()))
    si_b, el_b, st_b = res[-1]
    for si_a, el_a, st_a in res[:-1]:
      np.testing.assert_array_equal(si_a, si_b)
      np.testing.assert_array_equal(el_a, el_b)
      for k in ('si_idx', 'fov', 'ctrl_count', 'sim_time_us', 'n_events',
                'n_transitions', 'status'):
        np.testing.assert_array_equal(gh.np_(st_a[k]), gh.np_(st_b[k]),
                                      err_msg=k)
    assert gh.np_(st_b['n_transitions']).sum() > 0


def test_prepass_equals_exact(eng, monkeypatch):
  """The float32 pre-pass of the large-batch kernel (certainly_no_hop: controls
  whose waiting time certainly overshoots the dwell time skip the float64
  chain) never changes a result: rollouts, multi-control steps with ragged
  dwell times and material-frame apply_control agree bit for bit with
  PD_PREPASS=0, including event counts and the Philox control counter."""
  from putting_dune_b200 import _native as nat
  n, t_steps, seed = 90000, 7, 91
  rng = np.random.default_rng(12)
  acts = rng.uniform(-1.3, 1.3, size=(t_steps, n, 2))
  direct = 0.5 + rng.uniform(-0.12, 0.12, size=(t_steps, n, 2))
  direct[0, :50] = np.nan            # non-finite controls take the exact path
  direct[1, 50:100] = 1e9
  ctl3 = 0.5 + rng.uniform(-0.1, 0.1, size=(n, 3, 2))
  dwell3 = rng.integers(0, 4000000, size=(n, 3))
  beam_m = rng.uniform(-3, 3, size=(n, 1, 2))
  results = []
  for flag in (1, 0):
    # k_walk's float32 pre-pass on / off (with the fast path off: it would
    # take the rollouts otherwise)
    nat.check(nat.lib.pd_set_option(b'fast_path', 0))
    nat.check(nat.lib.pd_set_option(b'prepass', flag))
    out = []
    for rate_fn in (po.RATE_PRIOR, po.RATE_SIMPLE):
      spec = gh.rate_spec(rate_fn)
      b = eng.EnvBatch(n, seed=seed)
      b.reset()
      b.fov[::5] += 4.1
      for ctl, dwell, mode in ((acts, 1500000, nat.ACTION_RELATIVE_TO_SILICON),
                               (acts, 9000000, nat.ACTION_RELATIVE_TO_SILICON),
                               (direct, 2500000, 0)):
        si, el = b.rollout(ctl, dwell, spec, record=True, action_mode=mode,
                           max_distance_angstroms=1.42)
        out += [gh.np_(si), gh.np_(el)]
      o = b.step_and_image(ctl3, dwell3, spec)
      out += [gh.np_(o.elapsed_us), gh.np_(o.transitions), gh.np_(o.events)]
      b.apply_control(gh.np_(b.silicon_position()) + beam_m[:, 0, :],
                      3000000, spec)
      st = b.state_dict()
      out += [gh.np_(st[k]) for k in ('si_idx', 'fov', 'ctrl_count',
                                      'sim_time_us', 'n_events',
                                      'n_transitions', 'status')]
    results.append(out)
  for x, y in zip(*results):
    np.testing.assert_array_equal(x, y)
  assert results[0][-2].sum() > n  # transitions happened


def test_rollout_host_pipeline_matches_device_path(eng):
  """pd_rollout_actions_host (float64 actions, int64 elapsed) returns what the
  device-resident rollout computes: (4096, 37) through the chunked H2D / step
  / D2H pipeline, (4096, 100) and (2048, 130) through the streamed launch
  (k_rollout_pre<STREAM = 2>), the third with actions that carry the
  "not arrived" bit pattern, the last through pd_rollout_host (beam positions
  instead of adapter actions)."""
  import ctypes as C
  from putting_dune_b200 import _native as nat
  rel, direct = nat.ACTION_RELATIVE_TO_SILICON, nat.ACTION_DIRECT
  for n, t_steps, poison, mode in ((4096, 37, False, rel),
                                   (4096, 100, False, rel),
                                   (2048, 130, True, rel),
                                   (4096, 70, False, direct)):
    seed = 51
    rng = np.random.default_rng(9 + t_steps)
    acts_np = (rng.uniform(-1, 1, size=(t_steps, n, 2)) if mode == rel else
               0.5 + rng.uniform(-0.06, 0.06, size=(t_steps, n, 2)))
    if poison:
      bits = acts_np.view(np.uint64)
      for t, e in ((0, 0), (3, 11), (t_steps - 1, n - 1)):
        bits[t, e, :] = 0xFFFFFFFFFFFFFFFF
    acts = torch.as_tensor(acts_np).pin_memory()
    a = eng.EnvBatch(n, seed=seed)
    b = eng.EnvBatch(n, seed=seed)
    a.reset()
    b.reset()
    spec = gh.rate_spec(po.RATE_PRIOR)
    si, el = a.rollout(acts, 1500000, spec, record=True, action_mode=mode)
    dev = b.device
    d_ctl = torch.empty((t_steps, n, 2), dtype=torch.float64, device=dev)
    d_si = torch.empty((t_steps, n), dtype=torch.int32, device=dev)
    d_el = torch.empty((t_steps, n), dtype=torch.int64, device=dev)
    h_si = torch.zeros((t_steps, n), dtype=torch.int32).pin_memory()
    h_el = torch.zeros((t_steps, n), dtype=torch.int64).pin_memory()
    P = lambda t: C.c_void_p(t.data_ptr())
    for rep in range(2):  # second call re-uses the cached side streams
      if rep:
        b.load_state_dict(a.state_dict())
      stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
      if mode == direct:  # pd_rollout_host = the DIRECT form
        nat.check(nat.lib.pd_rollout_host(
            C.byref(b.lattice_tables.c), C.byref(b.c), C.byref(spec.c),
            P(acts), 1500000, t_steps, 2000000, P(d_ctl), P(d_si), P(d_el),
            P(h_si), P(h_el), stream))
      else:
        nat.check(nat.lib.pd_rollout_actions_host(
            C.byref(b.lattice_tables.c), C.byref(b.c), C.byref(spec.c),
            P(acts), mode, 1.42, 1500000, t_steps, 2000000, P(d_ctl), P(d_si),
            P(d_el), P(h_si), P(h_el), stream))
      if rep == 0:
        np.testing.assert_array_equal(h_si.numpy(), gh.np_(si))
        np.testing.assert_array_equal(h_el.numpy(), gh.np_(el))
        np.testing.assert_array_equal(gh.np_(b.si_idx), gh.np_(a.si_idx))
        np.testing.assert_array_equal(gh.np_(b.sim_time_us),
                                      gh.np_(a.sim_time_us))


def _host_f32_case(eng, n, t_steps, mode, rate_fn=po.RATE_PRIOR,
                   want_elapsed=True, poison=False, owned=False):
  """One pd_rollout_actions_host_f32 call against the device rollout on the
  widened actions."""
  import ctypes as C
  from putting_dune_b200 import _native as nat
  rng = np.random.default_rng(n + t_steps)
  if mode == nat.ACTION_RELATIVE_TO_SILICON:
    acts = rng.uniform(-1.1, 1.1, size=(t_steps, n, 2))
  else:
    acts = 0.5 + rng.uniform(-0.06, 0.06, size=(t_steps, n, 2))
  acts = acts.astype(np.float32)
  if poison:
    # actions whose two words are 0xFFFFFFFF (the streamed form's "not yet
    # arrived" pattern): consumed as data once the copy has ended
    bits = acts.view(np.uint32)
    for t, e in ((0, 0), (5, 7), (t_steps // 2, n // 2), (t_steps - 1, n - 1)):
      bits[t, e, :] = 0xFFFFFFFF
  acts32 = torch.as_tensor(acts).pin_memory()
  a = eng.EnvBatch(n, seed=52)
  b = eng.EnvBatch(n, seed=52)
  a.reset()
  b.reset()
  spec = gh.rate_spec(rate_fn)
  si, el = a.rollout(acts32.double(), 1500000, spec, record=True,
                     action_mode=mode)
  dev = b.device
  # one guard row behind every buffer the call writes: nothing may touch it
  GUARD = 0x5A5A5A5A
  d_a32 = torch.full((t_steps + 1, n, 2), 1.5, dtype=torch.float32, device=dev)
  d_ctl = torch.full((t_steps + 1, n, 2), 1.5, dtype=torch.float64, device=dev)
  d_si = torch.full((t_steps + 1, n), GUARD, dtype=torch.int32, device=dev)
  d_el = torch.full((t_steps + 1, n), GUARD, dtype=torch.int64, device=dev)
  d_el32 = torch.full((t_steps + 1, n), GUARD, dtype=torch.int32, device=dev)
  h_si_all = torch.full((t_steps + 1, n), GUARD, dtype=torch.int32).pin_memory()
  h_el_all = torch.full((t_steps + 1, n), GUARD, dtype=torch.int32).pin_memory()
  h_si, h_el = h_si_all[:t_steps], h_el_all[:t_steps]
  P = lambda t: C.c_void_p(t.data_ptr())
  stage = ((None,) * 5 if owned else
           (P(d_a32), P(d_ctl), P(d_si), P(d_el), P(d_el32)))
  args = lambda dwell, stream: (
      C.byref(b.lattice_tables.c), C.byref(b.c), C.byref(spec.c), P(acts32),
      mode, 1.42, dwell, t_steps, 2000000) + stage + (
          P(h_si), P(h_el) if want_elapsed else None, stream)
  nat.check(nat.lib.pd_rollout_actions_host_f32(*args(
      1500000, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))))
  np.testing.assert_array_equal(h_si.numpy(), gh.np_(si))
  if want_elapsed:
    np.testing.assert_array_equal(h_el.numpy().astype(np.int64), gh.np_(el))
  else:
    assert (h_el.numpy() == GUARD).all()
  assert (h_si_all[t_steps].numpy() == GUARD).all()
  assert (h_el_all[t_steps].numpy() == GUARD).all()
  assert (gh.np_(d_si[t_steps]) == GUARD).all()
  assert (gh.np_(d_el[t_steps]) == GUARD).all()
  assert (gh.np_(d_el32[t_steps]) == GUARD).all()
  assert (gh.np_(d_a32[t_steps]) == 1.5).all()
  assert (gh.np_(d_ctl[t_steps]) == 1.5).all()
  np.testing.assert_array_equal(gh.np_(b.si_idx), gh.np_(a.si_idx))
  np.testing.assert_array_equal(gh.np_(b.sim_time_us), gh.np_(a.sim_time_us))
  np.testing.assert_array_equal(gh.np_(b.n_events), gh.np_(a.n_events))
  np.testing.assert_array_equal(gh.np_(b.ctrl_count), gh.np_(a.ctrl_count))
  return args


def test_rollout_host_compact_formats(eng):
  """pd_rollout_actions_host_f32: float32 actions in (action_spec dtype,
  action_adapters.py:202-216), int32 elapsed microseconds out; equals the
  device rollout on the widened actions.  (4096, 300), (4112, 100), (2048,
  200) and the 4096-env cases after them take the streamed form (one launch
  that follows the H2D copy, writer CTAs; a ragged last result block; actions
  that look like the "not arrived" pattern), (700, 9) and (4100, 70) the
  chunked copy-engine pipeline."""
  from putting_dune_b200 import _native as nat
  rel = nat.ACTION_RELATIVE_TO_SILICON
  args = _host_f32_case(eng, 4096, 300, rel)
  _host_f32_case(eng, 4112, 100, rel)
  _host_f32_case(eng, 2048, 200, rel, rate_fn=po.RATE_SIMPLE)
  _host_f32_case(eng, 4096, 64, nat.ACTION_DIRECT)
  _host_f32_case(eng, 4096, 96, rel, want_elapsed=False)
  _host_f32_case(eng, 4096, 80, rel, poison=True)
  # stagings kept by the library (all five NULL): the first call fills them,
  # the following ones find them re-filled behind the previous call; sizes
  # going up and down, one output only, and the chunked form in between
  for n, t_steps, kw in ((4096, 128, {}), (4096, 128, {}), (2048, 160, {}),
                         (4096, 200, {}), (4096, 64, {'want_elapsed': False}),
                         (4096, 128, {}), (700, 9, {}), (4096, 128, {})):
    _host_f32_case(eng, n, t_steps, rel, owned=True, **kw)
  # a call with the caller's stagings in between uses (and dirties) the same
  # control words: the next library-owned call must not assume them ready
  _host_f32_case(eng, 4096, 128, rel, owned=True)
  _host_f32_case(eng, 4096, 128, rel)
  _host_f32_case(eng, 4096, 128, rel, owned=True)
  _host_f32_case(eng, 4096, 128, rel, owned=True)
  _host_f32_case(eng, 700, 9, rel)
  _host_f32_case(eng, 4100, 70, rel)
  with pytest.raises(nat.NativeError, match='int32'):
    nat.check(nat.lib.pd_rollout_actions_host_f32(*args(3000000000, None)))


def test_envbatch_rollout_host(eng):
  """EnvBatch.rollout_host / BatchedSimulator.rollout_host: host arrays in,
  pinned host tensors out, equal to the device rollout (float32 actions ->
  int32 elapsed, float64 -> int64; pageable and pinned inputs; `out`
  re-used)."""
  import datetime as dt
  from putting_dune_b200 import _native as nat
  from putting_dune_b200 import simulator as sim_lib
  rel = nat.ACTION_RELATIVE_TO_SILICON
  n, t_steps = 4096, 72
  spec = gh.rate_spec(po.RATE_PRIOR)
  rng = np.random.default_rng(4)
  out32 = None
  for dtype, pinned in ((np.float32, False), (np.float32, True),
                        (np.float64, False)):
    acts = rng.uniform(-1, 1, size=(t_steps, n, 2)).astype(dtype)
    a = eng.EnvBatch(n, seed=8)
    b = eng.EnvBatch(n, seed=8)
    a.reset()
    b.reset()
    si, el = a.rollout(torch.as_tensor(acts).double(), 1500000, spec,
                       record=True, action_mode=rel)
    src = torch.as_tensor(acts).pin_memory() if pinned else acts
    got = b.rollout_host(src, 1500000, spec, action_mode=rel,
                         out=out32 if dtype == np.float32 else None)
    if dtype == np.float32:
      out32 = got
    h_si, h_el = got
    assert h_si.is_pinned() and h_el.is_pinned()
    assert h_el.dtype == (torch.int32 if dtype == np.float32 else torch.int64)
    np.testing.assert_array_equal(h_si.numpy(), gh.np_(si))
    np.testing.assert_array_equal(h_el.numpy().astype(np.int64), gh.np_(el))
    np.testing.assert_array_equal(gh.np_(b.si_idx), gh.np_(a.si_idx))
  with pytest.raises(ValueError, match='T, E, 2'):
    b.rollout_host(np.zeros((3, n + 1, 2), np.float32), 1500000, spec)
  with pytest.raises(ValueError, match='host buffers'):
    b.rollout_host(torch.zeros((3, n, 2), device=b.device), 1500000, spec)
  s = sim_lib.BatchedSimulator(n, seed=8)
  with pytest.raises(RuntimeError, match='reset'):
    s.rollout_host(np.zeros((4, n, 2), np.float32), dt.timedelta(seconds=1.5))
  s.reset()
  h_si, h_el = s.rollout_host(
      rng.uniform(0.45, 0.55, size=(64, n, 2)).astype(np.float32),
      dt.timedelta(seconds=1.5))
  assert h_si.shape == (64, n) and (h_el.numpy() >= 3500000).all()


def _prof_e2e_digest(**env):
  import subprocess
  import sys
  script = os.path.join(os.path.dirname(__file__), '..', 'profiles',
                        'prof_e2e.py')
  out = subprocess.run([sys.executable, script],
                       env=dict(os.environ, **env), check=True,
                       capture_output=True, text=True, timeout=600).stdout
  assert 'streamed=' + env.get('PD_HOST_STREAMED', '0') in out
  return out.strip().rsplit('digest ', 1)[1]


def test_rollout_host_streamed_equals_chunked():
  """The two forms of pd_rollout_actions_host_f32 return the same bytes: the
  chunked copy-engine pipeline (default; prior / simple rates run the fast
  kernels on the float32 actions directly) and the opt-in streamed launch
  (PD_HOST_STREAMED=1: k_rollout_pre follows the H2D copy element by
  element), and so does the float64 code under the chunks (PD_FAST=0).  The
  environment variables are read once per process, hence the subprocesses."""
  digests = [_prof_e2e_digest(PD_HOST_STREAMED=s, PD_FAST=f, REPS='3')
             for s, f in (('0', '1'), ('1', '1'), ('0', '0'))]
  assert digests[0] == digests[1] == digests[2]


@pytest.mark.parametrize('n,t_steps,reps', [(4096, 256, 4000),
                                            (4112, 77, 3000),
                                            (8208, 33, 3000)])
def test_streamed_rollout_stress(n, t_steps, reps):
  """The streamed launch reads its action staging while a copy-engine copy is
  still writing it and recognises data by the absence of the 0xFF fill
  pattern in either word of an element -- behaviour of this hardware, not of
  the CUDA memory model (which is why the form is opt-in).  10^4 calls at
  three sizes, two of them with rows that are not whole cache lines: a
  checksum over the results of every call equals the chunked pipeline's."""
  env = dict(N_ENVS=str(n), N_STEPS=str(t_steps), REPS=str(reps),
             CHECKSUM='1')
  assert (_prof_e2e_digest(PD_HOST_STREAMED='1', **env) ==
          _prof_e2e_digest(PD_HOST_STREAMED='0', **env))


def test_race_sampling_matches_direct_method_in_distribution(eng):
  """pd_set_option('race_sampling', 1): events by competing exponentials
  (every neighbour draws Exp(rate_i), the smallest wins) instead of the
  reference's direct method.  Not draw for draw -- so never a parity path --
  but equal in distribution: over 2^20 envs the counts of (no hop, hop to
  neighbour 0 / 1 / 2) after one control, and the histogram of hops per
  control, pass a chi-square test against the direct method's, for the
  simple and the prior rates."""
  from putting_dune_b200 import _native as nat
  n = 1 << 20
  rng = np.random.default_rng(3)
  acts = rng.uniform(-1, 1, size=(1, n, 2))
  for rate_fn, dwell in ((po.RATE_SIMPLE, 1500000), (po.RATE_PRIOR, 5000000)):
    spec = gh.rate_spec(rate_fn)
    hists = []
    for race in (0, 1):
      nat.check(nat.lib.pd_set_option(b'race_sampling', race))
      b = eng.EnvBatch(n, seed=5 + race)  # independent draws
      b.reset()
      si0 = gh.np_(b.si_idx).copy()
      nbr = gh.np_(b.lattice_tables.nbr)[:, :3]
      b.rollout(acts, dwell, spec, action_mode=nat.ACTION_RELATIVE_TO_SILICON)
      hops = gh.np_(b.n_transitions)
      si1 = gh.np_(b.si_idx)
      one = hops == 1
      slot = np.argmax(nbr[si0[one]] == si1[one][:, None], axis=1)
      first = np.array([(hops == 0).sum()] +
                       [(slot == k).sum() for k in range(3)] +
                       [(hops >= 2).sum()], dtype=np.float64)
      per = np.bincount(np.minimum(hops, 4), minlength=5).astype(np.float64)
      hists.append((first, per))
    nat.check(nat.lib.pd_set_option(b'race_sampling', 0))
    for a, c in zip(hists[0], hists[1]):
      # two-sample chi-square, 4 degrees of freedom: 0.999 quantile = 18.5
      chi2 = (((a - c) ** 2) / np.maximum(a + c, 1.0)).sum()
      assert chi2 < 18.5, (rate_fn, chi2, a, c)
    assert hists[0][0][1:].sum() > 0.05 * n


GMM_PARAMS = {  # graphene_test.py:337-345 parameter set (as in make_golden)
    'max_rate': 5.0,
    'mixture_weights': np.asarray((0.3, 0.3, 0.2, 0.1, 0.1)),
    'loc_distances': np.asarray((0.0, 1.0, 0.5, 0.0, 1.5)),
    'variances': np.asarray(((0.1, 0.1), (1.0, 1.0), (2.0, 0.5), (0.5, 2.0),
                             (0.01, 0.01))),
}


def test_gmm_rate_function(eng, golden_dir):
  """graphene.py:279-390 GaussianMixtureRateFunction on the device: rates vs
  the reference's scipy evaluation, trajectories of the unmodified reference,
  and 4096-env parity with the oracle (float64 total rate path)."""
  spec = gh.rate_spec(po.RATE_GMM, gmm=GMM_PARAMS)
  fix = np.load(os.path.join(golden_dir, 'rates_reference.npz'))
  n = fix['beam'].shape[0]
  st = po.make_state(n, int(fix['seed']))
  po.reset(st)
  b = gh.batch_from_oracle(st)
  r, nb = b.rates(fix['beam'], spec)
  np.testing.assert_array_equal(gh.np_(nb), fix['succ_gmm'])
  np.testing.assert_allclose(gh.np_(r), fix['rates_gmm'].astype(np.float32),
                             rtol=2e-7, atol=1e-37)
  # reference golden trajectories
  fix = np.load(os.path.join(golden_dir, 'events_gmm.npz'))
  controls, dwell = fix['controls'], fix['dwell_us']
  b = eng.EnvBatch(controls.shape[1], seed=int(fix['seed']))
  b.reset()
  for t in range(controls.shape[0]):
    out = b.step_and_image(controls[t], dwell[t], spec)
    np.testing.assert_array_equal(gh.np_(b.si_idx), fix['si'][:, t])
    np.testing.assert_array_equal(gh.np_(out.elapsed_us),
                                  fix['elapsed_us'][:, t])
  # oracle parity at BASELINE config-2 size, through both kernels
  n, seed = 4096, 61
  st = po.make_state(n, seed)
  po.reset(st)
  b = gh.batch_from_oracle(st)
  rng = np.random.default_rng(3)
  for _ in range(25):
    ctl = gh.closed_loop_control(st, rng)[:, None, :]
    want = po.step_and_image(st, ctl, 1500000, rate_fn=po.RATE_GMM,
                             gmm=GMM_PARAMS)
    out = b.step_and_image(ctl, 1500000, spec)
    np.testing.assert_array_equal(gh.np_(b.si_idx), st.si_idx)
    np.testing.assert_array_equal(gh.np_(out.events), want['events'])
  assert st.n_transitions.sum() > 20000
  big = eng.EnvBatch(200000, seed=seed)
  small = eng.EnvBatch(4096, seed=seed)
  big.reset()
  small.reset()
  ctl = 0.5 + rng.uniform(-0.05, 0.05, size=(200000, 1, 2))
  big.step_and_image(ctl, 1500000, spec)
  small.step_and_image(ctl[:4096], 1500000, spec)
  np.testing.assert_array_equal(gh.np_(big.si_idx)[:4096], gh.np_(small.si_idx))


def test_gmm_facade_class():
  # graphene_test.py:312-330, :355-359
  import putting_dune_b200 as pd
  fn = pd.graphene.GaussianMixtureRateFunction(
      max_rate=1.0, mixture_weights=np.asarray((0.8, 0.2)),
      loc_distances=np.asarray((0.0, 1.0)),
      variances=np.asarray(((0.1, 0.1), (1.0, 1.0))))
  m = pd.graphene.PristineSingleDopedGraphene(rate_function=fn)
  m.reset(np.random.default_rng(0))
  rates = fn(m.grid, pd.geometry.Point(m.get_silicon_position()))
  assert abs(max(s.rate for s in rates.successor_states) - 1.0) < 0.05
  rng = np.random.default_rng(0)
  a = pd.graphene.GaussianMixtureRateFunction.sample_new(rng)
  b = pd.graphene.GaussianMixtureRateFunction.sample_new(rng)
  assert a != b and a == a
  import datetime as dt
  m.apply_control(np.random.default_rng(0), pd.microscope_utils.BeamControl(
      pd.geometry.Point(m.get_silicon_position()), dt.timedelta(seconds=5)))
