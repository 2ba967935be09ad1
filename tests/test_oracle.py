"""CPU tests: the oracle against the reference's golden vectors.

The fixtures under tests/golden/ were produced by the unmodified reference
(tests/golden/make_golden.py); the known-answer cases restate the goldens of
the reference's own tests.
"""

import os

import numpy as np
import pytest

from oracle import pdune_oracle as po


# -- Philox4x32-10 known-answer vectors (Random123 kat_vectors) ---------------
@pytest.mark.parametrize('ctr,key,want', [
    ([0, 0, 0, 0], [0, 0],
     [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2,
     [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
     [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
])
def test_philox_kat(ctr, key, want):
  got = po.philox4x32_10(*[np.uint32(c) for c in ctr], key[0], key[1])
  assert [int(g) for g in got] == want


def test_u53_range_and_exactness():
  assert po.u53(0, 0) == 0.0
  assert po.u53(0xffffffff, 0xffffffff) == 1.0 - 2.0**-53


# -- lattice facts --------------------------------------
def test_lattice_shape_and_bonds():
  base = po.base_lattice(50)
  assert base.shape == (1881, 2)
  tab = po.neighbor_table(50)
  d = np.linalg.norm(base[tab] - base[:, None, :], axis=2)
  interior = d[:, 2] < 1.43
  assert (~interior).sum() == 121  # edge sites with < 3 bonded neighbours
  # graphene_test.py:41-87: three neighbours at exactly one bond length.
  np.testing.assert_allclose(d[interior], 1.42, atol=1e-7)
  # canonical order: ascending index among equidistant neighbours
  assert (np.diff(tab[interior], axis=1) > 0).all()


def test_si_never_initialised_on_edge():
  # graphene_test.py:283-299 (grid_columns=10, many resets)
  st = po.make_state(100, seed=3, num_cols=10)
  po.reset(st)
  p = po.site_positions(st, st.nbr[st.si_idx], np.arange(100))
  c = po.site_positions(st, st.si_idx, np.arange(100))
  d = np.linalg.norm(p - c[:, None, :], axis=2)
  assert (d[:, 2] <= 1.42 + 1e-3).all()


GMM_PARAMS = {  # graphene_test.py:337-345 parameter set (as in make_golden)
    'max_rate': 5.0,
    'mixture_weights': np.asarray((0.3, 0.3, 0.2, 0.1, 0.1)),
    'loc_distances': np.asarray((0.0, 1.0, 0.5, 0.0, 1.5)),
    'variances': np.asarray(((0.1, 0.1), (1.0, 1.0), (2.0, 0.5), (0.5, 2.0),
                             (0.01, 0.01))),
}


PRIOR_CUSTOM = {  # as in make_golden: HumanPriorRatePredictor(mean, cov, max)
    'mean': (0.7, 0.15),
    'cov': ((0.12, 0.03), (0.03, 0.07)),
    'max_rate': 0.4,
}


# -- event-path golden vectors from the reference ------------------------------
def expand_controls(fix):
  """(controls float64 [T, E, C, 2], dwell_us int64 [T, E, C]) of an event
  fixture; the large ones store float32-valued controls and one dwell time
  per env."""
  controls = np.asarray(fix['controls'], dtype=np.float64)
  dwell = np.asarray(fix['dwell_us'])
  if dwell.ndim == 1:
    dwell = np.broadcast_to(dwell[None, :, None],
                            controls.shape[:3]).astype(np.int64)
  return controls, dwell


def _replay(fix):
  seed = int(fix['seed'])
  rate_fn = int(fix['rate_fn'])
  controls, dwell = expand_controls(fix)
  n_steps, n_envs = controls.shape[:2]
  st = po.make_state(n_envs, seed)
  po.reset(st)
  reset_state = dict(si=st.si_idx.copy(), fov=st.fov.copy(),
                     scale=st.fov_scale.copy(), ip=st.image_params.copy())
  log = po.EventLog([], [], [], [], [], [])
  si = np.zeros((n_envs, n_steps), np.int32)
  el = np.zeros((n_envs, n_steps), np.int64)
  fov = np.zeros((n_envs, n_steps, 4))
  for t in range(n_steps):
    out = po.step_and_image(st, controls[t], dwell[t], rate_fn=rate_fn,
                            log=log, gmm=GMM_PARAMS)
    si[:, t], el[:, t], fov[:, t] = st.si_idx, out['elapsed_us'], st.fov
  trans = sorted(zip(log.env, log.ctrl_seq, log.elapsed_us, log.new_si))
  return st, reset_state, si, el, fov, np.asarray(trans, dtype=np.int64)


@pytest.mark.parametrize('name', [
    'events_simple.npz', 'events_prior.npz', 'events_gmm.npz',
    # 256 envs x 100 steps per rate function; one env x 1000 steps (BASELINE
    # configs[0]) -- the unmodified reference, env by env
    'events_simple_large.npz', 'events_prior_large.npz',
    'events_prior_single1000.npz'])
def test_oracle_matches_reference_trajectories(golden_dir, name):
  fix = np.load(os.path.join(golden_dir, name))
  st, r0, si, el, fov, trans = _replay(fix)
  if 'fov_last' in fix:  # compact fixture
    np.testing.assert_array_equal(r0['si'], fix['si0'])
    np.testing.assert_allclose(r0['fov'], fix['fov0'], rtol=0, atol=1e-13)
    np.testing.assert_allclose(r0['ip'], fix['image_params'], rtol=0, atol=0)
    np.testing.assert_array_equal(si, fix['si'])
    np.testing.assert_array_equal(el, fix['elapsed_us'])
    want = fix['transitions']
    want = want[np.lexsort((want[:, 2], want[:, 1], want[:, 0]))]
    np.testing.assert_array_equal(trans, want)
    np.testing.assert_allclose(fov[:, -1], fix['fov_last'], rtol=0,
                               atol=1e-13)
    assert want.shape[0] >= 100
    return
  # reset: Si site, FOV, image parameters
  np.testing.assert_array_equal(r0['si'], fix['si0'])
  np.testing.assert_allclose(r0['fov'], fix['fov0'], rtol=0, atol=1e-13)
  np.testing.assert_allclose(r0['scale'], fix['fov_scale'], rtol=0, atol=0)
  np.testing.assert_allclose(r0['ip'], fix['image_params'], rtol=0, atol=0)
  # bit-exact: Si lattice index, microsecond clocks, transition log
  np.testing.assert_array_equal(si, fix['si'])
  np.testing.assert_array_equal(el, fix['elapsed_us'])
  want = fix['transitions']
  want = want[np.lexsort((want[:, 2], want[:, 1], want[:, 0]))]
  np.testing.assert_array_equal(trans, want)
  np.testing.assert_allclose(fov, fix['fov'], rtol=0, atol=1e-13)
  # lattice positions of a few sites (BLAS matmul rounding in the reference)
  for e in range(si.shape[0]):
    p = po.all_positions(st, e)[[0, 1, 940, 1880]]
    np.testing.assert_allclose(p, fix['sample_positions'][e], rtol=0,
                               atol=2e-14)
  # observed grid right after reset for the first envs
  st2 = po.make_state(4, int(fix['seed']))
  po.reset(st2)
  for e in range(4):
    q, z, _ = po.get_atoms_in_bounds(st2, e)
    np.testing.assert_array_equal(z, fix[f'obs0_numbers_{e}'])
    np.testing.assert_allclose(q, fix[f'obs0_positions_{e}'], rtol=0,
                               atol=1e-13)


def test_oracle_rates_match_reference(golden_dir):
  fix = np.load(os.path.join(golden_dir, 'rates_reference.npz'))
  n = fix['beam'].shape[0]
  st = po.make_state(n, int(fix['seed']))
  po.reset(st)
  mlp = po.MlpParams(**{k: fix[f'mlp_{k}'] for k in (
      'bn_scale', 'bn_offset', 'bn_mean', 'bn_var', 'w0', 'b0', 'w1', 'b1',
      'w2', 'b2')})
  envs = np.arange(n)
  r64, nbr = po.rates_for(st, envs, fix['beam'], po.RATE_GMM, gmm=GMM_PARAMS,
                          keep64=True)
  np.testing.assert_array_equal(nbr, fix['succ_gmm'])
  np.testing.assert_allclose(r64, fix['rates_gmm'], rtol=1e-12, atol=1e-300)
  for name, rate_fn, rtol in (('simple', po.RATE_SIMPLE, 0.0),
                              ('prior', po.RATE_PRIOR, 1e-6),
                              ('learned', po.RATE_LEARNED, 1e-6),
                              ('prior_custom', po.RATE_PRIOR, 1e-6)):
    r32, nbr = po.rates_for(
        st, envs, fix['beam'], rate_fn, mlp,
        prior=PRIOR_CUSTOM if name == 'prior_custom' else None)
    np.testing.assert_array_equal(nbr, fix[f'succ_{name}'])
    want = fix[f'rates_{name}']
    if rtol == 0.0:
      np.testing.assert_array_equal(r32, want)
    else:
      np.testing.assert_allclose(r32, want, rtol=rtol, atol=1e-30)


def test_standardize_matches_reference(golden_dir):
  fix = np.load(os.path.join(golden_dir, 'standardize_reference.npz'))
  nb, nn, order = po.standardize_beam_and_neighbors(fix['beam'], fix['nbr'])
  np.testing.assert_array_equal(order, fix['order'])
  np.testing.assert_allclose(nb, fix['new_beam'], rtol=0, atol=1e-14)
  np.testing.assert_allclose(nn, fix['new_nbr'], rtol=0, atol=1e-14)


# -- known-answer cases restated from the reference's own tests ----------------
@pytest.mark.parametrize('nbr,order,beam,new_beam,new_nbr', [
    # data_utils_test.py:149-198 'aligned'
    ([[1, 0], [-0.5, np.sqrt(3) / 2], [-0.5, -np.sqrt(3) / 2]], [0, 1, 2],
     [1, 0], [1.0, 0.0],
     [[1, 0], [-0.5, np.sqrt(3) / 2], [-0.5, -np.sqrt(3) / 2]]),
    # 'rotated'
    ([[0.5, np.sqrt(3) / 2], [0.5, -np.sqrt(3) / 2], [-1.0, 0.0]], [0, 2, 1],
     [0, 1], [np.sqrt(3) / 2, 0.5],
     [[1.0, 0.0], [-0.5, -np.sqrt(3) / 2], [-0.5, np.sqrt(3) / 2]]),
])
def test_standardize_goldens(nbr, order, beam, new_beam, new_nbr):
  nb, nn, od = po.standardize_beam_and_neighbors(
      np.asarray([beam], dtype=float), np.asarray([nbr], dtype=float))
  np.testing.assert_allclose(nb[0], new_beam, rtol=1e-6)
  np.testing.assert_allclose(nn[0], new_nbr, atol=1e-9)
  np.testing.assert_array_equal(od[0], order)


def test_geometry_goldens():
  # geometry_test.py:25-59
  xy = np.array([[1.0, 0.0], [0.0, -1.0], [-1.0, -1.0]])
  np.testing.assert_allclose(po.get_angles(xy),
                             [0.0, -np.pi / 2, -3 * np.pi / 4])
  np.testing.assert_allclose(po.rotate_coordinates(xy, np.pi / 2),
                             [[0.0, 1.0], [1.0, 0.0], [1.0, -1.0]],
                             atol=1e-12)


def test_frame_transform_goldens():
  # microscope_utils_test.py:121-288: FOV [(-5, 0), (5, 20)]
  fov = np.array([[-5.0, 0.0, 5.0, 20.0]])
  np.testing.assert_allclose(
      po.microscope_to_material(fov, np.array([[0.5, 1.0]])), [[0.0, 20.0]])
  np.testing.assert_allclose(
      po.microscope_to_material(fov, np.array([[-3.0, 2.5]])),
      [[-35.0, 50.0]])
  # simulator_test.py:86-115: (0, 1) in FOV [(-5.5, -6.3), (12, 9.1)]
  fov = np.array([[-5.5, -6.3, 12.0, 9.1]])
  np.testing.assert_allclose(
      po.microscope_to_material(fov, np.array([[0.0, 1.0]])), [[-5.5, 9.1]])
  back = po.material_to_microscope(fov, np.array([[-5.5, 9.1]]))
  np.testing.assert_allclose(back, [[0.0, 1.0]], atol=1e-15)


def test_simple_rate_closed_forms():
  # graphene.py:151-166: beam on a neighbour -> 1; beam on the Si -> 1/17.
  p_si = np.zeros((1, 2))
  p_n = po.BOND * np.array([[[1.0, 0.0], [-0.5, np.sqrt(3) / 2],
                             [-0.5, -np.sqrt(3) / 2]]])
  r = po.simple_rates(p_n[:, 0], p_si, p_n)
  assert r[0, 0] == 1.0
  r = po.simple_rates(p_si, p_si, p_n)
  np.testing.assert_allclose(r[0], 1 / 17, rtol=1e-14)


def test_zero_rates_elapsed_time_sum():
  # simulator_test.py:147-168: zero rates => elapsed == sum(dwell) + image.
  st = po.make_state(3, seed=1)
  po.reset(st)
  zero = lambda s, idx, beam, it: (np.zeros((idx.size, 3), np.float32),
                                   s.nbr[s.si_idx[idx]])
  dwell = np.array([1500000, 3000000, 7230000], dtype=np.int64)
  ctl = np.full((3, 3, 2), 0.5)
  e = st.num_envs
  elapsed = np.zeros(e, dtype=np.int64)
  for c in range(3):
    beam = po.microscope_to_material(st.fov, ctl[:, c])
    out = po.apply_control(st, beam, np.full(e, dwell[c]), rates_override=zero)
    assert (out['transitions'] == 0).all() and (out['events'] == 1).all()
    elapsed += dwell[c]
  assert (elapsed + 3500000 == 15230000).all()


@pytest.mark.parametrize('pct,expect', [
    ((0.2, 0.4), True), ((0.35, 0.95), True), ((0.751, 0.249), True),
    ((0.749, 0.250), False)])
def test_recentre_thresholds(pct, expect):
  # simulator_test.py:263-335
  st = po.make_state(1, seed=5)
  po.reset(st)
  p = po.site_positions(st, st.si_idx, np.arange(1))[0]
  ll = p - 10.0 * np.asarray(pct)
  st.fov[0] = np.concatenate((ll, ll + 10.0))
  out = po.step_and_image(st, np.array([[[1.0, 1.0]]]), np.array([[0]]))
  assert bool(out['recentred'][0]) == expect
  q = po.material_to_microscope(st.fov, p[None])[0]
  want = (0.5, 0.5) if expect else pct
  np.testing.assert_allclose(q, want, atol=1e-12)
  assert out['elapsed_us'][0] == (4000000 if expect else 2000000)


def test_multiple_transitions_at_high_rate():
  # graphene_test.py:228-281: rate 5 each => several hops within a dwell.
  st = po.make_state(64, seed=9)
  po.reset(st)
  five = lambda s, idx, beam, it: (np.full((idx.size, 3), 5.0, np.float32),
                                   s.nbr[s.si_idx[idx]])
  out = po.apply_control(st, np.zeros((64, 2)), np.full(64, 1500000),
                         rates_override=five)
  assert out['transitions'].mean() > 10
  assert (out['events'] == out['transitions'] + 1).all()


def test_seconds_to_us_rounding():
  # CPython timedelta: fractional microseconds round half to even.
  t = np.array([0.0000005, 0.0000015, 1.0000025, 3600.0, 2.9999999])
  import datetime as dt
  want = [dt.timedelta(seconds=float(x)) // dt.timedelta(microseconds=1)
          for x in t]
  np.testing.assert_array_equal(po.seconds_to_us(t), want)
  rng = np.random.default_rng(0)
  t = rng.exponential(2.0, size=20000)
  want = [dt.timedelta(seconds=float(x)) // dt.timedelta(microseconds=1)
          for x in t]
  np.testing.assert_array_equal(po.seconds_to_us(t), want)


# -- renderer oracle vs frames produced by the reference's imaging.py ----------
def test_imaging_oracle_matches_reference_frames(golden_dir):
  from oracle import pdune_oracle_imaging as oi
  fix = np.load(os.path.join(golden_dir, 'frames_reference.npz'))
  seed, size = int(fix['seed']), int(fix['size'])
  st = po.make_state(4, seed)
  po.reset(st)
  for e in range(4):
    got = oi.render_env(st, e, size=size, stages=True)
    for stage in ('clean', 'blur', 'poisson', 'jitter', 'final'):
      np.testing.assert_allclose(got[stage], fix[f'{stage}_{e}'], rtol=0,
                                 atol=2e-7, err_msg=f'{stage} env {e}')
    assert got['final'].min() >= 0 and got['final'].max() <= 1  # imaging_test
  st = po.make_state(1, seed)
  po.reset(st)
  full = oi.render_env(st, 0, size=512)
  assert full.shape == (512, 512)  # imaging_test.py:51-78
  np.testing.assert_allclose(full.mean(axis=1), fix['full512_rowmean'],
                             atol=1e-12)
  np.testing.assert_allclose(full[200:232, 300:332], fix['full512_patch'],
                             atol=1e-12)


def test_perception_oracle_matches_reference(golden_dir):
  """sample_noisy_image_parameters and generate_grid_mask vs the reference."""
  from oracle import pdune_oracle_imaging as oi
  fix = np.load(os.path.join(golden_dir, 'perception_reference.npz'))
  seed, size = int(fix['seed']), int(fix['size'])
  n = fix['noisy_params'].shape[0]
  st = po.make_state(n, seed)
  po.reset(st)
  for e in range(n):
    got = oi.mask_env(st, e, size, float(fix[f'mask_exponent_{e}']))
    np.testing.assert_array_equal(got, fix[f'mask_{e}'])
    assert set(got.reshape(-1).tolist()) <= {0, 6, 14}
  for e in (0, 1, 2):  # imaging.py:129-168 buffered clean image
    q, z = oi.grid_in_microscope_frame(st, e)
    f = st.fov[e]
    got = oi.clean_image(q, z, f[2] - f[0], f[3] - f[1],
                         float(st.image_params[e, 0]), size,
                         float(fix[f'buffer_{e}']))
    np.testing.assert_allclose(got, fix[f'buffered_clean_{e}'], rtol=0,
                               atol=2e-7)
  po.sample_noisy_image_parameters(st)
  np.testing.assert_array_equal(st.image_params, fix['noisy_params'])


def test_clahe_restatement_properties():
  from oracle import pdune_oracle_imaging as oi
  rng = np.random.default_rng(0)
  img = rng.random((128, 128)) ** 3
  out = oi.equalize_adapthist(img)
  assert out.shape == img.shape and out.min() == 0.0 and out.max() == 1.0
  # contrast-limited equalisation flattens the histogram
  assert np.abs(np.median(out) - 0.5) < np.abs(np.median(img) - 0.5)
  # clip_histogram conserves counts and respects the limit when it can
  hist = rng.integers(0, 80, size=256)
  clipped = oi.clip_histogram(hist, 40)
  assert clipped.sum() == hist.sum() or clipped.max() <= 40
  assert clipped.max() <= 40 + 1


# -- whole episodes (BASELINE configs[4]) vs the reference's eval_lib ----------
@pytest.mark.parametrize('name,rate_fn', [('simple', po.RATE_SIMPLE),
                                          ('prior', po.RATE_PRIOR)])
def test_episode_oracle_matches_reference_eval(golden_dir, name, rate_fn):
  from oracle import pdune_oracle_episode as oe
  fix = np.load(os.path.join(golden_dir, 'episodes_reference.npz'))
  n = fix[f'{name}_reached'].shape[0]
  st = po.make_state(n, int(fix[f'{name}_philox_seed']))
  got = oe.run_episodes(st, oe.EpisodeConfig(rate_fn=rate_fn))
  np.testing.assert_array_equal(got['reached'], fix[f'{name}_reached'])
  np.testing.assert_array_equal(got['num_actions'], fix[f'{name}_num_actions'])
  np.testing.assert_array_equal(got['env_seconds'],
                                fix[f'{name}_env_seconds'])
  np.testing.assert_allclose(got['total_reward'], fix[f'{name}_total_reward'],
                             rtol=1e-15)
  agg = oe.aggregate(got)
  np.testing.assert_allclose(
      [agg['average_num_times_reached_goal'], agg['average_num_actions_taken'],
       agg['average_environment_seconds_to_goal'],
       agg['average_total_reward']], fix[f'{name}_aggregate'], rtol=1e-12)


def test_aggregate_results_golden():
  # eval_lib_test.py:90-129 arithmetic: averages over successful episodes only
  from oracle import pdune_oracle_episode as oe
  res = {'reached': np.array([True, False, True]),
         'num_actions': np.array([10, 99, 20]),
         'env_seconds': np.array([70.0, np.nan, 150.0]),
         'total_reward': np.array([0.9, 0.0, 0.5])}
  agg = oe.aggregate(res)
  assert agg['average_num_times_reached_goal'] == 2 / 3
  assert agg['average_num_actions_taken'] == 15.0
  assert agg['average_environment_seconds_to_goal'] == 110.0
  assert agg['average_total_reward'] == 0.7


@pytest.mark.parametrize('si,action,want', [
    # action_adapters_test.py:124-186 (FOV 10 A wide, max distance 1.42 A)
    ((0.5, 0.75), (0.0, 0.2), (0.5, 0.7784)),
    ((0.31, 0.31), (-0.1, 0.0), (0.2958, 0.31)),
    ((0.92, 0.11), (-1.0, 0.75), (0.778, 0.2165)),
])
def test_relative_to_silicon_adapter_goldens(si, action, want):
  from oracle import pdune_oracle_episode as oe
  st = po.make_state(1, seed=2)
  po.reset(st)
  p = po.site_positions(st, st.si_idx, np.arange(1))[0]
  ll = p - 10.0 * np.asarray(si)
  st.fov[0] = np.concatenate((ll, ll + 10.0))
  got = oe.relative_to_silicon_controls(st, np.asarray([action]))
  np.testing.assert_allclose(got[0], want, atol=1e-9)


def test_gmm_max_rate_at_mode():
  # graphene_test.py:312-330: the dominant mode sits on the Si -> max ~ 1.0
  gmm = {'max_rate': 1.0, 'mixture_weights': np.asarray((0.8, 0.2)),
         'loc_distances': np.asarray((0.0, 1.0)),
         'variances': np.asarray(((0.1, 0.1), (1.0, 1.0)))}
  st = po.make_state(4, seed=1)
  po.reset(st)
  p_si = po.site_positions(st, st.si_idx, np.arange(4))
  r, _ = po.rates_for(st, np.arange(4), p_si, po.RATE_GMM, gmm=gmm,
                      keep64=True)
  np.testing.assert_allclose(r.max(axis=1), 1.0, atol=0.05)


# -- batched dm_env layer vs the reference PuttingDuneEnvironment --------------
ENV_CASES = ((2, 0, 1.5, 1.5, 1.42, 600), (2, 1, 1.0, 5.0, 2.84, 600),
             (3, 1, 5.0, 5.0, 2.84, 7), (0, 0, 1.5, 1.5, 1.42, 600),
             (1, 0, 1.5, 1.5, 1.42, 600))


@pytest.mark.parametrize('case', range(5))
def test_env_oracle_matches_reference_environment(golden_dir, case):
  from oracle import pdune_oracle_env as oenv
  fix = np.load(os.path.join(golden_dir, 'env_reference.npz'))
  ad, ft, d0, d1, md, lim = ENV_CASES[case]
  cfg = oenv.EnvConfig(adapter=ad, features=ft, min_dwell_s=d0, max_dwell_s=d1,
                       max_distance=md, step_limit=lim)
  acts = fix[f'actions_{case}']
  env = oenv.OracleEnv(acts.shape[1], int(fix['seed']), cfg)
  for t in range(acts.shape[0]):
    ts = env.step(acts[t])
    np.testing.assert_array_equal(ts['step_type'], fix[f'step_type_{case}'][t])
    np.testing.assert_allclose(ts['reward'], fix[f'reward_{case}'][t],
                               rtol=1e-6)
    np.testing.assert_allclose(ts['discount'], fix[f'discount_{case}'][t],
                               rtol=1e-6)
    np.testing.assert_allclose(ts['observation'],
                               fix[f'observation_{case}'][t], rtol=0,
                               atol=2e-6)
  # putting_dune_environment_test.py:110: the first step of a fresh env resets
  assert (fix[f'step_type_{case}'][0] == 0).all()


def test_synthetic_data_oracle_matches_reference_helpers(golden_dir):
  """oracle/pdune_oracle_synth.py against the reference's own
  get_all_position_rotations / single_silicon_prior_rates /
  jnp_rotate_coordinates / rotate_attributes / rotate_index
  (tests/golden/synth_reference.npz)."""
  from oracle import pdune_oracle_synth as osy
  ref = np.load(os.path.join(golden_dir, 'synth_reference.npz'))
  for ns in (3, 6):
    pos, rf, state = ref[f'pos_{ns}'], ref[f'rf_{ns}'], ref[f'state_{ns}']
    np.testing.assert_allclose(osy.prior_rates(pos, ns), ref[f'rates_{ns}'],
                               rtol=1e-12, atol=0)
    np.testing.assert_allclose(
        osy.rotate_coordinates(pos, 2 * rf * np.pi / ns), ref[f'pos_rot_{ns}'],
        rtol=1e-12, atol=1e-15)
    # the composition inside sample_from_draws: feed draws that reproduce the
    # fixture's position, state and rotation factor
    z = (pos - osy.MEAN) / np.sqrt(1.5 * osy.COV)
    rates = ref[f'rates_{ns}']
    cdf = np.cumsum(rates / rates.sum(-1, keepdims=True), -1)
    lo = np.concatenate((np.zeros((len(pos), 1)), cdf[:, :-1]), 1)
    u_state = (0.5 * (lo + cdf))[np.arange(len(pos)), state]
    out = osy.sample_from_draws(z, u_state, (rf + 0.5) / ns,
                                np.ones(len(pos)),  # next_time = 0: transitions
                                np.full(len(pos), 0.5),
                                np.zeros((len(pos), 2)), ns, (0.0, 5.0))
    np.testing.assert_array_equal(out['next_state'][:, 0],
                                  ref[f'state_rot_{ns}'] + 1)
    np.testing.assert_allclose(out['rates'], ref[f'rates_rot_{ns}'],
                               rtol=1e-6)
    np.testing.assert_allclose(out['position'], ref[f'pos_rot_{ns}'],
                               rtol=1e-6, atol=1e-7)


def test_synthetic_data_oracle_network_mode():
  """NETWORK mode of the synthetic-data oracle (data_utils.py:201-234) on a
  network whose output is known in closed form: zero weights, so rates =
  softplus(b2)[:num_states] for every sample (parity unpinned: the
  reference's Haiku / jax.random are not in this container)."""
  from oracle import pdune_oracle_synth as osy
  f32 = np.float32
  want = np.array((0.5, 1.0, 2.5), dtype=np.float64)
  b2 = np.concatenate((np.log(np.expm1(want)), (7.0,))).astype(f32)
  net = {'w0': np.zeros((4, 1), f32), 'b0': np.zeros(1, f32),
         'w1': np.zeros((1, 64), f32), 'b1': np.zeros(64, f32),
         'w2': np.zeros((64, 4), f32), 'b2': b2}
  n = 50000
  out = osy.generate_synthetic_data_network(n, 11, 0, net)
  assert out['rates'].shape == (n, 3) and out['next_state'].shape == (n, 1)
  assert out['context'].shape == (n, 2) and out['position'].shape == (n, 2)
  np.testing.assert_allclose(out['rates'], np.tile(want, (n, 1)), rtol=1e-6)
  # x ~ N(0, I): context / position are its halves
  x = np.concatenate((out['context'], out['position']), 1).astype(np.float64)
  assert abs(x.mean()) < 0.02 and abs(x.std() - 1.0) < 0.02
  # P(transition) = E_T[1 - exp(-total T)], T ~ U(0, 5); branch ratios = rates
  total = want.sum()
  p_tr = 1.0 - (1.0 - np.exp(-5.0 * total)) / (5.0 * total)
  ns = out['next_state'][:, 0]
  assert abs((ns > 0).mean() - p_tr) < 0.01
  frac = np.bincount(ns[ns > 0] - 1, minlength=3) / (ns > 0).sum()
  np.testing.assert_allclose(frac, want / total, atol=0.01)
  other = osy.generate_synthetic_data_network(n, 11, 1, net)
  assert not np.array_equal(other['position'], out['position'])


def test_mlp_restatement_against_independent_float64():
  """a9 / a10 (parity unpinned: Haiku and TF are absent): the oracle's float32
  NumPy network against an independent float64 torch.nn.functional one."""
  from tests import mlp_independent as mi
  rng = np.random.default_rng(3)
  for hidden in ((32, 32), (128, 128), (256, 256), (48, 128)):
    models = []
    for seed in (1, 2, 3):
      m = po.MlpParams.synthetic(seed, hidden=hidden)
      m.bn_mean = rng.normal(0, 0.3, 2).astype(np.float32)
      m.bn_var = rng.uniform(0.5, 2.0, 2).astype(np.float32)
      m.bn_scale = rng.uniform(0.5, 1.5, 2).astype(np.float32)
      m.bn_offset = rng.normal(0, 0.2, 2).astype(np.float32)
      m.b0 = rng.normal(0, 0.1, hidden[0]).astype(np.float32)
      m.b1 = rng.normal(0, 0.1, hidden[1]).astype(np.float32)
      m.b2 = rng.normal(0, 0.1, 4).astype(np.float32)
      models.append(m)
    x = rng.uniform(-2.0, 2.0, size=(500, 2)).astype(np.float32)
    o = po.mlp_forward(models[0], x)
    want = mi.forward(models[0], x)
    assert np.abs(o - want).max() <= 3e-6 * np.abs(want).max()
    r = po.apply_model(models, x)
    want = mi.apply_model(models, x)
    assert np.abs(r - want).max() <= 3e-6 * np.abs(want).max()
