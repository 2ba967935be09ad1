"""GPU tests of the learned rate model (K2) through the C ABI.

The network itself is "parity unpinned" (no Haiku/TF here, no reference test
covers predict); the oracle is a NumPy restatement of the Haiku forward, and
the frame canonicalisation around it is pinned by the reference's own code
(tests/golden/rates_reference.npz).  Tolerance: FP32 FMA vs NumPy float32
matmul, |d| <= 2e-5 * max(rate) + 1e-7.
"""

import os

import numpy as np
import pytest
import torch

from oracle import pdune_oracle as po
from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu


def _tol(want):
  return 2e-5 * np.abs(want).max() + 1e-7


@pytest.mark.parametrize('hidden', [(32, 32), (64, 64), (128, 128),
                                    (256, 256), (48, 128)])
def test_learned_rates_match_oracle(hidden):
  n, seed = 1000, 3
  st = po.make_state(n, seed)
  po.reset(st)
  mlp = po.MlpParams.synthetic(1, hidden=hidden)
  rng = np.random.default_rng(2)
  mlp.bn_mean = rng.normal(0, 0.3, 2).astype(np.float32)
  mlp.bn_var = rng.uniform(0.5, 2.0, 2).astype(np.float32)
  mlp.bn_scale = rng.uniform(0.5, 1.5, 2).astype(np.float32)
  mlp.bn_offset = rng.normal(0, 0.2, 2).astype(np.float32)
  mlp.b0 = rng.normal(0, 0.1, hidden[0]).astype(np.float32)
  mlp.b1 = rng.normal(0, 0.1, hidden[1]).astype(np.float32)
  mlp.b2 = rng.normal(0, 0.1, 4).astype(np.float32)
  beam = po.site_positions(st, st.si_idx, np.arange(n)) + rng.uniform(
      -2.5, 2.5, size=(n, 2))
  b = gh.batch_from_oracle(st)
  r, nb = b.rates(beam, gh.rate_spec(po.RATE_LEARNED, mlp))
  want, nbr = po.rates_for(st, np.arange(n), beam, po.RATE_LEARNED, mlp)
  np.testing.assert_array_equal(gh.np_(nb), nbr)
  assert np.abs(gh.np_(r) - want).max() <= _tol(want)


def test_learned_rates_match_oracle_on_every_site():
  """The frame canonicalisation takes a cross-product shortcut where the two
  far neighbours lie clearly on opposite sides of the nearest one and the
  reference's atan2 / argsort form elsewhere (sheet-edge sites, whose three
  nearest atoms are not at 120 degrees): every site of the sheet, several
  beams each, against the oracle's atan2 form."""
  st0 = po.make_state(1, 0)
  n_sites = st0.nbr.shape[0]
  reps = 6
  n = n_sites * reps
  st = po.make_state(n, 21)
  po.reset(st)
  st.si_idx[:] = np.tile(np.arange(n_sites), reps)
  mlp = po.MlpParams.synthetic(4, hidden=(64, 64))
  rng = np.random.default_rng(8)
  mlp.b2 = rng.normal(0, 0.3, 4).astype(np.float32)
  beam = po.site_positions(st, st.si_idx, np.arange(n)) + rng.uniform(
      -3.0, 3.0, size=(n, 2))
  # beams almost on a neighbour and almost on the Si atom itself
  beam[::7] = po.site_positions(st, st.si_idx, np.arange(n))[::7] + 1e-9
  b = gh.batch_from_oracle(st)
  r, nb = b.rates(beam, gh.rate_spec(po.RATE_LEARNED, mlp))
  want, nbr = po.rates_for(st, np.arange(n), beam, po.RATE_LEARNED, mlp)
  np.testing.assert_array_equal(gh.np_(nb), nbr)
  assert np.abs(gh.np_(r) - want).max() <= _tol(want)


def test_learned_rates_match_reference_predict(golden_dir):
  """rates_reference.npz: the reference's own predict() body around the
  NumPy network."""
  fix = np.load(os.path.join(golden_dir, 'rates_reference.npz'))
  n = fix['beam'].shape[0]
  st = po.make_state(n, int(fix['seed']))
  po.reset(st)
  mlp = po.MlpParams(**{k: fix[f'mlp_{k}'] for k in (
      'bn_scale', 'bn_offset', 'bn_mean', 'bn_var', 'w0', 'b0', 'w1', 'b1',
      'w2', 'b2')})
  b = gh.batch_from_oracle(st)
  r, nb = b.rates(fix['beam'], gh.rate_spec(po.RATE_LEARNED, mlp))
  np.testing.assert_array_equal(gh.np_(nb), fix['succ_learned'])
  assert np.abs(gh.np_(r) - fix['rates_learned']).max() <= _tol(
      fix['rates_learned'])


def test_tensor_core_split_matches_reference_predict(golden_dir):
  """The same golden through the tcgen05 fp16 hi + lo path (what
  LearnedTransitionRatePredictor(tensor_core='auto') selects for 32 x 32)."""
  from putting_dune_b200 import engine
  fix = np.load(os.path.join(golden_dir, 'rates_reference.npz'))
  n = fix['beam'].shape[0]
  st = po.make_state(n, int(fix['seed']))
  po.reset(st)
  w = engine.MlpWeights(**{k: fix[f'mlp_{k}'] for k in
                           engine.MlpWeights.NAMES})
  b = gh.batch_from_oracle(st)
  r, nb = b.rates(fix['beam'], engine.RateSpec(po.RATE_LEARNED, mlp=w,
                                               tensor_core=2))
  np.testing.assert_array_equal(gh.np_(nb), fix['succ_learned'])
  assert np.abs(gh.np_(r) - fix['rates_learned']).max() <= _tol(
      fix['rates_learned'])


def test_learned_step_event_machinery_is_exact():
  """With the oracle fed the device's own rates, the learned-rate step must
  reproduce sites, counters and the transition log bit-exactly."""
  n, seed = 700, 12
  st = po.make_state(n, seed)
  po.reset(st)
  mlp = po.MlpParams.synthetic(5, hidden=(64, 64))
  mlp.b2 = np.array([0.5, 0.2, 0.8, 0.0], np.float32)  # lively rates
  spec = gh.rate_spec(po.RATE_LEARNED, mlp)
  b = gh.batch_from_oracle(st, log_capacity=192)
  shadow = gh.batch_from_oracle(st)

  def device_rates(state, idx, beam, it):
    shadow.si_idx.copy_(torch.as_tensor(state.si_idx, device=shadow.device))
    full = np.zeros((n, 2))
    full[idx] = beam
    r, _ = shadow.rates(full, spec)
    return gh.np_(r)[idx], state.nbr[state.si_idx[idx]]

  rng = np.random.default_rng(3)
  for step in range(6):
    ctl = np.stack([gh.closed_loop_control(st, rng) for _ in range(2)], axis=1)
    dwell = rng.integers(0, 4000000, size=(n, 2))
    dwell[::9, 0] = 0
    log = po.EventLog([], [], [], [], [], [])
    elapsed = np.zeros(n, dtype=np.int64)
    tr = np.zeros(n, dtype=np.int64)
    ev = np.zeros(n, dtype=np.int64)
    for c in range(2):
      beam = po.microscope_to_material(st.fov, ctl[:, c])
      o = po.apply_control(st, beam, dwell[:, c], log=log,
                           rates_override=device_rates)
      elapsed += dwell[:, c]
      tr += o['transitions']
      ev += o['events']
    elapsed += 2000000
    rec = po.silicon_outside_safe_area(st)
    po.recenter_fov(st, np.nonzero(rec)[0])
    elapsed[rec] += 2000000
    out = b.step_and_image(ctl, dwell, spec)
    np.testing.assert_array_equal(gh.np_(b.si_idx), st.si_idx)
    np.testing.assert_array_equal(gh.np_(out.elapsed_us), elapsed)
    np.testing.assert_array_equal(gh.np_(out.transitions), tr)
    np.testing.assert_array_equal(gh.np_(out.events), ev)
    np.testing.assert_array_equal(gh.np_(out.recentred).astype(bool), rec)
    np.testing.assert_allclose(gh.np_(b.fov), st.fov, rtol=0, atol=1e-13)
    cnt = gh.np_(out.log_count)
    assert cnt.max() <= 192 and not (gh.np_(b.status) & 2).any()
    got = sorted((e, int(gh.np_(out.log_elapsed_us)[e, k]),
                  int(gh.np_(out.log_site)[e, k]))
                 for e in np.nonzero(cnt)[0] for k in range(cnt[e]))
    want = sorted(zip(log.env, log.elapsed_us, log.new_si))
    assert got == want
  assert st.n_transitions.sum() > 500
  np.testing.assert_array_equal(gh.np_(b.ctrl_count),
                                st.ctrl_count.astype(np.int32))


def test_learned_step_close_to_independent_oracle():
  """Against the oracle's own NumPy network the trajectories agree except
  where a float32 rounding difference moves a waiting time across the end of
  the dwell (expected ~1e-6 per event)."""
  n, seed = 4096, 4
  st = po.make_state(n, seed)
  po.reset(st)
  mlp = po.MlpParams.synthetic(9, hidden=(128, 128))
  spec = gh.rate_spec(po.RATE_LEARNED, mlp)
  b = gh.batch_from_oracle(st)
  rng = np.random.default_rng(5)
  for _ in range(8):
    ctl = gh.closed_loop_control(st, rng)[:, None, :]
    po.step_and_image(st, ctl, 5000000, rate_fn=po.RATE_LEARNED, mlp=mlp)
    b.step_and_image(ctl, 5000000, spec)
  # float32 sums of the network run in a different order on the device
  # (rates agree to 2e-5): an event changes its outcome with ~1e-5, so of the
  # ~5e4 events of this test fewer than one is expected to; at most 4 envs
  same = gh.np_(b.si_idx) == st.si_idx
  assert (~same).sum() <= 4, int((~same).sum())
  assert st.n_transitions.sum() > 1000


@pytest.mark.parametrize('tensor_core', [0, 2])
def test_learned_step_at_the_benchmarked_config(tensor_core):
  """BASELINE configs[2] as bench.py's measure_mlp runs it: 65,536 envs, one
  control of 1.5 s per step, beams near the frame centre, the 64 x 64 network
  with the seeded synthetic weights -- three steps through the FP32 kernel
  and through the tcgen05 fp16 hi + lo kernel against the oracle's float64
  network.  A rate rounding difference (<= 2e-5 of the largest rate) changes
  an event's outcome with ~1e-5, so of the ~8e5 events a handful of envs may
  end elsewhere; every other env must agree in site, event count and clock."""
  from putting_dune_b200 import engine
  n, seed = 65536, 11
  st = po.make_state(n, seed)
  po.reset(st)
  mlp = po.MlpParams.synthetic(7, hidden=(64, 64))
  w = engine.MlpWeights(**{k: getattr(mlp, k) for k in
                           engine.MlpWeights.NAMES})
  spec = engine.RateSpec(po.RATE_LEARNED, mlp=w, tensor_core=tensor_core)
  b = gh.batch_from_oracle(st)
  rng = np.random.default_rng(0)
  want_elapsed = np.zeros(n, dtype=np.int64)
  got_elapsed = np.zeros(n, dtype=np.int64)
  for _ in range(3):
    ctl = 0.5 + rng.uniform(-1.0, 1.0, size=(n, 1, 2)) * (1.42 / 22.5)
    o = po.step_and_image(st, ctl, 1500000, rate_fn=po.RATE_LEARNED, mlp=mlp)
    out = b.step_and_image(ctl, 1500000, spec)
    want_elapsed += o['elapsed_us']
    got_elapsed += gh.np_(out.elapsed_us)
  same = gh.np_(b.si_idx) == st.si_idx
  n_diff = int((~same).sum())
  print(f'tensor_core {tensor_core}: {n_diff} of {n} envs end elsewhere, '
        f'{int(st.n_events.sum())} events')
  assert n_diff <= 12, n_diff
  np.testing.assert_array_equal(gh.np_(b.n_events)[same], st.n_events[same])
  np.testing.assert_array_equal(got_elapsed[same], want_elapsed[same])
  assert st.n_events.sum() > 5e5 and st.n_transitions.sum() > 1e4
  assert not (gh.np_(b.status) & 1).any()


@pytest.mark.parametrize('tensor_core,hidden', [(2, (64, 64)), (2, (32, 64)),
                                                (1, (128, 128))])
def test_two_ctas_per_sm_form_is_the_same_computation(tensor_core, hidden):
  """pd_set_option("mlp_slim"): 256 threads, two CTAs per SM (default where
  two sets of tiles fit) against 512 threads, one CTA per SM -- identical
  sites, counters, clocks and transition logs."""
  from putting_dune_b200 import _native as nat
  from putting_dune_b200 import engine
  n, seed = 40000, 14
  st = po.make_state(n, seed)
  po.reset(st)
  mlp = po.MlpParams.synthetic(3, hidden=hidden)
  w = engine.MlpWeights(**{k: getattr(mlp, k) for k in
                           engine.MlpWeights.NAMES})
  spec = engine.RateSpec(po.RATE_LEARNED, mlp=w, tensor_core=tensor_core)
  rng = np.random.default_rng(4)
  ctl = [0.5 + rng.uniform(-1.0, 1.0, size=(n, 2, 2)) * (1.42 / 22.5)
         for _ in range(3)]
  dwell = rng.integers(0, 3000000, size=(n, 2))
  dwell[::11, 1] = 0
  res = []
  try:
    for slim in (1, 0):
      assert nat.lib.pd_set_option(b'mlp_slim', slim) == 0
      b = gh.batch_from_oracle(st, log_capacity=64)
      outs = [b.step_and_image(c, dwell, spec) for c in ctl]
      res.append([gh.np_(b.si_idx), gh.np_(b.n_events),
                  gh.np_(b.n_transitions), gh.np_(b.sim_time_us),
                  gh.np_(b.fov), gh.np_(b.status)] +
                 [gh.np_(o.elapsed_us) for o in outs] +
                 [gh.np_(outs[-1].log_count)] +
                 [np.where(np.arange(64)[None, :] <
                           gh.np_(outs[-1].log_count)[:, None], gh.np_(v), -1)
                  for v in (outs[-1].log_site, outs[-1].log_elapsed_us)])
  finally:
    nat.lib.pd_set_option(b'mlp_slim', 1)
  for x, y in zip(*res):
    np.testing.assert_array_equal(x, y)
  assert res[0][2].sum() > 1e5


def test_apply_model_ensemble():
  from putting_dune_b200 import engine
  from putting_dune_b200.rate_learning import learn_rates
  models = [po.MlpParams.synthetic(s, hidden=(64, 64)) for s in (1, 2, 3)]
  w = [engine.MlpWeights(**{k: getattr(m, k) for k in engine.MlpWeights.NAMES})
       for m in models]
  pred = learn_rates.LearnedTransitionRatePredictor(w)
  x = np.random.default_rng(0).uniform(-1.5, 1.5, size=(777, 2)).astype(
      np.float32)
  got = gh.np_(pred.apply_model(x))
  want = po.apply_model(models, x)
  np.testing.assert_allclose(got, want, rtol=2e-5, atol=1e-7)
  # and against the independent float64 network (tests/mlp_independent.py)
  from tests import mlp_independent as mi
  np.testing.assert_allclose(got, mi.apply_model(models, x), rtol=2e-5,
                             atol=1e-7)
  one = gh.np_(pred.apply_model(x, model_index=1))
  np.testing.assert_allclose(one, po.apply_model(models[1:2], x), rtol=2e-5,
                             atol=1e-7)
  with pytest.raises(ValueError):
    pred.rate_spec()


def test_unsupported_shapes_fail_loudly():
  from putting_dune_b200 import _native as nat
  st = po.make_state(8, 1)
  po.reset(st)
  b = gh.batch_from_oracle(st)
  bad = po.MlpParams.synthetic(1, hidden=(64, 40))
  with pytest.raises(nat.NativeError, match='hidden2'):
    b.rates(np.zeros((8, 2)), gh.rate_spec(po.RATE_LEARNED, bad))


@pytest.mark.parametrize('hidden', [(256, 256), (128, 128), (64, 64),
                                    (128, 256), (32, 32)])
def test_tensor_core_rates_close_to_fp32(hidden):
  """pd_mlp.tensor_core: tcgen05 BF16 contraction with FP32 accumulation.
  BF16 rounds the layer-1 activations and W1 to 8 bits of mantissa, so rates
  agree with the FP32 path to ~1e-2 of the largest rate (not a parity path)."""
  from putting_dune_b200 import engine
  n, seed = 3000, 6
  st = po.make_state(n, seed)
  po.reset(st)
  mlp = po.MlpParams.synthetic(4, hidden=hidden)
  rng = np.random.default_rng(1)
  mlp.b1 = rng.normal(0, 0.2, hidden[1]).astype(np.float32)
  mlp.b2 = rng.normal(0, 0.3, 4).astype(np.float32)
  beam = po.site_positions(st, st.si_idx, np.arange(n)) + rng.uniform(
      -2.0, 2.0, size=(n, 2))
  b = gh.batch_from_oracle(st)
  w = engine.MlpWeights(**{k: getattr(mlp, k) for k in
                           engine.MlpWeights.NAMES})
  fp32 = engine.RateSpec(po.RATE_LEARNED, mlp=w)
  tc = engine.RateSpec(po.RATE_LEARNED, mlp=w, tensor_core=True)
  r32, nb32 = b.rates(beam, fp32)
  rtc, nbtc = b.rates(beam, tc)
  np.testing.assert_array_equal(gh.np_(nb32), gh.np_(nbtc))
  r32, rtc = gh.np_(r32), gh.np_(rtc)
  err = np.abs(rtc - r32).max() / np.abs(r32).max()
  assert err <= 2e-2, err
  assert err > 0  # it really is a different arithmetic path
  # repeated calls (mbarrier phase, TMEM re-use) give identical results
  rtc2, _ = b.rates(beam, tc)
  np.testing.assert_array_equal(gh.np_(rtc2), rtc)


def test_tensor_core_step_statistics():
  from putting_dune_b200 import engine
  n, seed = 20000, 8
  mlp = po.MlpParams.synthetic(9, hidden=(128, 128))
  w = engine.MlpWeights(**{k: getattr(mlp, k) for k in
                           engine.MlpWeights.NAMES})
  rng = np.random.default_rng(2)
  ctl = 0.5 + rng.uniform(-0.05, 0.05, size=(n, 1, 2))
  tr = []
  for tensor_core in (False, True):
    b = engine.EnvBatch(n, seed=seed)
    b.reset()
    spec = engine.RateSpec(po.RATE_LEARNED, mlp=w, tensor_core=tensor_core)
    for _ in range(4):
      out = b.step_and_image(ctl, 1500000, spec)
    assert (gh.np_(out.elapsed_us) >= 3500000).all()
    tr.append(float(gh.np_(b.n_transitions).mean()))
  assert tr[0] > 0.5 and abs(tr[1] - tr[0]) <= 0.03 * tr[0], tr


@pytest.mark.parametrize('hidden', [(256, 256), (128, 128), (64, 64),
                                    (32, 32), (64, 128), (128, 256),
                                    (256, 128)])
def test_tensor_core_split_is_a_parity_path(hidden):
  """pd_mlp.tensor_core = 2: tcgen05 with both operands as fp16 hi + fp16 lo
  and three MMAs per K step.  The rates must meet the FP32 path's own bar --
  2e-5 of the largest rate against the float64 oracle -- and agree with the
  FP32 path to 2e-6; a step through it must leave almost every env where the
  FP32 path leaves it.  (H1 = 256 with H2 = 256: the W1 tiles do not fit
  beside the h1 tiles and are streamed, K in two halves per wave.)"""
  from putting_dune_b200 import engine
  n, seed = 3000, 6
  st = po.make_state(n, seed)
  po.reset(st)
  mlp = po.MlpParams.synthetic(4, hidden=hidden)
  rng = np.random.default_rng(1)
  mlp.b1 = rng.normal(0, 0.2, hidden[1]).astype(np.float32)
  mlp.b2 = rng.normal(0, 0.3, 4).astype(np.float32)
  beam = po.site_positions(st, st.si_idx, np.arange(n)) + rng.uniform(
      -2.0, 2.0, size=(n, 2))
  b = gh.batch_from_oracle(st)
  w = engine.MlpWeights(**{k: getattr(mlp, k) for k in
                           engine.MlpWeights.NAMES})
  fp32 = engine.RateSpec(po.RATE_LEARNED, mlp=w)
  split = engine.RateSpec(po.RATE_LEARNED, mlp=w, tensor_core=2)
  r32, nb32 = b.rates(beam, fp32)
  rsp, nbsp = b.rates(beam, split)
  np.testing.assert_array_equal(gh.np_(nb32), gh.np_(nbsp))
  r32, rsp = gh.np_(r32), gh.np_(rsp)
  scale = np.abs(r32).max()
  err = np.abs(rsp - r32).max() / scale
  print(f'hidden {hidden}: split vs FP32 path {err:.2e} of the largest rate')
  assert err <= 2e-6, err
  want, nbr = po.rates_for(st, np.arange(n), beam, po.RATE_LEARNED, mlp)
  np.testing.assert_array_equal(gh.np_(nbsp), nbr)
  assert np.abs(rsp - want).max() <= _tol(want)
  rsp2, _ = b.rates(beam, split)
  np.testing.assert_array_equal(gh.np_(rsp2), rsp)
  # a few steps: trajectories differ from the FP32 path's only where a rate
  # difference of 1e-6 flips a draw
  ctl = 0.5 + rng.uniform(-0.05, 0.05, size=(n, 1, 2))
  sites = []
  for spec in (fp32, split):
    bb = gh.batch_from_oracle(st)
    for _ in range(6):
      bb.step_and_image(ctl, 1500000, spec)
    sites.append(gh.np_(bb.si_idx))
  assert (sites[0] != sites[1]).sum() <= 3
