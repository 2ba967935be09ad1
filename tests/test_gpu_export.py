"""GPU parity tests of the device observation encoder (pd_encode_observations)
and the TrajectoryRecorder: the bytes the kernel writes equal the wire bytes
proto_wire.py produces from the oracle's observation of the same state --
and proto_wire.py is pinned to the reference's own to_proto bytes in
tests/test_proto_io.py."""

import ctypes as C
import datetime as dt

import numpy as np
import pytest
import torch

from oracle import pdune_oracle as po
from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
  if not torch.cuda.is_available():
    pytest.skip('needs a CUDA device')
  from putting_dune_b200 import engine
  return engine


def _want_observation(pw, st, e, ctl, dwell_us, elapsed_us, kv=60.0, na=0.1):
  q, z, _ = po.get_atoms_in_bounds(st, e)
  f = st.fov[e]
  return pw.observation(
      pw.atomic_grid(q, z), pw.field_of_view(*f),
      [pw.beam_control(c[0], c[1], dt.timedelta(
          microseconds=int(d)).total_seconds(), kv, na)
       for c, d in zip(ctl, dwell_us)],
      dt.timedelta(microseconds=int(elapsed_us)).total_seconds())


def test_device_encoder_matches_wire_codec(eng):
  from putting_dune_b200 import io as pio
  from putting_dune_b200 import proto_wire as pw
  n, seed = 300, 61
  st = po.make_state(n, seed)
  po.reset(st)
  b = gh.batch_from_oracle(st)
  rng = np.random.default_rng(2)
  spec = gh.rate_spec(po.RATE_SIMPLE)
  rec = pio.TrajectoryRecorder(b)
  # observation after reset: no controls
  rec.record(None, 0, elapsed_us=np.full(n, 2000000))
  by, off, ln, atoms = rec._steps[-1]
  for e in range(n):
    want = _want_observation(pw, st, e, [], [], 2000000)
    assert by[off[e]:off[e] + ln[e]].tobytes() == want, e
  # three steps with 2 ragged controls each; the oracle follows
  for t in range(3):
    ctl = 0.5 + rng.uniform(-0.1, 0.1, size=(n, 2, 2))
    dwell = rng.integers(1, 6000000, size=(n, 2))
    po.step_and_image(st, ctl, dwell, rate_fn=po.RATE_SIMPLE)
    b.step_and_image(ctl, dwell, spec)
    np.testing.assert_array_equal(gh.np_(b.si_idx), st.si_idx)
    rec.record(ctl, dwell, elapsed_us='sim_time')
    by, off, ln, atoms = rec._steps[-1]
    assert (off % 16 == 0).all()
    for e in range(n):
      want = _want_observation(pw, st, e, ctl[e], dwell[e], st.sim_time_us[e])
      assert ln[e] == len(want)
      assert by[off[e]:off[e] + ln[e]].tobytes() == want, (t, e)
  # per-env Trajectory records through the host assembler and back
  stream = rec.to_tfrecord_bytes()
  records = list(pw.tfrecord_iter(stream))
  assert len(records) == n
  from putting_dune_b200 import microscope_utils as mu
  for e in (0, 17, n - 1):
    traj = mu.Trajectory.from_proto_string(records[e])
    assert len(traj.observations) == 4
    assert len(traj.observations[0].controls) == 0
    assert len(traj.observations[3].controls) == 2
    q, z, _ = po.get_atoms_in_bounds(st, e)
    np.testing.assert_array_equal(traj.observations[3].grid.atom_positions,
                                  q.astype(np.float32))
    np.testing.assert_array_equal(traj.observations[3].grid.atomic_numbers, z)
    assert (traj.observations[3].grid.atomic_numbers == 14).sum() == 1


def test_encoder_overflow_and_empty_views(eng):
  from putting_dune_b200 import io as pio
  from putting_dune_b200 import proto_wire as pw
  n = 64
  st = po.make_state(n, 5)
  po.reset(st)
  b = gh.batch_from_oracle(st)
  # env 3 looks at empty space, env 4 at the whole sheet
  b.fov[3] = torch.tensor([500.0, 500.0, 520.0, 520.0], dtype=torch.float64)
  b.fov[4] = torch.tensor([-80.0, -80.0, 80.0, 80.0], dtype=torch.float64)
  st.fov[3] = [500.0, 500.0, 520.0, 520.0]
  st.fov[4] = [-80.0, -80.0, 80.0, 80.0]
  rec = pio.TrajectoryRecorder(b, max_atoms=b.lattice_tables.n_sites)
  rec.record(np.full((n, 1, 2), 0.5), 1500000, elapsed_us='sim_time')
  by, off, ln, atoms = rec._steps[-1]
  assert atoms[3] == 0 and atoms[4] == b.lattice_tables.n_sites
  for e in (3, 4, 5):
    want = _want_observation(pw, st, e, [[0.5, 0.5]], [1500000],
                             st.sim_time_us[e])
    assert by[off[e]:off[e] + ln[e]].tobytes() == want
  small = pio.TrajectoryRecorder(b, max_atoms=64)
  with pytest.raises(RuntimeError, match='max_atoms'):
    small.record(np.full((n, 1, 2), 0.5), 1500000, elapsed_us='sim_time')
  with pytest.raises(ValueError, match='elapsed_us is required'):
    rec.record(np.full((n, 1, 2), 0.5), 1500000)


@pytest.mark.parametrize('num_states,context_dim,time_range',
                         [(3, 2, (0.0, 5.0)), (6, 2, (0.0, 1.0)),
                          (3, 5, (0.5, 2.0)), (3, 0, (0.0, 5.0))])
def test_synthetic_rate_learning_data(eng, num_states, context_dim,
                                      time_range):
  """pd_generate_synthetic_data against the oracle on the same Philox draws
  (float32 tolerance 2e-5 relative; states exact away from decision
  boundaries) and the properties the reference's own test checks
  (data_utils_test.py:62-93: shapes, dt inside the window)."""
  from oracle import pdune_oracle_synth as osy
  from putting_dune_b200.rate_learning import data_utils
  n, seed = 20000, 1234
  train, test = data_utils.generate_synthetic_data(
      num_data=n, data_seed=seed, num_states=num_states,
      context_dim=context_dim, actual_time_range=time_range)
  for split, data in enumerate((train, test)):
    want = osy.generate_synthetic_data(n, seed, split, num_states,
                                       context_dim, time_range)
    got = {k: gh.np_(v) for k, v in data.items()}
    assert got['position'].shape == (n, 2)
    assert got['context'].shape == (n, context_dim)
    assert got['next_state'].shape == (n, 1) and got['dt'].shape == (n, 1)
    assert got['rates'].shape == (n, num_states)
    # float32 sums with cancellation (x c - y s): absolute 5e-5 on O(1) values
    np.testing.assert_allclose(got['position'], want['position'], rtol=2e-5,
                               atol=5e-5)
    np.testing.assert_allclose(got['rates'], want['rates'], rtol=2e-3,
                               atol=1e-9)
    np.testing.assert_allclose(got['context'], want['context'], rtol=2e-5,
                               atol=5e-5)
    np.testing.assert_allclose(got['dt'], want['dt'], rtol=1e-6, atol=1e-7)
    safe = (want['cdf_margin'] > 1e-5) & (want['time_margin'] > 1e-4)
    assert safe.mean() > 0.99
    np.testing.assert_array_equal(got['next_state'][safe],
                                  want['next_state'][safe])
    assert ((got['dt'] >= time_range[0]) & (got['dt'] <= time_range[1])).all()
    assert set(np.unique(got['next_state'])) <= set(range(num_states + 1))
  assert not np.array_equal(gh.np_(train['position']),
                            gh.np_(test['position']))
  # statistics of the generator: position ~ N(rotations of (0.85, 0), 0.15 I)
  p = gh.np_(train['position']).astype(np.float64)
  assert abs(np.hypot(p[:, 0], p[:, 1]).mean() - 0.93) < 0.05


@pytest.mark.parametrize('num_states,context_dim,position_dim,hidden',
                         [(3, 2, 2, (1, 64)), (5, 3, 2, (8, 32)),
                          (3, 0, 3, (4, 100))])
def test_synthetic_rate_learning_data_network(eng, num_states, context_dim,
                                              position_dim, hidden):
  """NETWORK mode of generate_synthetic_data (data_utils.py:196-234) against
  the oracle on the same Philox draws and the same weights (float32 MLP:
  2e-4 relative; states exact away from decision boundaries) and the
  properties data_utils_test.py:62-93 checks (shapes, dt inside the window)."""
  from oracle import pdune_oracle_synth as osy
  from putting_dune_b200.rate_learning import data_utils
  n, seed, time_range = 20000, 77, (0.0, 5.0)
  net = data_utils.init_network(seed, context_dim + position_dim, num_states,
                                hidden)
  net['b0'] += 0.1  # the biases take part
  net['b2'] -= 0.3
  train, test = data_utils.generate_synthetic_data(
      num_data=n, data_seed=seed, num_states=num_states,
      position_dim=position_dim, context_dim=context_dim,
      actual_time_range=time_range, mode='network', network=net)
  for split, data in enumerate((train, test)):
    want = osy.generate_synthetic_data_network(
        n, seed, split, net, num_states, context_dim, position_dim, time_range)
    got = {k: gh.np_(v) for k, v in data.items()}
    assert got['position'].shape == (n, position_dim)
    assert got['context'].shape == (n, context_dim)
    assert got['next_state'].shape == (n, 1) and got['dt'].shape == (n, 1)
    assert got['rates'].shape == (n, num_states)
    np.testing.assert_allclose(got['position'], want['position'], rtol=2e-5,
                               atol=5e-5)
    np.testing.assert_allclose(got['context'], want['context'], rtol=2e-5,
                               atol=5e-5)
    np.testing.assert_allclose(got['rates'], want['rates'], rtol=2e-4,
                               atol=1e-6)
    np.testing.assert_allclose(got['dt'], want['dt'], rtol=1e-6, atol=1e-7)
    assert (got['rates'] > 0).all()
    safe = (want['cdf_margin'] > 1e-4) & (want['time_margin'] > 1e-3)
    assert safe.mean() > 0.98
    np.testing.assert_array_equal(got['next_state'][safe],
                                  want['next_state'][safe])
    assert ((got['dt'] >= time_range[0]) & (got['dt'] <= time_range[1])).all()
    assert set(np.unique(got['next_state'])) <= set(range(num_states + 1))
  # default weights: drawn from the seed, same call twice = same data
  a, _ = data_utils.generate_synthetic_data(num_data=256, data_seed=5,
                                            mode='network')
  b, _ = data_utils.generate_synthetic_data(num_data=256, data_seed=5,
                                            mode='network')
  np.testing.assert_array_equal(gh.np_(a['rates']), gh.np_(b['rates']))
  with pytest.raises(ValueError, match='weights'):
    data_utils.generate_synthetic_data(
        num_data=8, data_seed=5, mode='network', num_states=num_states,
        context_dim=context_dim + 1, position_dim=position_dim, network=net)
