"""CPU tests of the host-side mirror of the reference's value types."""

import datetime as dt

import numpy as np

from putting_dune_b200 import geometry
from putting_dune_b200 import graphene
from putting_dune_b200 import microscope_utils as mu
from putting_dune_b200 import simulator_observers as so


def _fov(ll, ur):
  return mu.MicroscopeFieldOfView(geometry.Point(ll), geometry.Point(ur))


def test_point_construction_and_value_semantics():
  a, b, c = (geometry.Point(1.0, 2.0), geometry.Point((1.0, 2.0)),
             geometry.Point(np.array([1.0, 2.0])))
  assert a == b == c and hash(a) == hash(c)
  assert np.asarray(a.coords).shape == (1, 2)
  assert (a.x, a.y) == (1.0, 2.0)


def test_fov_transforms_goldens():
  # microscope_utils_test.py:121-288
  fov = _fov((-5.0, 0.0), (5.0, 20.0))
  p = fov.microscope_frame_to_material_frame(geometry.Point(0.5, 1.0))
  assert (p.x, p.y) == (0.0, 20.0)
  arr = fov.microscope_frame_to_material_frame(np.array([-3.0, 2.5]))
  np.testing.assert_allclose(arr, [-35.0, 50.0])
  q = fov.material_frame_to_microscope_frame(geometry.Point(0.0, 20.0))
  assert (q.x, q.y) == (0.5, 1.0)
  ctl = mu.BeamControl(geometry.Point(0.0, 1.0), dt.timedelta(seconds=1.5))
  m = _fov((-5.5, -6.3), (12.0, 9.1)).microscope_frame_to_material_frame(ctl)
  np.testing.assert_allclose((m.position.x, m.position.y), (-5.5, 9.1))
  assert m.dwell_time == ctl.dwell_time
  z = fov.zoom(2.0)
  assert (z.width, z.height) == (5.0, 10.0)
  assert z.offset == fov.offset
  grid = mu.AtomicGrid(np.array([[0.0, 0.0], [6.0, 5.0], [5.5, 5.0]]),
                       np.array([6, 6, 14]))
  assert fov.get_atoms_in_bounds(grid).atomic_numbers.tolist() == [6]
  assert fov.get_atoms_in_bounds(grid, 0.6).atomic_numbers.tolist() == [6, 14]
  g2 = fov.material_frame_to_microscope_frame(grid)
  np.testing.assert_allclose(g2.atom_positions[0], [0.5, 0.0])


def test_timedelta_to_us_exact():
  assert mu.timedelta_to_us(dt.timedelta(seconds=1.5)) == 1500000
  assert mu.timedelta_to_us(7.23) == 7230000
  assert mu.timedelta_to_us(dt.timedelta(hours=1)) == 3600000000


def test_event_observer_interface():
  obs = so.EventObserver()
  fov = _fov((0.0, 0.0), (1.0, 1.0))
  grid = mu.AtomicGrid(np.zeros((1, 2)), np.array([14]))
  obs.observe_take_image(dt.timedelta(seconds=2), fov)
  obs.observe_reset(grid, fov)
  obs.observe_apply_control(mu.BeamControl(geometry.Point(0, 0),
                                           dt.timedelta(seconds=1)))
  obs.observe_transition(dt.timedelta(seconds=0.5), grid)
  kinds = [e.event_type for e in obs.events]
  assert kinds == [so.SimulatorEventType.RESET,
                   so.SimulatorEventType.APPLY_CONTROL,
                   so.SimulatorEventType.TRANSITION]
  assert obs.events[1].event_data['dwell_time'] == dt.timedelta(seconds=1)


def test_silicon_lookup_helpers():
  grid = mu.AtomicGrid(np.array([[0.1, 0.2], [0.5, 0.6]]), np.array([6, 14]))
  np.testing.assert_allclose(graphene.get_single_silicon_position(grid),
                             [0.5, 0.6])
  import pytest
  with pytest.raises(graphene.SiliconNotFoundError):
    graphene.get_single_silicon_position(
        mu.AtomicGrid(np.zeros((2, 2)), np.array([6, 6])))


def test_unknown_rate_function_fails_loudly():
  import pytest
  pred = graphene.PristineSingleSiGrRatePredictor(lambda *a: np.ones(3))
  with pytest.raises(NotImplementedError):
    pred.rate_spec()
  # HumanPriorRatePredictor(mean, cov, max_rate) is supported on the device
  # (graphene.py:181-189): the defaults are recognised, anything else is kept
  p = graphene.HumanPriorRatePredictor(max_rate=1.0)
  assert p.max_rate == 1.0 and not p._is_default()
  assert graphene.HumanPriorRatePredictor()._is_default()
  with pytest.raises(ValueError):
    graphene.HumanPriorRatePredictor(mean=np.zeros(3))


def test_gmm_rate_function_serialisation(tmp_path):
  """GaussianMixtureRateFunction.serialize_to_directory /
  deserialize_from_directory (graphene.py:392-427): round trip, the
  msgpack-numpy array encoding on the wire, and -- when the reference is
  present -- files exchanged with the reference's own class in both
  directions (its `import msgpack_numpy` resolves to the same restated
  codec: the real package is not in the image)."""
  import msgpack
  rng = np.random.default_rng(0)
  fn = graphene.GaussianMixtureRateFunction.sample_new(rng)
  fn.serialize_to_directory(tmp_path / 'a')
  raw = (tmp_path / 'a' / 'gmm_parameters.mpk').read_bytes()
  plain = msgpack.unpackb(raw, raw=False, strict_map_key=False)
  assert plain['sem_ver'] == '1.0.0'
  w = plain['mixture_weights']
  assert w[b'nd'] is True and w[b'type'] == '<f8' and w[b'kind'] == b''
  assert list(w[b'shape']) == list(fn.mixture_weights.shape)
  assert w[b'data'] == fn.mixture_weights.tobytes()
  back = graphene.GaussianMixtureRateFunction.deserialize_from_directory(
      tmp_path / 'a')
  assert back == fn
  np.testing.assert_array_equal(back.variances, fn.variances)
  assert float(back.max_rate) == float(fn.max_rate)
  from oracle import refshim
  if not refshim.reference_available():
    return
  ref = refshim.load_reference().graphene.GaussianMixtureRateFunction
  theirs = ref.deserialize_from_directory(tmp_path / 'a')
  np.testing.assert_array_equal(theirs.loc_distances, fn.loc_distances)
  ref(max_rate=0.5, mixture_weights=np.asarray([0.25, 0.75]),
      loc_distances=np.asarray([0.0, 1.0]),
      variances=np.asarray([[0.1, 0.2], [1.0, 2.0]])).serialize_to_directory(
          tmp_path / 'b')
  ours = graphene.GaussianMixtureRateFunction.deserialize_from_directory(
      tmp_path / 'b')
  assert ours.max_rate == 0.5
  np.testing.assert_array_equal(ours.variances, [[0.1, 0.2], [1.0, 2.0]])
