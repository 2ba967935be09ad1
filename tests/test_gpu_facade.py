"""The reference's own simulator / graphene test scenarios, run against the
drop-in single-env classes (PuttingDuneSimulator, PristineSingleDopedGraphene)
on the GPU.  Each test names the reference test it restates."""

import datetime as dt

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def pd():
  import putting_dune_b200 as pd
  return pd


def _control(pd, x, y, seconds):
  return pd.microscope_utils.BeamControl(pd.geometry.Point(x, y),
                                         dt.timedelta(seconds=seconds))


def _fov(pd, ll, ur):
  return pd.microscope_utils.MicroscopeFieldOfView(pd.geometry.Point(ll),
                                                   pd.geometry.Point(ur))


def test_requires_reset(pd):
  # simulator.py:224-228, graphene.py:702-706
  sim = pd.PuttingDuneSimulator(pd.graphene.PristineSingleDopedGraphene())
  with pytest.raises(RuntimeError, match='Must call reset'):
    sim.step_and_image(np.random.default_rng(0), [_control(pd, 0.5, 0.5, 1)])
  with pytest.raises(RuntimeError, match='Must call reset'):
    pd.graphene.PristineSingleDopedGraphene().get_silicon_position()


@pytest.mark.parametrize('position,ll,ur,want', [
    ((0.3, 0.6), (0.0, 0.0), (1.0, 1.0), (0.3, 0.6)),
    ((0.0, 1.0), (-5.5, -6.3), (12.0, 9.1), (-5.5, 9.1)),
])
def test_control_position_is_converted_to_material_frame(pd, position, ll, ur,
                                                         want):
  # simulator_test.py:86-115
  obs = pd.simulator_observers.EventObserver()
  sim = pd.PuttingDuneSimulator(pd.graphene.PristineSingleDopedGraphene(),
                                observers=(obs,))
  sim.reset(np.random.default_rng(0))
  sim._fov = _fov(pd, ll, ur)
  sim.step_and_image(np.random.default_rng(0),
                     [_control(pd, *position, 1.5)])
  ev = [e for e in obs.events if e.event_type ==
        pd.simulator_observers.SimulatorEventType.APPLY_CONTROL]
  assert len(ev) == 1
  p = ev[0].event_data['position']
  np.testing.assert_allclose((p.x, p.y), want, atol=1e-12)


def test_elapsed_time_with_zero_rates(pd):
  # simulator_test.py:147-168: 1.5 + 3.0 + 7.23 + image 3.5 s
  material = pd.graphene.PristineSingleDopedGraphene(
      rate_function=pd.graphene.PristineSingleSiGrRatePredictor(
          pd.graphene.ConstantRatePredictor((0.0, 0.0, 0.0)).predict))
  sim = pd.PuttingDuneSimulator(material,
                                image_duration=dt.timedelta(seconds=3.5))
  sim.reset(np.random.default_rng(0))
  obs = sim.step_and_image(np.random.default_rng(0), [
      _control(pd, 0.5, 0.5, 1.5), _control(pd, 0.5, 0.5, 3.0),
      _control(pd, 0.5, 0.5, 7.23)])
  assert obs.elapsed_time == dt.timedelta(seconds=1.5 + 3.0 + 7.23 + 3.5)
  assert len(obs.controls) == 3


def test_observer_event_order(pd):
  # simulator_test.py:225-261: RESET, TAKE_IMAGE, APPLY_CONTROL,
  # TRANSITION..., TAKE_IMAGE
  T = pd.simulator_observers.SimulatorEventType
  material = pd.graphene.PristineSingleDopedGraphene(
      rate_function=pd.graphene.PristineSingleSiGrRatePredictor(
          pd.graphene.ConstantRatePredictor((5.0, 5.0, 5.0)).predict))
  obs = pd.simulator_observers.EventObserver()
  sim = pd.PuttingDuneSimulator(material, observers=(obs,))
  sim.reset(np.random.default_rng(0))
  sim.step_and_image(np.random.default_rng(0), [_control(pd, 0.5, 0.5, 1.5)])
  kinds = [e.event_type for e in obs.events]
  assert kinds[:3] == [T.RESET, T.TAKE_IMAGE, T.APPLY_CONTROL]
  n_tr = sum(k == T.TRANSITION for k in kinds)
  assert n_tr >= 2  # rate 15/s for 1.5 s
  assert kinds[3:3 + n_tr] == [T.TRANSITION] * n_tr
  assert kinds[3 + n_tr] == T.TAKE_IMAGE
  times = [e.event_data['time_since_control_was_applied'] for e in obs.events
           if e.event_type == T.TRANSITION]
  assert times == sorted(times) and times[-1] <= dt.timedelta(seconds=1.5)
  # graphene_test.py:199-219: exactly one Si in every successor grid, and
  # positions never change
  grids = [e.event_data['grid'] for e in obs.events
           if e.event_type == T.TRANSITION]
  for g in grids:
    assert (g.atomic_numbers == 14).sum() == 1
    np.testing.assert_array_equal(g.atom_positions, grids[0].atom_positions)


@pytest.mark.parametrize('pct,expect', [
    ((0.2, 0.4), True), ((0.35, 0.95), True), ((0.751, 0.249), True),
    ((0.749, 0.250), False)])
def test_simulator_correctly_updates_fov(pd, pct, expect):
  # simulator_test.py:263-335
  T = pd.simulator_observers.SimulatorEventType
  obs = pd.simulator_observers.EventObserver()
  sim = pd.PuttingDuneSimulator(pd.graphene.PristineSingleDopedGraphene(),
                                observers=(obs,))
  sim.reset(np.random.default_rng(0))
  si = sim.material.get_silicon_position()
  ll = si - 10.0 * np.asarray(pct)
  original = _fov(pd, ll, ll + 10.0)
  sim._fov = original
  o = sim.step_and_image(np.random.default_rng(0), [_control(pd, 1, 1, 0.0)])
  seen = pd.graphene.get_silicon_positions(o.grid).reshape(-1)
  assert seen.shape == (2,)
  np.testing.assert_allclose(seen, (0.5, 0.5) if expect else pct, atol=1e-9)
  images = [e for e in obs.events if e.event_type == T.TAKE_IMAGE]
  assert len(images) == (3 if expect else 2)
  assert images[1].event_data['fov'] == original


def test_seeding_gives_identical_trajectories(pd):
  # simulator_test.py:170-190
  def run():
    sim = pd.PuttingDuneSimulator(pd.graphene.PristineSingleDopedGraphene())
    rng = np.random.default_rng(0)
    out = [sim.reset(rng)]
    for _ in range(5):
      out.append(sim.step_and_image(rng, [_control(pd, 0.52, 0.5, 5.0)]))
    return out
  a, b = run(), run()
  for x, y in zip(a, b):
    np.testing.assert_array_equal(x.grid.atom_positions, y.grid.atom_positions)
    np.testing.assert_array_equal(x.grid.atomic_numbers, y.grid.atomic_numbers)
    assert x.fov == y.fov and x.elapsed_time == y.elapsed_time


def test_material_interface(pd):
  # graphene_test.py:41-87,139-144
  m = pd.graphene.PristineSingleDopedGraphene()
  m.reset(np.random.default_rng(3))
  grid = m.grid
  assert grid.atom_positions.shape == (1881, 2)
  assert (grid.atomic_numbers == 14).sum() == 1
  si = m.get_silicon_position()
  d = np.sort(np.linalg.norm(grid.atom_positions - si, axis=1))
  np.testing.assert_allclose(d[:4], [0.0, 1.42, 1.42, 1.42], atol=1e-7)
  # RateFunction seam on the material's own grid (graphene.py:238-276)
  rf = pd.graphene.PristineSingleSiGrRatePredictor(
      pd.graphene.simple_canonical_rate_function)
  nbr_pos = grid.atom_positions[np.argsort(np.linalg.norm(
      grid.atom_positions - si, axis=1))[1]]
  rates = rf(grid, pd.geometry.Point(nbr_pos))
  assert len(rates.successor_states) == 3
  assert max(s.rate for s in rates.successor_states) == pytest.approx(1.0)
  for s in rates.successor_states:
    assert (s.grid.atomic_numbers == 14).sum() == 1
  succ = np.array([int(np.argmax(s.grid.atomic_numbers == 14))
                   for s in rates.successor_states])
  prior = pd.graphene.HumanPriorRatePredictor().predict(
      grid, pd.geometry.Point(si), si, succ)  # smoke (:301-310)
  assert prior.shape == (3,)
  sub = m.get_atoms_in_bounds(pd.geometry.Point(si - 5.0),
                              pd.geometry.Point(si + 5.0))
  assert sub.atom_positions.min() >= 0 and sub.atom_positions.max() <= 1
  assert (sub.atomic_numbers == 14).sum() == 1


def test_batched_simulator_api(pd):
  sim = pd.BatchedSimulator(
      1000, rate_function=pd.graphene.PristineSingleSiGrRatePredictor(
          pd.graphene.HumanPriorRatePredictor().predict), seed=4)
  with pytest.raises(RuntimeError):
    sim.step_and_image(np.zeros((1000, 1, 2)), dt.timedelta(seconds=1.5))
  sim.reset()
  obs = sim.step_and_image(np.full((1000, 2), 0.5), dt.timedelta(seconds=1.5))
  assert obs.elapsed_us.shape == (1000,) and obs.fov.shape == (1000, 4)
  assert int(obs.elapsed_us.min()) >= 3500000
  obs = sim.step_and_image(np.full((1000, 2), 0.5), 1.5, return_image=False)
  assert obs.image is None
