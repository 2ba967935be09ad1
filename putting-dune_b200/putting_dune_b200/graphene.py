"""Graphene material and rate functions behind the reference's signatures.

Reference: putting_dune/graphene.py -- ``Material`` (:86-118),
``PristineSingleDopedGraphene`` (:562-706), ``PristineSingleSiGrRatePredictor``
(:232-276), ``simple_canonical_rate_function`` (:133-166),
``HumanPriorRatePredictor`` (:169-229), ``get_silicon_positions`` (:709-746).

The single-material classes here are views onto a batch of one environment in
``engine.EnvBatch``; all geometry, rates and event sampling run in the CUDA
library.  For many environments at once use ``simulator.BatchedSimulator``.
"""

from __future__ import annotations

import abc
import dataclasses
import datetime as dt
from typing import Iterable, Optional, Sequence

import numpy as np

from putting_dune_b200 import _native as nat
from putting_dune_b200 import constants
from putting_dune_b200 import engine
from putting_dune_b200 import geometry
from putting_dune_b200 import microscope_utils as mu


class SiliconNotFoundError(RuntimeError):
  """graphene.py:81."""


@dataclasses.dataclass(frozen=True)
class SuccessorState:
  grid: mu.AtomicGrid
  rate: float


@dataclasses.dataclass(frozen=True)
class Rates:
  successor_states: Sequence[SuccessorState]

  @property
  def total_rate(self) -> float:
    return sum([x.rate for x in self.successor_states])


class PhiloxKey:
  """Stands where the reference takes an ``np.random.Generator``.

  The device draws come from Philox4x32-10 keyed by ``seed`` with the env id in
  the counter, so the "generator" is just this key.  A ``np.random.Generator``
  is also accepted everywhere: ``reset`` derives the key from one
  ``rng.integers`` draw, so equal generator states give equal trajectories.
  """

  def __init__(self, seed: int, env_id: int = 0):
    self.seed = int(seed)
    self.env_id = int(env_id)


def _key_from_rng(rng) -> PhiloxKey:
  if isinstance(rng, PhiloxKey):
    return rng
  if hasattr(rng, 'seed') and hasattr(rng, 'env'):  # oracle InjectedRng
    return PhiloxKey(int(rng.seed), int(rng.env))
  return PhiloxKey(int(rng.integers(0, 2**63 - 1)), 0)


class BoundAtomicGrid(mu.AtomicGrid):
  """The material's grid, remembering which device env it mirrors."""

  def __init__(self, atom_positions, atomic_numbers, material=None):
    super().__init__(atom_positions, atomic_numbers)
    object.__setattr__(self, '_material', material)


# -- canonical rate prediction functions -------------------------------------
def simple_canonical_rate_function(grid, beam_position, silicon_position,
                                   neighbor_indices) -> np.ndarray:
  """graphene.py:133-166; evaluated on the device for the grid's material."""
  return _canonical_rates(engine.RateSpec.simple(), grid, beam_position,
                          neighbor_indices)


simple_canonical_rate_function.rate_spec = engine.RateSpec.simple


class HumanPriorRatePredictor:
  """graphene.py:169-229 with the constants of constants.py:26-28."""

  def __init__(self, mean: np.ndarray = constants.SIGR_PRIOR_RATE_MEAN,
               cov: np.ndarray = constants.SIGR_PRIOR_RATE_COV,
               max_rate: float = constants.SIGR_PRIOR_MAX_RATE):
    self.mean = np.asarray(mean, dtype=np.float64)
    self.cov = np.asarray(cov, dtype=np.float64)
    self.max_rate = float(max_rate)
    if self.mean.shape != (2,) or self.cov.shape != (2, 2):
      raise ValueError('mean must have shape (2,), cov (2, 2)')

  def _is_default(self) -> bool:
    return (np.array_equal(self.mean, constants.SIGR_PRIOR_RATE_MEAN) and
            np.array_equal(self.cov, constants.SIGR_PRIOR_RATE_COV) and
            self.max_rate == constants.SIGR_PRIOR_MAX_RATE)

  def rate_spec(self) -> engine.RateSpec:
    # the defaults take the specialised kernels (closed form, float32 fast
    # path); anything else the general float64 form of the same expression
    if self._is_default():
      return engine.RateSpec.prior()
    return engine.RateSpec.prior(self.mean, self.cov, self.max_rate)

  def predict(self, grid, beam_position, silicon_position,
              neighbor_indices) -> np.ndarray:
    return _canonical_rates(self.rate_spec(), grid, beam_position,
                            neighbor_indices)


class ConstantRatePredictor:
  """Fixed rates: the seam the reference's tests reach with
  ``mock.patch.object(graphene, 'simple_canonical_rate_function',
  return_value=...)`` (simulator_test.py:139-168, graphene_test.py:192-281)."""

  def __init__(self, rates: Sequence[float]):
    self.rates = tuple(float(r) for r in rates)

  def rate_spec(self) -> engine.RateSpec:
    return engine.RateSpec(nat.RATE_CONSTANT, constant=self.rates)

  def predict(self, grid, beam_position, silicon_position, neighbor_indices):
    return np.asarray(self.rates)


def _canonical_rates(spec, grid, beam_position, neighbor_indices):
  material = getattr(grid, '_material', None)
  if material is None:
    raise NotImplementedError(
        'rate functions are evaluated on the device for grids owned by a '
        'PristineSingleDopedGraphene (material.grid)')
  rates, nbr = material._device_rates(beam_position, spec)  # pylint: disable=protected-access
  order = {int(k): i for i, k in enumerate(nbr)}
  return np.asarray([rates[order[int(k)]] for k in neighbor_indices],
                    dtype=np.float64)


def _spec_of(canonical_fn) -> engine.RateSpec:
  """Maps a CanonicalRatePredictionFn to the device rate function."""
  spec = getattr(canonical_fn, 'rate_spec', None)
  if spec is None and hasattr(canonical_fn, '__self__'):
    spec = getattr(canonical_fn.__self__, 'rate_spec', None)
  if spec is None:
    raise NotImplementedError(
        f'{canonical_fn!r} has no device implementation; supported: '
        'simple_canonical_rate_function, HumanPriorRatePredictor().predict, '
        'LearnedTransitionRatePredictor.predict, ConstantRatePredictor')
  return spec()


@dataclasses.dataclass(frozen=True)
class PristineSingleSiGrRatePredictor:
  """graphene.py:232-276."""
  canonical_rate_prediction_fn: object

  def rate_spec(self) -> engine.RateSpec:
    return _spec_of(self.canonical_rate_prediction_fn)

  def __call__(self, grid: mu.AtomicGrid,
               beam_position: geometry.Point) -> Rates:
    material = getattr(grid, '_material', None)
    if material is None:
      raise NotImplementedError('grid must be a material.grid')
    rates, nbr = material._device_rates(beam_position, self.rate_spec())  # pylint: disable=protected-access
    assert (rates >= 0).all(), 'transition_rates were not positive.'
    states = []
    for k, r in zip(nbr, rates):
      numbers = np.full_like(grid.atomic_numbers, constants.CARBON)
      numbers[int(k)] = constants.SILICON
      states.append(SuccessorState(
          mu.AtomicGrid(grid.atom_positions, numbers), np.float32(r)))
    return Rates(states)


@dataclasses.dataclass(frozen=True, eq=False)
class GaussianMixtureRateFunction:
  """graphene.py:279-461: per-neighbour anisotropic Gaussian mixture placed
  along the Si->neighbour vector; a ``RateFunction`` evaluated on the device
  (PD_RATE_GMM)."""
  max_rate: float
  mixture_weights: np.ndarray  # [n_mixtures]
  loc_distances: np.ndarray  # [n_mixtures]
  variances: np.ndarray  # [n_mixtures, 2]

  def rate_spec(self) -> engine.RateSpec:
    return engine.RateSpec(nat.RATE_GMM, gmm={
        'max_rate': self.max_rate, 'mixture_weights': self.mixture_weights,
        'loc_distances': self.loc_distances, 'variances': self.variances})

  def __call__(self, grid: mu.AtomicGrid,
               beam_position: geometry.Point) -> Rates:
    material = getattr(grid, '_material', None)
    if material is None:
      raise NotImplementedError('grid must be a material.grid')
    rates, nbr = material._device_rates(beam_position, self.rate_spec())  # pylint: disable=protected-access
    states = []
    for k, r in zip(nbr, rates):
      numbers = np.full_like(grid.atomic_numbers, constants.CARBON)
      numbers[int(k)] = constants.SILICON
      states.append(SuccessorState(
          mu.AtomicGrid(grid.atom_positions, numbers), float(r)))
    return Rates(states)

  def serialize_to_directory(self, save_dir, /) -> None:
    """graphene.py:392-409: <save_dir>/gmm_parameters.mpk, a msgpack map of
    the four parameters with msgpack-numpy's array encoding."""
    import pathlib
    from putting_dune_b200 import msgpack_numpy_codec as codec
    path = pathlib.Path(save_dir)
    path.mkdir(parents=True, exist_ok=True)
    bundle = {
        'sem_ver': '1.0.0',
        'max_rate': self.max_rate,
        'mixture_weights': self.mixture_weights,
        'loc_distances': self.loc_distances,
        'variances': self.variances,
    }
    (path / 'gmm_parameters.mpk').write_bytes(codec.packb(bundle))

  @classmethod
  def deserialize_from_directory(cls, load_dir, /):
    """graphene.py:411-427."""
    import pathlib
    from putting_dune_b200 import msgpack_numpy_codec as codec
    bundle = codec.unpackb(
        (pathlib.Path(load_dir) / 'gmm_parameters.mpk').read_bytes())
    return cls(max_rate=bundle['max_rate'],
               mixture_weights=bundle['mixture_weights'],
               loc_distances=bundle['loc_distances'],
               variances=bundle['variances'])

  @classmethod
  def sample_new(cls, rng: np.random.Generator):
    """graphene.py:429-444 (domain randomisation; host draws)."""
    num_mixtures = rng.poisson(2.0) + 1
    max_rate = rng.uniform(0.01, 1.0)
    weights = rng.uniform(0.0, 10.0, size=(num_mixtures,))
    weights = weights / np.sum(weights)
    loc = rng.uniform(-2.0, 3.0, size=(num_mixtures,))
    variances = rng.uniform(0.1, 5.0, size=(num_mixtures, 2))
    return cls(max_rate=max_rate, mixture_weights=weights, loc_distances=loc,
               variances=variances)

  def __eq__(self, other) -> bool:
    """graphene.py:446-461: equal up to 1e-3."""
    a, b = self, other
    if (np.shape(a.mixture_weights) != np.shape(b.mixture_weights) or
        np.shape(a.loc_distances) != np.shape(b.loc_distances) or
        np.shape(a.variances) != np.shape(b.variances) or
        abs(a.max_rate - b.max_rate) > 1e-3):
      return False
    return not ((np.abs(np.asarray(a.mixture_weights) -
                        np.asarray(b.mixture_weights)) > 1e-3).any() or
                (np.abs(np.asarray(a.loc_distances) -
                        np.asarray(b.loc_distances)) > 1e-3).any() or
                (np.abs(np.asarray(a.variances) -
                        np.asarray(b.variances)) > 1e-3).any())


class Material(abc.ABC):
  """graphene.py:86-118."""

  @abc.abstractmethod
  def get_atoms_in_bounds(self, lower_left, upper_right) -> mu.AtomicGrid:
    ...

  @abc.abstractmethod
  def reset(self, rng) -> None:
    ...

  @abc.abstractmethod
  def apply_control(self, rng, control: mu.BeamControl,
                    observers: Iterable[mu.SimulatorObserver] = ()) -> None:
    ...


class PristineSingleDopedGraphene(Material):
  """graphene.py:562-706 on a device batch of one env."""

  LOG_CAPACITY = 256

  def __init__(self, *, rate_function=None, grid_columns: int = 50,
               device=None):
    if rate_function is None:
      rate_function = PristineSingleSiGrRatePredictor(
          canonical_rate_prediction_fn=simple_canonical_rate_function)
    self._grid_columns = grid_columns
    self._rate_function = rate_function
    self._device = device
    self._has_been_reset = False
    self._batch: Optional[engine.EnvBatch] = None
    self._lattice: Optional[engine.Lattice] = None

  # -- device plumbing ------------------------------------------------------
  @property
  def batch(self) -> engine.EnvBatch:
    self._assert_has_been_reset('batch')
    return self._batch

  def _rate_spec(self) -> engine.RateSpec:
    spec = getattr(self._rate_function, 'rate_spec', None)
    if spec is None:
      raise NotImplementedError(
          f'{self._rate_function!r} has no device implementation')
    return spec()

  def _device_rates(self, beam_position, spec):
    beam = np.array([[beam_position.x, beam_position.y]])
    r, nb = self._batch.rates(beam, spec)
    return r[0].cpu().numpy(), nb[0].cpu().numpy()

  # -- Material interface ---------------------------------------------------
  def reset(self, rng) -> None:
    key = _key_from_rng(rng)
    if self._lattice is None:
      self._lattice = engine.Lattice(self._grid_columns, self._device)
    keep_episode = (self._batch is not None and self._batch.seed ==
                    (key.seed & 0xFFFFFFFFFFFFFFFF) and
                    self._batch.env_offset == key.env_id)
    if not keep_episode:
      self._batch = engine.EnvBatch(
          1, seed=key.seed, env_offset=key.env_id, lattice=self._lattice,
          device=self._device, log_capacity=self.LOG_CAPACITY)
    self._batch.reset()
    self._has_been_reset = True

  @property
  def grid(self) -> mu.AtomicGrid:
    self._assert_has_been_reset('grid')
    pos = self._batch.grid_positions([0])[0].cpu().numpy()
    return self._grid_with_si(pos, int(self._batch.si_idx[0].item()))

  def _grid_with_si(self, pos: np.ndarray, si: int) -> mu.AtomicGrid:
    numbers = np.full(pos.shape[0], constants.CARBON)
    numbers[si] = constants.SILICON
    return BoundAtomicGrid(pos, numbers, self)

  def get_atoms_in_bounds(self, lower_left: geometry.Point,
                          upper_right: geometry.Point) -> mu.AtomicGrid:
    self._assert_has_been_reset('get_atoms_in_bounds')
    fov = np.array([[lower_left.x, lower_left.y, upper_right.x,
                     upper_right.y]])
    xy, z, count = self._batch.get_atoms_in_bounds(
        fov, max_atoms=self._lattice.n_sites)
    m = int(count[0].item())
    return mu.AtomicGrid(xy[0, :m].cpu().numpy(),
                         z[0, :m].cpu().numpy().astype(np.int64))

  def apply_control(self, rng, control: mu.BeamControl,
                    observers: Iterable[mu.SimulatorObserver] = ()) -> None:
    self._assert_has_been_reset('apply_control')
    beam = np.array([[control.position.x, control.position.y]])
    out = self._batch.apply_control(
        beam, mu.timedelta_to_us(control.dwell_time), self._rate_spec())
    self._raise_on_status()
    self._replay_transitions(out, 0, observers)

  def _replay_transitions(self, out, ctrl_index, observers) -> None:
    observers = list(observers)
    if not observers:
      return
    n = int(out.log_count[0].item())
    if n == 0:
      return
    n = min(n, self.LOG_CAPACITY)
    el = out.log_elapsed_us[0, :n].cpu().numpy()
    site = out.log_site[0, :n].cpu().numpy()
    ctrl = out.log_ctrl[0, :n].cpu().numpy()
    pos = None
    for k in range(n):
      if ctrl[k] != ctrl_index:
        continue
      if pos is None:
        pos = self._batch.grid_positions([0])[0].cpu().numpy()
      grid = self._grid_with_si(pos, int(site[k]))
      for observer in observers:
        observer.observe_transition(
            time_since_control_was_applied=dt.timedelta(
                microseconds=int(el[k])), grid=grid)

  def _raise_on_status(self) -> None:
    status = int(self._batch.status[0].item())
    if status & nat.ENV_BAD_RATE:
      raise AssertionError('transition_rates were not positive.')
    if status & nat.ENV_LOG_OVERFLOW:
      raise RuntimeError(
          f'more than {self.LOG_CAPACITY} transitions in one call')

  def get_silicon_position(self) -> np.ndarray:
    self._assert_has_been_reset('get_silicon_position')
    return self._batch.silicon_position()[0].cpu().numpy()

  def _assert_has_been_reset(self, fn_name: str) -> None:
    if not self._has_been_reset:
      raise RuntimeError(
          f'Must call reset on {self.__class__} before {fn_name}.')


def get_silicon_positions(grid: mu.AtomicGrid) -> np.ndarray:
  """graphene.py:709-710."""
  return grid.atom_positions[grid.atomic_numbers == constants.SILICON]


def get_single_silicon_position(grid: mu.AtomicGrid) -> np.ndarray:
  """graphene.py:713-746."""
  pos = get_silicon_positions(grid)
  n = pos.size // 2
  if n == 0:
    raise SiliconNotFoundError()
  if n > 1:
    d = np.linalg.norm(np.asarray([[0.5, 0.5]]) - pos, axis=1)
    pos = pos[np.argmin(d)]
  return pos.reshape(-1)
