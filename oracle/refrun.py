"""Runs the UNMODIFIED reference simulator under injected Philox draws.

TEST INFRASTRUCTURE ONLY (build container only: needs /root/reference).
Used by ``tests/golden/make_golden.py`` to produce the committed golden
vectors and by CPU tests (skipped when the reference is absent) to pin
``oracle/pdune_oracle.py``.

Seams used (the reference's own tests patch the same
ones): the ``rng`` argument (duck-typed ``InjectedRng``), the module attribute
``geometry.nearest_neighbors3`` (canonical neighbour order), and a
``SimulatorObserver`` to learn when each ``apply_control`` starts.
"""

from __future__ import annotations

import datetime as dt

import numpy as np

from oracle import pdune_oracle as po
from oracle import refshim


class InjectedRng:
  """Duck-typed ``np.random.Generator`` fed by the oracle's Philox streams."""

  def __init__(self, seed: int, env_id: int):
    self.seed = seed
    self.env = np.uint32(env_id)
    self.begin(po.STREAM_RESET, 0)

  def begin(self, stream: int, seq: int) -> None:
    self.stream = stream
    self.seq = np.uint32(seq)
    self.k = 0  # linear draw counter
    self.slot = -1  # KMC iteration

  def _linear(self, n: int) -> np.ndarray:
    ks = np.arange(self.k, self.k + n)
    self.k += n
    a, b = po.draw_pair(self.seed, self.env, self.seq, ks // 2, self.stream)
    return np.where(ks % 2 == 1, b, a)

  def _shape(self, u, size):
    return float(u[0]) if size is None else u.reshape(size)

  def uniform(self, low=0.0, high=1.0, size=None):
    n = 1 if size is None else int(np.prod(size))
    return self._shape(low + (high - low) * self._linear(n), size)

  def random(self, size=None):
    n = 1 if size is None else int(np.prod(size))
    return self._shape(self._linear(n), size)

  def exponential(self, scale=1.0, size=None):
    if self.stream == po.STREAM_KMC:
      assert size is None
      self.slot += 1
      u, _ = po.draw_pair(self.seed, self.env, self.seq, self.slot,
                          self.stream)
      with np.errstate(invalid='ignore'):
        return float(-np.log1p(-u) * np.float64(scale))
    n = 1 if size is None else int(np.prod(size))
    return self._shape(-np.log1p(-self._linear(n)) * np.float64(scale), size)

  def choice(self, n, p=None):
    if self.stream == po.STREAM_KMC:
      _, u = po.draw_pair(self.seed, self.env, self.seq, self.slot,
                          self.stream)
    else:
      u = self._linear(1)[0]
    if p is None:
      return int(np.floor(u * n))
    cdf = np.cumsum(np.asarray(p, dtype=np.float64))
    cdf = cdf / cdf[-1]
    return int(np.searchsorted(cdf, u, side='right'))


def install_canonical_neighbors(mods, table: np.ndarray):
  """Patches ``geometry.nearest_neighbors3`` to the canonical table order."""
  geometry = mods.geometry
  if getattr(geometry, '_pdune_original_nn3', None) is None:
    geometry._pdune_original_nn3 = geometry.nearest_neighbors3

  def nn3(atom_positions, query, *, include_self=False):
    assert not include_self
    q = np.asarray(query, dtype=np.float64).reshape(-1, 2)
    out_i, out_d = [], []
    for row in q:
      k = int(np.argmin(np.sum((atom_positions - row) ** 2, axis=1)))
      idx = table[k].astype(np.int64)
      out_i.append(idx)
      out_d.append(np.linalg.norm(atom_positions[idx] - row, axis=1))
    out_i, out_d = np.asarray(out_i), np.asarray(out_d)
    if np.ndim(query) == 1:
      out_i, out_d = out_i.reshape(-1), out_d.reshape(-1)
    return geometry.NearestNeighborsResult(out_d, out_i)

  geometry.nearest_neighbors3 = nn3


def uninstall_canonical_neighbors(mods):
  orig = getattr(mods.geometry, '_pdune_original_nn3', None)
  if orig is not None:
    mods.geometry.nearest_neighbors3 = orig


def make_rate_function(mods, rate_fn: int, mlp=None, gmm=None, prior=None):
  g = mods.graphene
  if rate_fn == po.RATE_GMM:
    return g.GaussianMixtureRateFunction(
        max_rate=gmm['max_rate'],
        mixture_weights=np.asarray(gmm['mixture_weights']),
        loc_distances=np.asarray(gmm['loc_distances']),
        variances=np.asarray(gmm['variances']))
  if rate_fn == po.RATE_SIMPLE:
    fn = g.simple_canonical_rate_function
  elif rate_fn == po.RATE_PRIOR:
    fn = (g.HumanPriorRatePredictor().predict if prior is None else
          g.HumanPriorRatePredictor(
              mean=np.asarray(prior['mean'], dtype=np.float64),
              cov=np.asarray(prior['cov'], dtype=np.float64),
              max_rate=float(prior['max_rate'])).predict)
  elif rate_fn == po.RATE_LEARNED:
    packaged = lambda ctx: po.mlp_forward(mlp, np.asarray(ctx, np.float32))
    fn = refshim.reference_learned_predict(packaged)
  else:
    raise ValueError(rate_fn)
  return g.PristineSingleSiGrRatePredictor(canonical_rate_prediction_fn=fn)


def run_reference_env(seed: int, env_id: int, controls: np.ndarray,
                      dwell_us: np.ndarray, rate_fn: int = po.RATE_SIMPLE,
                      mlp=None, image_duration_us: int = 2000000,
                      num_cols: int = 50, table=None, gmm=None) -> dict:
  """One env through ``PuttingDuneSimulator.reset`` + T ``step_and_image``.

  controls: [T, C, 2] microscope frame; dwell_us: [T, C].
  """
  mods = refshim.load_reference()
  if table is None:
    table = po.neighbor_table(num_cols)
  install_canonical_neighbors(mods, table)
  mu = mods.microscope_utils
  rng = InjectedRng(seed, env_id)
  transitions = []

  class Hook(mu.SimulatorObserver):
    ctrl_seq = 0

    def observe_apply_control(self, control):
      rng.begin(po.STREAM_KMC, Hook.ctrl_seq)
      Hook.ctrl_seq += 1

    def observe_transition(self, time_since_control_was_applied, grid):
      transitions.append((
          Hook.ctrl_seq - 1,
          time_since_control_was_applied // dt.timedelta(microseconds=1),
          int(np.argmax(grid.atomic_numbers == 14))))

  material = mods.graphene.PristineSingleDopedGraphene(
      rate_function=make_rate_function(mods, rate_fn, mlp, gmm),
      grid_columns=num_cols)
  sim = mods.simulator.PuttingDuneSimulator(
      material, image_duration=dt.timedelta(microseconds=image_duration_us),
      observers=[Hook()])
  rng.begin(po.STREAM_RESET, 0)
  obs = sim.reset(rng)
  ip = sim._image_parameters  # pylint: disable=protected-access
  out = {
      'positions': material.grid.atom_positions.copy(),
      'si0': int(np.argmax(material.grid.atomic_numbers == 14)),
      'fov0': np.array([obs.fov.lower_left.x, obs.fov.lower_left.y,
                        obs.fov.upper_right.x, obs.fov.upper_right.y]),
      'fov_scale': float(sim._fov_scale),  # pylint: disable=protected-access
      'image_params': np.array([getattr(ip, n)
                                for n in po.IMAGE_PARAM_NAMES]),
      'obs0_positions': obs.grid.atom_positions.copy(),
      'obs0_numbers': obs.grid.atomic_numbers.copy(),
  }
  t_steps = controls.shape[0]
  si = np.zeros(t_steps, dtype=np.int32)
  elapsed = np.zeros(t_steps, dtype=np.int64)
  fov = np.zeros((t_steps, 4))
  n_obs = np.zeros(t_steps, dtype=np.int32)
  with np.errstate(divide='ignore', over='ignore', invalid='ignore'):
    for t in range(t_steps):
      ctrls = [
          mu.BeamControlMicroscopeFrame(mu.BeamControl(
              mods.Point(float(controls[t, c, 0]), float(controls[t, c, 1])),
              dt.timedelta(microseconds=int(dwell_us[t, c]))))
          for c in range(controls.shape[1])
      ]
      obs = sim.step_and_image(rng, ctrls)
      si[t] = int(np.argmax(material.grid.atomic_numbers == 14))
      elapsed[t] = obs.elapsed_time // dt.timedelta(microseconds=1)
      fov[t] = (obs.fov.lower_left.x, obs.fov.lower_left.y,
                obs.fov.upper_right.x, obs.fov.upper_right.y)
      n_obs[t] = obs.grid.atomic_numbers.shape[0]
  out.update(si=si, elapsed_us=elapsed, fov=fov, n_observed=n_obs,
             transitions=np.asarray(transitions, dtype=np.int64).reshape(-1, 3),
             last_obs_positions=obs.grid.atom_positions.copy(),
             last_obs_numbers=obs.grid.atomic_numbers.copy())
  uninstall_canonical_neighbors(mods)
  return out


class _CanonicalKnn:
  """Stand-in for sklearn.neighbors.NearestNeighbors inside the reference's
  feature constructor (feature_constructors.py:200-208): same k-NN set,
  neighbours ordered by (distance rounded to 1e-6, index) instead of float
  noise -- the same canonical order the lattice table uses."""

  def __init__(self, n_neighbors=4, metric='l2', algorithm='brute'):
    self.k = n_neighbors

  def fit(self, x):
    self.x = np.asarray(x)
    return self

  def kneighbors(self, query):
    q = np.asarray(query).reshape(-1, 2)
    dist, idx = [], []
    for row in q:
      d = np.linalg.norm(self.x - row, axis=1)
      order = np.lexsort((np.arange(d.size), np.round(d, 6)))[:self.k]
      idx.append(order)
      dist.append(d[order])
    return np.asarray(dist), np.asarray(idx)


def run_reference_episodes(seed: int, env_ids, rate_fn: int = po.RATE_SIMPLE,
                           dwell_seconds: float = 5.0, step_limit: int = 600,
                           timeout_minutes: float = 10.0) -> dict:
  """`eval_lib.evaluate` of the unmodified reference for the
  greedy_on_neighbor experiment (registry.py:287-298), one episode per env id.

  The env's generator is replaced by InjectedRng(seed, env_id) (the reference
  seeds one generator per episode, putting_dune_environment.py:72-76) and
  `time.perf_counter` is frozen so that agent wall time is zero.
  """
  import time as _time
  mods = refshim.load_reference_env_stack()
  table = po.neighbor_table(50)
  install_canonical_neighbors(mods, table)
  fc = mods.feature_constructors
  orig_nn = fc.neighbors.NearestNeighbors
  fc.neighbors = type('N', (), {'NearestNeighbors': _CanonicalKnn})
  mu = mods.microscope_utils
  material = mods.graphene.PristineSingleDopedGraphene(
      rate_function=make_rate_function(mods, rate_fn))
  inner = mods.putting_dune_environment.PuttingDuneEnvironment(
      material=material,
      action_adapter=mods.action_adapters
      .RelativeToSiliconMaterialFrameActionAdapter(
          dwell_time_range=(dt.timedelta(seconds=dwell_seconds),
                            dt.timedelta(seconds=dwell_seconds)),
          max_distance_angstroms=2 * 1.42),
      feature_constructor=fc.SingleSiliconMaterialFrameFeatureConstructor(),
      goal=mods.goals.SingleSiliconGoalReaching(),
      image_duration=dt.timedelta(seconds=2.0))

  class Hook(mu.SimulatorObserver):
    ctrl_seq = 0

    def observe_reset(self, grid, fov):
      Hook.ctrl_seq = 0

    def observe_apply_control(self, control):
      inner._rng.begin(po.STREAM_KMC, Hook.ctrl_seq)  # pylint: disable=protected-access
      Hook.ctrl_seq += 1

  inner.sim.add_observer(Hook())
  inner.seed = lambda s: setattr(inner, '_rng', InjectedRng(seed, s))
  env = mods.run_helpers.StepLimitWrapper(inner, step_limit=step_limit)
  agent = mods.agent_lib.GreedyAgent(rng=np.random.default_rng(0),
                                     argmax=np.array([1.42, 0.0]))
  suite = mods.eval_lib.EvalSuite(seeds=tuple(int(e) for e in env_ids))
  real_counter = _time.perf_counter
  mods.eval_lib.time.perf_counter = lambda: 0.0
  try:
    with np.errstate(divide='ignore', over='ignore', invalid='ignore'):
      results = mods.eval_lib.evaluate(
          agent, env, suite, timeout=dt.timedelta(minutes=timeout_minutes))
  finally:
    mods.eval_lib.time.perf_counter = real_counter
    fc.neighbors = type('N', (), {'NearestNeighbors': orig_nn})
    uninstall_canonical_neighbors(mods)
  agg = mods.eval_lib.aggregate_results(results)
  return {
      'seed': np.asarray([r.seed for r in results]),
      'reached': np.asarray([r.reached_goal for r in results]),
      'num_actions': np.asarray([r.num_actions_taken for r in results]),
      'env_seconds': np.asarray([r.environment_seconds_to_goal
                                 for r in results]),
      'total_reward': np.asarray([r.total_reward for r in results]),
      'aggregate': np.asarray([agg.average_num_times_reached_goal,
                               agg.average_num_actions_taken,
                               agg.average_environment_seconds_to_goal,
                               agg.average_total_reward]),
  }


def run_reference_env_stack(seed: int, env_id: int, actions: np.ndarray,
                            adapter: int, features: int,
                            rate_fn: int = po.RATE_SIMPLE,
                            min_dwell_s: float = 1.5, max_dwell_s: float = 1.5,
                            max_distance: float = po.BOND,
                            step_limit: int = 600) -> dict:
  """T calls of `env.step(action)` on the unmodified reference
  PuttingDuneEnvironment + StepLimitWrapper (the first call resets, as do the
  calls after a LAST step).  Returns the TimeStep fields per call."""
  from oracle import pdune_oracle_env as oenv
  mods = refshim.load_reference_env_stack()
  table = po.neighbor_table(50)
  install_canonical_neighbors(mods, table)
  fc = mods.feature_constructors
  orig_nn = fc.neighbors.NearestNeighbors
  fc.neighbors = type('N', (), {'NearestNeighbors': _CanonicalKnn})
  mu, aa = mods.microscope_utils, mods.action_adapters
  dwell = (dt.timedelta(seconds=min_dwell_s), dt.timedelta(seconds=max_dwell_s))
  if adapter == oenv.ADAPTER_DIRECT:
    ad = aa.DirectActionAdapter()
  elif adapter == oenv.ADAPTER_DELTA:
    ad = aa.DeltaPositionActionAdapter(np.random.default_rng(0))
  elif adapter == oenv.ADAPTER_RELATIVE:
    ad = aa.RelativeToSiliconActionAdapter(
        dwell_time_range=dwell, max_distance_angstroms=max_distance)
  else:
    ad = aa.RelativeToSiliconMaterialFrameActionAdapter(
        dwell_time_range=dwell, max_distance_angstroms=max_distance)
  feat = (fc.SingleSiliconPristineGrapheneFeatureConstuctor()
          if features == oenv.FEATURES_MICROSCOPE else
          fc.SingleSiliconMaterialFrameFeatureConstructor())
  inner = mods.putting_dune_environment.PuttingDuneEnvironment(
      material=mods.graphene.PristineSingleDopedGraphene(
          rate_function=make_rate_function(mods, rate_fn)),
      action_adapter=ad, feature_constructor=feat,
      goal=mods.goals.SingleSiliconGoalReaching(),
      image_duration=dt.timedelta(seconds=2.0))
  rng = InjectedRng(seed, env_id)
  inner._rng = rng  # pylint: disable=protected-access
  if hasattr(ad, 'rng'):
    ad.rng = rng
  state = {'episode': 0, 'ctrl': 0}

  class Hook(mu.SimulatorObserver):
    def observe_apply_control(self, control):
      rng.begin(po.STREAM_KMC, state['ctrl'])
      state['ctrl'] += 1

  inner.sim.add_observer(Hook())
  sim_reset = inner.sim.reset

  def reset_with_stream(r, **kw):
    r.begin(po.STREAM_RESET, state['episode'])
    state['episode'] += 1
    return sim_reset(r, **kw)

  inner.sim.reset = reset_with_stream
  env = mods.run_helpers.StepLimitWrapper(inner, step_limit=step_limit)
  out = {'step_type': [], 'reward': [], 'discount': [], 'observation': []}
  try:
    with np.errstate(divide='ignore', over='ignore', invalid='ignore'):
      for a in actions:
        ts = env.step(np.asarray(a))
        out['step_type'].append(int(ts.step_type))
        out['reward'].append(0.0 if ts.reward is None else float(ts.reward))
        out['discount'].append(float(ts.discount))
        out['observation'].append(np.asarray(ts.observation, np.float32))
  finally:
    fc.neighbors = type('N', (), {'NearestNeighbors': orig_nn})
    uninstall_canonical_neighbors(mods)
  return {k: np.asarray(v) for k, v in out.items()}
