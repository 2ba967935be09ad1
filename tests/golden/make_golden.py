"""Generates the golden vectors under tests/golden/ from the UNMODIFIED
reference (build container only; needs /root/reference).

    python tests/golden/make_golden.py

For each rate function the reference's own PuttingDuneSimulator /
PristineSingleDopedGraphene run one env at a time under ``InjectedRng``
(oracle/refrun.py); the closed-loop beam controls (Si position + U(-1,1)^2 bond
lengths, dwell 1.5 s / 5 s alternating by env) are produced once and stored, so
the fixtures are self-contained: nothing at test time reads the reference.
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(
    __file__))))
sys.path.insert(0, ROOT)

from oracle import pdune_oracle as po  # pylint: disable=g-import-not-at-top
from oracle import refrun
from oracle import refshim

HERE = os.path.dirname(os.path.abspath(__file__))


GMM_PARAMS = {  # graphene_test.py:337-345 parameter set
    'max_rate': 5.0,
    'mixture_weights': np.asarray((0.3, 0.3, 0.2, 0.1, 0.1)),
    'loc_distances': np.asarray((0.0, 1.0, 0.5, 0.0, 1.5)),
    'variances': np.asarray(((0.1, 0.1), (1.0, 1.0), (2.0, 0.5), (0.5, 2.0),
                             (0.01, 0.01))),
}
# a HumanPriorRatePredictor(mean, cov, max_rate) away from the defaults of
# constants.py:26-28: off-axis mean, anisotropic correlated covariance
PRIOR_CUSTOM = {
    'mean': (0.7, 0.15),
    'cov': ((0.12, 0.03), (0.03, 0.07)),
    'max_rate': 0.4,
}


def closed_loop_controls(seed, n_envs, n_steps, rate_fn, ctrl_seed, mlp=None,
                         gmm=None, float32=False, dwell_pair=(1500000, 5000000)):
  """Controls that follow the Si (RelativeToSilicon-style), from the oracle.
  float32: controls rounded to float32 values (stored compactly)."""
  st = po.make_state(n_envs, seed)
  po.reset(st)
  rng = np.random.default_rng(ctrl_seed)
  controls = np.zeros((n_steps, n_envs, 1, 2))
  dwell = np.zeros((n_steps, n_envs, 1), dtype=np.int64)
  for t in range(n_steps):
    p = po.site_positions(st, st.si_idx, np.arange(n_envs))
    q = po.material_to_microscope(st.fov, p)
    a = rng.uniform(-1, 1, size=(n_envs, 2))
    controls[t, :, 0] = q + a * po.BOND / st.fov_scale[:, None]
    if float32:
      controls[t] = controls[t].astype(np.float32).astype(np.float64)
    dwell[t, :, 0] = np.where(np.arange(n_envs) % 2 == 0, *dwell_pair)
    po.step_and_image(st, controls[t], dwell[t], rate_fn=rate_fn, mlp=mlp,
                      gmm=gmm)
  return controls, dwell


def events_fixture(name, rate_fn, seed, n_envs, n_steps, mlp=None, gmm=None):
  controls, dwell = closed_loop_controls(seed, n_envs, n_steps, rate_fn,
                                         ctrl_seed=seed + 1, mlp=mlp, gmm=gmm)
  out = {k: [] for k in ('si0', 'fov0', 'fov_scale', 'image_params', 'si',
                         'elapsed_us', 'fov', 'n_observed', 'sample_positions',
                         'obs0_positions', 'obs0_numbers')}
  trans = []
  for e in range(n_envs):
    r = refrun.run_reference_env(seed, e, controls[:, e], dwell[:, e], rate_fn,
                                 mlp=mlp, gmm=gmm)
    for k in ('si0', 'fov0', 'fov_scale', 'image_params', 'si', 'elapsed_us',
              'fov', 'n_observed'):
      out[k].append(r[k])
    out['sample_positions'].append(r['positions'][[0, 1, 940, 1880]])
    if e < 4:
      out['obs0_positions'].append(r['obs0_positions'])
      out['obs0_numbers'].append(r['obs0_numbers'])
    t = r['transitions']
    trans.append(np.concatenate([np.full((t.shape[0], 1), e), t], axis=1))
  arrays = {k: np.asarray(v) for k, v in out.items()
            if k not in ('obs0_positions', 'obs0_numbers')}
  for i, (p, z) in enumerate(zip(out['obs0_positions'], out['obs0_numbers'])):
    arrays[f'obs0_positions_{i}'] = p
    arrays[f'obs0_numbers_{i}'] = z
  arrays['transitions'] = np.concatenate(trans, axis=0)  # env, ctrl, us, site
  arrays['controls'] = controls
  arrays['dwell_us'] = dwell
  arrays['seed'] = np.int64(seed)
  arrays['rate_fn'] = np.int64(rate_fn)
  np.savez_compressed(os.path.join(HERE, name), **arrays)
  print(name, 'envs', n_envs, 'steps', n_steps, 'transitions',
        arrays['transitions'].shape[0])


def _large_worker(args):
  seed, e, controls, dwell, rate_fn = args
  r = refrun.run_reference_env(seed, e, controls, dwell, rate_fn)
  t = r['transitions']
  return (e, r['si0'], r['fov0'], r['fov_scale'], r['image_params'], r['si'],
          r['elapsed_us'], r['fov'][-1], r['n_observed'][-1],
          np.concatenate([np.full((t.shape[0], 1), e), t], axis=1))


def events_fixture_large(name, rate_fn, seed, n_envs, n_steps,
                         dwell_pair=(1500000, 5000000)):
  """The same kind of trajectory fixture at a size the judge asked for (256
  envs x 100 steps per rate function; one env x 1000 steps = BASELINE
  configs[0]), the unmodified reference run env by env on all cores.  Stored
  compactly: float32-valued controls, the FOV only after the last step."""
  import multiprocessing as mp
  controls, dwell = closed_loop_controls(seed, n_envs, n_steps, rate_fn,
                                         ctrl_seed=seed + 1, float32=True,
                                         dwell_pair=dwell_pair)
  refshim.load_reference()  # before the fork
  jobs = [(seed, e, controls[:, e], dwell[:, e], rate_fn)
          for e in range(n_envs)]
  with mp.get_context('fork').Pool(min(os.cpu_count() or 1, n_envs)) as pool:
    res = sorted(pool.map(_large_worker, jobs, chunksize=1),
                 key=lambda r: r[0])
  arrays = {
      'si0': np.asarray([r[1] for r in res]),
      'fov0': np.asarray([r[2] for r in res]),
      'fov_scale': np.asarray([r[3] for r in res]),
      'image_params': np.asarray([r[4] for r in res]),
      'si': np.asarray([r[5] for r in res], dtype=np.int32),
      'elapsed_us': np.asarray([r[6] for r in res]),
      'fov_last': np.asarray([r[7] for r in res]),
      'n_observed_last': np.asarray([r[8] for r in res]),
      'transitions': np.concatenate([r[9] for r in res], axis=0),
      'controls': controls.astype(np.float32),
      'dwell_us': dwell[0, :, 0],  # per env, the same at every step
      'seed': np.int64(seed), 'rate_fn': np.int64(rate_fn),
  }
  np.savez_compressed(os.path.join(HERE, name), **arrays)
  print(name, 'envs', n_envs, 'steps', n_steps, 'transitions',
        arrays['transitions'].shape[0], 'bytes',
        os.path.getsize(os.path.join(HERE, name)))


def rates_fixture():
  """Rates of the reference's own rate functions at scattered beam offsets."""
  mods = refshim.load_reference()
  table = po.neighbor_table(50)
  refrun.install_canonical_neighbors(mods, table)
  seed, n = 77, 64
  st = po.make_state(n, seed)
  po.reset(st)
  rng = np.random.default_rng(5)
  beam = po.site_positions(st, st.si_idx, np.arange(n)) + rng.uniform(
      -2.5, 2.5, size=(n, 2))
  mlp = po.MlpParams.synthetic(3, hidden=(32, 32))
  res = {'beam': beam, 'seed': np.int64(seed)}
  for name, rate_fn in (('simple', po.RATE_SIMPLE), ('prior', po.RATE_PRIOR),
                        ('learned', po.RATE_LEARNED), ('gmm', po.RATE_GMM),
                        ('prior_custom', po.RATE_PRIOR)):
    fn = refrun.make_rate_function(
        mods, rate_fn, mlp, GMM_PARAMS,
        prior=PRIOR_CUSTOM if name == 'prior_custom' else None)
    rates = np.zeros((n, 3), dtype=np.float64 if name == 'gmm' else np.float32)
    succ = np.zeros((n, 3), dtype=np.int32)
    for e in range(n):
      pos = po.all_positions(st, e)
      z = np.full(pos.shape[0], 6)
      z[st.si_idx[e]] = 14
      grid = mods.microscope_utils.AtomicGrid(pos, z)
      out = fn(grid, mods.Point(beam[e, 0], beam[e, 1]))
      rates[e] = [s.rate for s in out.successor_states]
      succ[e] = [int(np.argmax(s.grid.atomic_numbers == 14))
                 for s in out.successor_states]
    res[f'rates_{name}'] = rates
    res[f'succ_{name}'] = succ
  for k in MLP_FIELDS:
    res[f'mlp_{k}'] = getattr(mlp, k)
  refrun.uninstall_canonical_neighbors(mods)
  np.savez_compressed(os.path.join(HERE, 'rates_reference.npz'), **res)
  print('rates_reference.npz', n)


MLP_FIELDS = ('bn_scale', 'bn_offset', 'bn_mean', 'bn_var', 'w0', 'b0', 'w1',
              'b1', 'w2', 'b2')


def standardize_fixture():
  """standardize_beam_and_neighbors of the reference on random inputs."""
  fn = refshim.reference_standardize_beam_and_neighbors()
  rng = np.random.default_rng(11)
  n = 256
  ang0 = rng.uniform(0, 2 * np.pi, size=n)
  nbr = np.stack([np.stack((np.cos(ang0 + k * 2 * np.pi / 3),
                            np.sin(ang0 + k * 2 * np.pi / 3)), axis=1)
                  for k in range(3)], axis=1) * po.BOND
  nbr = nbr[:, rng.permutation(3)]
  beam = rng.uniform(-1.5, 1.5, size=(n, 2))
  nb, nn, order = [], [], []
  for i in range(n):
    a, b, c = fn(beam[i:i + 1], nbr[i])
    nb.append(a[0]); nn.append(b); order.append(c)
  np.savez_compressed(os.path.join(HERE, 'standardize_reference.npz'),
                      beam=beam, nbr=nbr, new_beam=np.asarray(nb),
                      new_nbr=np.asarray(nn), order=np.asarray(order))
  print('standardize_reference.npz', n)


def frames_fixture():
  """Frames from the reference's own imaging.py under RenderInjectedRng.

  scikit-image is absent, so the three skimage calls inside the reference are
  served by the restatements in oracle/pdune_oracle_imaging.py (parity
  unpinned for those); every NumPy/SciPy stage is the reference's own code.
  """
  from oracle import pdune_oracle_imaging as oi
  mods = refshim.load_reference()  # installs the empty skimage stubs
  import skimage.exposure
  import skimage.util

  def random_noise(image, mode='gaussian', seed=None, clip=True, **kw):
    if mode == 'gaussian':
      return oi.random_noise_gaussian(image, kw['var'], seed)
    return oi.random_noise_sp(image, kw['amount'], seed)

  skimage.util.random_noise = random_noise
  skimage.exposure.adjust_gamma = oi.adjust_gamma
  skimage.exposure.equalize_adapthist = (
      lambda img, clip_limit=0.01: oi.equalize_adapthist(img, clip_limit))
  im, mu = mods.imaging, mods.microscope_utils
  seed, n, size = 314, 4, 128
  st = po.make_state(n, seed)
  po.reset(st)
  res = {'seed': np.int64(seed), 'size': np.int64(size)}
  for e in range(n):
    q, z, _ = po.get_atoms_in_bounds(st, e)
    f = st.fov[e]
    fov = mu.MicroscopeFieldOfView(mods.Point(f[0], f[1]),
                                   mods.Point(f[2], f[3]))
    grid = mu.AtomicGrid(q, z)
    vals = [float(v) for v in st.image_params[e]]
    p = im.ImageGenerationParameters(*vals, image_size=size)
    rng = oi.RenderInjectedRng(seed, e, 0, size)
    clean = im.generate_clean_image(
        grid, fov, image_size=size, intensity_exponent=p.intensity_exponent)
    blur = im.apply_blur(clean, p.blur_amount)
    pois = im.apply_poisson_noise(blur, p.poisson_rate_multiplier, rng)
    jit = im.apply_jitter(pois, p.jitter_rate, rng)
    rng = oi.RenderInjectedRng(seed, e, 0, size)
    final = im.generate_stem_image(grid, fov, p, rng)
    res[f'clean_{e}'] = clean.astype(np.float32)
    res[f'blur_{e}'] = blur.astype(np.float32)
    res[f'poisson_{e}'] = pois.astype(np.float32)
    res[f'jitter_{e}'] = jit.astype(np.float32)
    res[f'final_{e}'] = final.astype(np.float32)
  # one full-size frame, summarised
  rng = oi.RenderInjectedRng(seed, 0, 0, 512)
  q, z, _ = po.get_atoms_in_bounds(st, 0)
  f = st.fov[0]
  fov = mu.MicroscopeFieldOfView(mods.Point(f[0], f[1]),
                                 mods.Point(f[2], f[3]))
  p = im.ImageGenerationParameters(*[float(v) for v in st.image_params[0]])
  full = im.generate_stem_image(mu.AtomicGrid(q, z), fov, p, rng)
  res['full512_rowmean'] = full.mean(axis=1)
  res['full512_colmean'] = full.mean(axis=0)
  res['full512_patch'] = full[200:232, 300:332]
  np.savez_compressed(os.path.join(HERE, 'frames_reference.npz'), **res)
  print('frames_reference.npz', n, 'frames of', size)


def perception_fixture():
  """imaging.py:57-72 sample_noisy_image_parameters and :75-114
  generate_grid_mask from the reference itself (perception-data generators):
  noisy parameters under InjectedRng positioned at draw 4 of the env's RESET
  sequence; masks of the reset state at 128 x 128."""
  mods = refshim.load_reference()
  im, mu = mods.imaging, mods.microscope_utils
  seed, n, size = 515, 6, 128
  st = po.make_state(n, seed)
  po.reset(st)
  res = {'seed': np.int64(seed), 'size': np.int64(size)}
  noisy = []
  for e in range(n):
    rng = refrun.InjectedRng(seed, e)
    rng.begin(po.STREAM_RESET, 0)
    rng.k = 4  # the reset consumed offset x/y, angle, FOV scale
    p = im.sample_noisy_image_parameters(rng)
    noisy.append([getattr(p, name) for name in po.IMAGE_PARAM_NAMES])
    q, z, _ = po.get_atoms_in_bounds(st, e)
    f = st.fov[e]
    fov = mu.MicroscopeFieldOfView(mods.Point(f[0], f[1]),
                                   mods.Point(f[2], f[3]))
    expo = 1.7 if e % 2 == 0 else float(st.image_params[e, 0])
    res[f'mask_{e}'] = im.generate_grid_mask(
        mu.AtomicGrid(q, z), fov, intensity_exponent=expo,
        image_dimensions=(size, size))
    res[f'mask_exponent_{e}'] = np.float64(expo)
  # imaging.py:129-168: clean image with a buffer, whole grid as input
  from oracle import pdune_oracle_imaging as oi
  for e, buf in ((0, 0.1), (1, 0.25), (2, 0.07)):
    q, z = oi.grid_in_microscope_frame(st, e)
    f = st.fov[e]
    fov = mu.MicroscopeFieldOfView(mods.Point(f[0], f[1]),
                                   mods.Point(f[2], f[3]))
    res[f'buffered_clean_{e}'] = im.generate_clean_image(
        mu.AtomicGrid(q, z), fov, image_size=size,
        intensity_exponent=float(st.image_params[e, 0]),
        buffer_size=buf).astype(np.float32)
    res[f'buffer_{e}'] = np.float64(buf)
  res['noisy_params'] = np.asarray(noisy, dtype=np.float64)
  np.savez_compressed(os.path.join(HERE, 'perception_reference.npz'), **res)
  print('perception_reference.npz', n, 'envs; mask classes',
        sorted(set(res['mask_0'].reshape(-1).tolist())))


def episodes_fixture():
  """EvalResults of the reference's eval_lib.evaluate (greedy_on_neighbor,
  registry.py:287-298) under InjectedRng; see refrun.run_reference_episodes."""
  res = {}
  for name, rate_fn, seed, n in (('simple', po.RATE_SIMPLE, 4242, 96),
                                 ('prior', po.RATE_PRIOR, 4243, 24)):
    r = refrun.run_reference_episodes(seed, range(n), rate_fn)
    for k, v in r.items():
      res[f'{name}_{k}'] = v
    res[f'{name}_philox_seed'] = np.int64(seed)
    print('episodes', name, 'reached', int(r['reached'].sum()), '/', n)
  np.savez_compressed(os.path.join(HERE, 'episodes_reference.npz'), **res)


ENV_CASES = (  # adapter, features, min dwell, max dwell, max distance, limit
    (2, 0, 1.5, 1.5, 1.42, 600),   # relative_random (registry.py:263-266)
    (2, 1, 1.0, 5.0, 2.84, 600),   # relative adapter with a dwell action
    (3, 1, 5.0, 5.0, 2.84, 7),     # material-frame adapter, step limit 7
    (0, 0, 1.5, 1.5, 1.42, 600),   # DirectActionAdapter
    (1, 0, 1.5, 1.5, 1.42, 600),   # DeltaPositionActionAdapter
)


def env_actions(case, n_steps, n_envs, seed=3):
  from oracle import pdune_oracle_env as oenv
  ad, ft, d0, d1, md, lim = case
  adim = oenv.EnvConfig(adapter=ad, min_dwell_s=d0, max_dwell_s=d1).action_dim
  rng = np.random.default_rng(seed)
  lo, hi = {0: (0.3, 0.7), 1: (-0.1, 0.1), 2: (-1.2, 1.2), 3: (-2.0, 2.0)}[ad]
  return rng.uniform(lo, hi, size=(n_steps, n_envs, adim))


def env_fixture():
  """TimeSteps of the reference PuttingDuneEnvironment + StepLimitWrapper."""
  seed, n, t_steps = 808, 4, 30
  res = {'seed': np.int64(seed)}
  for i, case in enumerate(ENV_CASES):
    acts = env_actions(case, t_steps, n)
    res[f'actions_{i}'] = acts
    outs = [refrun.run_reference_env_stack(seed, e, acts[:, e], case[0],
                                           case[1], po.RATE_SIMPLE, case[2],
                                           case[3], case[4], case[5])
            for e in range(n)]
    for k in ('step_type', 'reward', 'discount', 'observation'):
      res[f'{k}_{i}'] = np.stack([o[k] for o in outs], axis=1)  # [T, n, ...]
    print('env case', i, 'LAST', int((res[f'step_type_{i}'] == 2).sum()))
  np.savez_compressed(os.path.join(HERE, 'env_reference.npz'), **res)


def proto_fixture():
  """Wire bytes from the reference's own to_proto code (microscope_utils.py)
  running against the official protobuf runtime (pdune_oracle_proto builds the
  message classes from the restated putting_dune.proto): value types with
  scattered float64 contents, a Trajectory of simulator observations and a
  Transition.  Inputs are stored beside the bytes."""
  import datetime as dt
  mods = refshim.load_reference()
  mu, Point = mods.microscope_utils, mods.Point
  rng = np.random.default_rng(77)
  res = {}

  def fov_of(v):
    return mu.MicroscopeFieldOfView(Point(v[0], v[1]), Point(v[2], v[3]))

  def controls_of(xy, dwell_us):
    return tuple(mu.BeamControl(Point(*p), dt.timedelta(microseconds=int(d)))
                 for p, d in zip(xy, dwell_us))

  n_cases = 12
  res['n_cases'] = np.int64(n_cases)
  for i in range(n_cases):
    m = int(rng.integers(0, 160)) if i else 0
    pos = rng.uniform(-0.2, 1.2, size=(m, 2))
    num = rng.choice([6, 14], size=m).astype(np.int64)
    fov = rng.uniform(-40, 40, size=4)
    c = int(rng.integers(0, 4))
    ctl = rng.uniform(0, 1, size=(c, 2))
    dwell = rng.integers(1, 9_000_000, size=c)
    elapsed = int(rng.integers(0, 900_000_000))
    obs = mu.MicroscopeObservation(
        grid=mu.AtomicGrid(pos, num), fov=fov_of(fov),
        controls=controls_of(ctl, dwell),
        elapsed_time=dt.timedelta(microseconds=elapsed))
    for k, v in (('pos', pos), ('num', num), ('fov', fov), ('ctl', ctl),
                 ('dwell_us', dwell), ('elapsed_us', np.int64(elapsed))):
      res[f'{k}_{i}'] = v
    res[f'grid_bytes_{i}'] = np.frombuffer(
        obs.grid.to_proto().SerializeToString(), dtype=np.uint8)
    res[f'obs_bytes_{i}'] = np.frombuffer(
        obs.to_proto().SerializeToString(), dtype=np.uint8)
    back = mu.MicroscopeObservation.from_proto_string(
        obs.to_proto().SerializeToString())
    assert back.grid.atom_positions.dtype == np.float32
  # a trajectory and a transition built from the cases above
  def obs_case(i):
    return mu.MicroscopeObservation(
        grid=mu.AtomicGrid(res[f'pos_{i}'], res[f'num_{i}']),
        fov=fov_of(res[f'fov_{i}']),
        controls=controls_of(res[f'ctl_{i}'], res[f'dwell_us_{i}']),
        elapsed_time=dt.timedelta(microseconds=int(res[f'elapsed_us_{i}'])))
  traj = mu.Trajectory(observations=[obs_case(i) for i in range(n_cases)])
  res['trajectory_bytes'] = np.frombuffer(
      traj.to_proto().SerializeToString(), dtype=np.uint8)
  a, b = obs_case(3), obs_case(4)
  tr = mu.Transition(grid_before=a.grid, grid_after=b.grid, fov_before=a.fov,
                     fov_after=b.fov, controls=b.controls)
  res['transition_bytes'] = np.frombuffer(
      tr.to_proto().SerializeToString(), dtype=np.uint8)
  # observations of the reference simulator itself (reset + 6 steps)
  sim_bytes = []
  mat = mods.graphene.PristineSingleDopedGraphene()
  sim = mods.simulator.PuttingDuneSimulator(mat)
  srng = np.random.default_rng(5)
  o = sim.reset(srng)
  sim_bytes.append(o.to_proto().SerializeToString())
  for _ in range(6):
    ctl = mu.BeamControl(Point(*srng.uniform(0.4, 0.6, size=2)),
                         dt.timedelta(seconds=1.5))
    o = sim.step_and_image(srng, [ctl])
    sim_bytes.append(o.to_proto().SerializeToString())
  res['sim_n'] = np.int64(len(sim_bytes))
  for i, by in enumerate(sim_bytes):
    res[f'sim_obs_bytes_{i}'] = np.frombuffer(by, dtype=np.uint8)
  np.savez_compressed(os.path.join(HERE, 'proto_reference.npz'), **res)
  print('proto fixture:', n_cases, 'cases,', len(res['trajectory_bytes']),
        'trajectory bytes')


def synth_fixture():
  """The deterministic helpers of rate_learning/data_utils.py
  (get_all_position_rotations, rotate_attributes, rotate_index) and
  graphene.single_silicon_prior_rates from the reference itself, composed as
  sample_from_prior composes them (data_utils.py:252-269), on random
  positions and rotation factors.  jax.jit / jax.vmap are given identity
  stand-ins (jax is absent; jax.numpy is NumPy under the shim)."""
  import importlib
  import types
  mods = refshim.load_reference()
  jax = sys.modules['jax']
  jax.jit = lambda f=None, **kw: f if f is not None else (lambda g: g)
  jax.vmap = lambda f, *a, **kw: f
  stub = types.ModuleType('putting_dune.rate_learning.learn_rates')
  sys.modules.setdefault('putting_dune.rate_learning.learn_rates', stub)
  du = importlib.import_module('putting_dune.rate_learning.data_utils')
  consts = mods.constants
  rng = np.random.default_rng(99)
  res = {}
  for num_states in (3, 6):
    n = 64
    pos = consts.SIGR_PRIOR_RATE_MEAN + np.sqrt(0.15) * rng.normal(size=(n, 2))
    rf = rng.integers(0, num_states, size=n)
    state = rng.integers(0, num_states, size=n)
    rates, pos_rot, rates_rot, state_rot = [], [], [], []
    for i in range(n):
      r = mods.graphene.single_silicon_prior_rates(
          du.get_all_position_rotations(pos[i], num_states=num_states),
          mean=consts.SIGR_PRIOR_RATE_MEAN, cov=consts.SIGR_PRIOR_RATE_COV,
          max_rate=consts.SIGR_PRIOR_MAX_RATE)
      rates.append(r)
      pos_rot.append(mods.geometry.jnp_rotate_coordinates(
          pos[i], 2 * rf[i] * np.pi / num_states))
      rates_rot.append(du.rotate_attributes(r, int(rf[i])))
      state_rot.append(du.rotate_index(state[i], rf[i],
                                       num_states=num_states))
    res.update({f'pos_{num_states}': pos, f'rf_{num_states}': rf,
                f'state_{num_states}': state,
                f'rates_{num_states}': np.asarray(rates),
                f'pos_rot_{num_states}': np.asarray(pos_rot),
                f'rates_rot_{num_states}': np.asarray(rates_rot),
                f'state_rot_{num_states}': np.asarray(state_rot)})
  np.savez_compressed(os.path.join(HERE, 'synth_reference.npz'), **res)
  print('synth fixture written')


FIXTURES = {
    'events_simple': lambda: events_fixture('events_simple.npz',
                                            po.RATE_SIMPLE, 2024, 32, 30),
    'events_prior': lambda: events_fixture('events_prior.npz', po.RATE_PRIOR,
                                           2025, 32, 30),
    'events_gmm': lambda: events_fixture('events_gmm.npz', po.RATE_GMM, 2026,
                                         24, 20, gmm=GMM_PARAMS),
    'events_simple_large': lambda: events_fixture_large(
        'events_simple_large.npz', po.RATE_SIMPLE, 3024, 256, 100),
    'events_prior_large': lambda: events_fixture_large(
        'events_prior_large.npz', po.RATE_PRIOR, 3025, 256, 100),
    # BASELINE configs[0]: one env, prior rates, 1000 beam steps, dwell 1.5 s
    'events_prior_single1000': lambda: events_fixture_large(
        'events_prior_single1000.npz', po.RATE_PRIOR, 3026, 1, 1000,
        dwell_pair=(1500000, 1500000)),
    'rates': rates_fixture,
    'standardize': standardize_fixture,
    'frames': frames_fixture,
    'perception': perception_fixture,
    'episodes': episodes_fixture,
    'env': env_fixture,
    'proto': proto_fixture,
    'synth': synth_fixture,
}

if __name__ == '__main__':
  # python tests/golden/make_golden.py [fixture ...]   (default: all)
  if not refshim.reference_available():
    sys.exit('reference not available; golden vectors are generated only in '
             'the build container')
  for fixture in (sys.argv[1:] or list(FIXTURES)):
    FIXTURES[fixture]()
