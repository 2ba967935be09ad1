"""CPU oracle: a NumPy restatement of the putting-dune simulator hot path.

TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA path in
``putting-dune_b200/``; it is never imported by the product.  Only ``tests/``,
``tests/golden/make_golden.py``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

Parity status: PINNED for the event path (reset, neighbour geometry, simple and
human-prior rates, ``apply_control``, ``step_and_image``, FOV transforms,
``get_atoms_in_bounds``) against the *unmodified* reference executed in the
build container through ``oracle/refshim.py`` with ``InjectedRng`` (vectors in
``tests/golden/``, generator ``tests/golden/make_golden.py``).  The learned-MLP
forward (Haiku/TF absent, and no reference test covers ``predict``) is
"parity unpinned" for the network itself; the surrounding ``predict`` frame
canonicalisation IS pinned against the reference's own code.

Every function cites the reference lines it restates (paths relative to
``/root/reference/putting_dune``).  All arithmetic follows the reference's
dtype at each step (float64 geometry, float32 rates, integer-microsecond
clock); arrays carry a leading env axis so 4096 envs run in seconds.

Random draws: the reference consumes a ``np.random.Generator``; parity is
defined under *injected* draws (BASELINE.json north_star).  Draws come from
Philox4x32-10 keyed by ``seed`` with counter ``(env, seq, slot, stream)``:

  stream 0  KMC     seq = apply_control calls seen by the env, slot = loop
                    iteration; words 0,1 -> exponential draw, words 2,3 ->
                    categorical draw of the same iteration.
  stream 1  RESET   seq = episode number; draw k lives in slot k//2, half k%2.
  streams 2.. are used by the renderer (see pdune_oracle_imaging.py).

A 53-bit uniform is ``((hi >> 5) * 2**26 + (lo >> 6)) / 2**53``.  Conventions
(SURVEY.md section 7 hard part 2): ``exponential(scale) = -log1p(-u) * scale``,
``choice(n, p) = searchsorted(cumsum(p64) / sum, u, 'right')``,
``uniform(lo, hi) = lo + (hi - lo) * u``.
"""

from __future__ import annotations

import dataclasses
from typing import Optional, Sequence

import numpy as np

CARBON = 6  # constants.py:20
SILICON = 14  # constants.py:21
BOND = 1.42  # constants.py:23 CARBON_BOND_DISTANCE_ANGSTROMS
PRIOR_MEAN = np.array((0.85, 0.0))  # constants.py:26
PRIOR_VAR = 0.1  # constants.py:27 (isotropic covariance diag)
PRIOR_MAX_RATE = np.log(2) / 3  # constants.py:28
GAMMA_PER_SECOND = 0.9967  # constants.py:35

STREAM_KMC = 0
STREAM_RESET = 1
STREAM_RENDER_POISSON = 2
STREAM_RENDER_SP = 3
STREAM_JITTER = 4
STREAM_GOAL = 5
STREAM_AGENT = 6
STREAM_RENDER_UNIFORM = 7
STREAM_RENDER_EXP = 8
STREAM_RENDER_GAUSS = 9

RATE_SIMPLE = 0
RATE_PRIOR = 1
RATE_LEARNED = 2
RATE_CONSTANT = 3  # fixed rates: the seam the reference's tests mock
RATE_GMM = 4  # graphene.py:279-390 GaussianMixtureRateFunction

MAX_TRANSITION_SECONDS = 3600.0  # graphene.py:668

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


# ----------------------------------------------------------------------------
# Philox4x32-10 (Random123); KAT vectors in tests/test_oracle_philox.py
# ----------------------------------------------------------------------------
def philox4x32_10(c0, c1, c2, c3, k0, k1):
  """Vectorised Philox4x32-10. Inputs broadcastable uint32; returns 4 uint32."""
  c0, c1, c2, c3 = np.broadcast_arrays(
      *[np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3)])
  k0 = int(k0) & 0xFFFFFFFF
  k1 = int(k1) & 0xFFFFFFFF
  for _ in range(10):
    p0 = _M0 * c0
    p1 = _M1 * c2
    hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
    hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
    c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0), lo1,
                      hi0 ^ c3 ^ np.uint64(k1), lo0)
    k0 = (k0 + _W0) & 0xFFFFFFFF
    k1 = (k1 + _W1) & 0xFFFFFFFF
  return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def u53(hi, lo):
  """53-bit uniform in [0, 1) from two uint32 words."""
  hi = np.asarray(hi, dtype=np.uint64) >> np.uint64(5)
  lo = np.asarray(lo, dtype=np.uint64) >> np.uint64(6)
  return (hi.astype(np.float64) * 67108864.0 + lo.astype(np.float64)) * (
      1.0 / 9007199254740992.0)


def draw_pair(seed, env, seq, slot, stream):
  """The two 53-bit uniforms of one Philox call (A = words 0,1; B = 2,3)."""
  w = philox4x32_10(env, seq, slot, stream, seed & 0xFFFFFFFF, seed >> 32)
  return u53(w[0], w[1]), u53(w[2], w[3])


def draw_linear(seed, env, seq, stream, k):
  """k-th draw of a linear stream: slot k//2, half k%2."""
  a, b = draw_pair(seed, env, seq, k // 2, stream)
  return b if (k % 2) else a


# ----------------------------------------------------------------------------
# Lattice (a1) and neighbour table (a5)
# ----------------------------------------------------------------------------
def hexagonal_grid(num_cols: int = 50) -> np.ndarray:
  """graphene.py:464-501 `_generate_hexagonal_grid`, restated.

  Rows j in [0, int(num_cols / (sqrt(3)/2))), columns i in [0, num_cols);
  x = i (+0.5 on odd rows), y = j*sqrt(3)/2; sites with i%3==0 on even rows and
  i%3==1 on odd rows are deleted; survivors are listed row-major.
  """
  ratio = np.sqrt(3) / 2
  num_rows = int(num_cols / ratio)
  pts = []
  for j in range(num_rows):
    y = np.float64(j) * ratio
    for i in range(num_cols):
      if (j % 2 == 0 and i % 3 == 0) or (j % 2 == 1 and i % 3 == 1):
        continue
      x = np.float64(i) + (0.5 if j % 2 else 0.0)
      pts.append((x, y))
  return np.asarray(pts, dtype=np.float64)


def lattice_int_coords(num_cols: int = 50) -> np.ndarray:
  """Integer coordinates (2x, j) of every site; 4*d^2 = dX^2 + 3*dj^2 exactly."""
  num_rows = int(num_cols / (np.sqrt(3) / 2))
  pts = []
  for j in range(num_rows):
    for i in range(num_cols):
      if (j % 2 == 0 and i % 3 == 0) or (j % 2 == 1 and i % 3 == 1):
        continue
      pts.append((2 * i + (j % 2), j))
  return np.asarray(pts, dtype=np.int64)


def base_lattice(num_cols: int = 50) -> np.ndarray:
  """graphene.py:537-543: scale by the bond length and centre on the mean.

  The mean over axis 0 is a sequential per-column sum (NumPy reduces the
  outer axis row by row), divided by N.
  """
  pos = hexagonal_grid(num_cols) * BOND
  acc = np.zeros(2, dtype=np.float64)
  for row in pos:  # sequential order == np.add.reduce(axis=0)
    acc = acc + row
  mean = acc / np.float64(pos.shape[0])
  return pos - mean


def neighbor_table(num_cols: int = 50) -> np.ndarray:
  """3 nearest neighbours of every site (geometry.py:93-111 semantics).

  The reference queries a KD-tree for the 4 nearest points and drops self; the
  three bonded neighbours are equidistant up to float noise, so their ORDER is
  arbitrary in the reference (SURVEY.md section 7 hard part 1).  Canonical
  order here: ascending (exact integer squared distance, site index).  Edge
  sites with fewer than three bonded neighbours pick up 2.46 A sites exactly
  as k-NN does; equidistant candidates are resolved by index.
  """
  ic = lattice_int_coords(num_cols)
  n = ic.shape[0]
  table = np.zeros((n, 3), dtype=np.int32)
  for k in range(n):
    d = ic - ic[k]
    d2 = d[:, 0] ** 2 + 3 * d[:, 1] ** 2
    d2[k] = np.iinfo(np.int64).max
    order = np.lexsort((np.arange(n), d2))[:3]
    table[k] = order
  return table


# ----------------------------------------------------------------------------
# State
# ----------------------------------------------------------------------------
@dataclasses.dataclass
class OracleState:
  """Per-env simulator state, struct of arrays."""
  seed: int
  env_ids: np.ndarray  # uint32 [E] global env ids (Philox c0)
  base: np.ndarray  # float64 [N, 2] shared lattice
  nbr: np.ndarray  # int32 [N, 3] shared neighbour table
  lattice: np.ndarray  # float64 [E, 4] off_x, off_y, cos, sin
  si_idx: np.ndarray  # int32 [E]
  fov: np.ndarray  # float64 [E, 4] ll_x, ll_y, ur_x, ur_y
  fov_scale: np.ndarray  # float64 [E]
  image_params: np.ndarray  # float64 [E, 9]
  episode: np.ndarray  # uint32 [E]
  ctrl_count: np.ndarray  # uint32 [E] apply_control calls (Philox seq)
  frame_count: np.ndarray  # uint32 [E] rendered frames (Philox seq)
  n_events: np.ndarray  # int64 [E] rate evaluations (KMC loop iterations)
  n_transitions: np.ndarray  # int64 [E]
  sim_time_us: np.ndarray  # int64 [E] cumulative simulated microseconds

  @property
  def num_envs(self) -> int:
    return self.si_idx.shape[0]


def make_state(num_envs: int, seed: int = 0, num_cols: int = 50,
               env_offset: int = 0) -> OracleState:
  e = num_envs
  return OracleState(
      seed=seed,
      env_ids=(np.arange(e, dtype=np.uint64) + np.uint64(env_offset)).astype(
          np.uint32),
      base=base_lattice(num_cols),
      nbr=neighbor_table(num_cols),
      lattice=np.zeros((e, 4)),
      si_idx=np.zeros(e, dtype=np.int32),
      fov=np.zeros((e, 4)),
      fov_scale=np.zeros(e),
      image_params=np.zeros((e, 9)),
      episode=np.zeros(e, dtype=np.uint32),
      ctrl_count=np.zeros(e, dtype=np.uint32),
      frame_count=np.zeros(e, dtype=np.uint32),
      n_events=np.zeros(e, dtype=np.int64),
      n_transitions=np.zeros(e, dtype=np.int64),
      sim_time_us=np.zeros(e, dtype=np.int64),
  )


def site_positions(state: OracleState, sites: np.ndarray,
                   envs: Optional[np.ndarray] = None) -> np.ndarray:
  """Material-frame position of lattice ``sites`` [E] or [E, K] per env.

  graphene.py:545-557: ``(base + offset) @ [[c, -s], [s, c]]`` i.e.
  ``x' = x*c + y*s``, ``y' = y*c - x*s`` (separately rounded products).
  """
  lat = state.lattice if envs is None else state.lattice[envs]
  sites = np.asarray(sites)
  shp = (-1,) + (1,) * (sites.ndim - 1)
  ox, oy = lat[:, 0].reshape(shp), lat[:, 1].reshape(shp)
  c, s = lat[:, 2].reshape(shp), lat[:, 3].reshape(shp)
  bx = state.base[sites, 0] + ox
  by = state.base[sites, 1] + oy
  return np.stack((bx * c + by * s, by * c - bx * s), axis=-1)


def all_positions(state: OracleState, env: int) -> np.ndarray:
  n = state.base.shape[0]
  return site_positions(state, np.arange(n)[None, :], np.array([env]))[0]


# ----------------------------------------------------------------------------
# reset (a2, a3, a16 reset half, a17)
# ----------------------------------------------------------------------------
IMAGE_PARAM_NAMES = (
    'intensity_exponent', 'gaussian_variance', 'jitter_rate',
    'poisson_rate_multiplier', 'salt_and_pepper_amount', 'blur_amount',
    'contrast_gamma', 'exponential_lambda', 'uniform_noise_scale')


def reset(state: OracleState, mask: Optional[np.ndarray] = None) -> None:
  """simulator.py:65-105 + graphene.py:533-559,584-598 + imaging.py:42-54.

  Draw order (RESET stream, linear): offset x, offset y, angle, fov scale,
  then the nine image parameters in dataclass order.
  """
  e = state.num_envs
  mask = np.ones(e, dtype=bool) if mask is None else np.asarray(mask, bool)
  idx = np.nonzero(mask)[0]
  if idx.size == 0:
    return
  env, ep = state.env_ids[idx], state.episode[idx]
  d = lambda k: draw_linear(state.seed, env, ep, STREAM_RESET, k)
  half = BOND / 2
  # rng.uniform(-0.71, 0.71, size=(1, 2)): low + (high - low) * u
  off_x = -half + (half - (-half)) * d(0)
  off_y = -half + (half - (-half)) * d(1)
  angle = 0.0 + (2 * np.pi - 0.0) * d(2)
  state.lattice[idx] = np.stack(
      (off_x, off_y, np.cos(angle), np.sin(angle)), axis=1)
  # Si = site nearest the origin (graphene.py:592-594), first index on ties.
  n = state.base.shape[0]
  for j, e_i in enumerate(idx):
    p = site_positions(state, np.arange(n)[None, :], np.array([e_i]))[0]
    dist = np.sqrt(p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1])
    state.si_idx[e_i] = np.argmin(dist)
  scale = 15.0 + (30.0 - 15.0) * d(3)  # simulator.py:77
  state.fov_scale[idx] = scale
  recenter_fov(state, idx)
  # imaging.py:42-54
  ip = np.empty((idx.size, 9))
  ip[:, 0] = 1.4 + (2.0 - 1.4) * d(4)
  ip[:, 1] = 0.0 + (5e-3 - 0.0) * d(5)
  ip[:, 2] = 0.0 + (5.0 - 0.0) * d(6)
  ip[:, 3] = -np.log1p(-d(7)) * 15.0 + 1.0
  ip[:, 4] = 0.0 + (1e-3 - 0.0) * d(8)
  ip[:, 5] = 0.0 + (1.0 - 0.0) * d(9)
  ip[:, 6] = 0.7 + (1.3 - 0.7) * d(10)
  ip[:, 7] = 0.0 + (0.2 - 0.0) * d(11)
  ip[:, 8] = 0.0 + (0.2 - 0.0) * d(12)
  state.image_params[idx] = ip
  state.episode[idx] += np.uint32(1)


def sample_noisy_image_parameters(state: OracleState,
                                  mask: Optional[np.ndarray] = None) -> None:
  """imaging.py:57-72 `sample_noisy_image_parameters`, drawn from the same
  nine RESET-stream uniforms (draws 4..12 of the env's current episode) the
  reset used for `sample_image_parameters`."""
  e = state.num_envs
  mask = np.ones(e, dtype=bool) if mask is None else np.asarray(mask, bool)
  idx = np.nonzero(mask)[0]
  if idx.size == 0:
    return
  env, ep = state.env_ids[idx], state.episode[idx] - np.uint32(1)
  d = lambda k: draw_linear(state.seed, env, ep, STREAM_RESET, k)
  ip = np.empty((idx.size, 9))
  ip[:, 0] = 1.4 + (2.0 - 1.4) * d(4)
  ip[:, 1] = 0.0 + (0.3 - 0.0) * d(5)
  ip[:, 2] = 0.0 + (5.0 - 0.0) * d(6)
  ip[:, 3] = -np.log1p(-d(7)) * 15.0 + 1.0
  ip[:, 4] = 0.0 + (1e-2 - 0.0) * d(8)
  ip[:, 5] = 0.0 + (0.25 - 0.0) * d(9)
  ip[:, 6] = 0.5 + (1.5 - 0.5) * d(10)
  ip[:, 7] = 0.0 + (0.25 - 0.0) * d(11)
  ip[:, 8] = 0.0 + (0.25 - 0.0) * d(12)
  state.image_params[idx] = ip


def recenter_fov(state: OracleState, idx: np.ndarray) -> None:
  """simulator.py:79-82,161-165: FOV = [P_si - s/2, P_si + s/2]."""
  p = site_positions(state, state.si_idx[idx], idx)
  h = state.fov_scale[idx] / 2.0
  state.fov[idx, 0] = p[:, 0] - h
  state.fov[idx, 1] = p[:, 1] - h
  state.fov[idx, 2] = p[:, 0] + h
  state.fov[idx, 3] = p[:, 1] + h


# ----------------------------------------------------------------------------
# Rate functions (a7, a8, a9, a10)
# ----------------------------------------------------------------------------
def simple_rates(beam: np.ndarray, p_si: np.ndarray,
                 p_nbr: np.ndarray) -> np.ndarray:
  """graphene.py:133-166 `simple_canonical_rate_function` -> float64 [E, 3]."""
  nbr_rel = p_nbr - p_si[:, None, :]
  beam_rel = beam - p_si
  diff = beam_rel[:, None, :] - nbr_rel
  # np.linalg.norm(axis=-1) == sqrt(add.reduce(x*x))
  dist = np.sqrt(diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1])
  dist = dist / BOND
  return 1.0 / (np.square(dist * 4) + 1.0)


def prior_rates(beam: np.ndarray, p_si: np.ndarray, p_nbr: np.ndarray,
                prior: Optional[dict] = None) -> np.ndarray:
  """graphene.py:191-229 `HumanPriorRatePredictor.predict` + :121-130.

  theta_i = atan2 of neighbour i; the mean (0.85, 0) is rotated by
  ``rotate_coordinates(mean, -theta)`` with the matrix [[c, s], [-s, c]] of
  angle -theta (geometry.py:51-66), i.e. mu = 0.85 * (cos theta, -sin theta):
  the peak is mirrored to -theta.  With an
  isotropic covariance 0.1*I, pdf(x)/pdf(mu) = exp(-|x-mu|^2 / (2*0.1)).
  """
  nbr_rel = p_nbr - p_si[:, None, :]
  theta = np.arctan2(nbr_rel[..., 1], nbr_rel[..., 0])
  ang = -theta
  c, s = np.cos(ang), np.sin(ang)
  mean = PRIOR_MEAN if prior is None else np.asarray(prior['mean'], float)
  # mean @ [[c, s], [-s, c]] (mean = (0.85, 0) by default)
  mu_x = mean[0] * c + mean[1] * (-s)
  mu_y = mean[0] * s + mean[1] * c
  x = (beam - p_si) / BOND
  dx = x[:, None, 0] - mu_x
  dy = x[:, None, 1] - mu_y
  if prior is None:
    maha = (dx * dx + dy * dy) / PRIOR_VAR
    return PRIOR_MAX_RATE * np.exp(-0.5 * maha)
  # HumanPriorRatePredictor(mean, cov, max_rate), graphene.py:181-229: only
  # the mean is rotated towards the neighbour, the covariance stays in the
  # material frame; max_rate * pdf(x) / pdf(mean) = max_rate * exp(-maha / 2)
  prec = np.linalg.inv(np.asarray(prior['cov'], float).reshape(2, 2))
  maha = (prec[0, 0] * dx * dx + (prec[0, 1] + prec[1, 0]) * dx * dy +
          prec[1, 1] * dy * dy)
  return float(prior['max_rate']) * np.exp(-0.5 * maha)


@dataclasses.dataclass
class MlpParams:
  """Haiku parameter tree of `get_mlp_fn` (learn_rates.py:80-99), eval mode."""
  bn_scale: np.ndarray  # [D]
  bn_offset: np.ndarray  # [D]
  bn_mean: np.ndarray  # [D]
  bn_var: np.ndarray  # [D]
  w0: np.ndarray  # [D, H1]
  b0: np.ndarray  # [H1]
  w1: np.ndarray  # [H1, H2]
  b1: np.ndarray  # [H2]
  w2: np.ndarray  # [H2, 4]
  b2: np.ndarray  # [4]
  batchnorm: bool = True

  @staticmethod
  def synthetic(seed: int, hidden=(128, 128), context_dim: int = 2):
    """Seeded weights: TruncatedNormal(stddev=1/sqrt(fan_in)), zero bias, the
    Haiku `hk.Linear` default; BN scale 1 / offset 0 / mean 0 / var 1."""
    rng = np.random.default_rng(seed)

    def trunc(shape):
      std = 1.0 / np.sqrt(shape[0])
      w = rng.standard_normal(shape)
      bad = np.abs(w) > 2
      while bad.any():
        w[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(w) > 2
      return (w * std).astype(np.float32)

    d, (h1, h2) = context_dim, hidden
    f32 = np.float32
    return MlpParams(
        bn_scale=np.ones(d, f32), bn_offset=np.zeros(d, f32),
        bn_mean=np.zeros(d, f32), bn_var=np.ones(d, f32),
        w0=trunc((d, h1)), b0=np.zeros(h1, f32),
        w1=trunc((h1, h2)), b1=np.zeros(h2, f32),
        w2=trunc((h2, 4)), b2=np.zeros(4, f32))


def mlp_forward(params: MlpParams, x: np.ndarray) -> np.ndarray:
  """learn_rates.py:80-99 `call_mlp(is_training=False)`; float32 [B, 4].

  hk.BatchNorm eval: (x - mean) * rsqrt(var + 1e-5) * scale + offset;
  hk.nets.MLP with swish between layers (not after the last); softplus.
  """
  f32 = np.float32
  x = np.asarray(x, dtype=f32)
  if params.batchnorm:
    inv = f32(1.0) / np.sqrt(params.bn_var + f32(1e-5))
    x = (x - params.bn_mean) * (inv * params.bn_scale) + params.bn_offset
  swish = lambda z: z / (f32(1.0) + np.exp(-z))
  h = swish(x @ params.w0 + params.b0)
  h = swish(h @ params.w1 + params.b1)
  o = h @ params.w2 + params.b2
  # softplus(z) = logaddexp(z, 0)
  return np.logaddexp(o, f32(0.0)).astype(f32)


def get_angles(xy: np.ndarray) -> np.ndarray:
  """geometry.py:33-48."""
  return np.arctan2(xy[..., 1], xy[..., 0])


def rotate_coordinates(xy: np.ndarray, theta: np.ndarray) -> np.ndarray:
  """geometry.py:51-66: right-multiply by [[c, s], [-s, c]]."""
  c, s = np.cos(theta), np.sin(theta)
  return np.stack((xy[..., 0] * c - xy[..., 1] * s,
                   xy[..., 0] * s + xy[..., 1] * c), axis=-1)


def standardize_beam_and_neighbors(beam: np.ndarray, nbr: np.ndarray):
  """rate_learning/data_utils.py:389-432, batched: beam [E,2], nbr [E,3,2]."""
  d = nbr - beam[:, None, :]
  dist = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])
  k = np.argmin(dist, axis=1)
  ang = get_angles(nbr)
  rot = -np.take_along_axis(ang, k[:, None], axis=1)[:, 0]
  new_nbr = rotate_coordinates(nbr, rot[:, None])
  new_beam = rotate_coordinates(beam, rot)
  positive = (ang + rot[:, None]) % (2 * np.pi)
  order = np.argsort(positive, axis=1, kind='stable')
  return new_beam, new_nbr, order


def learned_rates(params: MlpParams, beam: np.ndarray, p_si: np.ndarray,
                  p_nbr: np.ndarray) -> np.ndarray:
  """learn_rates.py:925-972 `predict` (context_dim=2 only; appendix B quirk 2).

  Neighbours stay in angstroms while the beam is divided by the bond length
  before the nearest-neighbour selection; the first three raw softplus heads
  are returned, un-permuted by argsort(order).
  """
  nbr_rel = p_nbr - p_si[:, None, :]
  b = (beam - p_si) / BOND
  new_beam, _, order = standardize_beam_and_neighbors(b, nbr_rel)
  out = mlp_forward(params, new_beam.astype(np.float32))
  inv = np.argsort(order, axis=1, kind='stable')
  return np.take_along_axis(out[:, :3], inv, axis=1).astype(np.float64)


def apply_model(params_list: Sequence[MlpParams], x: np.ndarray) -> np.ndarray:
  """learn_rates.py:704-732: softmax(out[:3]) * out[3], mean over ensemble."""
  acc = None
  for p in params_list:
    o = mlp_forward(p, x)
    z = o[:, :3] - o[:, :3].max(axis=1, keepdims=True)
    w = np.exp(z)
    w = w / w.sum(axis=1, keepdims=True)
    r = o[:, 3:4] * w
    acc = r if acc is None else acc + r
  return (acc / np.float32(len(params_list))).astype(np.float32)


def gmm_rates(beam: np.ndarray, p_si: np.ndarray, p_nbr: np.ndarray,
              gmm: dict) -> np.ndarray:
  """graphene.py:303-388 `GaussianMixtureRateFunction.__call__`, float64 [E, 3].

  The covariance `E diag(v) E^-1` (:352-361) has the unit Si->neighbour
  vector and its normal as eigenvectors, so the density is
  exp(-0.5 (d1^2/v1 + d2^2/v2)) / (2 pi sqrt(v1 v2)); the normalising factor
  (:289-301) makes the largest mixture mode equal `max_rate`."""
  w = np.asarray(gmm['mixture_weights'], dtype=np.float64)
  loc = np.asarray(gmm['loc_distances'], dtype=np.float64)
  var = np.asarray(gmm['variances'], dtype=np.float64).reshape(-1, 2)
  mode = w / (2 * np.pi * np.sqrt(var[:, 0] * var[:, 1]))
  norm = gmm['max_rate'] / mode.max()
  delta = p_nbr - p_si[:, None, :]
  e1 = delta / np.sqrt((delta ** 2).sum(axis=-1, keepdims=True))
  rates = np.zeros(delta.shape[:2])
  for m in range(w.size):
    mean = p_si[:, None, :] + delta * loc[m]
    d = beam[:, None, :] - mean
    d1 = d[..., 0] * e1[..., 0] + d[..., 1] * e1[..., 1]
    d2 = d[..., 1] * e1[..., 0] - d[..., 0] * e1[..., 1]
    pdf = np.exp(-0.5 * (d1 * d1 / var[m, 0] + d2 * d2 / var[m, 1])) / (
        2 * np.pi * np.sqrt(var[m, 0] * var[m, 1]))
    rates += pdf * norm * w[m]
  return rates


def rates_for(state: OracleState, envs: np.ndarray, beam: np.ndarray,
              rate_fn: int, mlp: Optional[MlpParams] = None, constant=None,
              gmm: Optional[dict] = None, keep64: bool = False,
              prior: Optional[dict] = None):
  """graphene.py:238-259: Si + 3-NN geometry -> canonical fn -> float32[3]."""
  si = state.si_idx[envs]
  nbr = state.nbr[si]
  p_si = site_positions(state, si, envs)
  p_nbr = site_positions(state, nbr, envs)
  if rate_fn == RATE_SIMPLE:
    r64 = simple_rates(beam, p_si, p_nbr)
  elif rate_fn == RATE_PRIOR:
    r64 = prior_rates(beam, p_si, p_nbr, prior)
  elif rate_fn == RATE_LEARNED:
    r64 = learned_rates(mlp, beam, p_si, p_nbr)
  elif rate_fn == RATE_CONSTANT:
    r64 = np.tile(np.asarray(constant, dtype=np.float64), (len(envs), 1))
  elif rate_fn == RATE_GMM:
    r64 = gmm_rates(beam, p_si, p_nbr, gmm)
  else:
    raise ValueError(rate_fn)
  if keep64:
    return r64, nbr
  r32 = r64.astype(np.float32)
  assert (r32 >= 0).all(), 'transition_rates were not positive.'
  return r32, nbr


# ----------------------------------------------------------------------------
# apply_control (a11, a12)
# ----------------------------------------------------------------------------
def seconds_to_us(t: np.ndarray) -> np.ndarray:
  """`dt.timedelta(seconds=t)` (graphene.py:669): whole seconds * 10^6 plus
  the fractional part * 10^6 rounded half-to-even (CPython delta_new)."""
  frac, whole = np.modf(t)
  return whole.astype(np.int64) * 1000000 + np.rint(frac * 1e6).astype(np.int64)


@dataclasses.dataclass
class EventLog:
  """Per-transition record (the `observe_transition` payload, a27)."""
  env: list
  ctrl_seq: list
  iteration: list
  elapsed_us: list
  slot: list
  new_si: list


def apply_control(state: OracleState, beam: np.ndarray, dwell_us: np.ndarray,
                  rate_fn: int = RATE_SIMPLE, mlp: Optional[MlpParams] = None,
                  log: Optional[EventLog] = None,
                  rates_override=None, gmm: Optional[dict] = None,
                  skip: Optional[np.ndarray] = None,
                  prior: Optional[dict] = None) -> dict:
  """graphene.py:646-694 for all envs at once.

  beam: float64 [E, 2] material frame; dwell_us: int64 [E].
  Returns per-call counters (events, transitions, last rates).
  """
  e = state.num_envs
  dwell_us = np.broadcast_to(np.asarray(dwell_us, dtype=np.int64), (e,))
  elapsed = np.zeros(e, dtype=np.int64)
  ev = np.zeros(e, dtype=np.int64)
  tr = np.zeros(e, dtype=np.int64)
  first_rates = np.zeros((e, 3), dtype=np.float32)
  it = 0
  live = np.ones(e, dtype=bool) if skip is None else ~np.asarray(skip, bool)
  active = (elapsed < dwell_us) & live  # graphene.py:658
  while active.any():
    idx = np.nonzero(active)[0]
    if rates_override is not None:
      r32, nbr = rates_override(state, idx, beam[idx], it)
    elif rate_fn == RATE_GMM:
      # GaussianMixtureRateFunction returns float64 rates (graphene.py:375):
      # the total and the exponential scale stay float64 (Python floats);
      # only the branch probabilities go through float32 (:679-683).
      r64, nbr = rates_for(state, idx, beam[idx], rate_fn, gmm=gmm,
                           keep64=True)
      r32 = r64.astype(np.float32)
    else:
      r32, nbr = rates_for(state, idx, beam[idx], rate_fn, mlp, prior=prior)
    if it == 0:
      first_rates[idx] = r32
    # Rates.total_rate: Python sum of np.float32 -> sequential float32 adds.
    tot = (r32[:, 0] + r32[:, 1]) + r32[:, 2]
    if rate_fn == RATE_GMM and rates_override is None:
      tot = (r64[:, 0] + r64[:, 1]) + r64[:, 2]
    with np.errstate(divide='ignore', over='ignore', invalid='ignore'):
      scale = tot.dtype.type(1.0) / tot  # float32 under NumPy 2 (NEP 50)
      u1, u2 = draw_pair(state.seed, state.env_ids[idx], state.ctrl_count[idx],
                         it, STREAM_KMC)
      t = -np.log1p(-u1) * scale.astype(np.float64)
      t = np.where(tot == 0, MAX_TRANSITION_SECONDS, t)
      t = np.minimum(t, MAX_TRANSITION_SECONDS)  # graphene.py:668
    elapsed[idx] += seconds_to_us(t)
    ev[idx] += 1
    hit = elapsed[idx] <= dwell_us[idx]  # graphene.py:677
    if hit.any():
      h = np.nonzero(hit)[0]
      with np.errstate(divide='ignore', invalid='ignore'):
        p32 = r32[h] / tot[h, None].astype(np.float32)  # float32
      cdf = np.cumsum(p32.astype(np.float64), axis=1)
      cdf = cdf / cdf[:, -1:]
      slot = (cdf <= u2[h, None]).sum(axis=1)  # searchsorted side='right'
      slot = np.minimum(slot, 2)
      new_si = nbr[h, slot]
      state.si_idx[idx[h]] = new_si
      tr[idx[h]] += 1
      if log is not None:
        for j, hh in enumerate(h):
          log.env.append(int(idx[hh]))
          log.ctrl_seq.append(int(state.ctrl_count[idx[hh]]))
          log.iteration.append(it)
          log.elapsed_us.append(int(elapsed[idx[hh]]))
          log.slot.append(int(slot[j]))
          log.new_si.append(int(new_si[j]))
    active = (elapsed < dwell_us) & live
    it += 1
  state.ctrl_count[live] += np.uint32(1)
  state.n_events += ev
  state.n_transitions += tr
  return {'events': ev, 'transitions': tr, 'first_rates': first_rates}


# ----------------------------------------------------------------------------
# Frame transforms and observation (a13, a14, a15, a16)
# ----------------------------------------------------------------------------
def microscope_to_material(fov: np.ndarray, p: np.ndarray) -> np.ndarray:
  """microscope_utils.py:362-369: p * (ur - ll) + ll per axis."""
  return np.stack((p[:, 0] * (fov[:, 2] - fov[:, 0]) + fov[:, 0],
                   p[:, 1] * (fov[:, 3] - fov[:, 1]) + fov[:, 1]), axis=1)


def material_to_microscope(fov: np.ndarray, p: np.ndarray) -> np.ndarray:
  """microscope_utils.py:421-428: (p - ll) / (ur - ll) per axis."""
  return np.stack(((p[:, 0] - fov[:, 0]) / (fov[:, 2] - fov[:, 0]),
                   (p[:, 1] - fov[:, 1]) / (fov[:, 3] - fov[:, 1])), axis=1)


def silicon_outside_safe_area(state: OracleState) -> np.ndarray:
  """simulator.py:230-250 on the observed grid of graphene.py:600-644."""
  e = np.arange(state.num_envs)
  p = site_positions(state, state.si_idx, e)
  f = state.fov
  in_view = ((f[:, 0] <= p[:, 0]) & (p[:, 0] <= f[:, 2]) &
             (f[:, 1] <= p[:, 1]) & (p[:, 1] <= f[:, 3]))
  q = material_to_microscope(f, p)
  near_edge = ((q < 0.25) | (q > 0.75)).any(axis=1)
  return (~in_view) | near_edge


def step_and_image(state: OracleState, controls: np.ndarray,
                   dwell_us: np.ndarray, image_duration_us: int = 2000000,
                   rate_fn: int = RATE_SIMPLE,
                   mlp: Optional[MlpParams] = None,
                   log: Optional[EventLog] = None,
                   gmm: Optional[dict] = None,
                   skip: Optional[np.ndarray] = None,
                   prior: Optional[dict] = None) -> dict:
  """simulator.py:107-182 for all envs (without rendering).

  controls: float64 [E, C, 2] microscope frame; dwell_us: int64 [E, C].
  Returns elapsed_us [E], recentred [E], transitions [E], events [E].
  """
  e = state.num_envs
  controls = np.asarray(controls, dtype=np.float64).reshape(e, -1, 2)
  dwell_us = np.broadcast_to(
      np.asarray(dwell_us, dtype=np.int64).reshape(-1, controls.shape[1])
      if np.ndim(dwell_us) else np.asarray(dwell_us, dtype=np.int64),
      (e, controls.shape[1]))
  elapsed = np.zeros(e, dtype=np.int64)
  ev = np.zeros(e, dtype=np.int64)
  tr = np.zeros(e, dtype=np.int64)
  for c in range(controls.shape[1]):
    beam = microscope_to_material(state.fov, controls[:, c])  # :137
    out = apply_control(state, beam, dwell_us[:, c], rate_fn, mlp, log,
                        gmm=gmm, skip=skip, prior=prior)  # :147
    ev += out['events']
    tr += out['transitions']
    elapsed += dwell_us[:, c]  # :149
  elapsed += image_duration_us  # :152-153
  recentre = silicon_outside_safe_area(state)  # :156
  if skip is not None:  # envs that sit this call out
    recentre &= ~np.asarray(skip, bool)
    elapsed[np.asarray(skip, bool)] = 0
  if recentre.any():
    recenter_fov(state, np.nonzero(recentre)[0])  # :161-165
    elapsed[recentre] += image_duration_us  # :168-169
  state.sim_time_us += elapsed
  return {'elapsed_us': elapsed, 'recentred': recentre, 'transitions': tr,
          'events': ev}


def get_atoms_in_bounds(state: OracleState, env: int,
                        fov: Optional[np.ndarray] = None):
  """graphene.py:600-644: inclusive box filter, lattice order, normalised."""
  f = state.fov[env] if fov is None else np.asarray(fov, dtype=np.float64)
  p = all_positions(state, env)
  keep = ((f[0] <= p[:, 0]) & (p[:, 0] <= f[2]) &
          (f[1] <= p[:, 1]) & (p[:, 1] <= f[3]))
  sel = p[keep]
  q = np.stack(((sel[:, 0] - f[0]) / (f[2] - f[0]),
                (sel[:, 1] - f[1]) / (f[3] - f[1])), axis=1)
  z = np.full(p.shape[0], CARBON, dtype=np.int64)
  z[state.si_idx[env]] = SILICON
  return q, z[keep], np.nonzero(keep)[0]
