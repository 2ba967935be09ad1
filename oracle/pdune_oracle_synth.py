"""TEST INFRASTRUCTURE -- NumPy restatement of
rate_learning/data_utils.py:158-303 generate_synthetic_data (PRIOR mode; the
NETWORK mode at the end of the file).  Only tests/ and
tests/golden/make_golden.py may import it.

The reference keys its draws by jax.random (absent here, and not
reproducible by another generator), so the sampling itself is **parity
unpinned**; the deterministic arithmetic between the draws and the outputs is
pinned: `sample_from_draws` is checked against the reference's own helper
functions (get_all_position_rotations, rotate_attributes, rotate_index,
geometry.jnp_rotate_coordinates, graphene.single_silicon_prior_rates) run
under oracle/refshim.py -- tests/golden/synth_reference.npz.
"""

from __future__ import annotations

import numpy as np

from oracle import pdune_oracle as po

STREAM_SYNTH = 10
MEAN = np.array((0.85, 0.0))       # constants.py:26
COV = 0.1                          # constants.py:27 (times the identity)
MAX_RATE = np.log(2) / 3           # constants.py:28


def rotate_coordinates(coord: np.ndarray, theta) -> np.ndarray:
  """geometry.py:69-84 jnp_rotate_coordinates: coord @ [[c, s], [-s, c]]."""
  c, s = np.cos(theta), np.sin(theta)
  return np.stack((coord[..., 0] * c - coord[..., 1] * s,
                   coord[..., 0] * s + coord[..., 1] * c), axis=-1)


def prior_rates(position: np.ndarray, num_states: int) -> np.ndarray:
  """data_utils.py:252-257: single_silicon_prior_rates of the num_states
  rotations of the position (graphene.py:121-130)."""
  out = []
  for k in range(num_states):
    x = rotate_coordinates(position, 2 * k * np.pi / num_states)
    d2 = ((x - MEAN) ** 2).sum(-1)
    out.append(MAX_RATE * np.exp(-0.5 * d2 / COV))
  return np.stack(out, axis=-1)


def sample_from_draws(z_pos, u_state, u_rot, u_time, u_win, z_ctx,
                      num_states, time_range):
  """data_utils.py:237-283 sample_from_prior given its random draws (normals
  z_pos [n, 2] and z_ctx [n, d]; uniforms in [0, 1), u_time in (0, 1])."""
  position = MEAN + np.sqrt(1.5 * COV) * z_pos
  rates = prior_rates(position, num_states)
  total = rates.sum(-1)
  cdf = np.cumsum(rates / total[:, None], axis=-1)
  state = np.minimum((u_state[:, None] >= cdf).sum(-1), num_states - 1)
  rf = np.minimum((u_rot * num_states).astype(np.int64), num_states - 1)
  position = rotate_coordinates(position, 2 * rf * np.pi / num_states)
  state = (state + rf) % num_states                      # rotate_index
  idx = (np.arange(num_states)[None, :] - rf[:, None]) % num_states
  rates = np.take_along_axis(rates, idx, axis=1)         # jnp.roll(rates, rf)
  next_time = -np.log(u_time) / total
  actual = time_range[0] + u_win * (time_range[1] - time_range[0])
  next_state = np.where(next_time < actual, state + 1, 0)
  return {'next_state': next_state.astype(np.int32)[:, None],
          'dt': actual.astype(np.float32)[:, None],
          'rates': rates.astype(np.float32),
          'context': z_ctx.astype(np.float32),
          'position': position.astype(np.float32),
          'cdf_margin': np.abs(u_state[:, None] - cdf).min(-1),
          'time_margin': np.abs(next_time - actual)}


def _normals(seed, ids, split, slot):
  w = po.philox4x32_10(ids, split, slot, STREAM_SYNTH, seed & 0xFFFFFFFF,
                       seed >> 32)
  u1 = (1.0 - po.u53(w[0], w[1])).astype(np.float32)
  u2 = po.u53(w[2], w[3]).astype(np.float32)
  r = np.sqrt(-2.0 * np.log(np.maximum(u1, np.float32(1e-37))))
  ang = 2.0 * np.pi * u2.astype(np.float64)
  return r * np.cos(ang), r * np.sin(ang)


def generate_synthetic_data(n, seed, split, num_states=3, context_dim=2,
                            time_range=(0.0, 5.0)):
  """The device kernel's draws (csrc/pd_synth.cu) + sample_from_draws."""
  ids = np.arange(n, dtype=np.uint32)
  zx, zy = _normals(seed, ids, split, 0)
  a, b = po.draw_pair(seed, ids, split, 1, STREAM_SYNTH)
  c, d = po.draw_pair(seed, ids, split, 2, STREAM_SYNTH)
  ctx = np.zeros((n, context_dim))
  for k in range(0, context_dim, 2):
    g0, g1 = _normals(seed, ids, split, 3 + k // 2)
    ctx[:, k] = g0
    if k + 1 < context_dim:
      ctx[:, k + 1] = g1
  f32 = lambda v: v.astype(np.float32).astype(np.float64)
  return sample_from_draws(np.stack((zx, zy), -1), f32(a), f32(b),
                           f32(1.0 - c), f32(d), ctx, num_states, time_range)


# ---------------------------------------------------------------------------
# NETWORK mode (data_utils.py:196-234).  The reference draws the MLP's weights
# and every sample with jax.random and runs a Haiku network (neither is in
# this container): **parity unpinned**.  The forward pass is po.mlp_forward,
# the restatement of learn_rates.py:80-99 that the learned-rate path uses.
# ---------------------------------------------------------------------------
def network_sample_from_draws(x, u_state, u_time, u_win, weights, num_states,
                              context_dim, time_range):
  """data_utils.py:201-234 sample_network_rates given its draws: x [n, D]
  standard normals, uniforms u_state / u_win in [0, 1), u_time in (0, 1]."""
  d = x.shape[1]
  f32 = np.float32
  params = po.MlpParams(
      bn_scale=np.ones(d, f32), bn_offset=np.zeros(d, f32),
      bn_mean=np.zeros(d, f32), bn_var=np.ones(d, f32),
      w0=weights['w0'], b0=weights['b0'], w1=weights['w1'], b1=weights['b1'],
      w2=weights['w2'], b2=weights['b2'], batchnorm=False)
  rates = po.mlp_forward(params, x)[:, :num_states].astype(np.float64)  # :213
  total = rates.sum(-1)
  cdf = np.cumsum(rates / total[:, None], axis=-1)
  state = np.minimum((u_state[:, None] >= cdf).sum(-1), num_states - 1)
  next_time = -np.log(u_time) / total
  actual = time_range[0] + u_win * (time_range[1] - time_range[0])
  next_state = np.where(next_time < actual, state + 1, 0)
  return {'next_state': next_state.astype(np.int32)[:, None],
          'dt': actual.astype(f32)[:, None],
          'rates': rates.astype(f32),
          'context': x[:, :context_dim].astype(f32),
          'position': x[:, context_dim:].astype(f32),
          'cdf_margin': np.abs(u_state[:, None] - cdf).min(-1),
          'time_margin': np.abs(next_time - actual)}


def generate_synthetic_data_network(n, seed, split, weights, num_states=3,
                                    context_dim=2, position_dim=2,
                                    time_range=(0.0, 5.0)):
  """The device kernel's draws (csrc/pd_synth.cu k_synthetic_network) +
  network_sample_from_draws."""
  ids = np.arange(n, dtype=np.uint32)
  d = context_dim + position_dim
  x = np.zeros((n, d))
  for k in range(0, d, 2):
    g0, g1 = _normals(seed, ids, split, 3 + k // 2)
    x[:, k] = g0
    if k + 1 < d:
      x[:, k + 1] = g1
  a, _ = po.draw_pair(seed, ids, split, 1, STREAM_SYNTH)
  c, dd = po.draw_pair(seed, ids, split, 2, STREAM_SYNTH)
  f32 = lambda v: v.astype(np.float32).astype(np.float64)
  return network_sample_from_draws(
      x.astype(np.float32), f32(a), f32(1.0 - c), f32(dd), weights,
      num_states, context_dim, time_range)
