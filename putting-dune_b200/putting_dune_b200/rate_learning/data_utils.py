"""Synthetic rate-learning datasets (reference:
putting_dune/rate_learning/data_utils.py:158-303 generate_synthetic_data),
generated on the device by pd_generate_synthetic_data (PRIOR mode) and
pd_generate_synthetic_data_network (NETWORK mode)."""

from __future__ import annotations

import ctypes as C
import enum
import time
from typing import Mapping, Optional, Tuple

import numpy as np
import torch

from putting_dune_b200 import _native as nat


class SyntheticDataType(str, enum.Enum):
  NETWORK = 'network'
  PRIOR = 'prior'


def init_network(seed: int, input_dim: int, num_states: int = 3,
                 hidden: Tuple[int, int] = (1, 64)) -> Mapping[str, np.ndarray]:
  """Weights of the NETWORK mode's MLP (learn_rates.py:80-99 get_mlp_fn((1,
  64), num_states, batchnorm=False), data_utils.py:196-201) with Haiku's
  hk.Linear defaults: TruncatedNormal(stddev=1/sqrt(fan_in)) weights, zero
  biases.  The reference draws them with jax.random; these come from
  numpy's Generator(seed)."""
  rng = np.random.default_rng(seed)

  def trunc(shape):
    w = rng.standard_normal(shape)
    bad = np.abs(w) > 2
    while bad.any():
      w[bad] = rng.standard_normal(int(bad.sum()))
      bad = np.abs(w) > 2
    return (w / np.sqrt(shape[0])).astype(np.float32)

  h0, h1 = hidden
  return {'w0': trunc((input_dim, h0)), 'b0': np.zeros(h0, np.float32),
          'w1': trunc((h0, h1)), 'b1': np.zeros(h1, np.float32),
          'w2': trunc((h1, num_states + 1)),
          'b2': np.zeros(num_states + 1, np.float32)}


def generate_synthetic_data(
    num_data: int = 100, data_seed: Optional[int] = None, num_states: int = 3,
    position_dim: int = 2, context_dim: int = 2,
    actual_time_range: Tuple[float, float] = (0, 5),
    mode=SyntheticDataType.PRIOR, device=None, network=None,
) -> Tuple[Mapping[str, torch.Tensor], Mapping[str, torch.Tensor]]:
  """(train_data, test_data): dicts with next_state [n, 1] int32, dt [n, 1],
  rates [n, num_states], context [n, context_dim], position [n, position_dim]
  (float32 device tensors).  Draws are keyed by Philox(data_seed), not
  jax.random.  NETWORK mode takes the MLP's weights from `network` (a mapping
  w0, b0, w1, b1, w2, b2, [in][out] like Haiku's) or draws them with
  init_network(data_seed, ...)."""
  mode = SyntheticDataType(mode)
  if mode == SyntheticDataType.PRIOR and position_dim != 2:
    raise ValueError('the prior is defined over 2-D beam positions')
  if not torch.cuda.is_available():
    raise RuntimeError('putting_dune_b200 needs a CUDA device: there is no '
                       'CPU fallback')
  if data_seed is None:
    data_seed = int(time.time())
  dev = torch.device(device if device is not None else 'cuda')
  weights = None
  if mode == SyntheticDataType.NETWORK:
    if network is None:
      network = init_network(data_seed, context_dim + position_dim, num_states)
    weights = {k: torch.as_tensor(np.ascontiguousarray(network[k],
                                                       dtype=np.float32),
                                  device=dev)
               for k in ('w0', 'b0', 'w1', 'b1', 'w2', 'b2')}
    if (weights['w0'].shape[0] != context_dim + position_dim or
        weights['w1'].shape[0] != weights['w0'].shape[1] or
        weights['w2'].shape != (weights['w1'].shape[1], num_states + 1)):
      raise ValueError('network weights do not match the requested sizes')
  out = []
  P = lambda t: C.c_void_p(t.data_ptr())
  for split in (0, 1):
    d = {'next_state': torch.empty((num_data, 1), dtype=torch.int32, device=dev),
         'dt': torch.empty((num_data, 1), dtype=torch.float32, device=dev),
         'rates': torch.empty((num_data, num_states), dtype=torch.float32,
                              device=dev),
         'context': torch.empty((num_data, context_dim), dtype=torch.float32,
                                device=dev),
         'position': torch.empty((num_data, position_dim),
                                 dtype=torch.float32, device=dev)}
    with torch.cuda.device(dev):
      stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
      if mode == SyntheticDataType.NETWORK:
        nat.check(nat.lib.pd_generate_synthetic_data_network(
            int(data_seed) & 0xFFFFFFFFFFFFFFFF, split, num_data, num_states,
            context_dim, position_dim, float(actual_time_range[0]),
            float(actual_time_range[1]), P(weights['w0']), P(weights['b0']),
            P(weights['w1']), P(weights['b1']), P(weights['w2']),
            P(weights['b2']), weights['w0'].shape[1], weights['w1'].shape[1],
            P(d['next_state']), P(d['dt']), P(d['rates']),
            P(d['context']) if context_dim else None,
            P(d['position']) if position_dim else None, stream))
      else:
        nat.check(nat.lib.pd_generate_synthetic_data(
            int(data_seed) & 0xFFFFFFFFFFFFFFFF, split, num_data, num_states,
            context_dim, float(actual_time_range[0]),
            float(actual_time_range[1]), P(d['next_state']), P(d['dt']),
            P(d['rates']), P(d['context']) if context_dim else None,
            P(d['position']), stream))
    out.append(d)
  return out[0], out[1]
