"""Synthetic rate-learning datasets (reference:
putting_dune/rate_learning/data_utils.py:158-303 generate_synthetic_data),
generated on the device by pd_generate_synthetic_data."""

from __future__ import annotations

import ctypes as C
import enum
import time
from typing import Mapping, Optional, Tuple

import torch

from putting_dune_b200 import _native as nat


class SyntheticDataType(str, enum.Enum):
  NETWORK = 'network'
  PRIOR = 'prior'


def generate_synthetic_data(
    num_data: int = 100, data_seed: Optional[int] = None, num_states: int = 3,
    position_dim: int = 2, context_dim: int = 2,
    actual_time_range: Tuple[float, float] = (0, 5),
    mode=SyntheticDataType.PRIOR, device=None,
) -> Tuple[Mapping[str, torch.Tensor], Mapping[str, torch.Tensor]]:
  """(train_data, test_data): dicts with next_state [n, 1] int32, dt [n, 1],
  rates [n, num_states], context [n, context_dim], position [n, 2] (float32
  device tensors).  Draws are keyed by Philox(data_seed), not jax.random."""
  if SyntheticDataType(mode) != SyntheticDataType.PRIOR:
    raise NotImplementedError(
        'only the informed-prior generator is provided (the network mode '
        'draws a random Haiku MLP)')
  if position_dim != 2:
    raise ValueError('the prior is defined over 2-D beam positions')
  if not torch.cuda.is_available():
    raise RuntimeError('putting_dune_b200 needs a CUDA device: there is no '
                       'CPU fallback')
  if data_seed is None:
    data_seed = int(time.time())
  dev = torch.device(device if device is not None else 'cuda')
  out = []
  P = lambda t: C.c_void_p(t.data_ptr())
  for split in (0, 1):
    d = {'next_state': torch.empty((num_data, 1), dtype=torch.int32, device=dev),
         'dt': torch.empty((num_data, 1), dtype=torch.float32, device=dev),
         'rates': torch.empty((num_data, num_states), dtype=torch.float32,
                              device=dev),
         'context': torch.empty((num_data, context_dim), dtype=torch.float32,
                                device=dev),
         'position': torch.empty((num_data, 2), dtype=torch.float32,
                                 device=dev)}
    with torch.cuda.device(dev):
      nat.check(nat.lib.pd_generate_synthetic_data(
          int(data_seed) & 0xFFFFFFFFFFFFFFFF, split, num_data, num_states,
          context_dim, float(actual_time_range[0]),
          float(actual_time_range[1]), P(d['next_state']), P(d['dt']),
          P(d['rates']), P(d['context']) if context_dim else None,
          P(d['position']),
          C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    out.append(d)
  return out[0], out[1]
