"""An independent float64 evaluation of the learned rate model.

The network of rate_learning/learn_rates.py:80-99 (hk.BatchNorm -> hk.nets.MLP
with swish -> softplus) and `apply_model` (:704-732) cannot be imported here
(Haiku / JAX / TensorFlow are absent), so a9 / a10 are pinned only by the
oracle's NumPy restatement.  This is a second restatement that shares no code
with it: torch.nn.functional building blocks (batch_norm in eval mode, linear,
silu, softplus, softmax) in float64.  Agreement of the two -- and of the
device kernels with this one -- is what stands in for the missing reference.
"""

import numpy as np
import torch
import torch.nn.functional as F


def _t(x):
  return torch.as_tensor(np.asarray(x), dtype=torch.float64)


def forward(params, x) -> np.ndarray:
  """call_mlp(is_training=False): float64 [B, 4]."""
  h = _t(x)
  if params.batchnorm:
    h = F.batch_norm(h, _t(params.bn_mean), _t(params.bn_var),
                     weight=_t(params.bn_scale), bias=_t(params.bn_offset),
                     training=False, eps=1e-5)
  h = F.silu(F.linear(h, _t(params.w0).T, _t(params.b0)))
  h = F.silu(F.linear(h, _t(params.w1).T, _t(params.b1)))
  return F.softplus(F.linear(h, _t(params.w2).T, _t(params.b2))).numpy()


def apply_model(params_list, x) -> np.ndarray:
  """softmax(out[:3]) * out[3], mean over the ensemble: float64 [B, 3]."""
  acc = 0.0
  for p in params_list:
    o = torch.as_tensor(forward(p, x))
    acc = acc + F.softmax(o[:, :3], dim=1) * o[:, 3:4]
  return (acc / len(params_list)).numpy()
