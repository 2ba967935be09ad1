"""Learned transition-rate predictor, inference only.

Reference: putting_dune/rate_learning/learn_rates.py --
``LearnedTransitionRatePredictor`` (:656), ``predict`` (:925-972),
``apply_model`` (:704-732), network ``get_mlp_fn`` (:80-99).

Weights come from the reference's Haiku parameter/state trees
(``engine.MlpWeights.from_haiku``) or any flat set of arrays; the forward pass
runs in the CUDA library (``csrc/pd_mlp.cu``): on the tensor cores with fp16
hi + lo operands (``tensor_core=2``: rates within 1e-6 of the FP32 kernel's,
inside its tolerance against the reference arithmetic) where the layer sizes
allow it, else as an FP32 FMA GEMM.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from putting_dune_b200 import _native as nat
from putting_dune_b200 import engine


class LearnedTransitionRatePredictor:
  """A distilled (one-model) predictor usable as a CanonicalRatePredictionFn,
  plus ``apply_model`` over an ensemble."""

  def __init__(self, models: Sequence[engine.MlpWeights] | engine.MlpWeights,
               device=None, tensor_core='auto'):
    """tensor_core: 'auto' (the split-precision tcgen05 kernel when the hidden
    sizes are multiples of 32 up to 256, else FP32), 0 (FP32 FMA), 2 (split
    precision, a parity path) or 1 (bf16 operands: 2e-2, throughput only)."""
    if isinstance(models, engine.MlpWeights):
      models = [models]
    self.models = list(models)
    self.num_models = len(self.models)
    self._device = device
    self._tensor_core = tensor_core
    self._specs = None

  @staticmethod
  def _tensor_mode(mlp: engine.MlpWeights, requested) -> int:
    if requested != 'auto':
      return int(requested)
    h1, h2 = np.asarray(mlp.w1).shape
    ok = (h1 % 32 == 0 and h1 <= 256 and h2 in (32, 64, 128, 256) and
          (h1 * (128 + h2) * 4 <= 190 * 1024 or h1 % 64 == 0))
    return 2 if ok else 0

  def _build(self):
    if self._specs is None:
      self._specs = [engine.RateSpec(
          nat.RATE_LEARNED, mlp=m, device=self._device,
          tensor_core=self._tensor_mode(m, self._tensor_core))
                     for m in self.models]
    return self._specs

  def rate_spec(self) -> engine.RateSpec:
    """The device rate function behind ``predict`` (learn_rates.py:925-972).
    Like the reference's packaged model (``jnp.squeeze(rates, axis=0)``,
    :900-903) this needs a one-model ensemble."""
    if self.num_models != 1:
      raise ValueError('predict() needs a distilled, single-model predictor '
                       '(learn_rates.py:900-903 squeezes the model axis)')
    return self._build()[0]

  def predict(self, grid, beam_pos, current_position, neighbor_indices,
              voltage_kv: float = 60, current_na: float = 0.1) -> np.ndarray:
    from putting_dune_b200 import graphene
    return graphene._canonical_rates(  # pylint: disable=protected-access
        self.rate_spec(), grid, beam_pos, neighbor_indices)

  def apply_model(self, x, key=None,
                  model_index: Optional[int] = None) -> torch.Tensor:
    """learn_rates.py:704-732: mean over models of softmax(o[:3]) * o[3].
    x: [B, 2] beam contexts -> float32 [B, 3] on the device."""
    specs = self._build()
    if model_index is not None:
      specs = [specs[model_index]]
    dev = next(iter(specs[0]._tensors.values())).device  # pylint: disable=protected-access
    xt = torch.as_tensor(x, dtype=torch.float32, device=dev).reshape(
        -1, 2).contiguous()
    out = torch.empty((xt.shape[0], 3), dtype=torch.float32, device=dev)
    arr = (nat.PdMlp * len(specs))(*[s._mlp_c for s in specs])  # pylint: disable=protected-access
    with torch.cuda.device(dev):
      nat.check(nat.lib.pd_mlp_apply_model(
          arr, len(specs), C.c_void_p(xt.data_ptr()), xt.shape[0],
          C.c_void_p(out.data_ptr()),
          C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out
