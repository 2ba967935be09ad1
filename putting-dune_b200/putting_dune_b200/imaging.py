"""STEM image generation (reference: putting_dune/imaging.py)."""

from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass(frozen=True)
class ImageGenerationParameters:
  """imaging.py:27-39."""
  intensity_exponent: float
  gaussian_variance: float
  jitter_rate: float
  poisson_rate_multiplier: float
  salt_and_pepper_amount: float
  blur_amount: float
  contrast_gamma: float
  exponential_lambda: float
  uniform_noise_scale: float
  image_size: int = 512

  def as_array(self) -> np.ndarray:
    return np.array([
        self.intensity_exponent, self.gaussian_variance, self.jitter_rate,
        self.poisson_rate_multiplier, self.salt_and_pepper_amount,
        self.blur_amount, self.contrast_gamma, self.exponential_lambda,
        self.uniform_noise_scale], dtype=np.float64)


def render_batch(batch, image_size: int = 512):
  raise NotImplementedError('renderer kernel not built yet')
