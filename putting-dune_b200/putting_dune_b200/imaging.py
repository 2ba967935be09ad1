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


_workspaces = {}


def _workspace(device, image_size: int):
  import ctypes as C
  import torch
  from putting_dune_b200 import _native as nat
  key = (str(device), image_size)
  if key not in _workspaces:
    need = C.c_int64()
    with torch.cuda.device(device):
      nat.check(nat.lib.pd_render_workspace_bytes(image_size, C.byref(need)))
    _workspaces[key] = torch.empty(need.value, dtype=torch.uint8,
                                   device=device)
  return _workspaces[key]


def render_batch(batch, env_ids=None, image_size: int = 512,
                 stop_stage: int = 7, advance_frame_count: bool = True,
                 out=None, buffer_size: float = 0.0):
  """Renders STEM frames for envs of an ``engine.EnvBatch`` on the device
  (imaging.py:239-265 ``generate_stem_image`` with each env's current FOV,
  Si position and image parameters).  Returns float32 [m, S, S] in [0, 1].
  ``buffer_size`` > 0 is imaging.py:129-168 with the whole lattice as the
  grid: atoms just outside the frame contribute their Gaussian tails."""
  import ctypes as C
  import torch
  from putting_dune_b200 import _native as nat
  dev = batch.device
  ids = None
  m = batch.num_envs
  if env_ids is not None:
    ids = torch.as_tensor(env_ids, dtype=torch.int32,
                          device=dev).reshape(-1).contiguous()
    m = ids.numel()
  if out is None:
    out = torch.empty((m, image_size, image_size), dtype=torch.float32,
                      device=dev)
  ws = _workspace(dev, image_size)
  P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
  with torch.cuda.device(dev):
    nat.check(nat.lib.pd_render(
        C.byref(batch.lattice_tables.c), C.byref(batch.c), P(ids), m,
        image_size, int(stop_stage), int(bool(advance_frame_count)),
        float(buffer_size), P(out), P(ws), ws.numel(),
        C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
  return out


def sample_image_parameters(rng, image_size: int = 512):
  """imaging.py:42-54 on the host, for callers that want to choose their own
  parameters (the simulator samples them on the device at reset)."""
  return ImageGenerationParameters(
      intensity_exponent=rng.uniform(1.4, 2.0),
      gaussian_variance=rng.uniform(0.0, 5e-3),
      jitter_rate=rng.uniform(0.0, 5.0),
      poisson_rate_multiplier=rng.exponential(15) + 1.0,
      salt_and_pepper_amount=rng.uniform(0.0, 1e-3),
      blur_amount=rng.uniform(0.0, 1.0),
      contrast_gamma=rng.uniform(0.7, 1.3),
      exponential_lambda=rng.uniform(0.0, 0.2),
      uniform_noise_scale=rng.uniform(0.0, 0.2),
      image_size=image_size)


def sample_noisy_image_parameters(rng, image_size: int = 512):
  """imaging.py:57-72 on the host (``EnvBatch.sample_image_params(noisy=True)``
  draws the same distribution on the device)."""
  return ImageGenerationParameters(
      intensity_exponent=rng.uniform(1.4, 2.0),
      gaussian_variance=rng.uniform(0.0, 0.3),
      jitter_rate=rng.uniform(0.0, 5.0),
      poisson_rate_multiplier=rng.exponential(15) + 1.0,
      salt_and_pepper_amount=rng.uniform(0.0, 1e-2),
      blur_amount=rng.uniform(0.0, 0.25),
      contrast_gamma=rng.uniform(0.5, 1.5),
      exponential_lambda=rng.uniform(0.0, 0.25),
      uniform_noise_scale=rng.uniform(0.0, 0.25),
      image_size=image_size)


def generate_grid_mask_batch(batch, env_ids=None, *,
                             intensity_exponent: float = 1.7,
                             image_size: int = 512, out=None):
  """imaging.py:75-114 ``generate_grid_mask`` for envs of an
  ``engine.EnvBatch`` (atoms in each env's current FOV): uint8 [m, S, S]
  holding 0, 6 (C) or 14 (Si)."""
  import ctypes as C
  import torch
  from putting_dune_b200 import _native as nat
  from putting_dune_b200 import constants
  dev = batch.device
  ids = None
  m = batch.num_envs
  if env_ids is not None:
    ids = torch.as_tensor(env_ids, dtype=torch.int32,
                          device=dev).reshape(-1).contiguous()
    m = ids.numel()
  if out is None:
    out = torch.empty((m, image_size, image_size), dtype=torch.uint8,
                      device=dev)
  # (atomic_number / CARBON) ** intensity_exponent * 0.1 in float64, as the
  # reference evaluates it (imaging.py:108)
  radius = [(np.int64(z) / constants.CARBON) ** intensity_exponent * 0.1
            for z in (constants.CARBON, constants.SILICON)]
  P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
  with torch.cuda.device(dev):
    nat.check(nat.lib.pd_render_mask(
        C.byref(batch.lattice_tables.c), C.byref(batch.c), P(ids), m,
        image_size, float(radius[0]), float(radius[1]), P(out),
        C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
  return out
