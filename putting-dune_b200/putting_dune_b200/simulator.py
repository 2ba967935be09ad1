"""Simulator front ends.

``PuttingDuneSimulator`` keeps the reference's call signatures
(putting_dune/simulator.py:27-250) on a device batch of one env;
``BatchedSimulator`` is the same step for many independent envs at once and is
what a throughput user calls.
"""

from __future__ import annotations

import dataclasses
import datetime as dt
from typing import Optional, Sequence

import numpy as np
import torch

from putting_dune_b200 import engine
from putting_dune_b200 import geometry
from putting_dune_b200 import graphene
from putting_dune_b200 import imaging
from putting_dune_b200 import microscope_utils as mu


def _fov_from_array(a: np.ndarray) -> mu.MicroscopeFieldOfView:
  return mu.MicroscopeFieldOfView(geometry.Point(a[0], a[1]),
                                  geometry.Point(a[2], a[3]))


class PuttingDuneSimulator:
  """simulator.py:27-250.  One env; see BatchedSimulator for many."""

  def __init__(self, material: graphene.PristineSingleDopedGraphene, *,
               image_duration: dt.timedelta = dt.timedelta(seconds=2.0),
               observers: Sequence[mu.SimulatorObserver] = ()):
    self.material = material
    self._observers = list(observers)
    self._image_duration = image_duration
    self._has_been_reset = False

  # The reference's tests poke these attributes (simulator_test.py:95,304);
  # they are views of the device state.
  @property
  def _fov(self) -> mu.MicroscopeFieldOfView:
    return _fov_from_array(self.material.batch.fov[0].cpu().numpy())

  @_fov.setter
  def _fov(self, fov: mu.MicroscopeFieldOfView) -> None:
    b = self.material.batch
    b.fov[0] = torch.as_tensor(fov.as_array(), device=b.device)

  @property
  def _fov_scale(self) -> float:
    return float(self.material.batch.fov_scale[0].item())

  @property
  def _image_parameters(self) -> imaging.ImageGenerationParameters:
    p = self.material.batch.image_params[0].cpu().numpy()
    return imaging.ImageGenerationParameters(*[float(v) for v in p])

  @_image_parameters.setter
  def _image_parameters(self, p: imaging.ImageGenerationParameters) -> None:
    b = self.material.batch
    b.image_params[0] = torch.as_tensor(p.as_array(), device=b.device)

  def reset(self, rng, return_image: bool = False) -> mu.MicroscopeObservation:
    """simulator.py:65-105 (material, FOV and image parameters are all drawn
    by the one reset kernel, in the reference's draw order)."""
    self.material.reset(rng)
    self._has_been_reset = True
    fov = self._fov
    if self._observers:
      grid = self.material.grid
      for observer in self._observers:
        observer.observe_reset(grid, fov)
        observer.observe_fov_change(fov)
    observed_grid = self._observe(fov)
    image = self._generate_image() if return_image else None
    return mu.MicroscopeObservation(
        grid=observed_grid, fov=fov, controls=(),
        elapsed_time=self._image_duration, image=image)

  def step_and_image(self, rng, controls: Sequence[mu.BeamControl],
                     return_image: bool = False) -> mu.MicroscopeObservation:
    """simulator.py:107-182."""
    self._assert_has_been_reset('step_and_image')
    controls = list(controls)
    batch = self.material.batch
    fov_before = self._fov
    xy = np.array([[[c.position.x, c.position.y] for c in controls]],
                  dtype=np.float64).reshape(1, len(controls), 2)
    dwell = np.array([[mu.timedelta_to_us(c.dwell_time) for c in controls]],
                     dtype=np.int64).reshape(1, len(controls))
    out = batch.step_and_image(
        xy, dwell, self.material._rate_spec(),  # pylint: disable=protected-access
        mu.timedelta_to_us(self._image_duration))
    self.material._raise_on_status()  # pylint: disable=protected-access
    recentred = bool(out.recentred[0].item())
    fov = self._fov if recentred else fov_before
    if self._observers:
      for i, control in enumerate(controls):
        material_control = fov_before.microscope_frame_to_material_frame(
            control)
        for observer in self._observers:
          observer.observe_apply_control(material_control)
        self.material._replay_transitions(out, i, self._observers)  # pylint: disable=protected-access
      for observer in self._observers:
        observer.observe_take_image(duration=self._image_duration,
                                    fov=fov_before)
      if recentred:
        for observer in self._observers:
          observer.observe_fov_change(fov)
          observer.observe_take_image(duration=self._image_duration, fov=fov)
    observed_grid = self.material.get_atoms_in_bounds(fov.lower_left,
                                                      fov.upper_right)
    image = self._generate_image() if return_image else None
    return mu.MicroscopeObservation(
        grid=observed_grid, fov=fov, controls=tuple(controls),
        elapsed_time=dt.timedelta(microseconds=int(out.elapsed_us[0].item())),
        image=image)

  def add_observer(self, observer: mu.SimulatorObserver) -> None:
    self._observers.append(observer)

  def remove_observer(self, observer: mu.SimulatorObserver) -> None:
    self._observers.remove(observer)

  def _observe(self, fov: mu.MicroscopeFieldOfView) -> mu.AtomicGrid:
    grid = self.material.get_atoms_in_bounds(fov.lower_left, fov.upper_right)
    for observer in self._observers:
      observer.observe_take_image(duration=self._image_duration, fov=fov)
    return grid

  def _generate_image(self) -> np.ndarray:
    image = imaging.render_batch(self.material.batch)[0].cpu().numpy()
    image = image.astype(np.float64)
    for observer in self._observers:
      observer.observe_generated_image(image)
    return image

  def _assert_has_been_reset(self, fn_name: str) -> None:
    if not self._has_been_reset:
      raise RuntimeError(
          f'Must call reset on {self.__class__} before {fn_name}.')


@dataclasses.dataclass
class BatchedObservation:
  """Device-resident observation of a batched step."""
  fov: torch.Tensor  # float64 [E, 4]
  elapsed_us: torch.Tensor  # int64 [E]
  silicon_xy: torch.Tensor  # float64 [E, 2] material frame
  transitions: torch.Tensor  # int32 [E]
  recentred: torch.Tensor  # uint8 [E]
  image: Optional[torch.Tensor] = None  # float32 [E, S, S]


class BatchedSimulator:
  """``num_envs`` independent PuttingDuneSimulators stepped by one launch.

  Semantics per env are those of simulator.py:65-182; inputs and outputs are
  device tensors so that an RL loop never leaves the GPU.
  """

  def __init__(self, num_envs: int, *, rate_function=None,
               grid_columns: int = 50,
               image_duration: dt.timedelta = dt.timedelta(seconds=2.0),
               seed: int = 0, device=None, env_offset: int = 0,
               log_capacity: int = 0):
    if rate_function is None:
      rate_function = graphene.PristineSingleSiGrRatePredictor(
          graphene.simple_canonical_rate_function)
    self.rate = (rate_function if isinstance(rate_function, engine.RateSpec)
                 else rate_function.rate_spec())
    self.batch = engine.EnvBatch(
        num_envs, seed=seed, grid_columns=grid_columns, device=device,
        env_offset=env_offset, log_capacity=log_capacity)
    self.image_duration_us = mu.timedelta_to_us(image_duration)
    self._has_been_reset = False

  @property
  def num_envs(self) -> int:
    return self.batch.num_envs

  def reset(self, mask=None, return_image: bool = False) -> BatchedObservation:
    self.batch.reset(mask)
    self._has_been_reset = True
    b = self.batch
    e = b.num_envs
    return BatchedObservation(
        fov=b.fov, silicon_xy=b.silicon_position(),
        elapsed_us=torch.full((e,), self.image_duration_us, dtype=torch.int64,
                              device=b.device),
        transitions=torch.zeros(e, dtype=torch.int32, device=b.device),
        recentred=torch.zeros(e, dtype=torch.uint8, device=b.device),
        image=imaging.render_batch(b) if return_image else None)

  def step_and_image(self, controls_xy, dwell_time,
                     return_image: bool = False) -> BatchedObservation:
    """controls_xy: [E, C, 2] (or [E, 2]) microscope frame; dwell_time: a
    timedelta / seconds for all envs or an int64 microsecond tensor [E, C]."""
    if not self._has_been_reset:
      raise RuntimeError(
          f'Must call reset on {self.__class__} before step_and_image.')
    dwell = (dwell_time if torch.is_tensor(dwell_time) or isinstance(
        dwell_time, np.ndarray) else mu.timedelta_to_us(dwell_time))
    out = self.batch.step_and_image(controls_xy, dwell, self.rate,
                                    self.image_duration_us)
    return BatchedObservation(
        fov=self.batch.fov, elapsed_us=out.elapsed_us, silicon_xy=out.si_xy,
        transitions=out.transitions, recentred=out.recentred,
        image=imaging.render_batch(self.batch) if return_image else None)

  def rollout(self, controls_xy, dwell_time, record: bool = False):
    if not self._has_been_reset:
      raise RuntimeError('Must call reset before rollout.')
    return self.batch.rollout(controls_xy, mu.timedelta_to_us(dwell_time),
                              self.rate, self.image_duration_us, record)

  def rollout_host(self, actions_xy, dwell_time, action_mode=None,
                   max_distance_angstroms: float = 1.42, out=None):
    """`rollout` with host buffers in and out (EnvBatch.rollout_host)."""
    if not self._has_been_reset:
      raise RuntimeError('Must call reset before rollout.')
    from putting_dune_b200 import _native as nat
    return self.batch.rollout_host(
        actions_xy, mu.timedelta_to_us(dwell_time), self.rate,
        self.image_duration_us,
        nat.ACTION_DIRECT if action_mode is None else action_mode,
        max_distance_angstroms, out)

  def get_atoms_in_bounds(self, fov=None, max_atoms=None):
    return self.batch.get_atoms_in_bounds(fov, max_atoms)
