"""Boundary value types of the simulator (reference:
putting_dune/microscope_utils.py:72-81,180-198,235-441,503-551).

Proto/TFRecord serialisation of the reference types is an offline data format
and is out of scope; these classes carry the same
fields, frame transforms and observer interface.
"""

from __future__ import annotations

import dataclasses
import datetime as dt
from typing import NewType, Optional, Tuple

import numpy as np

from putting_dune_b200 import geometry


@dataclasses.dataclass(frozen=True)
class AtomicGrid:
  """microscope_utils.py:72-87."""
  atom_positions: np.ndarray
  atomic_numbers: np.ndarray

  def shift(self, shift_vector: np.ndarray) -> 'AtomicGrid':
    return AtomicGrid(self.atom_positions + np.asarray(shift_vector).reshape(
        1, 2), self.atomic_numbers)


AtomicGridMaterialFrame = NewType('AtomicGridMaterialFrame', AtomicGrid)
AtomicGridMicroscopeFrame = NewType('AtomicGridMicroscopeFrame', AtomicGrid)


@dataclasses.dataclass(frozen=True)
class BeamControl:
  """microscope_utils.py:180-207."""
  position: geometry.Point
  dwell_time: dt.timedelta
  voltage_kv: Optional[float] = 60
  current_na: Optional[float] = 0.1

  def shift(self, shift: geometry.Point) -> 'BeamControl':
    return BeamControl(
        geometry.Point(self.position.x + shift.x, self.position.y + shift.y),
        self.dwell_time, self.voltage_kv, self.current_na)


BeamControlMaterialFrame = NewType('BeamControlMaterialFrame', BeamControl)
BeamControlMicroscopeFrame = NewType('BeamControlMicroscopeFrame', BeamControl)


def timedelta_to_us(t) -> int:
  """Exact integer microseconds of a timedelta (or float seconds)."""
  if isinstance(t, dt.timedelta):
    return t // dt.timedelta(microseconds=1)
  return dt.timedelta(seconds=float(t)) // dt.timedelta(microseconds=1)


@dataclasses.dataclass(frozen=True)
class MicroscopeFieldOfView:
  """microscope_utils.py:235-484: the scan window in material coordinates and
  the transforms between the unit microscope frame and angstroms."""
  lower_left: geometry.PointMaterialFrame
  upper_right: geometry.PointMaterialFrame

  def shift(self, shift: geometry.Point) -> 'MicroscopeFieldOfView':
    return MicroscopeFieldOfView(
        geometry.Point(self.lower_left.x + shift.x,
                       self.lower_left.y + shift.y),
        geometry.Point(self.upper_right.x + shift.x,
                       self.upper_right.y + shift.y))

  def _ll(self) -> np.ndarray:
    return self.lower_left.as_array()

  def _ur(self) -> np.ndarray:
    return self.upper_right.as_array()

  @property
  def offset(self) -> geometry.Point:
    return geometry.Point((self._ll() + self._ur()) / 2)

  @property
  def width(self) -> float:
    return self.upper_right.x - self.lower_left.x

  @property
  def height(self) -> float:
    return self.upper_right.y - self.lower_left.y

  def resize(self, new_width: float,
             new_height: float) -> 'MicroscopeFieldOfView':
    assert new_width > 0 and new_height > 0
    half = np.asarray([new_width, new_height]) / 2
    centre = (self._ll() + self._ur()) / 2
    return MicroscopeFieldOfView(geometry.Point(centre - half),
                                 geometry.Point(centre + half))

  def zoom(self, zoom_factor: float) -> 'MicroscopeFieldOfView':
    assert zoom_factor > 0
    return self.resize(self.width / zoom_factor, self.height / zoom_factor)

  def microscope_frame_to_material_frame(self, point):
    """p * (ur - ll) + ll (microscope_utils.py:344-383)."""
    ll, scale = self._ll(), self._ur() - self._ll()
    if isinstance(point, AtomicGrid):
      return AtomicGrid(point.atom_positions * scale + ll,
                        point.atomic_numbers)
    if isinstance(point, np.ndarray):
      shape = (2,) if point.ndim == 1 else (-1, 2)
      return (point.reshape(-1, 2) * scale + ll).reshape(shape)
    if isinstance(point, geometry.Point):
      return geometry.Point(point.x * scale[0] + ll[0],
                            point.y * scale[1] + ll[1])
    if isinstance(point, BeamControl):
      return BeamControl(
          geometry.Point(point.position.x * scale[0] + ll[0],
                         point.position.y * scale[1] + ll[1]),
          point.dwell_time, point.voltage_kv, point.current_na)
    raise NotImplementedError(f'Point of type {type(point)} is not supported.')

  def material_frame_to_microscope_frame(self, point):
    """(p - ll) / (ur - ll) (microscope_utils.py:409-441)."""
    ll, scale = self._ll(), self._ur() - self._ll()
    if isinstance(point, AtomicGrid):
      return AtomicGrid((point.atom_positions - ll) / scale,
                        point.atomic_numbers)
    if isinstance(point, np.ndarray):
      shape = (2,) if point.ndim == 1 else (-1, 2)
      return ((point.reshape(-1, 2) - ll) / scale).reshape(shape)
    if isinstance(point, geometry.Point):
      return geometry.Point((point.x - ll[0]) / scale[0],
                            (point.y - ll[1]) / scale[1])
    if isinstance(point, BeamControl):
      return BeamControl(
          geometry.Point((point.position.x - ll[0]) / scale[0],
                         (point.position.y - ll[1]) / scale[1]),
          point.dwell_time, voltage_kv=point.voltage_kv,
          current_na=point.current_na)
    raise NotImplementedError(f'Point of type {type(point)} is not supported.')

  def get_atoms_in_bounds(self, grid: AtomicGrid,
                          tolerance: float = 0) -> AtomicGrid:
    """Atoms of a material-frame grid inside the FOV (+/- tolerance),
    positions left in the material frame (microscope_utils.py:447-479)."""
    ll, ur = self._ll() - tolerance, self._ur() + tolerance
    pos = grid.atom_positions
    keep = np.all((ll <= pos) & (pos <= ur), axis=1)
    return AtomicGrid(pos[keep], grid.atomic_numbers[keep])

  def as_array(self) -> np.ndarray:
    return np.concatenate((self._ll(), self._ur()))

  def __str__(self) -> str:
    ll, ur = self.lower_left, self.upper_right
    return f'FOV [({ll.x:.2f}, {ll.y:.2f}), ({ur.x:.2f}, {ur.y:.2f})]'


class SimulatorObserver:
  """microscope_utils.py:503-535: callbacks on simulator events."""

  def observe_reset(self, grid: AtomicGrid,
                    fov: MicroscopeFieldOfView) -> None:
    pass

  def observe_apply_control(self, control: BeamControl) -> None:
    pass

  def observe_transition(self, time_since_control_was_applied: dt.timedelta,
                         grid: AtomicGrid) -> None:
    pass

  def observe_fov_change(self, fov: MicroscopeFieldOfView) -> None:
    pass

  def observe_take_image(self, duration: dt.timedelta,
                         fov: MicroscopeFieldOfView) -> None:
    pass

  def observe_generated_image(self, image: np.ndarray) -> None:
    pass


@dataclasses.dataclass(frozen=True)
class MicroscopeObservation:
  """microscope_utils.py:538-551."""
  grid: AtomicGrid
  fov: MicroscopeFieldOfView
  controls: Tuple[BeamControl, ...]
  elapsed_time: dt.timedelta
  image: Optional[np.ndarray] = None
  label_image: Optional[np.ndarray] = None
