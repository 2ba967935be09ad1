"""Boundary value types of the simulator (reference:
putting_dune/microscope_utils.py:72-81,180-198,235-441,503-551).

These classes carry the same fields, frame transforms and observer interface.
Serialisation (the reference's ProtoModel, microscope_utils.py:46-69, backed
by putting_dune_pb2 and TensorFlow) is provided through proto_wire.py, which
writes the same wire bytes without either dependency: `to_proto()` returns an
object with `SerializeToString()`, `from_proto_string()` parses.
"""

from __future__ import annotations

import dataclasses
import datetime as dt
from typing import NewType, Optional, Tuple

import numpy as np

from putting_dune_b200 import geometry
from putting_dune_b200 import proto_wire as pw


class SerializedProto:
  """Stands in for the generated message objects `to_proto()` returns in the
  reference: io.py:77 only ever calls SerializeToString() on them."""

  def __init__(self, data: bytes):
    self._data = data

  def SerializeToString(self) -> bytes:  # pylint: disable=invalid-name
    return self._data

  def __eq__(self, other):
    return isinstance(other, SerializedProto) and self._data == other._data

  def __hash__(self):
    return hash(self._data)


class ProtoModel:
  """microscope_utils.py:46-69."""

  def to_proto_string(self) -> bytes:
    raise NotImplementedError

  def to_proto(self) -> SerializedProto:
    return SerializedProto(self.to_proto_string())

  @classmethod
  def from_proto_string(cls, string: bytes):
    raise NotImplementedError

  @classmethod
  def from_proto(cls, message):
    return cls.from_proto_string(message.SerializeToString())


@dataclasses.dataclass(frozen=True)
class AtomicGrid(ProtoModel):
  """microscope_utils.py:72-131."""
  atom_positions: np.ndarray
  atomic_numbers: np.ndarray

  def to_proto_string(self) -> bytes:
    return pw.atomic_grid(self.atom_positions, self.atomic_numbers)

  @classmethod
  def from_proto_string(cls, string: bytes) -> 'AtomicGrid':
    return cls(*pw.parse_atomic_grid(string))

  def shift(self, shift_vector: np.ndarray) -> 'AtomicGrid':
    return AtomicGrid(self.atom_positions + np.asarray(shift_vector).reshape(
        1, 2), self.atomic_numbers)


AtomicGridMaterialFrame = NewType('AtomicGridMaterialFrame', AtomicGrid)
AtomicGridMicroscopeFrame = NewType('AtomicGridMicroscopeFrame', AtomicGrid)


@dataclasses.dataclass(frozen=True)
class BeamControl(ProtoModel):
  """microscope_utils.py:180-230."""
  position: geometry.Point
  dwell_time: dt.timedelta
  voltage_kv: Optional[float] = 60
  current_na: Optional[float] = 0.1

  def to_proto_string(self) -> bytes:
    return pw.beam_control(self.position.x, self.position.y,
                           self.dwell_time.total_seconds(), self.voltage_kv,
                           self.current_na)

  @classmethod
  def _from_fields(cls, d: dict) -> 'BeamControl':
    return cls(geometry.Point(d['position']),
               dt.timedelta(seconds=d['dwell_time_seconds']), d['voltage_kv'],
               d['current_na'])

  @classmethod
  def from_proto_string(cls, string: bytes) -> 'BeamControl':
    return cls._from_fields(pw.parse_beam_control(string))

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
class MicroscopeFieldOfView(ProtoModel):
  """microscope_utils.py:235-501: the scan window in material coordinates and
  the transforms between the unit microscope frame and angstroms."""
  lower_left: geometry.PointMaterialFrame
  upper_right: geometry.PointMaterialFrame

  def to_proto_string(self) -> bytes:
    return pw.field_of_view(self.lower_left.x, self.lower_left.y,
                            self.upper_right.x, self.upper_right.y)

  @classmethod
  def from_proto_string(cls, string: bytes) -> 'MicroscopeFieldOfView':
    ll, ur = pw.parse_field_of_view(string)
    return cls(geometry.Point(ll), geometry.Point(ur))

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
class MicroscopeObservation(ProtoModel):
  """microscope_utils.py:538-604."""
  grid: AtomicGrid
  fov: MicroscopeFieldOfView
  controls: Tuple[BeamControl, ...]
  elapsed_time: dt.timedelta
  image: Optional[np.ndarray] = None
  label_image: Optional[np.ndarray] = None

  def to_proto_string(self) -> bytes:
    return pw.observation(
        self.grid.to_proto_string(), self.fov.to_proto_string(),
        [c.to_proto_string() for c in self.controls],
        self.elapsed_time.total_seconds(), self.image, self.label_image)

  @classmethod
  def _from_fields(cls, d: dict) -> 'MicroscopeObservation':
    ll, ur = d['fov']
    return cls(
        grid=AtomicGrid(*d['grid']),
        fov=MicroscopeFieldOfView(geometry.Point(ll), geometry.Point(ur)),
        controls=tuple(BeamControl._from_fields(c) for c in d['controls']),
        elapsed_time=dt.timedelta(seconds=d['elapsed_time_seconds']),
        image=d['image'], label_image=d['label_image'])

  @classmethod
  def from_proto_string(cls, string: bytes) -> 'MicroscopeObservation':
    return cls._from_fields(pw.parse_observation(string))


@dataclasses.dataclass(frozen=True)
class Transition(ProtoModel):
  """microscope_utils.py:607-734."""
  grid_before: AtomicGrid
  grid_after: AtomicGrid
  fov_before: MicroscopeFieldOfView
  fov_after: MicroscopeFieldOfView
  controls: Tuple[BeamControl, ...]
  image_before: Optional[np.ndarray] = None
  image_after: Optional[np.ndarray] = None
  label_image_before: Optional[np.ndarray] = None
  label_image_after: Optional[np.ndarray] = None

  def to_proto_string(self) -> bytes:
    # microscope_utils.py:707-716: the label images are written only when the
    # corresponding image is present (reference quirk, kept)
    return pw.transition(
        self.grid_before.to_proto_string(), self.grid_after.to_proto_string(),
        self.fov_before.to_proto_string(), self.fov_after.to_proto_string(),
        [c.to_proto_string() for c in self.controls], self.image_before,
        self.image_after,
        self.label_image_before if self.image_before is not None else None,
        self.label_image_after if self.image_after is not None else None)

  @classmethod
  def from_proto_string(cls, string: bytes) -> 'Transition':
    d = pw.parse_transition(string)
    empty = (np.zeros((0, 2), np.float32), np.zeros(0, np.int32))
    zero = ((0.0, 0.0), (0.0, 0.0))
    fov = lambda v: MicroscopeFieldOfView(geometry.Point(v[0]),
                                          geometry.Point(v[1]))
    return cls(
        grid_before=AtomicGrid(*d.get('grid_before', empty)),
        grid_after=AtomicGrid(*d.get('grid_after', empty)),
        fov_before=fov(d.get('fov_before', zero)),
        fov_after=fov(d.get('fov_after', zero)),
        controls=tuple(BeamControl._from_fields(c) for c in d['controls']),
        image_before=d['image_before'], image_after=d['image_after'],
        label_image_before=d['label_image_before'],
        label_image_after=d['label_image_after'])


@dataclasses.dataclass(frozen=True)
class Trajectory(ProtoModel):
  """microscope_utils.py:737-770."""
  observations: Tuple[MicroscopeObservation, ...]

  def to_proto_string(self) -> bytes:
    return pw.trajectory([o.to_proto_string() for o in self.observations])

  @classmethod
  def from_proto_string(cls, string: bytes) -> 'Trajectory':
    return cls(tuple(MicroscopeObservation._from_fields(d)
                     for d in pw.parse_trajectory(string)))
