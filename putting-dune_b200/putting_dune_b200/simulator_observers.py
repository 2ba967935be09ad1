"""Event recorder with the reference's interface
(putting_dune/simulator_observers.py:26-105): an ordered list of
``SimulatorEvent(event_type, event_data)`` with the same event_data keys."""

from __future__ import annotations

import dataclasses
import enum
from typing import Any, Dict, List

from putting_dune_b200 import microscope_utils as mu


class SimulatorEventType(enum.Enum):
  RESET = enum.auto()
  TRANSITION = enum.auto()
  APPLY_CONTROL = enum.auto()
  TAKE_IMAGE = enum.auto()
  GENERATED_IMAGE = enum.auto()


@dataclasses.dataclass(frozen=True)
class SimulatorEvent:
  event_type: SimulatorEventType
  event_data: Dict[str, Any]


class EventObserver(mu.SimulatorObserver):
  """Keeps every observed simulator event; a reset starts a new list."""

  def __init__(self):
    self.grid = None
    self.events: List[SimulatorEvent] = []

  def _record(self, kind: SimulatorEventType, **data) -> None:
    self.events.append(SimulatorEvent(kind, data))

  def observe_reset(self, grid, fov) -> None:
    self.events = []
    self._record(SimulatorEventType.RESET, grid=grid, fov=fov)

  def observe_transition(self, time_since_control_was_applied, grid) -> None:
    self._record(SimulatorEventType.TRANSITION,
                 time_since_control_was_applied=time_since_control_was_applied,
                 grid=grid)

  def observe_apply_control(self, control) -> None:
    self._record(SimulatorEventType.APPLY_CONTROL,
                 dwell_time=control.dwell_time, position=control.position)

  def observe_take_image(self, duration, fov) -> None:
    self._record(SimulatorEventType.TAKE_IMAGE, duration=duration, fov=fov)

  def observe_generated_image(self, image) -> None:
    self._record(SimulatorEventType.GENERATED_IMAGE, image=image)
