"""Host-side geometry value types (reference: putting_dune/geometry.py).

Only the boundary types and the two small array helpers live on the host; the
3-NN query of ``geometry.py:93-111`` is a device-resident table built once by
``pd_build_lattice`` (see ``engine.Lattice``).
"""

from __future__ import annotations

from typing import NewType

import numpy as np


class Point:
  """2-D point with value semantics (stands where the reference uses
  ``shapely.geometry.Point``, geometry.py:26): ``Point(x, y)``,
  ``Point((x, y))`` or ``Point(ndarray[2])``."""

  __slots__ = ('_x', '_y')

  def __init__(self, *args):
    xy = np.asarray(args[0] if len(args) == 1 else args,
                    dtype=np.float64).reshape(-1)
    if xy.size != 2:
      raise ValueError(f'Point takes two coordinates, got {xy.size}')
    object.__setattr__(self, '_x', float(xy[0]))
    object.__setattr__(self, '_y', float(xy[1]))

  def __setattr__(self, name, value):
    raise AttributeError('Point is immutable')

  x = property(lambda self: self._x)
  y = property(lambda self: self._y)

  @property
  def coords(self):
    """[(x, y)] so that ``np.asarray(p.coords)`` has shape (1, 2)."""
    return [(self._x, self._y)]

  def as_array(self) -> np.ndarray:
    return np.array([self._x, self._y], dtype=np.float64)

  def __eq__(self, other):
    return (isinstance(other, Point) and self._x == other._x and
            self._y == other._y)

  def __hash__(self):
    return hash((self._x, self._y))

  def __repr__(self):
    return f'POINT ({self._x} {self._y})'


PointMicroscopeFrame = NewType('PointMicroscopeFrame', Point)
PointMaterialFrame = NewType('PointMaterialFrame', Point)


def get_angles(coordinates: np.ndarray) -> np.ndarray:
  """Angle of each row from +x, counter-clockwise (geometry.py:33-48)."""
  coordinates = np.asarray(coordinates)
  return np.arctan2(coordinates[:, 1], coordinates[:, 0])


def rotate_coordinates(coord: np.ndarray, theta: float) -> np.ndarray:
  """Rotates rows of ``coord`` by ``theta`` counter-clockwise
  (geometry.py:51-66; right-multiplication by [[c, s], [-s, c]])."""
  c, s = np.cos(theta), np.sin(theta)
  return np.asarray(coord) @ np.asarray([[c, s], [-s, c]])
