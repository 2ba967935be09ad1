"""Protobuf wire codec for the messages of putting_dune.proto (reference:
putting_dune/putting_dune.proto:7-62) without a generated module or
TensorFlow: the reference serialises its value types through
`putting_dune_pb2` and `tf.make_tensor_proto` (microscope_utils.py:46-757);
this module writes / reads the same bytes.

All messages are proto2 with `optional` fields that the reference always sets,
so every scalar field is written (explicit presence), in field-number order --
which is also what the protobuf runtime emits.  Field tables:

  Point2D                x=1 float, y=2 float
  Atom                   atomic_number=1 int32, position=2 Point2D
  AtomicGrid             atoms=1 repeated Atom
  BeamControl            position=1 Point2D, dwell_time_seconds=2 float,
                         voltage_kv=3 float, current_na=4 float
  FieldOfView            lower_left_angstroms=1, upper_right_angstroms=2
  MicroscopeObservation  grid=1, fov=2, controls=3 repeated,
                         elapsed_time_seconds=4 float, image=5 TensorProto,
                         label_image=6 TensorProto
  Trajectory             observations=1 repeated
  Transition             grid_before=1, grid_after=2, fov_before=3,
                         fov_after=4, controls=5 repeated, image_before=6,
                         image_after=7, label_image_before=8,
                         label_image_after=9
  tensorflow.TensorProto dtype=1 enum, tensor_shape=2 {dim=2 {size=1 int64}},
                         tensor_content=4 bytes  (what tf.make_tensor_proto
                         writes for an ndarray; TensorFlow is absent here, so
                         this one table follows the published tensor.proto and
                         is not pinned by a reference run)
"""

from __future__ import annotations

import struct
from typing import Iterator, List, Optional, Tuple

import numpy as np

_VARINT, _I64, _LEN, _I32 = 0, 1, 2, 5

# tensorflow/core/framework/types.proto
_TF_DTYPES = {np.dtype('float32'): 1, np.dtype('float64'): 2,
              np.dtype('int32'): 3, np.dtype('uint8'): 4,
              np.dtype('int16'): 5, np.dtype('int8'): 6,
              np.dtype('int64'): 9, np.dtype('bool'): 10}
_TF_DTYPES_INV = {v: k for k, v in _TF_DTYPES.items()}


def varint(v: int) -> bytes:
  if v < 0:
    v += 1 << 64  # int32 / int64 fields: two's complement, ten bytes
  out = bytearray()
  while v >= 0x80:
    out.append((v & 0x7F) | 0x80)
    v >>= 7
  out.append(v)
  return bytes(out)


def _tag(field: int, wire: int) -> bytes:
  return varint((field << 3) | wire)


def _f32(field: int, v: float) -> bytes:
  # float64 -> float32: round to nearest even (what the runtime's cast does)
  return _tag(field, _I32) + np.float32(v).tobytes()


def _msg(field: int, payload: bytes) -> bytes:
  return _tag(field, _LEN) + varint(len(payload)) + payload


def point(x: float, y: float) -> bytes:
  return _f32(1, x) + _f32(2, y)


def atomic_grid(positions: np.ndarray, numbers: np.ndarray) -> bytes:
  """AtomicGrid.to_proto (microscope_utils.py:105-122), vectorised: every
  atom is the same 16 bytes `0A 0E 08 Z 12 0A 0D x 15 y`."""
  m = int(np.asarray(numbers).shape[0])
  z = np.asarray(numbers).astype(np.int64)
  if m and (z.min() < 0 or z.max() > 127):
    return b''.join(
        _msg(1, _tag(1, _VARINT) + varint(int(z[i])) +
             _msg(2, point(*np.asarray(positions)[i])))
        for i in range(m))
  rec = np.zeros((m, 16), dtype=np.uint8)
  rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3] = 0x0A, 0x0E, 0x08, z
  rec[:, 4], rec[:, 5], rec[:, 6], rec[:, 11] = 0x12, 0x0A, 0x0D, 0x15
  xy = np.ascontiguousarray(np.asarray(positions, dtype=np.float64)
                            .reshape(m, 2).astype('<f4'))
  rec[:, 7:11] = xy[:, 0:1].view(np.uint8).reshape(m, 4)
  rec[:, 12:16] = xy[:, 1:2].view(np.uint8).reshape(m, 4)
  return rec.tobytes()


def beam_control(x, y, dwell_seconds, voltage_kv, current_na) -> bytes:
  return (_msg(1, point(x, y)) + _f32(2, dwell_seconds) +
          _f32(3, voltage_kv) + _f32(4, current_na))


def field_of_view(llx, lly, urx, ury) -> bytes:
  return _msg(1, point(llx, lly)) + _msg(2, point(urx, ury))


def tensor(array: np.ndarray) -> bytes:
  a = np.ascontiguousarray(array)
  if a.dtype not in _TF_DTYPES:
    raise TypeError(f'no TensorProto dtype for {a.dtype}')
  shape = b''.join(_msg(2, _tag(1, _VARINT) + varint(int(d)))
                   for d in a.shape)
  return (_tag(1, _VARINT) + varint(_TF_DTYPES[a.dtype]) + _msg(2, shape) +
          _msg(4, a.astype(a.dtype.newbyteorder('<')).tobytes()))


def observation(grid: bytes, fov: bytes, controls: List[bytes],
                elapsed_seconds: float, image: Optional[np.ndarray] = None,
                label_image: Optional[np.ndarray] = None) -> bytes:
  out = _msg(1, grid) + _msg(2, fov)
  out += b''.join(_msg(3, c) for c in controls)
  out += _f32(4, elapsed_seconds)
  if image is not None:
    out += _msg(5, tensor(image))
  if label_image is not None:
    out += _msg(6, tensor(label_image))
  return out


def trajectory(observations: List[bytes]) -> bytes:
  return b''.join(_msg(1, o) for o in observations)


def transition(grid_before: bytes, grid_after: bytes, fov_before: bytes,
               fov_after: bytes, controls: List[bytes], image_before=None,
               image_after=None, label_image_before=None,
               label_image_after=None) -> bytes:
  out = (_msg(1, grid_before) + _msg(2, grid_after) + _msg(3, fov_before) +
         _msg(4, fov_after) + b''.join(_msg(5, c) for c in controls))
  for field, img in ((6, image_before), (7, image_after),
                     (8, label_image_before), (9, label_image_after)):
    if img is not None:
      out += _msg(field, tensor(img))
  return out


# ---- decoding --------------------------------------------------------------
def _read_varint(buf: bytes, pos: int) -> Tuple[int, int]:
  shift = v = 0
  while True:
    b = buf[pos]
    pos += 1
    v |= (b & 0x7F) << shift
    if b < 0x80:
      return v, pos
    shift += 7
    if shift > 63:
      raise ValueError('varint too long')


def fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
  """Yields (field number, wire type, value) of one message."""
  pos, n = 0, len(buf)
  while pos < n:
    key, pos = _read_varint(buf, pos)
    field, wire = key >> 3, key & 7
    if wire == _VARINT:
      v, pos = _read_varint(buf, pos)
    elif wire == _I32:
      v, pos = buf[pos:pos + 4], pos + 4
    elif wire == _I64:
      v, pos = buf[pos:pos + 8], pos + 8
    elif wire == _LEN:
      ln, pos = _read_varint(buf, pos)
      v, pos = buf[pos:pos + ln], pos + ln
      if len(v) != ln:
        raise ValueError('truncated message')
    else:
      raise ValueError(f'unsupported wire type {wire}')
    yield field, wire, v


def _float(v) -> float:
  return struct.unpack('<f', v)[0]


def parse_point(buf: bytes) -> Tuple[float, float]:
  x = y = 0.0
  for f, _, v in fields(buf):
    if f == 1:
      x = _float(v)
    elif f == 2:
      y = _float(v)
  return x, y


def parse_atomic_grid(buf: bytes) -> Tuple[np.ndarray, np.ndarray]:
  """AtomicGrid.from_proto (microscope_utils.py:90-103): float32 positions,
  int32 numbers."""
  pos, num = [], []
  for f, _, atom in fields(buf):
    if f != 1:
      continue
    z, xy = 0, (0.0, 0.0)
    for g, _, v in fields(atom):
      if g == 1:
        z = v - (1 << 64) if v >= (1 << 63) else v
      elif g == 2:
        xy = parse_point(v)
    pos.append(xy)
    num.append(z)
  return (np.asarray(pos, dtype=np.float32).reshape(-1, 2),
          np.asarray(num, dtype=np.int32))


def parse_beam_control(buf: bytes) -> dict:
  out = {'position': (0.0, 0.0), 'dwell_time_seconds': 0.0, 'voltage_kv': 0.0,
         'current_na': 0.0}
  names = {2: 'dwell_time_seconds', 3: 'voltage_kv', 4: 'current_na'}
  for f, _, v in fields(buf):
    if f == 1:
      out['position'] = parse_point(v)
    elif f in names:
      out[names[f]] = _float(v)
  return out


def parse_field_of_view(buf: bytes) -> Tuple[Tuple[float, float],
                                              Tuple[float, float]]:
  ll = ur = (0.0, 0.0)
  for f, _, v in fields(buf):
    if f == 1:
      ll = parse_point(v)
    elif f == 2:
      ur = parse_point(v)
  return ll, ur


def parse_tensor(buf: bytes) -> Optional[np.ndarray]:
  dtype, shape, content = 0, [], b''
  for f, _, v in fields(buf):
    if f == 1:
      dtype = v
    elif f == 2:
      for g, _, dim in fields(v):
        if g == 2:
          size = 0
          for h, _, s in fields(dim):
            if h == 1:
              size = s
          shape.append(size)
    elif f == 4:
      content = v
  if dtype == 0:  # microscope_utils.py:568: dtype 0 means "no image"
    return None
  dt = _TF_DTYPES_INV[dtype].newbyteorder('<')
  return np.frombuffer(content, dtype=dt).reshape(shape).astype(
      _TF_DTYPES_INV[dtype])


def parse_observation(buf: bytes) -> dict:
  out = {'grid': (np.zeros((0, 2), np.float32), np.zeros(0, np.int32)),
         'fov': ((0.0, 0.0), (0.0, 0.0)), 'controls': [],
         'elapsed_time_seconds': 0.0, 'image': None, 'label_image': None}
  for f, _, v in fields(buf):
    if f == 1:
      out['grid'] = parse_atomic_grid(v)
    elif f == 2:
      out['fov'] = parse_field_of_view(v)
    elif f == 3:
      out['controls'].append(parse_beam_control(v))
    elif f == 4:
      out['elapsed_time_seconds'] = _float(v)
    elif f == 5:
      out['image'] = parse_tensor(v)
    elif f == 6:
      out['label_image'] = parse_tensor(v)
  return out


def parse_trajectory(buf: bytes) -> List[dict]:
  return [parse_observation(v) for f, _, v in fields(buf) if f == 1]


def parse_transition(buf: bytes) -> dict:
  out = {'controls': [], 'image_before': None, 'image_after': None,
         'label_image_before': None, 'label_image_after': None}
  names = {6: 'image_before', 7: 'image_after', 8: 'label_image_before',
           9: 'label_image_after'}
  for f, _, v in fields(buf):
    if f in (1, 2):
      out['grid_before' if f == 1 else 'grid_after'] = parse_atomic_grid(v)
    elif f in (3, 4):
      out['fov_before' if f == 3 else 'fov_after'] = parse_field_of_view(v)
    elif f == 5:
      out['controls'].append(parse_beam_control(v))
    elif f in names:
      out[names[f]] = parse_tensor(v)
  return out


# ---- TFRecord framing (io.py:30-82 uses tf.io.TFRecordWriter / tf.data) -----
def _crc_table():
  t = []
  for i in range(256):
    c = i
    for _ in range(8):
      c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
    t.append(c)
  return t


_CRC = _crc_table()


def crc32c(data: bytes) -> int:
  """CRC-32C (Castagnoli); the native library's pd_crc32c is the fast one."""
  c = 0xFFFFFFFF
  for b in data:
    c = (c >> 8) ^ _CRC[(c ^ b) & 0xFF]
  return c ^ 0xFFFFFFFF


def masked_crc(crc: int) -> int:
  return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def tfrecord_frame(payload: bytes, crc=crc32c) -> bytes:
  head = struct.pack('<Q', len(payload))
  return (head + struct.pack('<I', masked_crc(crc(head))) + payload +
          struct.pack('<I', masked_crc(crc(payload))))


def tfrecord_iter(stream: bytes, crc=crc32c, verify=True) -> Iterator[bytes]:
  pos, n = 0, len(stream)
  while pos < n:
    if pos + 12 > n:
      raise ValueError('truncated TFRecord header')
    head = stream[pos:pos + 8]
    (length,) = struct.unpack('<Q', head)
    (hcrc,) = struct.unpack('<I', stream[pos + 8:pos + 12])
    if verify and hcrc != masked_crc(crc(head)):
      raise ValueError('corrupted TFRecord length')
    data = stream[pos + 12:pos + 12 + length]
    if len(data) != length or pos + 16 + length > n:
      raise ValueError('truncated TFRecord')
    (dcrc,) = struct.unpack('<I', stream[pos + 12 + length:pos + 16 + length])
    if verify and dcrc != masked_crc(crc(data)):
      raise ValueError('corrupted TFRecord payload')
    yield data
    pos += 16 + length
