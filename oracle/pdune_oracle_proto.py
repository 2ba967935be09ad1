"""TEST INFRASTRUCTURE -- oracle for the trajectory-export row (SURVEY.md
section 8(f)4).  Only tests/, tests/golden/make_golden.py and refshim.py may
import it.

The reference serialises its value types through `putting_dune_pb2`, the
module protoc generates from putting_dune/putting_dune.proto; the generated
module is not in the tree and protoc is not in this image.  This file restates
the .proto (putting_dune.proto:7-62, field for field) as a FileDescriptorProto
and lets the *official protobuf runtime* (google.protobuf, in the image) build
the message classes from it -- the same classes protoc's output would create.
`refshim.load_reference()` installs them as `putting_dune.putting_dune_pb2`,
so the reference's own `to_proto` / `from_proto` code (microscope_utils.py)
runs unmodified against the real runtime; the bytes it produces are committed
as tests/golden/proto_reference.npz and pin both encoders of the product
(proto_wire.py and the device kernel).

`tensorflow.TensorProto` (images inside observations): TensorFlow is absent,
so tensor.proto / tensor_shape.proto / types.proto are restated from their
published definitions (the fields tf.make_tensor_proto fills for an ndarray:
dtype = 1, tensor_shape = 2, tensor_content = 4) and `make_tensor_proto`
below stands in for tf.make_tensor_proto: **parity unpinned** for that one
field.
"""

from __future__ import annotations

import types

import numpy as np
from google.protobuf import descriptor_pb2
from google.protobuf import descriptor_pool
from google.protobuf import message_factory

_F = descriptor_pb2.FieldDescriptorProto
_OPT, _REP = _F.LABEL_OPTIONAL, _F.LABEL_REPEATED


def _field(msg, name, number, ftype, label=_OPT, type_name=None):
  f = msg.field.add()
  f.name, f.number, f.type, f.label = name, number, ftype, label
  if type_name:
    f.type_name = type_name


def _tensorflow_file() -> descriptor_pb2.FileDescriptorProto:
  fd = descriptor_pb2.FileDescriptorProto()
  fd.name = 'tensorflow/core/framework/tensor.proto'
  fd.package = 'tensorflow'
  fd.syntax = 'proto3'
  shape = fd.message_type.add()
  shape.name = 'TensorShapeProto'
  dim = shape.nested_type.add()
  dim.name = 'Dim'
  _field(dim, 'size', 1, _F.TYPE_INT64)
  _field(dim, 'name', 2, _F.TYPE_STRING)
  _field(shape, 'dim', 2, _F.TYPE_MESSAGE, _REP,
         '.tensorflow.TensorShapeProto.Dim')
  _field(shape, 'unknown_rank', 3, _F.TYPE_BOOL)
  tensor = fd.message_type.add()
  tensor.name = 'TensorProto'
  _field(tensor, 'dtype', 1, _F.TYPE_INT32)  # enum DataType on the wire
  _field(tensor, 'tensor_shape', 2, _F.TYPE_MESSAGE, _OPT,
         '.tensorflow.TensorShapeProto')
  _field(tensor, 'version_number', 3, _F.TYPE_INT32)
  _field(tensor, 'tensor_content', 4, _F.TYPE_BYTES)
  return fd


def _putting_dune_file() -> descriptor_pb2.FileDescriptorProto:
  """putting_dune.proto:1-72."""
  fd = descriptor_pb2.FileDescriptorProto()
  fd.name = 'putting_dune/putting_dune.proto'
  fd.package = 'putting_dune.google'
  fd.syntax = 'proto2'
  fd.dependency.append('tensorflow/core/framework/tensor.proto')
  pkg = '.putting_dune.google.'
  tensor = '.tensorflow.TensorProto'

  def message(name):
    m = fd.message_type.add()
    m.name = name
    return m

  m = message('Point2D')  # :7-10
  _field(m, 'x', 1, _F.TYPE_FLOAT)
  _field(m, 'y', 2, _F.TYPE_FLOAT)
  m = message('Atom')  # :12-16
  _field(m, 'atomic_number', 1, _F.TYPE_INT32)
  _field(m, 'position', 2, _F.TYPE_MESSAGE, _OPT, pkg + 'Point2D')
  m = message('AtomicGrid')  # :18-20
  _field(m, 'atoms', 1, _F.TYPE_MESSAGE, _REP, pkg + 'Atom')
  m = message('BeamControl')  # :22-27
  _field(m, 'position', 1, _F.TYPE_MESSAGE, _OPT, pkg + 'Point2D')
  _field(m, 'dwell_time_seconds', 2, _F.TYPE_FLOAT)
  _field(m, 'voltage_kv', 3, _F.TYPE_FLOAT)
  _field(m, 'current_na', 4, _F.TYPE_FLOAT)
  m = message('FieldOfView')  # :29-32
  _field(m, 'lower_left_angstroms', 1, _F.TYPE_MESSAGE, _OPT, pkg + 'Point2D')
  _field(m, 'upper_right_angstroms', 2, _F.TYPE_MESSAGE, _OPT,
         pkg + 'Point2D')
  m = message('MicroscopeObservation')  # :34-42
  _field(m, 'grid', 1, _F.TYPE_MESSAGE, _OPT, pkg + 'AtomicGrid')
  _field(m, 'fov', 2, _F.TYPE_MESSAGE, _OPT, pkg + 'FieldOfView')
  _field(m, 'controls', 3, _F.TYPE_MESSAGE, _REP, pkg + 'BeamControl')
  _field(m, 'elapsed_time_seconds', 4, _F.TYPE_FLOAT)
  _field(m, 'image', 5, _F.TYPE_MESSAGE, _OPT, tensor)
  _field(m, 'label_image', 6, _F.TYPE_MESSAGE, _OPT, tensor)
  m = message('Trajectory')  # :44-46
  _field(m, 'observations', 1, _F.TYPE_MESSAGE, _REP,
         pkg + 'MicroscopeObservation')
  m = message('Transition')  # :48-62
  _field(m, 'grid_before', 1, _F.TYPE_MESSAGE, _OPT, pkg + 'AtomicGrid')
  _field(m, 'grid_after', 2, _F.TYPE_MESSAGE, _OPT, pkg + 'AtomicGrid')
  _field(m, 'fov_before', 3, _F.TYPE_MESSAGE, _OPT, pkg + 'FieldOfView')
  _field(m, 'fov_after', 4, _F.TYPE_MESSAGE, _OPT, pkg + 'FieldOfView')
  _field(m, 'controls', 5, _F.TYPE_MESSAGE, _REP, pkg + 'BeamControl')
  _field(m, 'image_before', 6, _F.TYPE_MESSAGE, _OPT, tensor)
  _field(m, 'image_after', 7, _F.TYPE_MESSAGE, _OPT, tensor)
  _field(m, 'label_image_before', 8, _F.TYPE_MESSAGE, _OPT, tensor)
  _field(m, 'label_image_after', 9, _F.TYPE_MESSAGE, _OPT, tensor)
  m = message('Drift')  # :64-67
  _field(m, 'jitter', 1, _F.TYPE_MESSAGE, _REP, pkg + 'Point2D')
  _field(m, 'drift', 2, _F.TYPE_MESSAGE, _OPT, pkg + 'Point2D')
  m = message('LabeledAlignmentTrajectory')  # :69-72
  _field(m, 'trajectory', 1, _F.TYPE_MESSAGE, _OPT, pkg + 'Trajectory')
  _field(m, 'drifts', 2, _F.TYPE_MESSAGE, _REP, pkg + 'Drift')
  return fd


_cache = None


def build_pb2() -> types.ModuleType:
  """A module with the message classes of putting_dune.proto (what the
  generated putting_dune_pb2 exposes) plus `TensorProto`."""
  global _cache
  if _cache is not None:
    return _cache
  pool = descriptor_pool.DescriptorPool()
  pool.Add(_tensorflow_file())
  pool.Add(_putting_dune_file())
  mod = types.ModuleType('putting_dune.putting_dune_pb2')
  for name in ('Point2D', 'Atom', 'AtomicGrid', 'BeamControl', 'FieldOfView',
               'MicroscopeObservation', 'Trajectory', 'Transition', 'Drift',
               'LabeledAlignmentTrajectory'):
    desc = pool.FindMessageTypeByName('putting_dune.google.' + name)
    setattr(mod, name, message_factory.GetMessageClass(desc))
  mod.TensorProto = message_factory.GetMessageClass(
      pool.FindMessageTypeByName('tensorflow.TensorProto'))
  _cache = mod
  return mod


_TF_DTYPES = {np.dtype('float32'): 1, np.dtype('float64'): 2,
              np.dtype('int32'): 3, np.dtype('uint8'): 4,
              np.dtype('int64'): 9, np.dtype('bool'): 10}


def make_tensor_proto(array):
  """Stand-in for tf.make_tensor_proto(ndarray): dtype, shape, raw content."""
  a = np.ascontiguousarray(array)
  t = build_pb2().TensorProto()
  t.dtype = _TF_DTYPES[a.dtype]
  t.tensor_shape.SetInParent()
  for d in a.shape:
    t.tensor_shape.dim.add().size = int(d)
  t.tensor_content = a.tobytes()
  return t


def make_ndarray(t):
  inv = {v: k for k, v in _TF_DTYPES.items()}
  shape = [d.size for d in t.tensor_shape.dim]
  return np.frombuffer(t.tensor_content, dtype=inv[t.dtype]).reshape(shape)
