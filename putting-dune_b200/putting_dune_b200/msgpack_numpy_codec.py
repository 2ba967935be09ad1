"""The msgpack-numpy wire encoding of NumPy values, restated.

The reference serialises ``GaussianMixtureRateFunction`` with ``import
msgpack_numpy as msgpack`` (putting_dune/graphene.py:28,392-427).  That
package (an unpinned third-party dependency, pyproject.toml:17-33) is not in
this image, so its published encoding (msgpack-numpy >= 0.4.4) is restated on
top of the ``msgpack`` package that is:

  ndarray     -> {b'nd': True, b'type': dtype.str (or dtype.descr for
                  structured arrays, b'kind': b'V'), b'kind': b'',
                  b'shape': shape, b'data': C-order bytes}
  NumPy scalar -> {b'nd': False, b'type': dtype.str, b'data': bytes}
  complex      -> {b'complex': True, b'data': repr}

Files written here are readable by the reference and vice versa.  Byte parity
with the real package is unpinned (it cannot be imported here).
"""

from __future__ import annotations

import msgpack
import numpy as np


def _encode(obj):
  if isinstance(obj, np.ndarray):
    if obj.dtype.kind == 'V':
      kind, descr = b'V', obj.dtype.descr
    else:
      kind, descr = b'', obj.dtype.str
    return {b'nd': True, b'type': descr, b'kind': kind,
            b'shape': obj.shape, b'data': np.ascontiguousarray(obj).tobytes()}
  if isinstance(obj, (np.bool_, np.number)):
    return {b'nd': False, b'type': obj.dtype.str, b'data': obj.tobytes()}
  if isinstance(obj, complex):
    return {b'complex': True, b'data': repr(obj)}
  return obj


def _dtype(descr):
  if isinstance(descr, (list, tuple)):
    return np.dtype([tuple(t.decode() if isinstance(t, bytes) else t
                           for t in d) for d in descr])
  return np.dtype(descr.decode() if isinstance(descr, bytes) else descr)


def _decode(obj):
  if b'nd' in obj:
    if obj[b'nd'] is True:
      return np.ndarray(buffer=bytearray(obj[b'data']),
                        dtype=_dtype(obj[b'type']), shape=obj[b'shape'])
    return np.frombuffer(obj[b'data'], dtype=_dtype(obj[b'type']))[0]
  if b'complex' in obj:
    return complex(obj[b'data'].decode() if isinstance(obj[b'data'], bytes)
                   else obj[b'data'])
  return obj


def packb(obj) -> bytes:
  return msgpack.packb(obj, default=_encode, use_bin_type=True)


def unpackb(data: bytes):
  return msgpack.unpackb(data, object_hook=_decode, raw=False,
                         strict_map_key=False)
