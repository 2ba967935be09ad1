"""CPU tests of the trajectory-export row: the wire
codec (proto_wire.py), the value types' to_proto / from_proto_string, the
TFRecord framing and the native host assembler, against

* tests/golden/proto_reference.npz -- bytes produced by the reference's own
  to_proto code running on the official protobuf runtime
  (tests/golden/make_golden.py proto_fixture), and
* the protobuf runtime itself on fresh random messages (oracle/
  pdune_oracle_proto.py builds the classes from the restated .proto).
"""

import ctypes as C
import datetime as dt
import os

import numpy as np
import pytest

from oracle import pdune_oracle_proto as op
from putting_dune_b200 import _native as nat
from putting_dune_b200 import geometry
from putting_dune_b200 import io as pio
from putting_dune_b200 import microscope_utils as mu
from putting_dune_b200 import proto_wire as pw

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden',
                      'proto_reference.npz')


@pytest.fixture(scope='module')
def ref():
  return np.load(GOLDEN)


def _obs(ref, i) -> mu.MicroscopeObservation:
  f = ref[f'fov_{i}']
  return mu.MicroscopeObservation(
      grid=mu.AtomicGrid(ref[f'pos_{i}'], ref[f'num_{i}']),
      fov=mu.MicroscopeFieldOfView(geometry.Point(f[0], f[1]),
                                   geometry.Point(f[2], f[3])),
      controls=tuple(
          mu.BeamControl(geometry.Point(*p),
                         dt.timedelta(microseconds=int(d)))
          for p, d in zip(ref[f'ctl_{i}'], ref[f'dwell_us_{i}'])),
      elapsed_time=dt.timedelta(microseconds=int(ref[f'elapsed_us_{i}'])))


def test_value_types_reproduce_reference_bytes(ref):
  n = int(ref['n_cases'])
  for i in range(n):
    o = _obs(ref, i)
    assert o.grid.to_proto().SerializeToString() == ref[f'grid_bytes_{i}'].tobytes()
    assert o.to_proto().SerializeToString() == ref[f'obs_bytes_{i}'].tobytes()
    # the fixed-size layout the device kernel relies on
    assert len(ref[f'obs_bytes_{i}']) == nat.lib.pd_observation_bytes(
        len(ref[f'num_{i}']), len(ref[f'dwell_us_{i}']))
  traj = mu.Trajectory(tuple(_obs(ref, i) for i in range(n)))
  assert traj.to_proto_string() == ref['trajectory_bytes'].tobytes()
  a, b = _obs(ref, 3), _obs(ref, 4)
  tr = mu.Transition(a.grid, b.grid, a.fov, b.fov, b.controls)
  assert tr.to_proto_string() == ref['transition_bytes'].tobytes()


def test_parse_matches_protobuf_runtime(ref):
  pb = op.build_pb2()
  for i in range(int(ref['sim_n'])):
    raw = ref[f'sim_obs_bytes_{i}'].tobytes()
    want = pb.MicroscopeObservation.FromString(raw)
    got = mu.MicroscopeObservation.from_proto_string(raw)
    assert got.grid.atom_positions.dtype == np.float32  # microscope_utils.py:94
    assert got.grid.atomic_numbers.dtype == np.int32
    np.testing.assert_array_equal(
        got.grid.atom_positions,
        np.asarray([(a.position.x, a.position.y) for a in want.grid.atoms],
                   dtype=np.float32).reshape(-1, 2))
    np.testing.assert_array_equal(
        got.grid.atomic_numbers, [a.atomic_number for a in want.grid.atoms])
    assert got.fov.lower_left.x == want.fov.lower_left_angstroms.x
    assert got.fov.upper_right.y == want.fov.upper_right_angstroms.y
    assert len(got.controls) == len(want.controls)
    for c, w in zip(got.controls, want.controls):
      assert c.position.x == w.position.x
      assert c.dwell_time == dt.timedelta(seconds=w.dwell_time_seconds)
      assert c.voltage_kv == w.voltage_kv and c.current_na == w.current_na
    assert got.elapsed_time == dt.timedelta(seconds=want.elapsed_time_seconds)
    # parse -> serialise is the identity on reference bytes
    assert got.to_proto_string() == raw
  traj = mu.Trajectory.from_proto_string(ref['trajectory_bytes'].tobytes())
  assert len(traj.observations) == int(ref['n_cases'])
  assert traj.to_proto_string() == ref['trajectory_bytes'].tobytes()
  tr = mu.Transition.from_proto_string(ref['transition_bytes'].tobytes())
  assert tr.to_proto_string() == ref['transition_bytes'].tobytes()


def test_codec_against_protobuf_runtime_random_messages():
  pb = op.build_pb2()
  rng = np.random.default_rng(5)
  for trial in range(20):
    m = int(rng.integers(0, 40))
    pos = rng.normal(size=(m, 2)) * 10.0 ** rng.integers(-3, 4)
    z = rng.choice([1, 6, 14, 79, 200, -3], size=m)  # multi-byte varints too
    g = pb.AtomicGrid()
    for i in range(m):
      g.atoms.append(pb.Atom(atomic_number=int(z[i]),
                             position=pb.Point2D(x=pos[i, 0], y=pos[i, 1])))
    mine = pw.atomic_grid(pos, z)
    assert mine == g.SerializeToString()
    back_pos, back_z = pw.parse_atomic_grid(mine)
    np.testing.assert_array_equal(back_z, z)
    np.testing.assert_array_equal(back_pos, pos.astype(np.float32))
    img = (rng.uniform(size=(4, 6)) if trial % 2 else
           rng.integers(0, 3, size=(3, 5)).astype(np.int32))
    label = rng.integers(0, 2, size=(4, 6)).astype(np.uint8)
    ctl = rng.uniform(size=(trial % 3, 2))
    obs = pb.MicroscopeObservation(
        grid=g,
        fov=pb.FieldOfView(lower_left_angstroms=pb.Point2D(x=-1.5, y=2.25),
                           upper_right_angstroms=pb.Point2D(x=20.1, y=23.7)),
        controls=[pb.BeamControl(position=pb.Point2D(x=c[0], y=c[1]),
                                 dwell_time_seconds=1.5, voltage_kv=60,
                                 current_na=0.1) for c in ctl],
        elapsed_time_seconds=trial * 3.5,
        image=op.make_tensor_proto(img),
        label_image=op.make_tensor_proto(label))
    mine = pw.observation(
        pw.atomic_grid(pos, z), pw.field_of_view(-1.5, 2.25, 20.1, 23.7),
        [pw.beam_control(c[0], c[1], 1.5, 60, 0.1) for c in ctl], trial * 3.5,
        img, label)
    assert mine == obs.SerializeToString()
    d = pw.parse_observation(mine)
    np.testing.assert_array_equal(d['image'], img)
    np.testing.assert_array_equal(d['label_image'], label)
    assert d['image'].dtype == img.dtype


def test_crc32c_known_answers():
  # RFC 3720 appendix B.4 / the common check value
  cases = [(b'123456789', 0xE3069283), (bytes(32), 0x8A9136AA),
           (b'\xff' * 32, 0x62A8AB43), (bytes(range(32)), 0x46DD794E),
           (b'', 0)]
  for data, want in cases:
    assert pw.crc32c(data) == want
    assert nat.lib.pd_crc32c(data, len(data)) == want
  rng = np.random.default_rng(1)
  for n in (1, 7, 8, 9, 63, 64, 65, 1000, 4099):
    data = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
    for skew in (0, 1, 3):  # unaligned starts exercise the byte prologue
      buf = bytes(skew) + data
      arr = np.frombuffer(buf, dtype=np.uint8)
      got = nat.lib.pd_crc32c(C.c_void_p(arr.ctypes.data + skew), n)
      assert got == pw.crc32c(data)


def test_tfrecord_framing_round_trip_and_corruption():
  payloads = [b'', b'a', b'hello world' * 50, bytes(range(256)) * 9]
  stream = b''.join(pw.tfrecord_frame(p) for p in payloads)
  assert list(pw.tfrecord_iter(stream)) == payloads
  # length 11 little endian, masked crc of the length, payload, masked crc
  one = pw.tfrecord_frame(b'hello world')
  assert one[:8] == (11).to_bytes(8, 'little') and len(one) == 8 + 4 + 11 + 4
  assert pw.masked_crc(0) == 0xA282EAD8
  bad = bytearray(stream)
  bad[20] ^= 1
  with pytest.raises(ValueError):
    list(pw.tfrecord_iter(bytes(bad)))
  with pytest.raises(ValueError):
    list(pw.tfrecord_iter(stream[:-3]))


def test_write_and_read_records(tmp_path, ref):
  trajs = [mu.Trajectory(tuple(_obs(ref, i) for i in range(k, k + 3)))
           for k in range(4)]
  f = tmp_path / 'run.tfrecords'
  pio.write_records(f, trajs)
  back = list(pio.read_records(f, mu.Trajectory))
  assert [t.to_proto_string() for t in back] == [t.to_proto_string()
                                                 for t in trajs]
  raw = list(pio.read_records(f))
  assert raw[0] == trajs[0].to_proto_string()
  pio.write_records(f, raw)  # io.py:74: plain strings are written as they are
  assert list(pio.read_records(f)) == raw
  with pytest.raises(ValueError):
    pio.write_records(tmp_path / 'run.txt', trajs)
  with pytest.raises(ValueError):
    list(pio.read_records(tmp_path / 'run.array_record'))


def test_native_trajectory_assembler_matches_python(ref):
  """pd_tfrecord_trajectories (host C++) on fake per-step buffers laid out as
  pd_encode_observations leaves them (16-byte aligned records)."""
  n_envs, n_steps = 3, 4
  obs = [[_obs(ref, (t * n_envs + e) % int(ref['n_cases'])).to_proto_string()
          for e in range(n_envs)] for t in range(n_steps)]
  bufs, offs, lens = [], [], []
  for t in range(n_steps):
    off, pos, chunks = [], 0, []
    for e in range(n_envs):
      off.append(pos)
      pad = (-len(obs[t][e])) % 16
      chunks.append(obs[t][e] + bytes(pad))
      pos += len(obs[t][e]) + pad
    bufs.append(np.frombuffer(b''.join(chunks), dtype=np.uint8).copy())
    offs.append(np.asarray(off, dtype=np.int64))
    lens.append(np.asarray([len(o) for o in obs[t]], dtype=np.int32))
  PB, PO, PL = ((C.c_void_p * n_steps)(), (C.c_void_p * n_steps)(),
                (C.c_void_p * n_steps)())
  for t in range(n_steps):
    PB[t], PO[t], PL[t] = (bufs[t].ctypes.data, offs[t].ctypes.data,
                           lens[t].ctypes.data)
  size = C.c_int64()
  nat.check(nat.lib.pd_tfrecord_trajectories(n_steps, n_envs, PB, PO, PL, None,
                                             0, C.byref(size)))
  out = np.empty(size.value, dtype=np.uint8)
  nat.check(nat.lib.pd_tfrecord_trajectories(
      n_steps, n_envs, PB, PO, PL, C.c_void_p(out.ctypes.data), out.size,
      C.byref(size)))
  want = b''.join(
      pw.tfrecord_frame(pw.trajectory([obs[t][e] for t in range(n_steps)]))
      for e in range(n_envs))
  assert out.tobytes() == want
  small = np.empty(10, dtype=np.uint8)
  with pytest.raises(nat.NativeError):
    nat.check(nat.lib.pd_tfrecord_trajectories(
        n_steps, n_envs, PB, PO, PL, C.c_void_p(small.ctypes.data), small.size,
        C.byref(size)))
