"""Record IO with the reference's interface (putting_dune/io.py:30-82):
`.tfrecords` files of serialised ProtoModel records, written and read without
TensorFlow (proto_wire.py holds the TFRecord framing; the native library's
pd_crc32c does the checksums), plus the batched exporter that turns device
rollouts into per-env Trajectory records.
"""

from __future__ import annotations

import ctypes as C
import os
import pathlib
from typing import Iterable, Iterator, List, Optional, Type

import numpy as np
import torch

from putting_dune_b200 import _native as nat
from putting_dune_b200 import microscope_utils
from putting_dune_b200 import proto_wire as pw

PathLike = os.PathLike | str


def _crc(data: bytes) -> int:
  return nat.lib.pd_crc32c(data, len(data))


def read_records(file: PathLike,
                 record_type: Optional[Type[microscope_utils.ProtoModel]] = None
                 ) -> Iterator:
  """io.py:46-63: yields the records of a .tfrecords file, parsed into
  `record_type` when one is given."""
  file = pathlib.Path(file)
  if file.suffix != '.tfrecords':
    raise ValueError(f'File {file} has unknown extension {file.suffix}')
  stream = file.read_bytes()
  for record in pw.tfrecord_iter(stream, crc=_crc):
    if record_type and issubclass(record_type, microscope_utils.ProtoModel):
      yield record_type.from_proto_string(record)
    else:
      yield record


def write_records(file: PathLike, records: Iterable) -> None:
  """io.py:65-82: writes serialised records (bytes or ProtoModel)."""
  file = pathlib.Path(file)
  if file.suffix != '.tfrecords':
    raise ValueError(f'File {file} has unknown extension {file.suffix}')
  with open(file, 'wb') as f:
    for record in records:
      if isinstance(record, microscope_utils.ProtoModel):
        record = record.to_proto().SerializeToString()
      f.write(pw.tfrecord_frame(bytes(record), crc=_crc))


class TrajectoryRecorder:
  """Collects the observation of every env after each step of an `EnvBatch`
  (serialised on the device by pd_encode_observations) and frames one
  Trajectory record per env -- what the reference obtains by appending each
  step's MicroscopeObservation to a list and calling
  Trajectory(...).to_proto() (microscope_utils.py:737-757, io.py:65-82).

    rec = TrajectoryRecorder(batch)
    rec.record(None, 0, elapsed_us=image_duration_us)     # after reset
    out = batch.step_and_image(ctl, dwell, rate)
    rec.record(ctl, dwell, elapsed_us=out.elapsed_us)
    rec.write('run.tfrecords')

  ``elapsed_us`` is the observation's ``elapsed_time``: per step in the
  reference (simulator.py:100-105,131-182: the image duration after reset,
  dwell + image time(s) after a step), so it has to be passed -- the
  ``StepResult.elapsed_us`` of the step, an int for all envs, or 'sim_time'
  for the cumulative simulated clock of the batch.
  """

  def __init__(self, batch, voltage_kv: float = 60.0, current_na: float = 0.1,
               max_atoms: Optional[int] = None):
    self.batch = batch
    self.voltage_kv, self.current_na = voltage_kv, current_na
    self.max_atoms = max_atoms
    self._steps: List[tuple] = []

  def record(self, controls_xy=None, dwell_us=0, elapsed_us=None) -> None:
    b = self.batch
    n, dev = b.num_envs, b.device
    if elapsed_us is None:
      raise ValueError(
          "elapsed_us is required: the step's StepResult.elapsed_us, the "
          "image duration (int, microseconds) for the observation after "
          "reset, or 'sim_time' for the batch's cumulative clock")
    if isinstance(elapsed_us, str):
      if elapsed_us != 'sim_time':
        raise ValueError(f'unknown elapsed_us {elapsed_us!r}')
      elapsed_us = None  # the encoder reads st.sim_time_us
    elif isinstance(elapsed_us, (int, np.integer)):
      elapsed_us = torch.full((n,), int(elapsed_us), dtype=torch.int64,
                              device=dev)
    if controls_xy is None:
      ctl = torch.zeros((n, 0, 2), dtype=torch.float64, device=dev)
    else:
      ctl = torch.as_tensor(controls_xy, dtype=torch.float64,
                            device=dev).reshape(n, -1, 2).contiguous()
    n_controls = ctl.shape[1]
    if isinstance(dwell_us, (int, np.integer)):
      d, scalar = None, int(dwell_us)
    else:
      d = torch.as_tensor(dwell_us, dtype=torch.int64,
                          device=dev).reshape(n, n_controls).contiguous()
      scalar = 0
    el = None if elapsed_us is None else torch.as_tensor(
        elapsed_us, dtype=torch.int64, device=dev).reshape(n).contiguous()
    max_atoms = self.max_atoms or min(b.max_atoms_in_view(),
                                      b.lattice_tables.n_sites)
    slot = (int(nat.lib.pd_observation_bytes(max_atoms, n_controls)) + 15) & ~15
    out = torch.empty(n * slot, dtype=torch.uint8, device=dev)
    offsets = torch.empty(n + 1, dtype=torch.int64, device=dev)
    length = torch.empty(n, dtype=torch.int32, device=dev)
    atoms = torch.empty(n, dtype=torch.int32, device=dev)
    overflow = torch.zeros(n, dtype=torch.uint8, device=dev)
    P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):
      nat.check(nat.lib.pd_encode_observations(
          C.byref(b.lattice_tables.c), C.byref(b.c), P(ctl), P(d), scalar,
          n_controls, P(el), self.voltage_kv, self.current_na, max_atoms,
          P(out), out.numel(), P(offsets), P(length), P(atoms), P(overflow),
          C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    h_off = offsets.cpu()
    if int(overflow.max().item()) if n else 0:
      raise RuntimeError('an observation exceeded max_atoms '
                         f'({max_atoms}); pass a larger TrajectoryRecorder'
                         '(max_atoms=...)')
    used = int(h_off[n].item())
    self._steps.append((out[:used].cpu().numpy(), h_off.numpy()[:n].copy(),
                        length.cpu().numpy(), atoms.cpu().numpy()))

  @property
  def num_steps(self) -> int:
    return len(self._steps)

  def to_tfrecord_bytes(self) -> bytes:
    t, n = len(self._steps), self.batch.num_envs
    PB, PO, PL = (C.c_void_p * max(t, 1))(), (C.c_void_p * max(t, 1))(), \
        (C.c_void_p * max(t, 1))()
    for i, (by, off, ln, _) in enumerate(self._steps):
      PB[i], PO[i], PL[i] = (by.ctypes.data, off.ctypes.data, ln.ctypes.data)
    size = C.c_int64(0)
    nat.check(nat.lib.pd_tfrecord_trajectories(
        t, n, PB, PO, PL, None, 0, C.byref(size)))
    out = np.empty(size.value, dtype=np.uint8)
    nat.check(nat.lib.pd_tfrecord_trajectories(
        t, n, PB, PO, PL, C.c_void_p(out.ctypes.data), out.size,
        C.byref(size)))
    return out.tobytes()

  def write(self, file: PathLike) -> None:
    file = pathlib.Path(file)
    if file.suffix != '.tfrecords':
      raise ValueError(f'File {file} has unknown extension {file.suffix}')
    file.write_bytes(self.to_tfrecord_bytes())
