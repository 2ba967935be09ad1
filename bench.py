#!/usr/bin/env python
"""Benchmark of the batched putting-dune simulator hot path on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's arm
  python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm

A *step* is one pass of the hot path over one batch of synthetic input: a
beam-action stream of ``--beam-steps`` controls for each of ``--envs``
environments per GPU (BASELINE.json configs[1]: 4096 batched envs, prior
rates, event sampling only, no rendering), i.e. envs * beam_steps simulated
env-steps per launch.  Environments shard across GPUs by global env id with no
data-path collective (weak scaling: per-GPU work fixed).

Printed line (rank 0): see DESIGN.md "Measurement".  `value` is device-timed
whole-job env-steps/s with the action stream resident in HBM and the per-step
observation (Si site, elapsed microseconds) written for every env-step; `e2e`
is the same metric through the host-buffer C-ABI call (pinned host actions
copied in, per-step results copied out, every step); `roofline` uses SURVEY.md
section 8(d)'s 64 algorithmic bytes per env-step against the measured HBM copy
bandwidth in MEASURED_PEAKS.json and carries the issue-slot figure beside it;
`cpu_baseline` is the UNMODIFIED reference simulator (baseline/_ref) timed on
this box's host cores, with the NumPy port of the oracle next to it.
"""

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'putting-dune_b200')):
  if _p not in sys.path:
    sys.path.insert(0, _p)

import numpy as np  # noqa: E402

ALGORITHMIC_BYTES_PER_ENV_STEP = 64
DWELL_US = 1500000  # registry.py:263-266 relative_random: dwell 1.5 s
IMAGE_US = 2000000  # simulator.py:37
METRIC = 'simulated env-steps/sec'
UNIT = 'env-steps/s'


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=20)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--workload', default='config2',
                  choices=['config2', 'config5'])
  ap.add_argument('--envs', type=int, default=None,
                  help='envs per GPU (default: 4096 config2, 1Mi/N config5)')
  ap.add_argument('--beam-steps', type=int, default=None,
                  help='controls per env per launch')
  ap.add_argument('--rate', default='prior', choices=['prior', 'simple'])
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-at-scale', action='store_true')
  ap.add_argument('--no-frames', action='store_true')
  ap.add_argument('--no-mlp', action='store_true')
  ap.add_argument('--no-export', action='store_true')
  ap.add_argument('--frames', type=int, default=16384,
                  help='frames per render launch (512x512; BASELINE '
                       'configs[3]: 16384 envs per step = 17.2 GB of frames)')
  ap.add_argument('--episodes', type=int, default=1 << 20,
                  help='total envs of the greedy-controller episode run, '
                       'sharded over the GPUs, stats all-gathered (BASELINE '
                       'configs[4]: 1048576); 0 = skip')
  ap.add_argument('--cpu-seconds', type=float, default=12.0)
  return ap.parse_args()


def workload_config(args):
  n = max(1, args.gpus)
  if args.workload == 'config2':
    envs = args.envs or 4096
    beam_steps = args.beam_steps or 256
    name = ('configs[1]: 4096 batched envs, prior rates, event sampling only '
            '(no rendering)')
    scaling = 'weak'
  else:
    envs = args.envs or (1 << 20) // n
    beam_steps = args.beam_steps or 8
    name = 'configs[4]-sized: 1Mi envs sharded over the GPUs, open-loop beam'
    scaling = 'weak' if args.envs else 'strong'
  return dict(workload=name, envs_per_gpu=envs, beam_steps_per_launch=beam_steps,
              rate_function=args.rate, dwell_s=DWELL_US / 1e6,
              image_duration_s=IMAGE_US / 1e6, grid_columns=50,
              sharding=f'env-index x{n}', l2='flushed between timed steps',
              scaling=scaling)


def synthetic_controls(n_envs, beam_steps, seed):
  """The `relative_random` workload (experiments/registry.py:263-266): agent
  actions U(-1, 1)^2, turned into beam positions within one bond length of the
  Si by RelativeToSiliconActionAdapter (action_adapters.py:131-216).  The
  adapter runs on the device (pd_rollout_actions); the CPU arm applies the
  oracle's restatement of the same adapter."""
  rng = np.random.default_rng(seed)
  return rng.uniform(-1.0, 1.0, size=(beam_steps, n_envs, 2))


def direct_controls(n_envs, beam_steps, seed):
  """Microscope-frame beam positions near the frame centre (Direct adapter)."""
  rng = np.random.default_rng(seed)
  return 0.5 + rng.uniform(-1.0, 1.0, size=(beam_steps, n_envs, 2)) * (
      1.42 / 22.5)


# ----------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------
class ClockSampler:
  FIELDS = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
            'clocks_event_reasons.hw_thermal_slowdown,'
            'clocks_event_reasons.sw_thermal_slowdown,'
            'clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.samples, self.proc = [], None
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', f'--query-gpu={self.FIELDS}',
           '--format=csv,noheader,nounits', '-lms', '100', '-i', str(index)],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.samples.append((time.perf_counter(), line.strip()))

  def stop(self, t0, t1):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
    time.sleep(0.15)
    self.proc.terminate()
    rows = [s for t, s in self.samples if t0 - 0.1 <= t <= t1 + 0.1]
    rows = rows or [s for _, s in self.samples]
    mhz, mx, reasons = [], None, set()
    names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
             'sw_power_cap')
    for r in rows:
      parts = [p.strip() for p in r.split(',')]
      try:
        mhz.append(float(parts[0]))
        mx = float(parts[1])
      except (ValueError, IndexError):
        continue
      for nme, v in zip(names, parts[2:]):
        if v.lower().startswith('active'):
          reasons.add(nme)
    return {'sm_mhz': float(np.median(mhz)) if mhz else None,
            'sm_max_mhz': mx, 'reasons': sorted(reasons),
            'samples': len(mhz)}


# ----------------------------------------------------------------------------
# CPU arm: the oracle port on host cores
# ----------------------------------------------------------------------------
def _cpu_worker(args):
  seed, env_offset, n_envs, beam_steps, rate_fn, ctrl_seed = args
  from oracle import pdune_oracle as po
  from oracle import pdune_oracle_episode as oe
  st = po.make_state(n_envs, seed, env_offset=env_offset)
  po.reset(st)
  act = synthetic_controls(n_envs, beam_steps, ctrl_seed)
  t0 = time.perf_counter()
  for t in range(beam_steps):
    ctl = oe.relative_to_silicon_controls(st, act[t])
    po.step_and_image(st, ctl[:, None, :], DWELL_US, IMAGE_US,
                      rate_fn=rate_fn)
  return time.perf_counter() - t0, n_envs * beam_steps


def cpu_port_throughput(n_envs, rate_fn, seconds, cores):
  """Times oracle.step_and_image on `cores` processes, each owning a shard of
  the env batch, for about `seconds`; returns (env-steps/s, sample text)."""
  import multiprocessing as mp
  shard = max(1, n_envs // cores)
  # calibrate on one short run
  dt, work = _cpu_worker((0, 0, shard, 2, rate_fn, 1))
  per_step = dt / 2
  beam_steps = int(max(2, min(20000, seconds / max(per_step, 1e-6))))
  jobs = [(0, i * shard, shard, beam_steps, rate_fn, 1 + i)
          for i in range(cores)]
  t0 = time.perf_counter()
  if cores == 1:
    res = [_cpu_worker(jobs[0])]
  else:
    with mp.get_context('fork').Pool(cores) as pool:
      res = pool.map(_cpu_worker, jobs)
  wall = time.perf_counter() - t0
  total = sum(w for _, w in res)
  busy = max(d for d, _ in res)
  sample = (f'{shard * cores} envs x {beam_steps} beam steps, NumPy oracle '
            f'port (vectorised over envs), {cores} process(es), '
            f'{busy:.1f} s busy / {wall:.1f} s wall')
  return total / busy, sample, wall


def reference_throughput(seconds, cores, rate):
  """env-steps/s of the UNMODIFIED reference (oracle/refbench.py: its own
  PuttingDuneSimulator, KD-tree, rate functions, adapter and
  np.random.default_rng) on `cores` processes, one env each, for about
  `seconds`; None if the reference sources are not on this box."""
  from oracle import refbench
  if not refbench.reference_available():
    return None
  refbench.run_steps(20, 0, rate)  # imports + first-call costs, once
  n_steps = int(max(50, min(5000, 230 * seconds)))  # ~230-300 steps/s/core
  v, busy, hops, sample = refbench.throughput(n_steps, cores, rate)
  return {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'reference',
          'sample': sample}


def run_reference_arm(args, cfg):
  """The reference's own CPU implementation of the path on all host cores:
  the unmodified PuttingDuneSimulator from baseline/_ref, one env per process
  (it has no batch dimension: configs[1]'s 4096 envs are 4096 such
  simulators, so env-steps/s is the same per-core figure), each timed step a
  bounded sample.  The NumPy port of the oracle, vectorised over the 4096
  envs, is reported beside it."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  from oracle import pdune_oracle as po
  from oracle import refbench
  rate_fn = po.RATE_PRIOR if args.rate == 'prior' else po.RATE_SIMPLE
  cores = os.cpu_count() or 1
  n_envs = cfg['envs_per_gpu']
  have_ref = refbench.reference_available()
  vals, samples, walls = [], [], []
  t_all = time.perf_counter()
  for i in range(args.warmup + args.steps):
    t0 = time.perf_counter()
    if have_ref:
      r = reference_throughput(2.0, cores, args.rate)
      v, smp = r['value'], r['sample']
    else:
      v, smp, _ = cpu_port_throughput(n_envs, rate_fn, 6.0, cores)
    if i >= args.warmup:
      vals.append(v)
      samples.append(smp)
      walls.append(time.perf_counter() - t0)
    if time.perf_counter() - t_all > 240:
      break
  value = float(np.mean(vals)) if vals else 0.0
  port_v, port_sample, _ = cpu_port_throughput(n_envs, rate_fn, 6.0, cores)
  kind = 'reference' if have_ref else 'port'
  line = {
      'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
      'n_gpus': args.gpus, 'steps': len(vals), 'warmup': args.warmup,
      'ms_per_step': 1e3 * float(np.mean(walls)) if walls else None,
      'higher_is_better': True,
      'scaling': cfg['scaling'], 'vs_baseline': None, 'dtype': 'f64',
      'data': 'synthetic', 'config': cfg,
      'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores,
                       'kind': kind, 'sample': samples[-1] if samples else ''},
      'cpu_port': {'value': port_v, 'unit': UNIT, 'cores': cores,
                   'kind': 'port', 'sample': port_sample},
      'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
              'd2h_bytes_per_step': 0},
      'note': ('unmodified reference simulator from baseline/_ref on all host '
               'cores (oracle/refbench.py); cpu_port = oracle/pdune_oracle.py, '
               'the NumPy restatement vectorised over the envs'
               if have_ref else
               'reference sources not on this box: oracle/pdune_oracle.py '
               '(NumPy port, pinned against the unmodified reference)'),
  }
  print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------
def ncu_traffic(summary_name):
  """dram read+write bytes per launch from a committed ncu --set full summary
  (profiles/*.ncu.json), or None."""
  path = os.path.join(ROOT, 'profiles', summary_name)
  try:
    d = json.load(open(path))
  except (OSError, ValueError):
    return None
  scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
  total = 0.0
  for k, v in d.items():
    if k.startswith('dram__bytes_read.sum') or k.startswith(
        'dram__bytes_write.sum'):
      unit = k[k.index('[') + 1:k.index(']')]
      total += float(v) * scale.get(unit, 1.0)
  return total or None


def ncu_value(summary_name, key):
  """One metric of a committed ncu summary (profiles/*.ncu.json), or None."""
  path = os.path.join(ROOT, 'profiles', summary_name)
  try:
    d = json.load(open(path))
  except (OSError, ValueError):
    return None
  for k, v in d.items():
    if k.startswith(key):
      try:
        return float(v)
      except ValueError:
        return None
  return None


def issue_slots(warp_instructions, seconds, sm_mhz, sms=148):
  """Issue-slot roofline: warp-instructions of one launch over the issue
  slots the launch had (4 schedulers per SM, one issue per cycle each)."""
  if not warp_instructions or not sm_mhz:
    return None
  slots = 4 * sms * sm_mhz * 1e6 * seconds
  return {'warp_instructions_per_launch': warp_instructions,
          'schedulers': 4 * sms, 'sm_mhz': sm_mhz,
          'frac': warp_instructions / slots,
          'source': 'smsp__inst_executed.sum of the committed ncu summary / '
                    '(schedulers x SM clock x measured launch time)'}


def bind_to_gpu_numa(index):
  """Pins this process (and so the pinned buffers it first touches) to the
  CPUs next to GPU `index` (NVML affinity); returns the CPU count or None."""
  try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(index)
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    cpus = [64 * i + b for i, w in enumerate(words) for b in range(64)
            if (w >> b) & 1]
    if cpus:
      os.sched_setaffinity(0, cpus)
      return len(cpus)
  except Exception:  # pylint: disable=broad-except
    return None
  return None


def measure_link(dev, mb=64):
  """Pinned-memory copy rates of this box's host link in the same run: H2D
  alone, D2H alone, and both at once (GB/s per direction)."""
  import torch
  n = mb << 20
  h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
  h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
  d_in = torch.empty(n, dtype=torch.uint8, device=dev)
  d_out = torch.zeros(n, dtype=torch.uint8, device=dev)
  s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

  def timed(h2d, d2h):
    best = 1e9
    for _ in range(4):
      torch.cuda.synchronize()
      t0 = time.perf_counter()
      if h2d:
        with torch.cuda.stream(s1):
          d_in.copy_(h_in, non_blocking=True)
      if d2h:
        with torch.cuda.stream(s2):
          h_out.copy_(d_out, non_blocking=True)
      torch.cuda.synchronize()
      best = min(best, time.perf_counter() - t0)
    return n / best / 1e9

  return {'h2d_gbs': timed(True, False), 'd2h_gbs': timed(False, True),
          'both_gbs_per_direction': timed(True, True), 'mbytes': mb,
          'how': 'pinned cudaMemcpyAsync, best of 4, host wall clock'}


def measured_peak():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  try:
    return float(json.load(open(path))['hbm_gbs']), 'measured'
  except (OSError, KeyError, ValueError):
    return 6650.0, 'fallback'


def measure_export(pd, batch, dev, peak):
  """Observation records/s of pd_encode_observations (protobuf wire bytes of
  MicroscopeObservation for every env): byte work
  bounded by HBM -- the records written plus the 76 B of env state read."""
  import torch
  import ctypes as C
  from putting_dune_b200 import _native as nat
  n = 65536
  eb = pd.EnvBatch(n, seed=5, device=dev, lattice=batch.lattice_tables)
  eb.reset()
  max_atoms = min(eb.max_atoms_in_view(), eb.lattice_tables.n_sites)
  slot = (int(nat.lib.pd_observation_bytes(max_atoms, 1)) + 15) & ~15
  out = torch.empty(n * slot, dtype=torch.uint8, device=dev)
  offsets = torch.empty(n + 1, dtype=torch.int64, device=dev)
  length = torch.empty(n, dtype=torch.int32, device=dev)
  atoms = torch.empty(n, dtype=torch.int32, device=dev)
  ctl = torch.full((n, 1, 2), 0.5, dtype=torch.float64, device=dev)
  P = lambda t: C.c_void_p(t.data_ptr())
  stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
  def launch():
    nat.check(nat.lib.pd_encode_observations(
        C.byref(eb.lattice_tables.c), C.byref(eb.c), P(ctl), None, DWELL_US, 1,
        None, 60.0, 0.1, max_atoms, P(out), out.numel(), P(offsets), P(length),
        P(atoms), None, stream))
  for _ in range(3):
    launch()
  torch.cuda.synchronize()
  reps = 10
  a, b = (torch.cuda.Event(enable_timing=True),
          torch.cuda.Event(enable_timing=True))
  a.record()
  for _ in range(reps):
    launch()
  b.record()
  torch.cuda.synchronize()
  ms = a.elapsed_time(b) / reps
  nbytes = int(offsets[n].item())
  gbs = nbytes / (ms / 1e3) / 1e9
  return {'metric': 'observation records/sec', 'value': n / (ms / 1e3),
          'records_per_launch': n, 'bytes_per_launch': nbytes,
          'mean_atoms': float(atoms.float().mean().item()), 'launch_ms': ms,
          'kernels': 'pd::k_obs_sizes + k_obs_offsets + k_obs_encode',
          'roofline': {'bound': 'hbm', 'achieved': gbs, 'peak': peak,
                       'unit': 'GB/s', 'frac': gbs / peak,
                       'algorithmic_bytes_per_record': nbytes / n}}


def measure_frames(pd, batch, dev, peak, args):
  """Frames/s of pd_render (512x512, all noise stages + CLAHE)."""
  import torch
  from putting_dune_b200 import imaging
  import ctypes as C
  from putting_dune_b200 import _native as nat
  n_cl = C.c_int32()
  nat.check(nat.lib.pd_render_clusters(512, C.byref(n_cl)))
  m = max(n_cl.value, args.frames)
  free, _ = torch.cuda.mem_get_info(dev)
  m = int(min(m, (free - (2 << 30)) // (512 * 512 * 4)))
  fb = pd.EnvBatch(m, seed=3, device=dev, lattice=batch.lattice_tables)
  fb.reset()
  out = torch.empty((m, 512, 512), dtype=torch.float32, device=dev)
  for _ in range(2):
    imaging.render_batch(fb, out=out)
  torch.cuda.synchronize()
  ev = [(torch.cuda.Event(enable_timing=True),
         torch.cuda.Event(enable_timing=True)) for _ in range(5)]
  for a, b in ev:
    a.record()
    imaging.render_batch(fb, out=out)
    b.record()
  torch.cuda.synchronize()
  ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
  del out, fb
  torch.cuda.empty_cache()
  fps = m / (ms / 1e3)
  gbs = fps * 512 * 512 * 4 / 1e9
  return {'metric': 'STEM frames/sec', 'value': fps, 'unit': 'frames/s',
          'frames_per_launch': m, 'image_size': 512, 'launch_ms': ms,
          'kernel': 'pd::k_render_cluster',
          'clusters_resident': n_cl.value, 'ctas_per_frame': 8,
          'workload': 'configs[3]: batched STEM image rendering 512x512 with '
                      'noise, %d envs per step' % m,
          'roofline': {'bound': 'hbm', 'achieved': gbs, 'peak': peak,
                       'unit': 'GB/s', 'frac': gbs / peak,
                       'algorithmic_bytes_per_frame': 512 * 512 * 4,
                       # ncu --set full of 120 frames, per frame
                       'traffic_per_frame': (ncu_traffic(
                           'r01_k_render_cluster_120frames.ncu.json') or 0) /
                                            120 or None}}


def measure_mlp(pd, batch, dev, args):
  """BASELINE configs[2]: 65536 envs stepping with the learned rate-model MLP
  (seeded synthetic weights) -- FP32 FMA GEMM path."""
  import torch
  from oracle import pdune_oracle as po  # weights generator only
  out = {}
  n = 65536
  rng = np.random.default_rng(0)
  # tensor_core: False = FP32 FMA (parity path), 1 = tcgen05 with bf16
  # operands (2e-2 of the largest rate), 2 = tcgen05 with fp16 hi + lo
  # operands, three MMAs per K step (3e-7 .. 6e-7: a parity path; at H = 256
  # the W1 tiles are streamed from L2, K in two halves per wave)
  for hidden, tensor_core in (((64, 64), False), ((128, 128), False),
                              ((256, 256), False), ((64, 64), 2),
                              ((128, 128), 2), ((256, 256), 2),
                              ((128, 128), True), ((256, 256), True)):
    mlp = po.MlpParams.synthetic(7, hidden=hidden)
    w = pd.MlpWeights(**{k: getattr(mlp, k) for k in pd.MlpWeights.NAMES})
    rate = pd.RateSpec(2, mlp=w, device=dev, tensor_core=tensor_core)
    b = pd.EnvBatch(n, seed=11, device=dev, lattice=batch.lattice_tables)
    b.reset()
    ctl = torch.as_tensor(direct_controls(n, 1, 3)[0][:, None, :]).to(dev)
    for _ in range(3):
      b.step_and_image(ctl, DWELL_US, rate, IMAGE_US)
    torch.cuda.synchronize()
    ev0 = int(b.n_events.sum().item())
    evs = [(torch.cuda.Event(enable_timing=True),
            torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, c in evs:
      a.record()
      b.step_and_image(ctl, DWELL_US, rate, IMAGE_US)
      c.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(c) for a, c in evs) / len(evs)
    evals = (int(b.n_events.sum().item()) - ev0) / len(evs)
    flop = 2 * (2 * hidden[0] + hidden[0] * hidden[1] + 4 * hidden[1])
    tag = {0: '_fp32', 1: '_tcgen05_bf16', 2: '_tcgen05_f16x3'}[
        int(tensor_core)]
    out[f'H{hidden[0]}' + tag] = {
        'envs': n, 'launch_ms': ms, 'env_steps_per_s': n / (ms / 1e3),
        'rate_evals_per_step': evals / n, 'flop_per_eval': flop,
        'tflops': evals * flop / (ms / 1e3) / 1e12,
        'parity_path': int(tensor_core) != 1,
        'kernel': {
            0: 'pd::k_step_learned (FP32 FMA, queue-batched GEMM, cp.async '
               'double-buffered W1 chunks; 512 threads)',
            1: 'pd::k_step_learned<TC> (tcgen05.mma kind::f16, BF16 '
               'operands, FP32 accumulate in TMEM; two CTAs of 256 threads per '
               'SM where two sets of tiles fit, else one of 512)',
            2: 'pd::k_step_learned<TC> (tcgen05.mma kind::f16, FP16 hi + lo '
               'operands, three MMAs per K step into one TMEM accumulator; '
               'rates within 3e-7 of the FP32 path; two CTAs of 256 threads per SM '
               'at H <= 64, else one of 512)'}[int(tensor_core)]}
  return out


def measure_episodes(pd, args, world, rank, dev, barrier):
  """BASELINE configs[4]: `--episodes` envs sharded over the ranks, greedy
  controller to the end of every episode, stats all-gathered with NCCL."""
  import torch
  from putting_dune_b200 import episodes as ep
  lo, n = ep.shard_bounds(args.episodes, rank, world)
  b = pd.EnvBatch(n, seed=5, env_offset=lo, device=dev)
  rate = pd.RateSpec.simple()
  stats, _, _ = ep.run_greedy_episodes(b, rate)  # warm-up
  ep.gather_episode_stats(stats)
  barrier()
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
  ev[0].record()
  stats, _, _ = ep.run_greedy_episodes(b, rate)
  ev[1].record()
  full = ep.gather_episode_stats(stats)
  ev[2].record()
  barrier()
  times = [(ev[0].elapsed_time(ev[2]), ev[1].elapsed_time(ev[2]))]
  # four more runs (new goals each: the episode counter moves on), timed the
  # same way: a single ~5 ms sample was seen to vary by a millisecond with
  # whatever the host was doing between its six launches.  The records and
  # their digest are those of the first timed run.
  for _ in range(4):
    barrier()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    st2, _, _ = ep.run_greedy_episodes(b, rate)
    e[1].record()
    ep.gather_episode_stats(st2)
    e[2].record()
    barrier()
    times.append((e[0].elapsed_time(e[2]), e[1].elapsed_time(e[2])))
  times.sort()
  tm = torch.tensor(list(times[len(times) // 2]), dtype=torch.float64,
                    device=dev)
  if world > 1:
    import torch.distributed as dist
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
  agg = ep.aggregate_results(full)
  ms, gather_ms = float(tm[0].item()), float(tm[1].item())
  import hashlib
  digest = hashlib.sha256(full.cpu().numpy().tobytes()).hexdigest()[:16]
  agg.update(launch_ms=ms, env_steps_per_s=agg['total_actions'] / (ms / 1e3),
             timing='median of 5 runs (reset + goal selection + episodes + '
                    'all-gather), CUDA events, max over ranks',
             launch_ms_all=[round(t[0], 3) for t in times],
             allgather_us=gather_ms * 1e3 if world > 1 else 0.0,
             collective=('NCCL all_gather of 16 B/env episode records '
                         f'({16 * args.episodes / 1e6:.1f} MB total)'
                         if world > 1 else 'none (one rank)'),
             records_sha256_16=digest,
             note='records_sha256_16 is over the gathered per-env records: '
                  'identical for every number of GPUs (Philox is keyed by '
                  'the global env id)',
             rate_function='simple', controller='GreedyAgent argmax (1.42, 0)',
             workload='configs[4]: %d envs sharded over %d GPU(s), greedy '
                      'goal-reaching controller, step limit 600, 10 simulated '
                      'minutes' % (args.episodes, world))
  return agg


def run_ours(args, cfg):
  import torch
  import torch.distributed as dist
  import putting_dune_b200 as pd
  from putting_dune_b200 import _native as nat

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if not torch.cuda.is_available():
    raise SystemExit('bench.py needs a GPU: there is no CPU fallback')
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  numa_cpus = bind_to_gpu_numa(local)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  n, t_steps = cfg['envs_per_gpu'], cfg['beam_steps_per_launch']
  rate = pd.RateSpec.prior() if args.rate == 'prior' else pd.RateSpec.simple()
  batch = pd.EnvBatch(n, seed=0, env_offset=rank * n, device=dev)
  batch.reset()
  pool = 4
  h_ctl = [torch.as_tensor(synthetic_controls(n, t_steps, 100 + rank * pool + i)
                           ).pin_memory() for i in range(pool)]
  d_ctl = [c.to(dev) for c in h_ctl]
  flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
  lat_c, st_c = C.byref(batch.lattice_tables.c), C.byref(batch.c)
  stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
  P = lambda t: C.c_void_p(t.data_ptr())

  # the per-step observation of every env-step is part of the timed call:
  # Si site (int32) and MicroscopeObservation.elapsed_time (int64 us)
  d_si = torch.empty((t_steps, n), dtype=torch.int32, device=dev)
  d_el = torch.empty((t_steps, n), dtype=torch.int64, device=dev)

  def launch(i):
    nat.check(nat.lib.pd_rollout_actions(
        lat_c, st_c, C.byref(rate.c), P(d_ctl[i % pool]),
        nat.ACTION_RELATIVE_TO_SILICON, 1.42, DWELL_US, t_steps, IMAGE_US,
        P(d_si), P(d_el), stream))

  # -- device-resident: value -----------------------------------------------
  for i in range(args.warmup):
    flush.zero_()
    launch(i)
  barrier()
  sampler = ClockSampler(local) if rank == 0 else None
  clocks = None
  ev = [(torch.cuda.Event(enable_timing=True),
         torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
  t0 = time.perf_counter()
  for i in range(args.steps):
    flush.zero_()
    ev[i][0].record()
    launch(i)
    ev[i][1].record()
  barrier()
  t1 = time.perf_counter()
  dev_ms = sum(a.elapsed_time(b) for a, b in ev)
  tm = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
  dev_ms = float(tm.item())
  env_steps_per_launch = n * t_steps
  total_env_steps = env_steps_per_launch * args.steps * world
  value = total_env_steps / (dev_ms / 1e3)

  # -- end to end through the host-buffer C ABI -------------------------------
  d_stage = torch.empty((t_steps, n, 2), dtype=torch.float64, device=dev)
  h_si = torch.empty((t_steps, n), dtype=torch.int32).pin_memory()
  h_el = torch.empty((t_steps, n), dtype=torch.int64).pin_memory()

  def launch_host(i):
    nat.check(nat.lib.pd_rollout_actions_host(
        lat_c, st_c, C.byref(rate.c), P(h_ctl[i % pool]),
        nat.ACTION_RELATIVE_TO_SILICON, 1.42, DWELL_US, t_steps, IMAGE_US,
        P(d_stage), P(d_si), P(d_el), P(h_si), P(h_el), stream))

  # odd, and long enough (~0.1 s) to span the clock sampler's 100 ms period:
  # an nvidia-smi query in flight delays launches for a few milliseconds
  E2E_REPEATS = 15

  def time_host(fn):
    """Median over E2E_REPEATS timings of exactly args.steps calls each (a
    call is ~0.25 ms, so one timing is a few milliseconds of wall clock and a
    single scheduling hiccup of the host would otherwise set the figure)."""
    for i in range(args.warmup):
      fn(i)
    rates = []
    for _ in range(E2E_REPEATS):
      barrier()
      e0 = time.perf_counter()
      for i in range(args.steps):
        fn(i)
      barrier()
      tm = torch.tensor([time.perf_counter() - e0], dtype=torch.float64,
                        device=dev)
      if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
      rates.append(total_env_steps / float(tm.item()))
    return float(np.median(rates))

  # float64 actions in, int64 elapsed out (the reference's in-memory dtypes)
  e2e_f64 = time_host(launch_host)
  # compact formats: float32 actions (action_spec dtype), int32 elapsed us
  h_ctl32 = [c.float().pin_memory() for c in h_ctl]
  d_a32 = torch.empty((t_steps, n, 2), dtype=torch.float32, device=dev)
  d_el32 = torch.empty((t_steps, n), dtype=torch.int32, device=dev)
  h_el32 = torch.empty((t_steps, n), dtype=torch.int32).pin_memory()

  def launch_host32(i):
    nat.check(nat.lib.pd_rollout_actions_host_f32(
        lat_c, st_c, C.byref(rate.c), P(h_ctl32[i % pool]),
        nat.ACTION_RELATIVE_TO_SILICON, 1.42, DWELL_US, t_steps, IMAGE_US,
        P(d_a32), P(d_stage), P(d_si), P(d_el), P(d_el32), P(h_si), P(h_el32),
        stream))

  e2e_f32 = time_host(launch_host32)
  # packed results: uint16 Si site | re-centred << 15 (the elapsed time is
  # dwell + image * (1 + re-centred)): 8 B in + 2 B out per env-step
  h_packed = torch.empty((t_steps, n), dtype=torch.uint16).pin_memory()

  def launch_packed(i):
    nat.check(nat.lib.pd_rollout_actions_host_packed(
        lat_c, st_c, C.byref(rate.c), P(h_ctl32[i % pool]),
        nat.ACTION_RELATIVE_TO_SILICON, 1.42, DWELL_US, t_steps, IMAGE_US,
        P(h_packed), stream))

  e2e_value = time_host(launch_packed)
  h2d = h_ctl32[0].numel() * 4
  d2h = h_packed.numel() * 2
  link = measure_link(dev) if rank == 0 else None
  if world > 1:
    # every rank copies at once: what the ranks' shared host path (PCIe
    # switches / root ports / DRAM) gives each of them, summed
    barrier()
    mine = measure_link(dev, mb=32)
    tsum = torch.tensor([mine['h2d_gbs'], mine['d2h_gbs'],
                         mine['both_gbs_per_direction']],
                        dtype=torch.float64, device=dev)
    dist.all_reduce(tsum)
    if rank == 0:
      link['all_ranks_at_once'] = {
          'h2d_gbs_sum': float(tsum[0]), 'd2h_gbs_sum': float(tsum[1]),
          'both_gbs_per_direction_sum': float(tsum[2]), 'ranks': world,
          'note': 'the same three copies issued by every rank after a '
                  'barrier; sums over ranks (ceiling of the shared host '
                  'path for the e2e call)'}

  # -- roofline of the dominant kernel ----------------------------------------
  peak, peak_kind = measured_peak()
  launch_s = dev_ms / 1e3 / args.steps
  achieved = ALGORITHMIC_BYTES_PER_ENV_STEP * env_steps_per_launch / launch_s / 1e9
  roofline = {
      'kernel': 'pd::k_rollout_fast', 'bound': 'hbm', 'achieved': achieved,
      'peak': peak, 'peak_source': f'{peak_kind} (MEASURED_PEAKS.json hbm_gbs)',
      'unit': 'GB/s', 'frac': achieved / peak,
      'algorithmic_bytes_per_env_step': ALGORITHMIC_BYTES_PER_ENV_STEP,
      'env_steps_per_launch': env_steps_per_launch,
      'launch_ms': launch_s * 1e3,
      # ncu --set full of the same workload (profiles/
      # r02_k_rollout_fast_config2): the state stays in registers across the
      # 256 steps, so DRAM traffic is the action stream in (16 B/env-step)
      # and the observations out (12 B/env-step), below the 64 B figure
      'traffic': ncu_traffic('r02_k_rollout_fast_config2.ncu.json')
      if cfg['envs_per_gpu'] == 4096 and t_steps == 256 else None,
  }
  issue_src = ('r02_k_rollout_fast_config2.ncu.json'
               if cfg['envs_per_gpu'] == 4096 and t_steps == 256 else None)

  # -- the same kernel family with every SM filled (1Mi envs, one step) -------
  at_scale = None
  if rank == 0 and not args.no_at_scale:
    big_n = 1 << 20
    big = pd.EnvBatch(big_n, seed=1, device=dev,
                      lattice=batch.lattice_tables)
    big.reset()
    acts = [torch.as_tensor(synthetic_controls(big_n, 1, 7 + i)).to(dev)
            for i in range(pool)]
    b_si = torch.empty((8, big_n), dtype=torch.int32, device=dev)
    b_el = torch.empty((8, big_n), dtype=torch.int64, device=dev)
    def big_launch(i):
      nat.check(nat.lib.pd_rollout_actions(
          C.byref(big.lattice_tables.c), C.byref(big.c), C.byref(rate.c),
          P(acts[i % pool]), nat.ACTION_RELATIVE_TO_SILICON, 1.42, DWELL_US,
          1, IMAGE_US, P(b_si), P(b_el), stream))
    for i in range(3):
      big_launch(i)
    torch.cuda.synchronize()
    bev = [(torch.cuda.Event(enable_timing=True),
            torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for i in range(10):
      flush.zero_()
      bev[i][0].record()
      big_launch(i)
      bev[i][1].record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in bev) / 10
    a_gbs = ALGORITHMIC_BYTES_PER_ENV_STEP * big_n / (ms / 1e3) / 1e9
    at_scale = {
        'workload': '1Mi envs x 1 step per launch, same actions/adapter, '
                    '1 GPU',
        'kernel': 'pd::k_walk_fast', 'value': big_n / (ms / 1e3), 'unit': UNIT,
        'outputs': 'Si site int32 + elapsed us int64 per env-step',
        'launch_ms': ms, 'achieved': a_gbs, 'peak': peak, 'frac': a_gbs / peak,
        'traffic': ncu_traffic('r02_k_walk_fast_1Mi_1step.ncu.json')}
    # the same batch in 8-step rollouts (state read and written once per 8)
    acts8 = [torch.as_tensor(synthetic_controls(big_n, 8, 17 + i)).to(dev)
             for i in range(2)]
    def big_launch8(i):
      nat.check(nat.lib.pd_rollout_actions(
          C.byref(big.lattice_tables.c), C.byref(big.c), C.byref(rate.c),
          P(acts8[i % 2]), nat.ACTION_RELATIVE_TO_SILICON, 1.42, DWELL_US,
          8, IMAGE_US, P(b_si), P(b_el), stream))
    for i in range(2):
      big_launch8(i)
    torch.cuda.synchronize()
    bev = [(torch.cuda.Event(enable_timing=True),
            torch.cuda.Event(enable_timing=True)) for _ in range(6)]
    for i in range(6):
      flush.zero_()
      bev[i][0].record()
      big_launch8(i)
      bev[i][1].record()
    torch.cuda.synchronize()
    ms8 = sum(a.elapsed_time(b) for a, b in bev) / 6
    g8 = ALGORITHMIC_BYTES_PER_ENV_STEP * big_n * 8 / (ms8 / 1e3) / 1e9
    at_scale['rollout8'] = {
        'workload': '1Mi envs x 8 steps per launch',
        'kernel': 'pd::k_walk_plan (+ pd::k_walk_fast<LIST> over the envs it '
                  'hands over)',
        'value': big_n * 8 / (ms8 / 1e3), 'unit': UNIT, 'launch_ms': ms8,
        'achieved': g8, 'peak': peak, 'frac': g8 / peak,
        'traffic': ncu_traffic('r02_k_walk_plan_1Mi_8step.ncu.json'),
        'warp_instructions': ncu_value('r02_k_walk_plan_1Mi_8step.ncu.json',
                                       'smsp__inst_executed.sum')}
    del acts8, b_si, b_el
    # ... and in 64-step rollouts (the per-launch costs -- env state in and
    # out, the second launch -- spread over 64 steps)
    acts64 = torch.as_tensor(synthetic_controls(big_n, 64, 29)).to(dev)
    b_si = torch.empty((64, big_n), dtype=torch.int32, device=dev)
    b_el = torch.empty((64, big_n), dtype=torch.int64, device=dev)
    def big_launch64(i):
      nat.check(nat.lib.pd_rollout_actions(
          C.byref(big.lattice_tables.c), C.byref(big.c), C.byref(rate.c),
          P(acts64), nat.ACTION_RELATIVE_TO_SILICON, 1.42, DWELL_US,
          64, IMAGE_US, P(b_si), P(b_el), stream))
    big_launch64(0)
    torch.cuda.synchronize()
    bev = [(torch.cuda.Event(enable_timing=True),
            torch.cuda.Event(enable_timing=True)) for _ in range(4)]
    for i in range(4):
      flush.zero_()
      bev[i][0].record()
      big_launch64(i)
      bev[i][1].record()
    torch.cuda.synchronize()
    ms64 = sum(a.elapsed_time(b) for a, b in bev) / 4
    g64 = ALGORITHMIC_BYTES_PER_ENV_STEP * big_n * 64 / (ms64 / 1e3) / 1e9
    at_scale['rollout64'] = {
        'workload': '1Mi envs x 64 steps per launch',
        'kernel': 'pd::k_walk_plan (+ pd::k_walk_fast<LIST>)',
        'value': big_n * 64 / (ms64 / 1e3), 'unit': UNIT, 'launch_ms': ms64,
        'achieved': g64, 'peak': peak, 'frac': g64 / peak,
        'traffic': ncu_traffic('r02_k_walk_plan_1Mi_64step.ncu.json'),
        'warp_instructions': ncu_value('r02_k_walk_plan_1Mi_64step.ncu.json',
                                       'smsp__inst_executed.sum')}
    del big, acts64, b_si, b_el

  # -- STEM frames/s (the second half of BASELINE.json's metric) ---------------
  frames = None
  if rank == 0 and not args.no_frames:
    frames = measure_frames(pd, batch, dev, peak, args)

  # -- goal-reaching episodes with the greedy controller (configs[4]) ---------
  episodes = None
  if args.episodes:
    episodes = measure_episodes(pd, args, world, rank, dev, barrier)

  export = None
  if rank == 0 and not args.no_export:
    export = measure_export(pd, batch, dev, peak)

  mlp = None
  if rank == 0 and not args.no_mlp:
    mlp = measure_mlp(pd, batch, dev, args)

  cpu = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    if sampler:  # the CPU leg is not a GPU region
      clocks = sampler.stop(t0, time.perf_counter())
      sampler = None
    from oracle import pdune_oracle as po
    rate_fn = po.RATE_PRIOR if args.rate == 'prior' else po.RATE_SIMPLE
    v, sample, _ = cpu_port_throughput(n, rate_fn, args.cpu_seconds, 1)
    port = {'value': v, 'unit': UNIT, 'cores': 1, 'kind': 'port',
            'sample': sample}
    # the unmodified reference simulator (BASELINE configs[0]) on one core
    cpu = reference_throughput(args.cpu_seconds, 1, args.rate)
    if cpu is None:
      cpu = port
    else:
      cpu['port'] = port

  # clocks over every timed GPU region of this run (value, e2e, at_scale ...)
  if sampler:
    clocks = sampler.stop(t0, time.perf_counter())
  if rank == 0:
    sm_mhz = (clocks or {}).get('sm_mhz') or (clocks or {}).get('sm_max_mhz')
    roofline['issue_slots'] = issue_slots(
        ncu_value(issue_src, 'smsp__inst_executed.sum') if issue_src else None,
        launch_s, sm_mhz)
    for key in ('rollout8', 'rollout64'):
      if at_scale and at_scale.get(key):
        r8 = at_scale[key]
        r8['issue_slots'] = issue_slots(r8.pop('warp_instructions'),
                                        r8['launch_ms'] / 1e3, sm_mhz)
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dev_ms / args.steps, 'higher_is_better': True,
        'scaling': cfg['scaling'], 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': cfg, 'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h,
                'timing': f'median of {E2E_REPEATS} timings of {args.steps} '
                          'calls each, host wall clock, max over ranks',
                'api': 'pd_rollout_actions_host_packed (pinned host buffers: '
                       'float32 actions in = the adapters\' action_spec '
                       'dtype; one uint16 per env-step out = Si site | '
                       're-centred << 15, from which elapsed = dwell + image '
                       '* (1 + re-centred); copy-engine chunks of whole '
                       'steps that the k_rollout_fast launches and the '
                       'result copies follow)',
                'link': link, 'numa_bound_cpus': numa_cpus,
                'f32_int32_io': {
                    'value': e2e_f32, 'unit': UNIT,
                    'h2d_bytes_per_step': h_ctl32[0].numel() * 4,
                    'd2h_bytes_per_step': h_si.numel() * 4 +
                                          h_el32.numel() * 4,
                    'api': 'pd_rollout_actions_host_f32 (int32 Si site + '
                           'int32 elapsed us out; the same copy-engine '
                           'pipeline of chunks)'},
                'float64_io': {
                    'value': e2e_f64, 'unit': UNIT,
                    'h2d_bytes_per_step': h_ctl[0].numel() * 8,
                    'd2h_bytes_per_step': h_si.numel() * 4 + h_el.numel() * 8,
                    'api': 'pd_rollout_actions_host (float64 actions, int64 '
                           'elapsed us; copy-engine pipeline of 4 MiB '
                           'chunks)'}},
        'gpu_launches': args.steps,
        'value_outputs': 'Si site int32 + elapsed us int64 written for every '
                         'env-step inside the timed launch',
        'roofline': roofline,
        'cpu_baseline': cpu, 'at_scale': at_scale, 'frames': frames,
        'episodes': episodes, 'learned_mlp': mlp, 'export': export,
        'wall_s_timed_region': t1 - t0,
    }
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()


def main():
  args = parse_args()
  cfg = workload_config(args)
  if args.impl == 'reference':
    run_reference_arm(args, cfg)
  else:
    run_ours(args, cfg)


if __name__ == '__main__':
  main()
