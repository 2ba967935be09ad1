"""Batched environment state and the calls into the CUDA library.

PyTorch is used for device memory, streams and (in ``bench.py``)
``torch.distributed`` only; every computation on the simulator path happens in
``libpdune_b200.so``.  There is no CPU fallback: constructing an ``EnvBatch``
without a CUDA device raises.
"""

from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from putting_dune_b200 import _native as nat

IMAGE_PARAM_NAMES = (
    'intensity_exponent', 'gaussian_variance', 'jitter_rate',
    'poisson_rate_multiplier', 'salt_and_pepper_amount', 'blur_amount',
    'contrast_gamma', 'exponential_lambda', 'uniform_noise_scale')


def _require_cuda(device) -> torch.device:
  if not torch.cuda.is_available():
    raise RuntimeError(
        'putting_dune_b200 needs a CUDA device (sm_100a); there is no CPU '
        'fallback for the simulator path.')
  device = torch.device('cuda' if device is None else device)
  if device.type != 'cuda':
    raise RuntimeError(f'device must be a CUDA device, got {device}')
  if device.index is None:
    device = torch.device('cuda', torch.cuda.current_device())
  return device


def _ptr(t: Optional[torch.Tensor]):
  return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
  return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Lattice:
  """Shared graphene lattice tables on the device (graphene.py:464-559 and the
  canonical 3-NN table replacing geometry.py:93-111)."""

  def __init__(self, grid_columns: int = 50, device=None):
    self.device = _require_cuda(device)
    n_sites, n_rows = C.c_int32(), C.c_int32()
    nat.check(nat.lib.pd_lattice_size(grid_columns, C.byref(n_sites),
                                      C.byref(n_rows)))
    self.grid_columns = grid_columns
    self.n_sites = n_sites.value
    self.n_rows = n_rows.value
    with torch.cuda.device(self.device):
      self.base_xy = torch.empty((self.n_sites, 2), dtype=torch.float64,
                                 device=self.device)
      self.nbr = torch.empty((self.n_sites, 4), dtype=torch.int32,
                             device=self.device)
      nat.check(nat.lib.pd_build_lattice(grid_columns, _ptr(self.base_xy),
                                         _ptr(self.nbr),
                                         _stream(self.device)))
    self.c = nat.PdLattice(grid_columns, self.n_sites,
                           self.base_xy.data_ptr(), self.nbr.data_ptr())


@dataclasses.dataclass
class MlpWeights:
  """Learned rate model parameters (Haiku tree of learn_rates.py:80-99)."""
  bn_scale: np.ndarray
  bn_offset: np.ndarray
  bn_mean: np.ndarray
  bn_var: np.ndarray
  w0: np.ndarray
  b0: np.ndarray
  w1: np.ndarray
  b1: np.ndarray
  w2: np.ndarray
  b2: np.ndarray
  batchnorm: bool = True

  NAMES = ('bn_scale', 'bn_offset', 'bn_mean', 'bn_var', 'w0', 'b0', 'w1',
           'b1', 'w2', 'b2')

  @classmethod
  def from_haiku(cls, params: dict, state: Optional[dict] = None,
                 batchnorm: bool = True) -> 'MlpWeights':
    """Builds weights from flat Haiku trees (names in SURVEY.md appendix
    A.4): params['batch_norm']['scale'], state['batch_norm/~/mean_ema']
    ['average'], params['mlp/~/linear_0']['w'] ..."""
    d = np.asarray(params['mlp/~/linear_0']['w']).shape[0]
    one, zero = np.ones(d, np.float32), np.zeros(d, np.float32)
    bn = params.get('batch_norm', {})
    st = state or {}
    return cls(
        bn_scale=np.asarray(bn.get('scale', one)).reshape(-1),
        bn_offset=np.asarray(bn.get('offset', zero)).reshape(-1),
        bn_mean=np.asarray(st.get('batch_norm/~/mean_ema', {}).get(
            'average', zero)).reshape(-1),
        bn_var=np.asarray(st.get('batch_norm/~/var_ema', {}).get(
            'average', one)).reshape(-1),
        w0=np.asarray(params['mlp/~/linear_0']['w']),
        b0=np.asarray(params['mlp/~/linear_0']['b']),
        w1=np.asarray(params['mlp/~/linear_1']['w']),
        b1=np.asarray(params['mlp/~/linear_1']['b']),
        w2=np.asarray(params['mlp/~/linear_2']['w']),
        b2=np.asarray(params['mlp/~/linear_2']['b']),
        batchnorm=batchnorm)


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
  """float32 -> bfloat16 bit patterns (uint16), round to nearest even."""
  u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
  rounded = u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))
  return (rounded >> np.uint32(16)).astype(np.uint16)


def umma_k_major_bf16(mat: np.ndarray) -> np.ndarray:
  """[rows, K] float32 -> bf16 in the canonical no-swizzle K-major UMMA layout
  (tcgen05 shared-memory descriptor, layout type INTERLEAVE): 8-row x 16-byte
  core matrices; consecutive core matrices along K are 128 B apart (LBO) and
  each group of 8 rows owns K/8 of them (SBO = K/8 * 128 B).  Returned as the
  flat uint16 array to copy into shared memory verbatim."""
  return _umma_k_major(to_bf16_bits(mat))


def _umma_k_major(bits: np.ndarray) -> np.ndarray:
  rows, k = bits.shape
  if rows % 8 or k % 8:
    raise ValueError('rows and K must be multiples of 8')
  bits = bits.reshape(rows // 8, 8, k // 8, 8)
  # (row group, row in group, k block, k in block) ->
  # (row group, k block, row in group, k in block)
  return np.ascontiguousarray(bits.transpose(0, 2, 1, 3)).reshape(-1)


def umma_k_major_f16_split(mat: np.ndarray) -> np.ndarray:
  """[rows, K] float32 -> fp16 hi tile followed by fp16 lo tile (mat = hi + lo
  to ~2^-22 relative), each in the layout of `umma_k_major_bf16`: the operand
  of pd_mlp.tensor_core = 2."""
  mat = np.asarray(mat, dtype=np.float32)
  hi = mat.astype(np.float16)
  lo = (mat - hi.astype(np.float32)).astype(np.float16)
  return np.concatenate([_umma_k_major(hi.view(np.uint16)),
                         _umma_k_major(lo.view(np.uint16))])


class RateSpec:
  """Selects the rate function for a stepping call (the reference's
  ``RateFunction`` seam, graphene.py:52-78) and owns device copies of the
  learned model's weights."""

  def __init__(self, kind: int, *, mlp: Optional[MlpWeights] = None,
               constant: Optional[Sequence[float]] = None, device=None,
               tensor_core: bool = False, gmm: Optional[dict] = None,
               prior: Optional[dict] = None):
    self.kind = int(kind)
    self.c = nat.PdRateConfig()
    self.c.rate_fn = self.kind
    self._mlp_c = None
    self._tensors = {}
    if prior is not None:
      # HumanPriorRatePredictor(mean, cov, max_rate), graphene.py:181-189
      if self.kind != nat.RATE_PRIOR:
        raise ValueError('prior parameters need RATE_PRIOR')
      mean = np.asarray(prior['mean'], dtype=np.float64).reshape(2)
      cov = np.asarray(prior['cov'], dtype=np.float64).reshape(2, 2)
      self._prior_c = nat.PdPrior()
      for i in range(2):
        self._prior_c.mean[i] = mean[i]
        for j in range(2):
          self._prior_c.cov[i][j] = cov[i, j]
      self._prior_c.max_rate = float(prior['max_rate'])
      self.c.prior = C.pointer(self._prior_c)
    if self.kind == nat.RATE_CONSTANT:
      if constant is None or len(constant) != 3:
        raise ValueError('RATE_CONSTANT needs three rates')
      for i, r in enumerate(constant):
        self.c.constant_rates[i] = float(r)
    if self.kind == nat.RATE_GMM:
      if gmm is None:
        raise ValueError('RATE_GMM needs the mixture parameters')
      w = np.asarray(gmm['mixture_weights'], dtype=np.float64).reshape(-1)
      loc = np.asarray(gmm['loc_distances'], dtype=np.float64).reshape(-1)
      var = np.asarray(gmm['variances'], dtype=np.float64).reshape(-1, 2)
      n = w.size
      if not (1 <= n <= nat.GMM_MAX_MIXTURES) or loc.size != n or len(var) != n:
        raise ValueError('inconsistent mixture parameters')
      self._gmm_c = nat.PdGmm()
      self._gmm_c.n_mixtures = n
      self._gmm_c.max_rate = float(gmm['max_rate'])
      for i in range(n):
        self._gmm_c.mixture_weights[i] = w[i]
        self._gmm_c.loc_distances[i] = loc[i]
        self._gmm_c.variances[i][0] = var[i, 0]
        self._gmm_c.variances[i][1] = var[i, 1]
      self.c.gmm = C.pointer(self._gmm_c)
    if self.kind == nat.RATE_LEARNED:
      if mlp is None:
        raise ValueError('RATE_LEARNED needs MlpWeights')
      device = _require_cuda(device)
      d, h1 = mlp.w0.shape
      h2 = mlp.w1.shape[1]
      if mlp.w1.shape[0] != h1 or mlp.w2.shape != (h2, 4):
        raise ValueError('inconsistent MLP shapes')
      for name in MlpWeights.NAMES:
        self._tensors[name] = torch.as_tensor(
            np.ascontiguousarray(getattr(mlp, name), dtype=np.float32),
            device=device)
      umma_ptr = None
      # tensor_core: False / 0 = FP32 FMA (the parity path); True / 1 = tcgen05
      # with bf16 operands (2e-2 of the largest rate); 2 = tcgen05 with fp16
      # hi + lo operands, three MMAs per K step (3e-7; hidden sizes <= 128)
      tc_mode = int(tensor_core)
      if tc_mode:
        w1t = np.asarray(mlp.w1, dtype=np.float32).T
        self._tensors['w1_umma'] = torch.as_tensor(
            umma_k_major_f16_split(w1t) if tc_mode == 2
            else umma_k_major_bf16(w1t), device=device)
        umma_ptr = self._tensors['w1_umma'].data_ptr()
      self._mlp_c = nat.PdMlp(d, h1, h2, int(mlp.batchnorm),
                              *[self._tensors[n].data_ptr()
                                for n in MlpWeights.NAMES],
                              tc_mode, 0, umma_ptr)
      self.c.mlp = C.pointer(self._mlp_c)
      self.mlp = mlp

  @classmethod
  def simple(cls):
    return cls(nat.RATE_SIMPLE)

  @classmethod
  def prior(cls, mean=None, cov=None, max_rate=None):
    """The human prior; with arguments, HumanPriorRatePredictor(mean, cov,
    max_rate) (graphene.py:181-189; evaluated by the float64 kernels)."""
    if mean is None and cov is None and max_rate is None:
      return cls(nat.RATE_PRIOR)
    return cls(nat.RATE_PRIOR, prior={
        'mean': (0.85, 0.0) if mean is None else mean,
        'cov': ((0.1, 0.0), (0.0, 0.1)) if cov is None else cov,
        'max_rate': np.log(2) / 3 if max_rate is None else max_rate})


@dataclasses.dataclass
class StepResult:
  """Device-resident outputs of one batched step."""
  elapsed_us: torch.Tensor  # int64 [E]
  transitions: torch.Tensor  # int32 [E]
  events: torch.Tensor  # int32 [E]
  recentred: torch.Tensor  # uint8 [E]
  si_xy: torch.Tensor  # float64 [E, 2] material frame
  log_count: Optional[torch.Tensor] = None  # int32 [E]
  log_elapsed_us: Optional[torch.Tensor] = None  # int64 [E, K]
  log_site: Optional[torch.Tensor] = None  # int32 [E, K]
  log_ctrl: Optional[torch.Tensor] = None  # int32 [E, K]


class EnvBatch:
  """Struct-of-arrays state of ``num_envs`` independent simulators in HBM."""

  STATE_FIELDS = ('si_idx', 'lattice', 'fov', 'fov_scale', 'image_params',
                  'episode', 'ctrl_count', 'frame_count', 'sim_time_us',
                  'n_events', 'n_transitions', 'status')

  def __init__(self, num_envs: int, *, seed: int = 0, grid_columns: int = 50,
               device=None, env_offset: int = 0,
               lattice: Optional[Lattice] = None, log_capacity: int = 0):
    self.device = _require_cuda(device)
    self.num_envs = int(num_envs)
    self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    self.env_offset = int(env_offset)
    self.lattice_tables = lattice or Lattice(grid_columns, self.device)
    self.log_capacity = int(log_capacity)
    e, dev = self.num_envs, self.device
    z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
    self.si_idx = z((e,), torch.int32)
    self.lattice = z((e, 4), torch.float64)
    self.fov = z((e, 4), torch.float64)
    self.fov_scale = z((e,), torch.float64)
    self.image_params = z((e, 9), torch.float64)
    self.episode = z((e,), torch.int32)  # bit pattern of uint32
    self.ctrl_count = z((e,), torch.int32)
    self.frame_count = z((e,), torch.int32)
    self.sim_time_us = z((e,), torch.int64)
    self.n_events = z((e,), torch.int64)
    self.n_transitions = z((e,), torch.int64)
    self.status = torch.full((e,), nat.ENV_NOT_RESET, dtype=torch.uint8,
                             device=dev)
    # Per-call outputs, allocated once.
    self._out = StepResult(
        elapsed_us=z((e,), torch.int64), transitions=z((e,), torch.int32),
        events=z((e,), torch.int32), recentred=z((e,), torch.uint8),
        si_xy=z((e, 2), torch.float64))
    if self.log_capacity > 0:
      k = self.log_capacity
      self._out.log_count = z((e,), torch.int32)
      self._out.log_elapsed_us = z((e, k), torch.int64)
      self._out.log_site = z((e, k), torch.int32)
      self._out.log_ctrl = z((e, k), torch.int32)
    self._refresh_c()

  # -- plumbing -------------------------------------------------------------
  def _refresh_c(self) -> None:
    self.c = nat.PdState(
        self.num_envs, self.seed, self.env_offset, 0,
        *[getattr(self, f).data_ptr() for f in self.STATE_FIELDS])
    o = self._out
    self._out_c = nat.PdStepOut(
        o.elapsed_us.data_ptr(), o.transitions.data_ptr(),
        o.events.data_ptr(), o.recentred.data_ptr(), o.si_xy.data_ptr(),
        self.log_capacity, 0,
        *[(t.data_ptr() if t is not None else None)
          for t in (o.log_count, o.log_elapsed_us, o.log_site, o.log_ctrl)])

  def _f64(self, x, shape) -> torch.Tensor:
    t = torch.as_tensor(x, dtype=torch.float64, device=self.device)
    return t.reshape(shape).contiguous()

  def _dwell(self, dwell_us, shape):
    """Returns (device int64 tensor or None, scalar)."""
    if isinstance(dwell_us, (int, np.integer)):
      return None, int(dwell_us)
    t = torch.as_tensor(dwell_us, dtype=torch.int64, device=self.device)
    return t.expand(shape).contiguous(), 0

  # -- simulator calls ------------------------------------------------------
  def reset(self, mask=None) -> None:
    m = None
    if mask is not None:
      m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
    with torch.cuda.device(self.device):
      nat.check(nat.lib.pd_reset(C.byref(self.lattice_tables.c),
                                 C.byref(self.c), _ptr(m),
                                 _stream(self.device)))

  def sample_image_params(self, noisy: bool = False, mask=None) -> None:
    """Re-draws the image parameters of the current episode on the device:
    imaging.py:42-54 ``sample_image_parameters`` (what reset applies) or
    :57-72 ``sample_noisy_image_parameters``."""
    m = None
    if mask is not None:
      m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
    with torch.cuda.device(self.device):
      nat.check(nat.lib.pd_sample_image_params(
          C.byref(self.c), _ptr(m), 1 if noisy else 0, _stream(self.device)))

  def rates(self, beam_xy, rate: RateSpec):
    beam = self._f64(beam_xy, (self.num_envs, 2))
    r = torch.empty((self.num_envs, 3), dtype=torch.float32,
                    device=self.device)
    nb = torch.empty((self.num_envs, 3), dtype=torch.int32,
                     device=self.device)
    with torch.cuda.device(self.device):
      nat.check(nat.lib.pd_rates(C.byref(self.lattice_tables.c),
                                 C.byref(self.c), C.byref(rate.c), _ptr(beam),
                                 _ptr(r), _ptr(nb), _stream(self.device)))
    return r, nb

  def apply_control(self, beam_xy, dwell_us, rate: RateSpec) -> StepResult:
    beam = self._f64(beam_xy, (self.num_envs, 2))
    d, scalar = self._dwell(dwell_us, (self.num_envs,))
    with torch.cuda.device(self.device):
      nat.check(nat.lib.pd_apply_control(
          C.byref(self.lattice_tables.c), C.byref(self.c), C.byref(rate.c),
          _ptr(beam), _ptr(d), scalar, C.byref(self._out_c),
          _stream(self.device)))
    return self._out

  def step_and_image(self, controls_xy, dwell_us, rate: RateSpec,
                     image_duration_us: int = 2000000) -> StepResult:
    ctl = torch.as_tensor(controls_xy, dtype=torch.float64,
                          device=self.device)
    if ctl.ndim == 2:
      ctl = ctl[:, None, :]
    if ctl.ndim != 3 or ctl.shape[0] != self.num_envs or ctl.shape[2] != 2:
      raise ValueError(f'controls must be [E, C, 2], got {tuple(ctl.shape)}')
    ctl = ctl.contiguous()
    n_controls = ctl.shape[1]
    d, scalar = self._dwell(dwell_us, (self.num_envs, n_controls))
    with torch.cuda.device(self.device):
      nat.check(nat.lib.pd_step_and_image(
          C.byref(self.lattice_tables.c), C.byref(self.c), C.byref(rate.c),
          _ptr(ctl), _ptr(d), scalar, n_controls, int(image_duration_us),
          C.byref(self._out_c), _stream(self.device)))
    return self._out

  def rollout(self, controls_xy, dwell_us: int, rate: RateSpec,
              image_duration_us: int = 2000000, record: bool = False,
              action_mode: int = nat.ACTION_DIRECT,
              max_distance_angstroms: float = 1.42):
    """``n_steps`` single-control steps fused in one launch.

    controls_xy: [T, E, 2]: beam positions in the microscope frame
    (ACTION_DIRECT) or actions in [-1, 1]^2 relative to the Si
    (ACTION_RELATIVE_TO_SILICON, action_adapters.py:131-216).  Returns
    (si_idx [T, E], elapsed_us [T, E]) if ``record`` else None.
    """
    ctl = torch.as_tensor(controls_xy, dtype=torch.float64,
                          device=self.device)
    if ctl.ndim != 3 or ctl.shape[1] != self.num_envs or ctl.shape[2] != 2:
      raise ValueError(f'controls must be [T, E, 2], got {tuple(ctl.shape)}')
    ctl = ctl.contiguous()
    t = ctl.shape[0]
    si = el = None
    if record:
      si = torch.empty((t, self.num_envs), dtype=torch.int32,
                       device=self.device)
      el = torch.empty((t, self.num_envs), dtype=torch.int64,
                       device=self.device)
    with torch.cuda.device(self.device):
      nat.check(nat.lib.pd_rollout_actions(
          C.byref(self.lattice_tables.c), C.byref(self.c), C.byref(rate.c),
          _ptr(ctl), int(action_mode), float(max_distance_angstroms),
          int(dwell_us), t, int(image_duration_us), _ptr(si), _ptr(el),
          _stream(self.device)))
    return (si, el) if record else None

  def rollout_host(self, actions_xy, dwell_us: int, rate: RateSpec,
                   image_duration_us: int = 2000000,
                   action_mode: int = nat.ACTION_DIRECT,
                   max_distance_angstroms: float = 1.42,
                   out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """`rollout` with HOST buffers in and out (the C ABI's
    pd_rollout_actions_host_f32 / pd_rollout_actions_host).

    actions_xy: [T, E, 2] NumPy array or CPU tensor; float32 (the dtype the
    action adapters' action_spec declares: 8 B per env-step over PCIe, int32
    elapsed microseconds back) or float64 (int64 elapsed back).  Page-locked
    input (``tensor.pin_memory()``) is used as is, anything else is copied
    into a pinned buffer first.  Returns (si_idx [T, E] int32, elapsed_us
    [T, E]) as pinned CPU tensors (``out`` re-uses a previous result).  Small
    batches on the prior / simple rates run as one launch that steps while
    the actions arrive and the results leave over PCIe.
    """
    act = torch.as_tensor(actions_xy)
    if act.device.type != 'cpu':
      raise ValueError('rollout_host takes host buffers; use rollout() for '
                       'device tensors')
    if act.dtype not in (torch.float32, torch.float64):
      act = act.to(torch.float64)
    if act.ndim != 3 or act.shape[1] != self.num_envs or act.shape[2] != 2:
      raise ValueError(f'actions must be [T, E, 2], got {tuple(act.shape)}')
    # int32 microseconds hold dwell + 2 image durations up to ~35 minutes;
    # beyond that the float32 actions are widened (exactly) and the call takes
    # the float64 / int64 entry point
    if (act.dtype == torch.float32 and
        int(dwell_us) + 2 * int(image_duration_us) >= 2 ** 31):
      act = act.to(torch.float64)
    act = act.contiguous()
    if not act.is_pinned():
      # one page-locked input buffer per (shape, dtype), re-used across calls
      key = (tuple(act.shape), act.dtype)
      if getattr(self, '_pinned_in_key', None) != key:
        self._pinned_in = torch.empty(act.shape, dtype=act.dtype).pin_memory()
        self._pinned_in_key = key
      self._pinned_in.copy_(act)
      act = self._pinned_in
    t, e = act.shape[0], self.num_envs
    wide = act.dtype == torch.float64
    el_dtype = torch.int64 if wide else torch.int32
    if out is not None:
      h_si, h_el = out
      if (h_si.shape != (t, e) or h_el.shape != (t, e) or
          h_si.dtype != torch.int32 or h_el.dtype != el_dtype or
          not h_si.is_pinned() or not h_el.is_pinned()):
        raise ValueError('out does not match this call')
    else:
      h_si = torch.empty((t, e), dtype=torch.int32).pin_memory()
      h_el = torch.empty((t, e), dtype=el_dtype).pin_memory()
    if t == 0:
      return h_si, h_el
    with torch.cuda.device(self.device):
      if wide:
        key = ('wide', t)
        if getattr(self, '_host_stage_key', None) != key:
          self._host_stage = (
              torch.empty((t, e, 2), dtype=torch.float64, device=self.device),
              torch.empty((t, e), dtype=torch.int32, device=self.device),
              torch.empty((t, e), dtype=torch.int64, device=self.device))
          self._host_stage_key = key
        d_ctl, d_si, d_el = self._host_stage
        nat.check(nat.lib.pd_rollout_actions_host(
            C.byref(self.lattice_tables.c), C.byref(self.c), C.byref(rate.c),
            _ptr(act), int(action_mode), float(max_distance_angstroms),
            int(dwell_us), t, int(image_duration_us), _ptr(d_ctl), _ptr(d_si),
            _ptr(d_el), _ptr(h_si), _ptr(h_el), _stream(self.device)))
      else:
        # stagings kept (and prepared for the next call) by the library
        nat.check(nat.lib.pd_rollout_actions_host_f32(
            C.byref(self.lattice_tables.c), C.byref(self.c), C.byref(rate.c),
            _ptr(act), int(action_mode), float(max_distance_angstroms),
            int(dwell_us), t, int(image_duration_us), None, None, None, None,
            None, _ptr(h_si), _ptr(h_el), _stream(self.device)))
    return h_si, h_el

  def rollout_host_packed(self, actions_xy, dwell_us: int, rate: RateSpec,
                          image_duration_us: int = 2000000,
                          action_mode: int = nat.ACTION_DIRECT,
                          max_distance_angstroms: float = 1.42,
                          out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`rollout_host` in the packed result format of
    pd_rollout_actions_host_packed: float32 actions [T, E, 2] in, one uint16
    per env-step out (Si site | re-centred << 15); `unpack_rollout` widens it
    to the (si_idx, elapsed_us) pair of `rollout_host`.  Prior / simple rates
    with one positive dwell time below 3000 s."""
    act = torch.as_tensor(actions_xy)
    if act.device.type != 'cpu':
      raise ValueError('rollout_host_packed takes host buffers')
    act = act.to(torch.float32)
    if act.ndim != 3 or act.shape[1] != self.num_envs or act.shape[2] != 2:
      raise ValueError(f'actions must be [T, E, 2], got {tuple(act.shape)}')
    act = act.contiguous()
    if not act.is_pinned():
      act = act.pin_memory()
    t, e = act.shape[0], self.num_envs
    if out is None:
      out = torch.empty((t, e), dtype=torch.uint16).pin_memory()
    elif out.shape != (t, e) or out.dtype != torch.uint16:
      raise ValueError('out does not match this call')
    if t == 0:
      return out
    with torch.cuda.device(self.device):
      nat.check(nat.lib.pd_rollout_actions_host_packed(
          C.byref(self.lattice_tables.c), C.byref(self.c), C.byref(rate.c),
          _ptr(act), int(action_mode), float(max_distance_angstroms),
          int(dwell_us), t, int(image_duration_us), _ptr(out),
          _stream(self.device)))
    return out

  @staticmethod
  def unpack_rollout(packed: torch.Tensor, dwell_us: int,
                     image_duration_us: int = 2000000):
    """(si_idx int32, elapsed_us int64) of a packed rollout result:
    elapsed = dwell + image duration * (1 + re-centred)
    (simulator.py:131-169)."""
    p = packed.numpy().astype(np.int32)
    si = p & 0x7FFF
    el = (int(dwell_us) + int(image_duration_us) * (1 + (p >> 15))).astype(
        np.int64)
    return torch.from_numpy(si), torch.from_numpy(el)

  # -- queries --------------------------------------------------------------
  def max_atoms_in_view(self) -> int:
    """Upper bound on atoms inside any current FOV (density 0.382 / A^2
    with margin for the boundary rows)."""
    w = (self.fov[:, 2] - self.fov[:, 0]).max().item()
    h = (self.fov[:, 3] - self.fov[:, 1]).max().item()
    return int(0.382 * (w + 3.0) * (h + 3.0)) + 16

  def get_atoms_in_bounds(self, fov=None, max_atoms: Optional[int] = None,
                          with_sites: bool = False):
    f = None if fov is None else self._f64(fov, (self.num_envs, 4))
    if max_atoms is None:
      if f is None:
        max_atoms = self.max_atoms_in_view()
      else:
        w = (f[:, 2] - f[:, 0]).max().item()
        h = (f[:, 3] - f[:, 1]).max().item()
        max_atoms = int(0.382 * (w + 3.0) * (h + 3.0)) + 16
      max_atoms = min(max_atoms, self.lattice_tables.n_sites)
    e, dev = self.num_envs, self.device
    xy = torch.zeros((e, max_atoms, 2), dtype=torch.float64, device=dev)
    z = torch.zeros((e, max_atoms), dtype=torch.uint8, device=dev)
    site = (torch.zeros((e, max_atoms), dtype=torch.int32, device=dev)
            if with_sites else None)
    count = torch.zeros((e,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
      nat.check(nat.lib.pd_get_atoms_in_bounds(
          C.byref(self.lattice_tables.c), C.byref(self.c), _ptr(f), max_atoms,
          _ptr(xy), _ptr(z), _ptr(site), _ptr(count), _stream(dev)))
    return (xy, z, count, site) if with_sites else (xy, z, count)

  def silicon_position(self) -> torch.Tensor:
    out = torch.empty((self.num_envs, 2), dtype=torch.float64,
                      device=self.device)
    with torch.cuda.device(self.device):
      nat.check(nat.lib.pd_get_silicon_position(
          C.byref(self.lattice_tables.c), C.byref(self.c), _ptr(out),
          _stream(self.device)))
    return out

  def grid_positions(self, env_ids) -> torch.Tensor:
    ids = torch.as_tensor(env_ids, dtype=torch.int32,
                          device=self.device).reshape(-1).contiguous()
    out = torch.empty((ids.numel(), self.lattice_tables.n_sites, 2),
                      dtype=torch.float64, device=self.device)
    with torch.cuda.device(self.device):
      nat.check(nat.lib.pd_get_grid(
          C.byref(self.lattice_tables.c), C.byref(self.c), _ptr(ids),
          ids.numel(), _ptr(out), _stream(self.device)))
    return out

  # -- checkpoint / resume --------------------------------------------------
  def state_dict(self) -> dict:
    d = {f: getattr(self, f).detach().cpu().clone() for f in self.STATE_FIELDS}
    d.update(seed=self.seed, env_offset=self.env_offset,
             grid_columns=self.lattice_tables.grid_columns)
    return d

  def load_state_dict(self, d: dict) -> None:
    if d['grid_columns'] != self.lattice_tables.grid_columns:
      raise ValueError('grid_columns mismatch')
    for f in self.STATE_FIELDS:
      getattr(self, f).copy_(d[f])
    self.seed, self.env_offset = int(d['seed']), int(d['env_offset'])
    self._refresh_c()
