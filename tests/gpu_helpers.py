"""Helpers shared by the GPU parity tests (all go through the C ABI)."""

import numpy as np
import torch

from oracle import pdune_oracle as po
from putting_dune_b200 import engine


def batch_from_oracle(st: po.OracleState, **kw) -> engine.EnvBatch:
  """A device batch whose state is a copy of the oracle's state."""
  b = engine.EnvBatch(st.num_envs, seed=st.seed,
                      env_offset=int(st.env_ids[0]), **kw)
  load_from_oracle(b, st)
  return b


def load_from_oracle(b: engine.EnvBatch, st: po.OracleState) -> None:
  dev = b.device
  b.si_idx.copy_(torch.as_tensor(st.si_idx, device=dev))
  b.lattice.copy_(torch.as_tensor(st.lattice, device=dev))
  b.fov.copy_(torch.as_tensor(st.fov, device=dev))
  b.fov_scale.copy_(torch.as_tensor(st.fov_scale, device=dev))
  b.image_params.copy_(torch.as_tensor(st.image_params, device=dev))
  b.episode.copy_(torch.as_tensor(st.episode.astype(np.int32), device=dev))
  b.ctrl_count.copy_(torch.as_tensor(st.ctrl_count.astype(np.int32),
                                     device=dev))
  b.frame_count.copy_(torch.as_tensor(st.frame_count.astype(np.int32),
                                      device=dev))
  b.sim_time_us.copy_(torch.as_tensor(st.sim_time_us, device=dev))
  b.n_events.copy_(torch.as_tensor(st.n_events, device=dev))
  b.n_transitions.copy_(torch.as_tensor(st.n_transitions, device=dev))
  b.status.zero_()


def rate_spec(rate_fn: int, mlp=None, constant=None,
              gmm=None) -> engine.RateSpec:
  if rate_fn == po.RATE_GMM:
    return engine.RateSpec(rate_fn, gmm=gmm)
  if rate_fn == po.RATE_LEARNED:
    w = engine.MlpWeights(**{k: getattr(mlp, k) for k in
                             engine.MlpWeights.NAMES},
                          batchnorm=mlp.batchnorm)
    return engine.RateSpec(rate_fn, mlp=w)
  return engine.RateSpec(rate_fn, constant=constant)


def closed_loop_control(st: po.OracleState, rng) -> np.ndarray:
  """Beam = Si position + U(-1, 1)^2 bond lengths, in the microscope frame."""
  e = st.num_envs
  p = po.site_positions(st, st.si_idx, np.arange(e))
  q = po.material_to_microscope(st.fov, p)
  return q + rng.uniform(-1, 1, size=(e, 2)) * po.BOND / st.fov_scale[:, None]


def np_(t: torch.Tensor) -> np.ndarray:
  return t.detach().cpu().numpy()
