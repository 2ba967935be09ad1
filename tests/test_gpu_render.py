"""GPU parity tests of the STEM renderer (through the C ABI) against the
imaging oracle and the frames the reference's imaging.py produced.

Tolerances (float32 pipeline vs the reference's float64):
  * clean / blur: |d| <= 2e-6 absolute on [0, 1] images.
  * poisson / jitter: integer counts / max -- identical except where the
    float32 rate sits on an inverse-CDF threshold: <= 0.1 % of pixels may
    differ, the rest agree to 1e-6.
  * uniform / exponential / gaussian: <= 0.1 % of pixels outside 2e-5.
  * final (CLAHE, parity unpinned): 14-bit quantisation + tile histograms
    amplify float32 rounding into bin flips; mean |d| <= 2e-3 and 99 % of
    pixels within 2e-2.
"""

import os

import numpy as np
import pytest
import torch

from oracle import pdune_oracle as po
from oracle import pdune_oracle_imaging as oi
from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu

STAGES = ['clean', 'blur', 'poisson', 'jitter', 'uniform', 'exponential',
          'gaussian', 'final']


def _check_stage(name, got, want):
  d = np.abs(got.astype(np.float64) - want)
  if name in ('clean', 'blur'):
    assert d.max() <= 2e-6, (name, d.max())
  elif name in ('poisson', 'jitter'):
    assert (d > 1e-6).mean() <= 1e-3, (name, (d > 1e-6).mean())
  elif name in ('uniform', 'exponential', 'gaussian'):
    assert (d > 2e-5).mean() <= 1e-3, (name, (d > 2e-5).mean(), d.max())
  else:
    # CLAHE itself agrees with the oracle on every pixel to 7e-8 when both
    # quantise the same input (test_clahe_matches_oracle_on_the_device_input).
    # Here the inputs are the float32 and the float64 chain: a pixel whose
    # 14-bit level sits on a bin edge lands in the other histogram bin, which
    # moves that tile's map by a few grey levels.  Measured over the frames
    # of this file: mean <= 6e-6, <= 4e-3 of the pixels beyond 1e-4, <= 1.1e-3
    # beyond 2e-3, none beyond 2e-2.
    assert d.mean() <= 5e-5, (name, d.mean())
    assert (d > 1e-4).mean() <= 1e-2, (name, (d > 1e-4).mean())
    assert (d > 2e-3).mean() <= 3e-3, (name, (d > 2e-3).mean())
    assert (d > 2e-2).mean() <= 1e-4, (name, (d > 2e-2).mean(), d.max())
    assert got.min() >= 0.0 and got.max() <= 1.0  # imaging_test.py:75-78


@pytest.mark.parametrize('size', [128, 512])
def test_render_stages_match_oracle(size):
  from putting_dune_b200 import imaging
  n, seed = (6 if size == 128 else 2), 77
  st = po.make_state(n, seed)
  po.reset(st)
  # exercise the parameter corners
  st.image_params[0, 5] = 0.0  # no blur at all
  st.image_params[1, 5] = 0.1  # radius-0 blur
  st.image_params[0, 3] = 120.0  # large Poisson rates
  b = gh.batch_from_oracle(st)
  want = [oi.render_env(st, e, size=size, stages=True) for e in range(n)]
  for k, name in enumerate(STAGES):
    got = gh.np_(imaging.render_batch(b, image_size=size, stop_stage=k,
                                      advance_frame_count=False))
    assert got.shape == (n, size, size)
    for e in range(n):
      _check_stage(name, got[e], want[e][name])
  # the frame counter selects fresh noise
  f0 = gh.np_(imaging.render_batch(b, image_size=size))
  f1 = gh.np_(imaging.render_batch(b, image_size=size))
  assert (gh.np_(b.frame_count) == 2).all()
  assert np.abs(f0 - f1).mean() > 1e-3
  st2 = po.make_state(n, seed)
  po.reset(st2)
  st2.image_params[:] = st.image_params
  st2.frame_count[:] = 1
  _check_stage('final', f1[n - 1], oi.render_env(st2, n - 1, size=size))


def test_benchmarked_render_config_sampled_vs_oracle():
  """BASELINE configs[3] as bench.py runs it: one pd_render launch over
  16,384 frames of 512 x 512; 16 randomly chosen frames of it against the
  imaging oracle, stage by stage."""
  from putting_dune_b200 import imaging
  n, seed = 16384, 3
  st = po.make_state(n, seed)
  po.reset(st)
  b = gh.batch_from_oracle(st)
  pick = np.sort(np.random.default_rng(11).choice(n, size=16, replace=False))
  want = {int(e): oi.render_env(st, int(e), size=512, stages=True)
          for e in pick}
  out = torch.empty((n, 512, 512), dtype=torch.float32, device=b.device)
  for k, name in enumerate(STAGES):
    imaging.render_batch(b, image_size=512, stop_stage=k, out=out,
                         advance_frame_count=False)
    got = gh.np_(out[torch.as_tensor(pick, device=b.device)])
    for j, e in enumerate(pick):
      _check_stage(name, got[j], want[int(e)][name])
  del out
  torch.cuda.empty_cache()


def test_render_matches_reference_golden_frames(golden_dir):
  from putting_dune_b200 import engine, imaging
  fix = np.load(os.path.join(golden_dir, 'frames_reference.npz'))
  seed, size = int(fix['seed']), int(fix['size'])
  b = engine.EnvBatch(4, seed=seed)
  b.reset()
  for k, name in enumerate(STAGES):
    if name not in ('clean', 'blur', 'poisson', 'jitter', 'final'):
      continue
    got = gh.np_(imaging.render_batch(b, image_size=size, stop_stage=k,
                                      advance_frame_count=False))
    for e in range(4):
      _check_stage(name, got[e], fix[f'{name}_{e}'].astype(np.float64))
  full = gh.np_(imaging.render_batch(b, env_ids=[0], image_size=512))[0]
  assert full.shape == (512, 512)
  np.testing.assert_allclose(full.mean(axis=1), fix['full512_rowmean'],
                             atol=5e-3)


def test_render_subset_and_after_steps():
  """Frames follow the simulator state: FOV re-centre and the Si site."""
  from putting_dune_b200 import imaging
  n, seed = 40, 5
  st = po.make_state(n, seed)
  po.reset(st)
  b = gh.batch_from_oracle(st)
  rng = np.random.default_rng(0)
  spec = gh.rate_spec(po.RATE_SIMPLE)
  for _ in range(6):
    ctl = gh.closed_loop_control(st, rng)[:, None, :]
    po.step_and_image(st, ctl, 5000000)
    b.step_and_image(ctl, 5000000, spec)
  ids = [3, 17, 39]
  got = gh.np_(imaging.render_batch(b, env_ids=ids, image_size=128,
                                    stop_stage=1, advance_frame_count=False))
  for j, e in enumerate(ids):
    want = oi.render_env(st, e, size=128, stages=True)['blur']
    _check_stage('blur', got[j], want)
  assert (gh.np_(b.frame_count) == 0).all()


def test_simulator_facade_returns_image():
  # simulator_test.py:337-354: image returned, mean in [0, 1].
  import datetime as dt
  import putting_dune_b200 as pd
  from putting_dune_b200 import geometry, graphene, microscope_utils as mu
  sim = pd.PuttingDuneSimulator(graphene.PristineSingleDopedGraphene())
  obs = sim.reset(np.random.default_rng(0), return_image=True)
  assert obs.image.shape == (512, 512)
  assert 0.0 <= obs.image.mean() <= 1.0
  ctl = mu.BeamControl(geometry.Point(0.5, 0.5), dt.timedelta(seconds=1.5))
  obs = sim.step_and_image(np.random.default_rng(0), [ctl], return_image=True)
  assert obs.image.shape == (512, 512) and obs.image.min() >= 0.0


@pytest.mark.parametrize('size', [128, 512])
def test_clahe_matches_oracle_on_the_device_input(size):
  """CLAHE alone (imaging.py:264 exposure.equalize_adapthist): the oracle's
  restatement applied to the device's own stage-6 frame (float32, read back)
  against the device's final frame.  Both sides then quantise the same
  numbers, so the 14-bit levels, tile histograms, clip redistribution, maps
  and the bilinear blend have to agree pixel for pixel; what is left is the
  rounding of the final value to float32 (measured: <= 7e-8 on every
  pixel)."""
  from putting_dune_b200 import imaging
  n, seed = (6 if size == 128 else 3), 91
  st = po.make_state(n, seed)
  po.reset(st)
  st.image_params[0, 3] = 120.0  # large Poisson rates
  b = gh.batch_from_oracle(st)
  src = gh.np_(imaging.render_batch(b, image_size=size, stop_stage=6,
                                    advance_frame_count=False))
  got = gh.np_(imaging.render_batch(b, image_size=size, stop_stage=7,
                                    advance_frame_count=False))
  for e in range(n):
    want = oi.equalize_adapthist(src[e].astype(np.float64))
    d = np.abs(got[e].astype(np.float64) - want)
    print(f'size {size} env {e}: max {d.max():.3g} mean {d.mean():.3g} '
          f'>1e-6 {(d > 1e-6).mean():.3g} >1e-4 {(d > 1e-4).mean():.3g}')
    assert d.max() <= 2.5e-7, (e, d.max())
