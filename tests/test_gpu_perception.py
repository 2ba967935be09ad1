"""GPU parity tests (through the C ABI) of the perception-data pieces:
imaging.py:75-114 generate_grid_mask (bit-exact uint8 masks),
imaging.py:57-72 sample_noisy_image_parameters (bit-exact float64 values) and
frames rendered with the noisy parameter ranges."""

import os

import numpy as np
import pytest
import torch

from oracle import pdune_oracle as po
from oracle import pdune_oracle_imaging as oi
from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu


def test_masks_match_reference_golden(golden_dir):
  from putting_dune_b200 import engine, imaging
  fix = np.load(os.path.join(golden_dir, 'perception_reference.npz'))
  seed, size = int(fix['seed']), int(fix['size'])
  n = fix['noisy_params'].shape[0]
  b = engine.EnvBatch(n, seed=seed)
  b.reset()
  for e in range(n):
    got = gh.np_(imaging.generate_grid_mask_batch(
        b, env_ids=[e], image_size=size,
        intensity_exponent=float(fix[f'mask_exponent_{e}'])))[0]
    np.testing.assert_array_equal(got, fix[f'mask_{e}'])
  # the reference's own test (imaging_test.py:80-106): C, Si and background
  full = gh.np_(imaging.generate_grid_mask_batch(b, env_ids=[0]))[0]
  assert full.shape == (512, 512) and full.dtype == np.uint8
  assert set(full.reshape(-1).tolist()) == {0, 6, 14}


@pytest.mark.parametrize('size', [64, 512])
def test_masks_match_oracle_after_steps(size):
  from putting_dune_b200 import imaging
  n, seed = 12, 91
  st = po.make_state(n, seed)
  po.reset(st)
  b = gh.batch_from_oracle(st)
  rng = np.random.default_rng(1)
  spec = gh.rate_spec(po.RATE_SIMPLE)
  for _ in range(5):  # move the Si and the FOV
    ctl = gh.closed_loop_control(st, rng)[:, None, :]
    po.step_and_image(st, ctl, 5000000)
    b.step_and_image(ctl, 5000000, spec)
  ids = [0, 5, 11]
  got = gh.np_(imaging.generate_grid_mask_batch(b, env_ids=ids,
                                                image_size=size))
  for j, e in enumerate(ids):
    np.testing.assert_array_equal(got[j], oi.mask_env(st, e, size))
  # batch form, per-env exponent of the image parameters
  allm = gh.np_(imaging.generate_grid_mask_batch(b, image_size=size,
                                                 intensity_exponent=1.4))
  np.testing.assert_array_equal(allm[7], oi.mask_env(st, 7, size, 1.4))


def test_noisy_image_parameters(golden_dir):
  from putting_dune_b200 import engine, imaging
  fix = np.load(os.path.join(golden_dir, 'perception_reference.npz'))
  n = fix['noisy_params'].shape[0]
  b = engine.EnvBatch(n, seed=int(fix['seed']))
  b.reset()
  default = gh.np_(b.image_params).copy()
  mask = np.zeros(n, dtype=bool)
  mask[::2] = True
  b.sample_image_params(noisy=True, mask=mask)
  got = gh.np_(b.image_params)
  np.testing.assert_array_equal(got[::2], fix['noisy_params'][::2])
  np.testing.assert_array_equal(got[1::2], default[1::2])
  b.sample_image_params(noisy=True)
  np.testing.assert_array_equal(gh.np_(b.image_params), fix['noisy_params'])
  b.sample_image_params(noisy=False)
  np.testing.assert_array_equal(gh.np_(b.image_params), default)
  # host mirror: same ranges (imaging.py:57-72)
  p = imaging.sample_noisy_image_parameters(np.random.default_rng(0))
  assert 0.0 <= p.gaussian_variance <= 0.3 and 0.5 <= p.contrast_gamma <= 1.5


def test_noisy_frames_match_oracle():
  """The renderer with the noisy ranges (gaussian variance up to 0.3, s&p up
  to 1e-2): stage tolerances as in test_gpu_render.py."""
  from putting_dune_b200 import imaging
  from tests.test_gpu_render import STAGES, _check_stage
  n, seed, size = 3, 23, 128
  st = po.make_state(n, seed)
  po.reset(st)
  po.sample_noisy_image_parameters(st)
  b = gh.batch_from_oracle(st)
  want = [oi.render_env(st, e, size=size, stages=True) for e in range(n)]
  for k, name in enumerate(STAGES):
    got = gh.np_(imaging.render_batch(b, image_size=size, stop_stage=k,
                                      advance_frame_count=False))
    for e in range(n):
      _check_stage(name, got[e], want[e][name])


def test_buffered_clean_image(golden_dir):
  """imaging.py:129-168 generate_clean_image(buffer_size > 0): atoms outside
  the frame contribute their tails.  |d| <= 2e-6 like the unbuffered stage."""
  from putting_dune_b200 import engine, imaging
  fix = np.load(os.path.join(golden_dir, 'perception_reference.npz'))
  seed, size = int(fix['seed']), int(fix['size'])
  b = engine.EnvBatch(6, seed=seed)
  b.reset()
  for e in (0, 1, 2):
    buf = float(fix[f'buffer_{e}'])
    got = gh.np_(imaging.render_batch(b, env_ids=[e], image_size=size,
                                      stop_stage=0, buffer_size=buf,
                                      advance_frame_count=False))[0]
    d = np.abs(got.astype(np.float64) - fix[f'buffered_clean_{e}'])
    assert d.max() <= 2e-6, (e, d.max())
  # full size, all stages, against the oracle; and the buffer matters
  st = po.make_state(2, 19)
  po.reset(st)
  bb = gh.batch_from_oracle(st)
  want = oi.render_env(st, 1, size=512, stages=True, buffer_size=0.1)
  from tests.test_gpu_render import STAGES, _check_stage
  for k, name in enumerate(STAGES):
    got = gh.np_(imaging.render_batch(bb, env_ids=[1], image_size=512,
                                      stop_stage=k, buffer_size=0.1,
                                      advance_frame_count=False))[0]
    _check_stage(name, got, want[name])
  plain = gh.np_(imaging.render_batch(bb, env_ids=[1], image_size=512,
                                      stop_stage=0,
                                      advance_frame_count=False))[0]
  assert np.abs(plain - want['clean']).max() > 1e-3
