"""CPU oracle for the STEM frame renderer (reference: putting_dune/imaging.py).

TEST INFRASTRUCTURE ONLY (same rules as pdune_oracle.py).

Parity status
  * PINNED against the unmodified reference (run through oracle/refshim.py):
    `generate_clean_image` (:117-173), `apply_blur` (:212-214),
    `apply_poisson_noise` (:199-203), `apply_jitter` (:188-196),
    `apply_uniform_noise` (:231-236), `apply_exponential_noise` (:221-228),
    under `RenderInjectedRng` (tests/golden/frames_reference.npz).
  * PARITY UNPINNED: the three skimage calls -- `random_noise` (gaussian and
    s&p, imaging.py:179-184,209), `exposure.adjust_gamma` (:218) and
    `exposure.equalize_adapthist` (:264).  scikit-image is not in this image
    and not in /root/reference; they are restated here from the published
    algorithm (scikit-image >= 0.19 `util/noise.py`, `exposure/exposure.py`,
    `exposure/_adapthist.py`).  The reference's own tests pin only shape and
    range for these stages (imaging_test.py:51-78).

Injected noise convention (one Philox4x32-10 call = four 32-bit words; counter
= (env, frame, index, stream)); u24(w) = (w >> 8) * 2**-24.  Every per-pixel
field is laid out the same way: pixel p = row * S + col (of the array at that
stage) takes word p % 4 of the call with index p // 4, so one call serves four
consecutive pixels of a row:

  stream 2 RENDER_POISSON: Poisson(image * mult) by inverse CDF on u24(w)
  stream 3 RENDER_SP:      salt&pepper flip: u24(w) <= amount; salt vs pepper
                           from the low byte of the same word (independent of
                           the 24 bits above): ((w & 255) + 0.5) / 256 <= 0.5
  stream 7 RENDER_UNIFORM: uniform noise: scale * u24(w)
  stream 8 RENDER_EXP:     exponential noise: -log1p(-u24(w)) * scale
  stream 9 RENDER_GAUSS:   Box-Muller pairs: words (0, 1) serve pixels 4g and
                           4g+1, words (2, 3) pixels 4g+2 and 4g+3:
                           r = sqrt(-2 ln(((wa >> 8) + 1) / 2**24)),
                           t = 2 pi u24(wb); first pixel r cos t, second r sin t
  stream 4 JITTER, index = row: w0 -> Poisson(jitter_rate) by inverse CDF.
"""

from __future__ import annotations

import numpy as np

from oracle import pdune_oracle as po

NR_OF_GRAY = 2 ** 14  # skimage exposure/_adapthist.py


def u24(w):
  return (np.asarray(w, dtype=np.uint32) >> np.uint32(8)).astype(
      np.float64) * (1.0 / 16777216.0)


def u24_open(w):
  """(0, 1] variant for logarithms."""
  return ((np.asarray(w, dtype=np.uint32) >> np.uint32(8)).astype(np.float64)
          + 1.0) * (1.0 / 16777216.0)


def poisson_icdf(lam, u):
  """Inverse-CDF Poisson: smallest k with CDF(k) > u (float64 recurrence)."""
  lam = np.asarray(lam, dtype=np.float64)
  u = np.broadcast_to(np.asarray(u, dtype=np.float64), lam.shape)
  k = np.zeros(lam.shape, dtype=np.int64)
  p = np.exp(-lam)
  cdf = p.copy()
  active = u >= cdf
  while active.any():
    k[active] += 1
    p[active] = p[active] * lam[active] / k[active]
    cdf[active] = cdf[active] + p[active]
    active = active & (u >= cdf) & (k < 100000)
  return k


class RenderInjectedRng:
  """Duck-typed rng for one `generate_stem_image` call (call-order keyed)."""

  def __init__(self, seed: int, env_id: int, frame: int, size: int = 512):
    self.seed, self.env, self.frame, self.size = seed, env_id, frame, size
    self._random_calls = 0
    self._cache = {}

  def _words(self, stream, n):
    idx = np.arange(n, dtype=np.uint64)
    return po.philox4x32_10(np.uint32(self.env), np.uint32(self.frame), idx,
                            stream, self.seed & 0xFFFFFFFF, self.seed >> 32)

  def _field(self, stream):
    """uint32 word per pixel, flat [S*S]: word p % 4 of call p // 4."""
    if stream not in self._cache:
      w = self._words(stream, self.size * self.size // 4)
      self._cache[stream] = np.stack(
          [np.asarray(x, dtype=np.uint32) for x in w], axis=1).reshape(-1)
    return self._cache[stream]

  def poisson(self, lam, size=None):
    if np.ndim(lam) == 2:  # apply_poisson_noise, imaging.py:202
      u = u24(self._field(po.STREAM_RENDER_POISSON)).reshape(lam.shape)
      return poisson_icdf(lam, u)
    # apply_jitter, imaging.py:192
    w = self._words(po.STREAM_JITTER, int(size))
    return poisson_icdf(np.full(int(size), float(lam)), u24(w[0]))

  def random(self, size=None):
    # skimage random_noise 's&p': two fields, flip then salt
    w = self._field(po.STREAM_RENDER_SP)
    first = self._random_calls % 2 == 0
    self._random_calls += 1
    if first:
      return u24(w).reshape(size)
    return (((w & np.uint32(255)).astype(np.float64) + 0.5) /
            256.0).reshape(size)

  def uniform(self, low=0.0, high=1.0, size=None):
    return low + (high - low) * u24(
        self._field(po.STREAM_RENDER_UNIFORM)).reshape(size)

  def exponential(self, scale=1.0, size=None):
    u = u24(self._field(po.STREAM_RENDER_EXP)).reshape(size)
    return -np.log1p(-u) * np.float64(scale)

  def normal(self, loc=0.0, scale=1.0, size=None):
    w = self._field(po.STREAM_RENDER_GAUSS).reshape(-1, 2)
    r = np.sqrt(-2.0 * np.log(u24_open(w[:, 0])))
    t = 2.0 * np.pi * u24(w[:, 1])
    z = np.stack((r * np.cos(t), r * np.sin(t)), axis=1).reshape(size)
    return loc + scale * z


# ----------------------------------------------------------------------------
# scipy.ndimage.gaussian_filter restated (imaging.py:161-163, :213)
# ----------------------------------------------------------------------------
def gaussian_kernel1d(sigma: float) -> np.ndarray:
  """scipy `_gaussian_kernel1d(order=0)`, radius int(4*sigma + 0.5)."""
  radius = int(4.0 * float(sigma) + 0.5)
  x = np.arange(-radius, radius + 1)
  phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
  return phi / phi.sum()


def correlate1d(img: np.ndarray, w: np.ndarray, axis: int, mode: str):
  r = (w.size - 1) // 2
  if r == 0:
    return img * w[0]
  pad = [(0, 0), (0, 0)]
  pad[axis] = (r, r)
  if mode == 'constant':
    p = np.pad(img, pad, mode='constant')
  else:  # scipy 'reflect' == numpy 'symmetric' (d c b a | a b c d | d c b a)
    p = np.pad(img, pad, mode='symmetric')
  out = np.zeros_like(img)
  n = img.shape[axis]
  for k in range(w.size):
    sl = [slice(None), slice(None)]
    sl[axis] = slice(k, k + n)
    out = out + w[k] * p[tuple(sl)]
  return out


def gaussian_filter(img: np.ndarray, sigma, mode: str) -> np.ndarray:
  sig = (sigma, sigma) if np.ndim(sigma) == 0 else tuple(sigma)
  out = img
  for axis in (0, 1):
    if sig[axis] > 1e-15:
      out = correlate1d(out, gaussian_kernel1d(sig[axis]), axis, mode)
  return out


# ----------------------------------------------------------------------------
# imaging.py stages
# ----------------------------------------------------------------------------
def clean_image(q: np.ndarray, z: np.ndarray, fov_w: float, fov_h: float,
                intensity_exponent: float, size: int = 512,
                buffer_size: float = 0.0) -> np.ndarray:
  """imaging.py:117-173.  q: positions [M, 2] in the microscope frame (atoms
  outside [-buffer, 1 + buffer] are dropped, like np.histogram2d does)."""
  bw = int(buffer_size * size)
  n = size + 2 * bw
  img = np.zeros((n, n))
  lo, hi = -buffer_size, 1 + buffer_size
  edges = np.linspace(lo, hi, n + 1)
  for number in sorted(set(int(v) for v in z)):
    sel = q[z == number]
    ok = ((sel[:, 0] >= lo) & (sel[:, 0] <= hi) & (sel[:, 1] >= lo) &
          (sel[:, 1] <= hi))
    sel = sel[ok]
    bx = np.searchsorted(edges, sel[:, 0], side='right') - 1
    by = np.searchsorted(edges, sel[:, 1], side='right') - 1
    bx[sel[:, 0] == edges[-1]] = n - 1
    by[sel[:, 1] == edges[-1]] = n - 1
    counts = np.zeros((n, n))
    np.add.at(counts, (bx, by), 1.0)
    img = img + counts * (np.int64(number) ** np.float64(intensity_exponent))
  img = np.flipud(np.transpose(img))
  sigma = (size / (2.15 * fov_w), size / (2.15 * fov_h))
  img = gaussian_filter(img, sigma, 'constant')
  img = img[bw:bw + size, bw:bw + size]
  return img / np.max(img)


def apply_blur(img, amount):
  img = gaussian_filter(img, amount, 'reflect')
  return img / np.max(img)


def apply_poisson_noise(img, mult, rng):
  img = rng.poisson(img * mult)
  return img / np.max(img)


def apply_jitter(img, rate, rng):
  roll = rng.poisson(rate, size=img.shape[0])
  return np.stack([np.roll(img[i], roll[i]) for i in range(img.shape[0])])


def random_noise_sp(img, amount, rng):
  """skimage.util.random_noise(mode='s&p', salt_vs_pepper=0.5, clip=True)."""
  img = np.asarray(img, dtype=np.float64)
  low_clip = -1.0 if img.min() < 0 else 0.0
  out = img.copy()

  def bernoulli(p):
    if p == 0:
      return np.zeros(img.shape, dtype=bool)
    if p == 1:
      return np.ones(img.shape, dtype=bool)
    return rng.random(img.shape) <= p

  flipped = bernoulli(amount)
  salted = bernoulli(0.5)
  out[flipped & salted] = 1.0
  out[flipped & ~salted] = low_clip
  return np.clip(out, low_clip, 1.0)


def adjust_gamma(img, gamma):
  """skimage.exposure.adjust_gamma for float images: scale 1, gain 1."""
  assert img.min() >= 0
  return (img / 1.0) ** gamma * 1.0 * 1.0


def apply_uniform_noise(img, scale, rng):
  img = img + rng.uniform(0.0, scale, size=img.shape)
  return img / np.max(img)


def apply_exponential_noise(img, scale, rng):
  img = img + rng.exponential(scale, size=img.shape)
  return img / np.max(img)


def random_noise_gaussian(img, var, rng):
  """skimage.util.random_noise(mode='gaussian', mean=0, var, clip=True)."""
  low_clip = -1.0 if img.min() < 0 else 0.0
  out = img + rng.normal(0.0, var ** 0.5, img.shape)
  return np.clip(out, low_clip, 1.0)


# ----------------------------------------------------------------------------
# skimage.exposure.equalize_adapthist restated (PARITY UNPINNED)
# ----------------------------------------------------------------------------
def rescale_intensity(img, out_max):
  imin, imax = img.min(), img.max()
  img = np.clip(img, imin, imax)
  if imin != imax:
    img = (img - imin) / (imax - imin)
    return img * out_max
  return np.clip(img, 0.0, out_max)


def clip_histogram(hist: np.ndarray, clip_limit: int) -> np.ndarray:
  hist = hist.copy()
  excess_mask = hist > clip_limit
  excess = hist[excess_mask]
  n_excess = int(excess.sum() - excess.size * clip_limit)
  hist[excess_mask] = clip_limit
  bin_incr = n_excess // hist.size
  upper = clip_limit - bin_incr
  low_mask = hist < upper
  n_excess -= int(hist[low_mask].size * bin_incr)
  hist[low_mask] += bin_incr
  mid_mask = np.logical_and(hist >= upper, hist < clip_limit)
  mid = hist[mid_mask]
  n_excess += int(mid.sum() - mid.size * clip_limit)
  hist[mid_mask] = clip_limit
  while n_excess > 0:
    prev = n_excess
    for index in range(hist.size):
      under = hist < clip_limit
      step = max(1, int(np.count_nonzero(under)) // n_excess)
      under = under[index::step]
      hist[index::step][under] += 1
      n_excess -= int(np.count_nonzero(under))
      if n_excess <= 0:
        break
    if prev == n_excess:
      break
  return hist


def map_histogram(hist: np.ndarray, n_pixels: int) -> np.ndarray:
  out = np.cumsum(hist, axis=-1).astype(float)
  out *= (NR_OF_GRAY - 1) / n_pixels
  np.clip(out, None, NR_OF_GRAY - 1, out=out)
  return out.astype(int)


def equalize_adapthist(img: np.ndarray, clip_limit: float = 0.01,
                       nbins: int = 256) -> np.ndarray:
  """CLAHE as scikit-image computes it for a 2-D float image with the default
  kernel (shape // 8).  Supports shapes divisible by the kernel."""
  size = img.shape[0]
  assert img.shape == (size, size)
  k = max(size // 8, 1)
  assert size % k == 0
  q = np.round(rescale_intensity(img.astype(np.float64),
                                 NR_OF_GRAY - 1)).astype(np.uint16)
  # pad: half a kernel before, ceil(k/2) after (shape is a kernel multiple)
  p0, p1 = k // 2, int(np.ceil(k / 2.0))
  q = np.pad(q, ((p0, p1), (p0, p1)), mode='reflect')
  bin_size = 1 + NR_OF_GRAY // nbins
  bins = (q // bin_size).astype(np.int64)
  n_t = size // k  # tiles per side of the unpadded image
  clim = int(np.clip(clip_limit * k * k, 1, None)) if clip_limit > 0 else k * k
  maps = np.zeros((n_t, n_t, nbins), dtype=np.int64)
  for ti in range(n_t):
    for tj in range(n_t):
      blk = bins[p0 + ti * k:p0 + (ti + 1) * k, p0 + tj * k:p0 + (tj + 1) * k]
      hist = np.bincount(blk.reshape(-1), minlength=nbins)
      maps[ti, tj] = map_histogram(clip_histogram(hist, clim), k * k)
  map_array = np.pad(maps, ((1, 1), (1, 1), (0, 0)), mode='edge')
  n_p = (size + p0 + p1) // k  # processing blocks per side
  result = np.zeros(bins.shape, dtype=np.float32)
  coef = np.arange(k) / k
  for bi in range(n_p):
    for bj in range(n_p):
      blk = bins[bi * k:(bi + 1) * k, bj * k:(bj + 1) * k]
      acc = np.zeros(blk.shape, dtype=np.float32)
      for er in (0, 1):
        for ec in (0, 1):
          mapped = map_array[bi + er, bj + ec][blk]
          wr = coef if er else 1 - coef
          wc = coef if ec else 1 - coef
          w = wr[:, None] * wc[None, :]
          acc += (mapped * w).astype(np.float32)
      result[bi * k:(bi + 1) * k, bj * k:(bj + 1) * k] = acc
  result = result.astype(np.uint16)[p0:p0 + size, p0:p0 + size]
  return rescale_intensity(result.astype(np.float64), 1.0)


# ----------------------------------------------------------------------------
# imaging.py:75-114 generate_grid_mask
# ----------------------------------------------------------------------------
def grid_mask(q: np.ndarray, z: np.ndarray, fov: np.ndarray,
              intensity_exponent: float = 1.7, size: int = 512) -> np.ndarray:
  """Semantic mask of the atoms in view.  q: normalised positions [M, 2]
  (microscope frame), fov: (ll_x, ll_y, ur_x, ur_y).  Quirk kept: the SQUARED
  distance in angstrom^2 is compared with the radius (imaging.py:109-110)."""
  xs = np.linspace(fov[0], fov[2], size + 1, endpoint=True)
  xs = (xs[:-1] + xs[1:]) / 2
  ys = np.linspace(fov[1], fov[3], size + 1, endpoint=True)
  ys = (ys[:-1] + ys[1:]) / 2
  xx, yy = np.meshgrid(xs, ys)
  # microscope_utils.py:362-369 (grid branch): p * (ur - ll) + ll
  pos = q * np.array([fov[2] - fov[0], fov[3] - fov[1]]) + np.array(
      [fov[0], fov[1]])
  mask = np.zeros((size, size), dtype=np.uint8)
  for p, number in zip(pos, z):
    radius = (number / 6) ** intensity_exponent * 0.1
    distance = (xx - p[0]) ** 2.0 + (yy - p[1]) ** 2.0
    mask[distance < radius] = number
  return np.flipud(mask)


def mask_env(state: po.OracleState, env: int, size: int = 512,
             intensity_exponent: float = 1.7) -> np.ndarray:
  q, z, _ = po.get_atoms_in_bounds(state, env)
  return grid_mask(q, z, state.fov[env], intensity_exponent, size)


# ----------------------------------------------------------------------------
# imaging.py:239-265 generate_stem_image
# ----------------------------------------------------------------------------
def generate_stem_image(q, z, fov_w, fov_h, params, rng, size: int = 512,
                        stages: bool = False, buffer_size: float = 0.0):
  """params: the 9 values in dataclass order (po.IMAGE_PARAM_NAMES)."""
  (exponent, gauss_var, jitter_rate, poisson_mult, sp_amount, blur_amount,
   gamma, exp_lambda, uniform_scale) = [float(v) for v in params]
  out = {}
  img = clean_image(q, z, fov_w, fov_h, exponent, size, buffer_size)
  out['clean'] = img
  img = apply_blur(img, blur_amount)
  out['blur'] = img
  img = apply_poisson_noise(img, poisson_mult, rng)
  out['poisson'] = img
  img = apply_jitter(img, jitter_rate, rng)
  out['jitter'] = img
  img = random_noise_sp(img, sp_amount, rng)
  img = adjust_gamma(img, gamma)
  img = apply_uniform_noise(img, uniform_scale, rng)
  out['uniform'] = img
  img = apply_exponential_noise(img, exp_lambda, rng)
  out['exponential'] = img
  img = random_noise_gaussian(img, gauss_var, rng)
  out['gaussian'] = img
  img = equalize_adapthist(img, clip_limit=0.01)
  out['final'] = img
  return out if stages else img


def grid_in_microscope_frame(state: po.OracleState, env: int):
  """The whole material grid in the microscope frame of the env's FOV
  (microscope_utils.py:421-428): what a caller hands generate_stem_image when
  it renders with a buffer."""
  f = state.fov[env]
  p = po.all_positions(state, env)
  q = np.stack(((p[:, 0] - f[0]) / (f[2] - f[0]),
                (p[:, 1] - f[1]) / (f[3] - f[1])), axis=1)
  z = np.full(p.shape[0], po.CARBON, dtype=np.int64)
  z[state.si_idx[env]] = po.SILICON
  return q, z


def render_env(state: po.OracleState, env: int, size: int = 512,
               stages: bool = False, buffer_size: float = 0.0):
  """simulator.py:206-221 `_generate_image` for one env of the oracle state;
  advances the env's frame counter.  buffer_size > 0: imaging.py:129-168 with
  the whole grid as input."""
  if buffer_size > 0:
    q, z = grid_in_microscope_frame(state, env)
  else:
    q, z, _ = po.get_atoms_in_bounds(state, env)
  f = state.fov[env]
  rng = RenderInjectedRng(state.seed, int(state.env_ids[env]),
                          int(state.frame_count[env]), size)
  res = generate_stem_image(q, z, f[2] - f[0], f[3] - f[1],
                            state.image_params[env], rng, size, stages,
                            buffer_size)
  state.frame_count[env] += np.uint32(1)
  return res
