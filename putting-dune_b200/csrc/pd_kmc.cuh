// Device-side building blocks of one kinetic-Monte-Carlo control
// (graphene.py:646-694 PristineSingleDopedGraphene.apply_control and the rate
// functions it calls).  Included by pd_step.cu and pd_mlp.cu.
//
// Arithmetic contract: geometry -> rate in float64
// with every operation individually rounded (translation units including
// this header are compiled with -fmad=false and use explicit *_rn
// intrinsics); rates cast to float32; total rate = sequential float32 sum;
// waiting time in float64; integer-microsecond clock.
#pragma once

#include "pd_common.cuh"

namespace pd {

// ---------------------------------------------------------------------------
// Lattice tables: either straight from global memory through L1 (small
// batches: staging 45 KB per CTA would cost more than the step) or staged in
// shared memory (large batches: random 16-byte gathers are ~3x cheaper from
// shared memory than as 32 separate L1 wavefronts).
// ---------------------------------------------------------------------------
struct GlobalTables {
  const double2* base;
  const int4* nbr;
  __device__ __forceinline__ double2 position(int k) const {
    return __ldg(base + k);
  }
  __device__ __forceinline__ void neighbors(int k, int out[3]) const {
    const int4 v = __ldg(nbr + k);
    out[0] = v.x;
    out[1] = v.y;
    out[2] = v.z;
  }
  // neighbours + geometry class of site k (pd_lattice.cu)
  __device__ __forceinline__ int neighbors_class(int k, int out[3]) const {
    const int4 v = __ldg(nbr + k);
    out[0] = v.x;
    out[1] = v.y;
    out[2] = v.z;
    return (v.w >> kSiteClassShift) & 3;
  }
};

struct SharedTables {
  const double2* base;   // shared memory
  const ushort4* nbr;    // shared memory
  __device__ __forceinline__ double2 position(int k) const { return base[k]; }
  __device__ __forceinline__ void neighbors(int k, int out[3]) const {
    const ushort4 v = nbr[k];
    out[0] = v.x;
    out[1] = v.y;
    out[2] = v.z;
  }
  __device__ __forceinline__ int neighbors_class(int k, int out[3]) const {
    const ushort4 v = nbr[k];
    out[0] = v.x;
    out[1] = v.y;
    out[2] = v.z;
    return v.w;
  }
};

__device__ __forceinline__ size_t shared_tables_bytes(int n_sites) {
  return static_cast<size_t>(n_sites) * (sizeof(double2) + sizeof(ushort4));
}

// Cooperative stage of the lattice tables into dynamic shared memory.
__device__ __forceinline__ SharedTables stage_tables(const pd_lattice& lat,
                                                     unsigned char* smem) {
  double2* sbase = reinterpret_cast<double2*>(smem);
  ushort4* snbr = reinterpret_cast<ushort4*>(sbase + lat.n_sites);
  const double2* gbase = reinterpret_cast<const double2*>(lat.base_xy);
  const int4* gnbr = reinterpret_cast<const int4*>(lat.nbr);
  for (int k = threadIdx.x; k < lat.n_sites; k += blockDim.x) {
    sbase[k] = __ldg(gbase + k);
    const int4 v = __ldg(gnbr + k);
    snbr[k] = make_ushort4(static_cast<unsigned short>(v.x),
                           static_cast<unsigned short>(v.y),
                           static_cast<unsigned short>(v.z),
                           static_cast<unsigned short>(
                               (v.w >> kSiteClassShift) & 3));
  }
  __syncthreads();
  return SharedTables{sbase, snbr};
}

// Rate-function parameters that travel by value to the kernels.
// Internal rate id: PD_RATE_PRIOR with caller-supplied parameters
// (pd_rate_config.prior); the float64 kernels only.
constexpr int kRatePriorGeneral = 100;

struct RateArgs {
  float constant_rates[3];
  int32_t race_sampling;  // 1: kmc_event_race instead of the direct method
  // HumanPriorRatePredictor(mean, cov, max_rate): mean, the precision matrix
  // cov^-1 as (p00, p01 + p10, p11), max_rate
  int32_t prior_general;
  double prior_mean[2];
  double prior_prec[3];
  double prior_max_rate;
  // PD_RATE_GMM, precomputed on the host: coef = normalising factor * weight
  // / (2 pi sqrt(v1 v2)); nh_inv_v = -0.5 / variance
  int32_t gmm_n;
  double gmm_coef[PD_GMM_MAX_MIXTURES];
  double gmm_loc[PD_GMM_MAX_MIXTURES];
  double gmm_nh_inv_v[PD_GMM_MAX_MIXTURES][2];
};

// ---------------------------------------------------------------------------
// Rate functions -> float32[3]
//
// The reference evaluates the rate in float64 and casts to float32
// (graphene.py:256); everything downstream (total rate, waiting-time scale,
// branch probabilities) is a function of those float32 values only.  The
// default forms below drop the square root and all but one division; the
// simple rate falls back to the reference's operation sequence whenever the
// short form's float32 cast could differ from it (cast_margin_ulps), so its
// float32 value is the reference's in every case; -DPD_EXACT_RATE_OPS uses
// the operation sequence everywhere.  pd_rate_ops_audit compares the forms.
// ---------------------------------------------------------------------------
// One neighbour of graphene.py:133-166 simple_canonical_rate_function:
//   r = 1 / ((4 |beam - nbr| / 1.42)^2 + 1) = 1 / (|beam - nbr|^2 16/1.42^2 + 1)
// Op for op as NumPy evaluates it (np.linalg.norm(axis=-1) = sqrt(dx*dx +
// dy*dy)), in float64.
__device__ __forceinline__ double rate_simple_ops(const double2 beam,
                                                  const double2 psi,
                                                  const double2 p) {
  const double bx = __dsub_rn(beam.x, psi.x);
  const double by = __dsub_rn(beam.y, psi.y);
  const double nx = __dsub_rn(p.x, psi.x);
  const double ny = __dsub_rn(p.y, psi.y);
  const double dx = __dsub_rn(bx, nx);
  const double dy = __dsub_rn(by, ny);
  double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
  d = __ddiv_rn(d, kBond);
  const double a = __dmul_rn(d, 4.0);
  return __ddiv_rn(1.0, __dadd_rn(__dmul_rn(a, a), 1.0));
}

// The same without the square root and with one division.
__device__ __forceinline__ double rate_simple_short(const double2 beam,
                                                    const double2 p) {
  const double kScale = 16.0 / (kBond * kBond);
  const double dx = beam.x - p.x;
  const double dy = beam.y - p.y;
  return 1.0 / fma(fma(dx, dx, dy * dy), kScale, 1.0);
}

// Distance, in units of the float64 ulp, of a float64 value from the nearest
// point where its float32 cast changes (the midpoints of float32's grid).
__device__ __forceinline__ unsigned cast_margin_ulps(double r) {
  const unsigned low =
      static_cast<unsigned>(__double_as_longlong(r)) & 0x1FFFFFFFu;
  return low > 0x10000000u ? low - 0x10000000u : 0x10000000u - low;
}

// The two float64 forms differ by the roundings of the coordinate
// differences: <= ~2e-13 relative for coordinates up to a few hundred
// angstrom (measured by pd_rate_ops_audit: < 2^11 ulp).  Unless the short
// form lands within kCastGuardUlps of a float32 rounding boundary (6e-5 of
// the evaluations) its cast IS the cast of the op-for-op value; otherwise the
// op-for-op form decides.  The simple rate is pure float64 NumPy upstream, so
// this is the reference's float32 value bit for bit.
constexpr unsigned kCastGuardUlps = 1u << 14;

__device__ __forceinline__ float rate_simple_one(const double2 beam,
                                                 const double2 psi,
                                                 const double2 p) {
#ifdef PD_EXACT_RATE_OPS
  return __double2float_rn(rate_simple_ops(beam, psi, p));
#else
  const double r = rate_simple_short(beam, p);
  if (cast_margin_ulps(r) < kCastGuardUlps || !(r > 1e-37))
    return __double2float_rn(rate_simple_ops(beam, psi, p));
  return __double2float_rn(r);
#endif
}

__device__ __forceinline__ void rates_simple(const double2 beam,
                                             const double2 psi,
                                             const double2 pn[3], float r[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) r[i] = rate_simple_one(beam, psi, pn[i]);
}

// One neighbour of graphene.py:191-229 HumanPriorRatePredictor.predict with
// the constants of constants.py:26-28.  The reference rotates the mean
// (0.85, 0) by rotate_coordinates(mean, -theta) with theta = atan2 of the
// neighbour, which places the peak at 0.85*(cos theta, -sin theta) (mirror
// quirk); cos/sin of atan2 are taken directly from the
// neighbour vector.  With covariance 0.1*I,
//   max_rate * pdf(x)/pdf(mu) = (ln 2 / 3) * exp(-5 |x - mu|^2).
__device__ __forceinline__ double rate_prior_ops(const double2 beam,
                                                 const double2 psi,
                                                 const double2 p) {
  // operation for operation (float64; the reference evaluates the pdf ratio
  // through jax.scipy.stats in float32, which no float64 form reproduces)
  const double kMaxRate = 0.23104906018664842;  // np.log(2) / 3
  const double x = __ddiv_rn(__dsub_rn(beam.x, psi.x), kBond);
  const double y = __ddiv_rn(__dsub_rn(beam.y, psi.y), kBond);
  const double nx = __dsub_rn(p.x, psi.x);
  const double ny = __dsub_rn(p.y, psi.y);
  const double inv = __ddiv_rn(
      0.85, __dsqrt_rn(__dadd_rn(__dmul_rn(nx, nx), __dmul_rn(ny, ny))));
  const double dx = __dsub_rn(x, __dmul_rn(nx, inv));
  const double dy = __dadd_rn(y, __dmul_rn(ny, inv));
  const double maha =
      __ddiv_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), 0.1);
  return __dmul_rn(kMaxRate, exp(__dmul_rn(-0.5, maha)));
}

__device__ __forceinline__ double rate_prior_short(const double2 beam,
                                                   const double2 psi,
                                                   const double2 p) {
  const double kMaxRate = 0.23104906018664842;  // np.log(2) / 3
  const double kInvBond = 1.0 / kBond;
  const double x = (beam.x - psi.x) * kInvBond;
  const double y = (beam.y - psi.y) * kInvBond;
  const double nx = p.x - psi.x;
  const double ny = p.y - psi.y;
  const double inv = 0.85 * rsqrt(fma(nx, nx, ny * ny));
  const double dx = fma(-nx, inv, x);
  const double dy = fma(ny, inv, y);
  return kMaxRate * exp(-5.0 * fma(dx, dx, dy * dy));
}

// The human prior is Gaussian in float32 upstream (JAX): its float32 value is
// defined up to the tolerance north_star states for rates, not to the bit,
// so the short float64 form stands as is; pd_rate_ops_audit reports how often
// its cast differs from the op-for-op form's (~1e-7 of the evaluations).
__device__ __forceinline__ float rate_prior_one(const double2 beam,
                                                const double2 psi,
                                                const double2 p) {
#ifdef PD_EXACT_RATE_OPS
  return __double2float_rn(rate_prior_ops(beam, psi, p));
#else
  return __double2float_rn(rate_prior_short(beam, psi, p));
#endif
}

// graphene.py:191-229 with caller-supplied mean / cov / max_rate:
//   mu_i = rotate_coordinates(mean, -theta_i)
//        = (mx cos t + my sin t, -mx sin t + my cos t),
//   r_i = max_rate * pdf(x; mu_i, cov) / pdf(mu_i; mu_i, cov)
//       = max_rate * exp(-(x - mu_i)^T cov^-1 (x - mu_i) / 2).
__device__ __forceinline__ void rates_prior_general(const RateArgs& ra,
                                                    const double2 beam,
                                                    const double2 psi,
                                                    const double2 pn[3],
                                                    float r[3]) {
  const double x = __ddiv_rn(__dsub_rn(beam.x, psi.x), kBond);
  const double y = __ddiv_rn(__dsub_rn(beam.y, psi.y), kBond);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double nx = pn[i].x - psi.x, ny = pn[i].y - psi.y;
    const double inv = rsqrt(fma(nx, nx, ny * ny));
    const double c = nx * inv, s = ny * inv;  // cos, sin of theta_i
    const double mux = fma(ra.prior_mean[0], c, ra.prior_mean[1] * s);
    const double muy = fma(ra.prior_mean[1], c, -ra.prior_mean[0] * s);
    const double dx = x - mux, dy = y - muy;
    const double maha = fma(ra.prior_prec[0] * dx, dx,
                            fma(ra.prior_prec[1] * dx, dy,
                                ra.prior_prec[2] * dy * dy));
    r[i] = __double2float_rn(ra.prior_max_rate * exp(-0.5 * maha));
  }
}

__device__ __forceinline__ void rates_prior(const double2 beam,
                                            const double2 psi,
                                            const double2 pn[3], float r[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) r[i] = rate_prior_one(beam, psi, pn[i]);
}

// graphene.py:303-388 GaussianMixtureRateFunction.__call__ (float64 rates).
// The covariance E diag(v) E^-1 has the unit Si->neighbour vector and its
// normal as eigenvectors, so pdf = exp(-0.5 (d1^2/v1 + d2^2/v2)) /
// (2 pi sqrt(v1 v2)) with d1, d2 the beam offset from the mean along them.
__device__ __forceinline__ void rates_gmm(const RateArgs& ra,
                                          const double2 beam,
                                          const double2 psi,
                                          const double2 pn[3], double r[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double nx = pn[i].x - psi.x, ny = pn[i].y - psi.y;
    const double inv_len = rsqrt(fma(nx, nx, ny * ny));
    const double e1x = nx * inv_len, e1y = ny * inv_len;
    double acc = 0.0;
    for (int m = 0; m < ra.gmm_n; ++m) {
      const double dx = beam.x - fma(nx, ra.gmm_loc[m], psi.x);
      const double dy = beam.y - fma(ny, ra.gmm_loc[m], psi.y);
      const double d1 = fma(dx, e1x, dy * e1y);
      const double d2 = fma(dy, e1x, -dx * e1y);
      acc = fma(ra.gmm_coef[m],
                exp(fma(d1 * d1, ra.gmm_nh_inv_v[m][0],
                        d2 * d2 * ra.gmm_nh_inv_v[m][1])),
                acc);
    }
    r[i] = acc;
  }
}

// ---------------------------------------------------------------------------
// One event of the direct-method loop (graphene.py:658-694).
// Returns true if a transition was accepted; *slot is the successor index.
// ---------------------------------------------------------------------------
__device__ __forceinline__ long long seconds_to_us(double t) {
  // dt.timedelta(seconds=t): whole seconds * 10^6 + fractional part * 10^6
  // rounded half-to-even (CPython datetime delta_new).
  const double whole = trunc(t);
  const double frac = __dsub_rn(t, whole);
  return static_cast<long long>(whole) * 1000000LL +
         static_cast<long long>(rint(__dmul_rn(frac, 1e6)));
}

__device__ __forceinline__ bool kmc_event_drawn(const float r[3], double draw,
                                                double u_choice,
                                                long long dwell_us,
                                                long long* elapsed_us,
                                                int* slot, bool* bad_rate);

__device__ __forceinline__ bool kmc_event(const float r[3], double u_exp,
                                          double u_choice, long long dwell_us,
                                          long long* elapsed_us, int* slot,
                                          bool* bad_rate) {
  return kmc_event_drawn(r, -log1p(-u_exp), u_choice, dwell_us, elapsed_us,
                         slot, bad_rate);
}

// `draw` is the unit exponential variate -log1p(-u) (state independent, so
// it can be produced by another lane or ahead of time).
__device__ __forceinline__ bool kmc_event_drawn(const float r[3], double draw,
                                                double u_choice,
                                                long long dwell_us,
                                                long long* elapsed_us,
                                                int* slot, bool* bad_rate) {
  // assert (transition_rates >= 0).all()  (graphene.py:258)
  *bad_rate = !(r[0] >= 0.f) || !(r[1] >= 0.f) || !(r[2] >= 0.f);
  // Rates.total_rate: sequential float32 sum (graphene.py:47-49).
  const float tot = __fadd_rn(__fadd_rn(r[0], r[1]), r[2]);
  double t = kMaxTransitionSeconds;
  if (tot > 0.f) {
    // rng.exponential(scale=1.0 / total): float32 scale under NumPy 2.
    const float scale = __fdiv_rn(1.0f, tot);
    t = __dmul_rn(draw, static_cast<double>(scale));
    t = fmin(t, kMaxTransitionSeconds);  // graphene.py:668
  }
  *elapsed_us += seconds_to_us(t);
  if (*elapsed_us > dwell_us) return false;  // graphene.py:677
  // rng.choice(3, p=rates/total): float32 p, float64 normalised CDF,
  // searchsorted(side='right').
  const double c0 = static_cast<double>(__fdiv_rn(r[0], tot));
  const double c1 = __dadd_rn(c0, static_cast<double>(__fdiv_rn(r[1], tot)));
  const double c2 = __dadd_rn(c1, static_cast<double>(__fdiv_rn(r[2], tot)));
  *slot = (__ddiv_rn(c0, c2) <= u_choice) + (__ddiv_rn(c1, c2) <= u_choice);
  return true;
}

// Variant for rate functions that return float64 rates
// (GaussianMixtureRateFunction): the total and the waiting-time scale stay
// float64 (Python floats, graphene.py:47-49,666); the branch probabilities
// are float32(rate) / float32(total) (graphene.py:679-683 under NumPy 2).
__device__ __forceinline__ bool kmc_event_drawn64(const double r64[3],
                                                  double draw, double u_choice,
                                                  long long dwell_us,
                                                  long long* elapsed_us,
                                                  int* slot) {
  const double tot = __dadd_rn(__dadd_rn(r64[0], r64[1]), r64[2]);
  double t = kMaxTransitionSeconds;
  if (tot > 0.0) {
    t = __dmul_rn(draw, __ddiv_rn(1.0, tot));
    t = fmin(t, kMaxTransitionSeconds);
  }
  *elapsed_us += seconds_to_us(t);
  if (*elapsed_us > dwell_us) return false;
  const float tot32 = __double2float_rn(tot);
  const double c0 =
      static_cast<double>(__fdiv_rn(__double2float_rn(r64[0]), tot32));
  const double c1 = __dadd_rn(
      c0, static_cast<double>(__fdiv_rn(__double2float_rn(r64[1]), tot32)));
  const double c2 = __dadd_rn(
      c1, static_cast<double>(__fdiv_rn(__double2float_rn(r64[2]), tot32)));
  *slot = (__ddiv_rn(c0, c2) <= u_choice) + (__ddiv_rn(c1, c2) <= u_choice);
  return true;
}

// ---------------------------------------------------------------------------
// Float32 pre-pass: "does this iteration certainly end its control without a
// hop?"
//
// An iteration whose waiting time overshoots the dwell time changes nothing
// but the event and control counters (graphene.py:677: elapsed > dwell ->
// break), and with the prior / simple rates that is ~70 % of all iterations.
// Deciding it does not need the float64 chain: a float32 evaluation with
// *one-sided* bounds -- an upper bound of the total rate and a lower bound of
// the unit-exponential draw -- gives a lower bound of the waiting time; if
// even that overshoots, the exact computation does too.  Everything else
// (hops, near misses, non-finite inputs) falls through to the float64
// iteration, so results are unchanged bit for bit
// (tests: test_prepass_equals_exact).
//
// Error budget (relative): beam offset 2e-6 A absolute -> 1e-4 on a rate that
// matters; __expf 1e-5; the sum 2e-7; folded into tot * 1.002 + 1e-30.  The
// draw uses the top 24 bits of the Philox word the exact path turns into u53
// (truncation: u24 <= u53, and 1 - u24 is exact in float32), __logf absolute
// error 4e-7: draw * 0.999 - 1e-6.  Waiting time * 0.998 against the remaining
// microseconds + 2 covers the division, the float32 scale of the exact path
// and the microsecond rounding.
// ---------------------------------------------------------------------------
template <int RATE>
struct PrepassGeo {
  // PD_RATE_PRIOR: peak positions 0.85 * (cos, -sin) of the three neighbour
  // directions (bond units); PD_RATE_SIMPLE: neighbour offsets (angstrom).
  float cx[3], cy[3];
};

template <int RATE, class Tables>
__device__ __forceinline__ void prepass_geometry(const Tables& tab, int si,
                                                 const Lattice4& lat,
                                                 PrepassGeo<RATE>* g) {
  int nb[3];
  tab.neighbors(si, nb);
  const double2 b0 = tab.position(si);
  const float c = static_cast<float>(lat.c), s = static_cast<float>(lat.s);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double2 bi = tab.position(nb[i]);
    const float dx = static_cast<float>(bi.x - b0.x);
    const float dy = static_cast<float>(bi.y - b0.y);
    const float nx = dx * c + dy * s;  // graphene.py:545-557
    const float ny = dy * c - dx * s;
    if (RATE == PD_RATE_PRIOR) {
      const float inv = 0.85f * rsqrtf(nx * nx + ny * ny);
      g->cx[i] = nx * inv;
      g->cy[i] = -ny * inv;  // mirror quirk, see rates_prior
    } else {
      g->cx[i] = nx;
      g->cy[i] = ny;
    }
  }
}

// bx, by: beam - Si in angstrom; word: Philox word x of the iteration;
// rem_us: dwell - elapsed (> 0).
template <int RATE>
__device__ __forceinline__ bool certainly_no_hop(const PrepassGeo<RATE>& g,
                                                 float bx, float by,
                                                 uint32_t word,
                                                 long long rem_us) {
  float tot = 0.f;
  if (RATE == PD_RATE_PRIOR) {
    const float x = bx * (1.0f / 1.42f), y = by * (1.0f / 1.42f);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float dx = x - g.cx[i], dy = y - g.cy[i];
      tot += 0.23104906f * __expf(-5.0f * (dx * dx + dy * dy));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float dx = bx - g.cx[i], dy = by - g.cy[i];
      tot += __fdividef(
          1.0f, (dx * dx + dy * dy) * (16.0f / (1.42f * 1.42f)) + 1.0f);
    }
  }
  const float u24 = static_cast<float>(word >> 8) * (1.0f / 16777216.0f);
  const float draw_lb = -__logf(1.0f - u24) * 0.999f - 1e-6f;
  const float t_lb =
      fminf(__fdividef(draw_lb, tot * 1.002f + 1e-30f), 3600.0f);
  // tot >= 0 is false for NaN inputs: those take the exact path (and its
  // PD_ENV_BAD_RATE flag).
  return (tot >= 0.f) && (t_lb * 0.998e6f > __ll2float_ru(rem_us) + 2.0f);
}

// Frame transforms: microscope_utils.py:362-369 / :421-428.
__device__ __forceinline__ double2 microscope_to_material(const Fov4& f,
                                                          double px,
                                                          double py) {
  return make_double2(
      __dadd_rn(__dmul_rn(px, __dsub_rn(f.urx, f.llx)), f.llx),
      __dadd_rn(__dmul_rn(py, __dsub_rn(f.ury, f.lly)), f.lly));
}

// simulator.py:230-250 _silicon_outside_of_safe_area on the observed grid of
// graphene.py:600-644 (inclusive bounds, normalised coordinates).
__device__ __forceinline__ bool silicon_outside_safe_area(const Fov4& f,
                                                          const double2 p) {
  const bool in_view = (f.llx <= p.x) && (p.x <= f.urx) && (f.lly <= p.y) &&
                       (p.y <= f.ury);
  const double qx = __ddiv_rn(__dsub_rn(p.x, f.llx), __dsub_rn(f.urx, f.llx));
  const double qy = __ddiv_rn(__dsub_rn(p.y, f.lly), __dsub_rn(f.ury, f.lly));
  const bool near_edge = (qx < 0.25) || (qx > 0.75) || (qy < 0.25) ||
                         (qy > 0.75);
  return !in_view || near_edge;
}

// Site ids are row-major over the horizontal rows of the base lattice
// (graphene.py:483-499), so the rows that can intersect a FOV's circumscribed
// circle form one contiguous id range: [*c_lo, *c_hi) in 32-site chunks.  A
// conservative bound (margin 0.05 A); callers still test every site.
__device__ __forceinline__ void fov_chunk_range(const pd_lattice& lat,
                                                const Lattice4& t,
                                                const Fov4& f, int* c_lo,
                                                int* c_hi) {
  const double w = f.urx - f.llx, h = f.ury - f.lly;
  const double cx = 0.5 * (f.llx + f.urx), cy = 0.5 * (f.lly + f.ury);
  const double by = cx * t.s + cy * t.c - t.oy;  // base-frame y of the centre
  const double rad = 0.5 * sqrt(w * w + h * h) + 0.05;
  const double y0 = __ldg(reinterpret_cast<const double2*>(lat.base_xy)).y;
  const double hrow = 0.8660254037844386 * kBond;
  const int ce = lat.n_cols - (lat.n_cols + 2) / 3;  // sites on even rows
  const int co = lat.n_cols - (lat.n_cols + 1) / 3;  // sites on odd rows
  const int n_chunks = (lat.n_sites + 31) / 32;
  const double lo = floor((by - rad - y0) / hrow);
  const double hi = ceil((by + rad - y0) / hrow) + 1.0;
  if (!(lo == lo) || !(hi == hi)) {  // non-finite FOV: scan everything
    *c_lo = 0;
    *c_hi = n_chunks;
    return;
  }
  const double rows = 1.0e6;  // ids are clamped to n_sites below
  const int j_lo = static_cast<int>(fmin(fmax(lo, 0.0), rows));
  const int j_hi = static_cast<int>(fmin(fmax(hi, 0.0), rows));
  long long k_lo =
      static_cast<long long>(j_lo / 2) * (ce + co) + (j_lo & 1) * ce;
  long long k_hi =
      static_cast<long long>(j_hi / 2) * (ce + co) + (j_hi & 1) * ce;
  if (k_lo > lat.n_sites) k_lo = lat.n_sites;
  if (k_hi > lat.n_sites) k_hi = lat.n_sites;
  *c_lo = static_cast<int>(k_lo / 32);
  *c_hi = static_cast<int>((k_hi + 31) / 32);
}

// In-view run of one lattice row.  Along a row the base y is fixed and the
// site position (graphene.py:545-557) is linear in the base x, so the four
// inclusive bounds of graphene.py:600-644 cut out one interval of base x; the
// sites of a row are 1.23 A or more apart, so only the first and the last
// site of the run can sit within rounding distance of a bound and those two
// are tested with the exact expression; the run itself comes from the column
// arithmetic of the lattice, widened by 1e-6 A.  Returns the run [*m_lo, *m_hi] as
// positions inside the row (empty if *m_lo > *m_hi).  Requires |c|, |s| >=
// 1e-6 (the caller falls back to the exhaustive scan otherwise).
__device__ __forceinline__ void row_run_in_view(const double2* base, int k0,
                                                int cnt, int row, int n_cols,
                                                const Lattice4& t,
                                                const Fov4& f, int* m_lo,
                                                int* m_hi) {
  const double Y = __ldg(base + k0).y + t.oy;
  // llx <= X c + Y s <= urx ; lly <= Y c - X s <= ury, X = base x + ox
  // (only an estimate, widened by kSlack: reciprocals instead of divisions)
  const double inv_c = 1.0 / t.c, inv_s = 1.0 / t.s;  // loop invariant
  double a0 = (f.llx - Y * t.s) * inv_c, a1 = (f.urx - Y * t.s) * inv_c;
  double b0 = (Y * t.c - f.ury) * inv_s, b1 = (Y * t.c - f.lly) * inv_s;
  if (a0 > a1) { const double tmp = a0; a0 = a1; a1 = tmp; }
  if (b0 > b1) { const double tmp = b0; b0 = b1; b1 = tmp; }
  const double kSlack = 1e-6;
  const double xa = fmax(a0, b0) - t.ox - kSlack;
  const double xb = fmin(a1, b1) - t.ox + kSlack;
  // Column i of a row sits at base x = (i + 0.5 [odd rows]) * 1.42 + x0 with
  // x0 = base[0].x - 1.42 (site 0 is column 1 of row 0); even rows lack the
  // columns i % 3 == 0, odd rows i % 3 == 1 (graphene.py:483-499).  The table
  // differs from this formula by ~1e-14, far inside the slack.
  const int odd = row & 1;
  const double x0 = __ldg(base).x - kBond + (odd ? 0.5 * kBond : 0.0);
  const double kInvBond = 1.0 / kBond;
  const double ia = ceil((xa - x0) * kInvBond);
  const double ib = floor((xb - x0) * kInvBond);
  int first = cnt, last = -1;
  if (ia <= ib && ib >= 0.0 && ia <= static_cast<double>(n_cols - 1)) {
    int i_lo = ia < 0.0 ? 0 : static_cast<int>(ia);
    int i_hi = ib > static_cast<double>(n_cols - 1) ? n_cols - 1
                                                    : static_cast<int>(ib);
    const int gap = odd ? 1 : 0;  // residue of the missing columns
    if (i_lo % 3 == gap) ++i_lo;
    if (i_hi % 3 == gap) --i_hi;
    if (i_lo <= i_hi) {
      // position inside the row of a present column
      first = odd ? i_lo - (i_lo + 1) / 3 : (i_lo - 1) - (i_lo - 1) / 3;
      last = odd ? i_hi - (i_hi + 1) / 3 : (i_hi - 1) - (i_hi - 1) / 3;
      if (last > cnt - 1) last = cnt - 1;
    }
  }
  auto in_view = [&](int m) {
    const double2 p = site_position(__ldg(base + k0 + m), t);
    return f.llx <= p.x && p.x <= f.urx && f.lly <= p.y && p.y <= f.ury;
  };
  if (first <= last && !in_view(first)) ++first;
  if (first <= last && !in_view(last)) --last;
  *m_lo = first;
  *m_hi = last;
}

// simulator.py:161-165: FOV = [P_si - s/2, P_si + s/2].
__device__ __forceinline__ Fov4 centred_fov(const double2 p, double scale) {
  const double h = __ddiv_rn(scale, 2.0);
  return Fov4{__dsub_rn(p.x, h), __dsub_rn(p.y, h), __dadd_rn(p.x, h),
              __dadd_rn(p.y, h)};
}


// ---------------------------------------------------------------------------
// Arguments and per-env registers shared by the stepping kernels.
// ---------------------------------------------------------------------------
#ifndef PD_STEP_THREADS
#define PD_STEP_THREADS 128
#endif
constexpr int kStepThreads = PD_STEP_THREADS;


// Per-env registers carried through a call.
struct EnvRegs {
  int si;
  double2 psi;
  Lattice4 lat;
  uint32_t env_id;
  uint32_t ctrl_count;
  int transitions;
  int events;
  int log_n;
  uint8_t status;
};

struct LogSink {
  int capacity;
  int64_t* elapsed_us;
  int32_t* site;
  int32_t* ctrl;
};

// Philox4x32-10 round keys: they depend on the seed only, so the host computes
// them once per launch and the kernels read them as constant-bank operands.
struct PhiloxKeys {
  uint32_t k0[10], k1[10];
};

inline PhiloxKeys philox_keys_host(uint64_t seed) {
  PhiloxKeys k;
  uint32_t a = static_cast<uint32_t>(seed);
  uint32_t b = static_cast<uint32_t>(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    k.k0[r] = a;
    k.k1[r] = b;
    a += 0x9E3779B9u;
    b += 0xBB67AE85u;
  }
  return k;
}

struct StepArgs {
  pd_lattice lat;
  pd_state st;
  RateArgs ra;
  PhiloxKeys keys;            // of st.seed (filled by launch_step)
  const double* controls_xy;  // [n][C][2] or [T][n][2] (rollout)
  const int64_t* dwell_us;    // [n][C] or null
  int64_t dwell_us_scalar;
  int32_t n_controls;
  int32_t n_steps;            // rollout only
  int64_t image_duration_us;
  int32_t material_frame;     // 1: apply_control (no observe phase)
  const uint8_t* skip;        // [n] or null: envs that sit this call out
  int32_t lane_stride;        // 1 thread in `lane_stride` owns an env (k_step /
                              // k_rollout; small batches trade idle lanes for
                              // more warps and less intra-warp divergence)
  int32_t action_mode;        // pd_action_mode (rollouts)
  int32_t plan_envs_per_cta;  // k_rollout_plan: envs a CTA owns (<= 16)
  // k_walk_plan -> k_walk_fast<LIST>: (env, first step still to do) of the
  // envs that left the plan, and how many there are
  int2* defer_list;
  uint32_t* defer_count;
  uint32_t* list_hint;        // mapped host word: the list's length, or null
  float fast_dwell_s, fast_margin;  // FastTimes of dwell_us_scalar (pd_fast.cuh)
  int32_t fast_episode;       // k_walk<EPISODE>: guarded float32 iterations
  int32_t prepass;            // 1: float32 pre-pass (certainly_no_hop) enabled
  int32_t walk_min_ready;     // k_walk: lanes with an exact iteration pending
  int32_t walk_max_reps;      //   that end the bookkeeping repeats / their cap
  int32_t walk_controls_per_pass;  // controls a lane may settle per pass
  int32_t walk_lockstep;      // episodes: bookkeeping only between controls
  double max_distance;        // RelativeToSilicon adapter, angstroms
  pd_step_out out;
  int32_t* si_idx_out;        // rollout [T][n]
  int64_t* elapsed_us_out;    // rollout [T][n]
  uint16_t* packed_out;       // rollout [T][n]: Si site | re-centred << 15
                              // (fast kernels, with actions_f32 as the input)
  // streamed host rollout (k_rollout_pre<.., STREAM>): the float32 actions
  // are still arriving (one copy-engine H2D copy into a staging pre-filled
  // with 0xFF) while the launch runs, and the CTAs of a few SMs write the
  // int32 results back to the caller's pinned buffers row by row
  const float2* actions_f32;  // [T][n] device staging; an element whose words
                              // are not both 0xFFFFFFFF has arrived
  const uint32_t* copy_done;  // set behind the H2D copy: whatever is in the
                              // staging now is data
  int32_t* elapsed32_out;     // [T][n] int32 microseconds (device staging);
                              // both result stagings start as 0xFF bytes and
                              // a word is final once it is not -1
  int32_t* h_si_idx_out;      // [T][n] pinned host memory or null
  int32_t* h_elapsed32_out;   // [T][n] pinned host memory or null
  int64_t* h_elapsed64_out;   // stream_mode 2: [T][n] pinned host memory or null
  int32_t stream_mode;        // 0 off; 1 float32 actions (actions_f32) / int32
                              // elapsed; 2 float64 actions (controls_xy is the
                              // staging) / int64 elapsed (elapsed_us_out)
  uint32_t* sm_ctl;           // role election and work tickets (pd_step.cu)
  int32_t copy_sms;           // SMs that only run writer CTAs
  int32_t step_ctas;          // blocks of kStepThreads stepping lanes
  int32_t stream_wave;        // CTAs of the launch (one full wave)
  unsigned long long* trace;  // PD_HOST_TRACE: globaltimer marks, else null
  // episode mode (pd_run_episodes): controls come from the greedy controller
  pd_episode_config ep;
  const double* goal_xy;      // [n][2] material frame
  pd_episode_stats* stats;    // [n]
};

__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// Pulls the HBM lines an environment's step will read into L1/L2 ahead of
// time (no register, no scoreboard): si_idx, lattice, fov, ctrl_count.
__device__ __forceinline__ void prefetch_env(const StepArgs& a, int64_t e) {
  prefetch_l1(a.st.si_idx + e);
  prefetch_l1(a.st.lattice + 4 * e);
  prefetch_l1(a.st.fov + 4 * e);
  prefetch_l1(a.st.ctrl_count + e);
}

template <class Tables>
__device__ __forceinline__ EnvRegs load_env(const Tables& tab,
                                            const StepArgs& a, int64_t e) {
  EnvRegs r;
  r.si = a.st.si_idx[e];
  r.lat = load_lattice4(a.st.lattice, e);
  r.psi = site_position(tab.position(r.si), r.lat);
  r.env_id = a.st.env_offset + static_cast<uint32_t>(e);
  r.ctrl_count = a.st.ctrl_count[e];
  r.transitions = 0;
  r.events = 0;
  r.log_n = 0;
  r.status = a.st.status[e];
  return r;
}

// ---------------------------------------------------------------------------
// The exact (float64) control: shared by every stepping kernel.
// ---------------------------------------------------------------------------
template <int RATE>
__device__ __forceinline__ void eval_rates(const RateArgs& ra,
                                           const double2 beam,
                                           const double2 psi,
                                           const double2 pn[3], float r[3]) {
  if (RATE == PD_RATE_SIMPLE) {
    rates_simple(beam, psi, pn, r);
  } else if (RATE == PD_RATE_PRIOR) {
    rates_prior(beam, psi, pn, r);
  } else if (RATE == kRatePriorGeneral) {
    rates_prior_general(ra, beam, psi, pn, r);
  } else if (RATE == PD_RATE_GMM) {
    double r64[3];
    rates_gmm(ra, beam, psi, pn, r64);
    r[0] = __double2float_rn(r64[0]);
    r[1] = __double2float_rn(r64[1]);
    r[2] = __double2float_rn(r64[2]);
  } else {
    r[0] = ra.constant_rates[0];
    r[1] = ra.constant_rates[1];
    r[2] = ra.constant_rates[2];
  }
}

// One event by the race of competing exponentials (first-reaction method):
// every neighbour draws its own waiting time Exp(rate_i), the smallest one
// happens.  Equal in distribution to the direct method above (waiting time
// Exp(total), neighbour i with probability rate_i / total) but not draw for
// draw, so this is an opt-in sampling mode (pd_set_option "race_sampling"),
// never the parity path.  The three uniforms are the 53-bit waiting-time
// uniform of the direct method and the two 26-bit halves of its choice
// uniform (+ half a step, so that none is 0).
__device__ __forceinline__ bool kmc_event_race(const double r[3], double u_exp,
                                               double u_choice,
                                               long long dwell_us,
                                               long long* elapsed_us,
                                               int* slot) {
  const double scaled = u_choice * 67108864.0;  // 2^26
  const double hi = floor(scaled);
  const double u[3] = {u_exp, (hi + 0.5) * (1.0 / 67108864.0),
                       (scaled - hi) * (1.0 - 1.0 / 134217728.0) +
                           1.0 / 268435456.0};
  double t = kMaxTransitionSeconds * 4.0;
  int best = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double ti = r[i] > 0.0 ? -log1p(-u[i]) / r[i]
                                 : kMaxTransitionSeconds * 4.0;
    if (ti < t) {
      t = ti;
      best = i;
    }
  }
  *elapsed_us += seconds_to_us(fmin(t, kMaxTransitionSeconds));
  if (*elapsed_us > dwell_us) return false;
  *slot = best;
  return true;
}

// Rates + one event of the direct method.  Float64-rate functions (GMM) keep
// the total in float64 (see kmc_event_drawn64).
template <int RATE>
__device__ __forceinline__ bool rate_event(const RateArgs& ra,
                                           const double2 beam,
                                           const double2 psi,
                                           const double2 pn[3], double u_exp,
                                           double u_choice, long long dwell_us,
                                           long long* elapsed_us, int* slot,
                                           bool* bad) {
  if constexpr (RATE == PD_RATE_GMM) {
    double r64[3];
    rates_gmm(ra, beam, psi, pn, r64);
    *bad = false;
    if (ra.race_sampling)
      return kmc_event_race(r64, u_exp, u_choice, dwell_us, elapsed_us, slot);
    return kmc_event_drawn64(r64, -log1p(-u_exp), u_choice, dwell_us,
                             elapsed_us, slot);
  } else {
    float r[3];
    eval_rates<RATE>(ra, beam, psi, pn, r);
    if (ra.race_sampling) {
      *bad = !(r[0] >= 0.f) || !(r[1] >= 0.f) || !(r[2] >= 0.f);
      const double r64[3] = {r[0], r[1], r[2]};
      return kmc_event_race(r64, u_exp, u_choice, dwell_us, elapsed_us, slot);
    }
    return kmc_event(r, u_exp, u_choice, dwell_us, elapsed_us, slot, bad);
  }
}

// graphene.py:646-694 for one env.
template <int RATE, class Tables>
__device__ __forceinline__ void run_control(const Tables& tab,
                                            const RateArgs& ra, uint64_t seed,
                                            const double2 beam,
                                            long long dwell_us, int ctrl_index,
                                            int64_t env_local,
                                            const LogSink& log, EnvRegs* e) {
  long long elapsed = 0;
  uint32_t it = 0;
  while (elapsed < dwell_us) {  // graphene.py:658
    int nb[3];
    tab.neighbors(e->si, nb);
    double2 pn[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
      pn[i] = site_position(tab.position(nb[i]), e->lat);
    const uint4 w =
        philox4x32_10(e->env_id, e->ctrl_count, it, PD_STREAM_KMC, seed);
    int slot = 0;
    bool bad = false;
    const bool hit =
        rate_event<RATE>(ra, beam, e->psi, pn, u53(w.x, w.y), u53(w.z, w.w),
                         dwell_us, &elapsed, &slot, &bad);
    if (bad) e->status |= PD_ENV_BAD_RATE;
    e->events += 1;
    if (hit) {
      e->si = nb[slot];
      e->psi = slot == 0 ? pn[0] : (slot == 1 ? pn[1] : pn[2]);
      e->transitions += 1;
      if (log.capacity > 0) {
        if (e->log_n < log.capacity) {
          const int64_t o = env_local * log.capacity + e->log_n;
          log.elapsed_us[o] = elapsed;
          log.site[o] = e->si;
          if (log.ctrl) log.ctrl[o] = ctrl_index;
        } else {
          e->status |= PD_ENV_LOG_OVERFLOW;
        }
        e->log_n += 1;
      }
    }
    ++it;
  }
  e->ctrl_count += 1;
}

__device__ __forceinline__ double shfl_double(unsigned mask, double v,
                                              int src) {
  const int lo = __shfl_sync(mask, __double2loint(v), src);
  const int hi = __shfl_sync(mask, __double2hiint(v), src);
  return __hiloint2double(hi, lo);
}

}  // namespace pd
