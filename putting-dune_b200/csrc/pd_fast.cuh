// Guarded float32 form of one kinetic-Monte-Carlo iteration
// (graphene.py:658-694) for the prior / simple rate functions.
//
// What an iteration decides is discrete: does the waiting time fit into the
// rest of the dwell time (graphene.py:677) and, if so, which neighbour the Si
// hops to (graphene.py:679-688).  The float64 -> float32 -> float64 chain of
// the reference (pd_kmc.cuh kmc_event_drawn) is ~900 dependent instructions;
// the same decisions follow from a float32 evaluation (~60) whenever the
// float32 result is further from the deciding threshold than its own error.
// fast_event evaluates the iteration in float32 together with a bound of
// that error and answers NO_HOP / HOP(slot) only when the answer cannot
// depend on the error; otherwise it answers UNSURE and the caller replays the
// whole control with the exact float64 code (run_control).  Results are
// therefore those of the exact path, decision for decision.
//
// Error budget (relative, float32 ulp = 6e-8).  The inputs are the beam
// offset from the Si and the three neighbour directions, each good to
// ~4e-7 absolute in bond units (one or two float32 roundings of float64
// values); a squared distance d^2 then carries 2 d 5e-7 + 2e-7 d^2, the prior's
// exponent -5 d^2 five times that: 1.3e-5 at d = 1.6, the furthest the
// dominant neighbour can be under the relative adapter; ex2.approx adds
// 2.4e-7 + 9e-8 |arg|.  kFastEps0 + kFastEps1 * |arg| is >= 3.5x that
// everywhere (and the simple rate, a rational function, sits below 2e-6).
// The reference's own float32 steps (cast of the rates, sequential sum,
// reciprocal: 3e-7) are inside the same bound.  The unit-exponential draw
// uses the top 24 bits u24 of the 53-bit uniform: 1 - u24 is exact,
// -ln 2 lg2.approx(1 - u24) is good to 3e-7 absolute (measured over all 2^24
// values by pd_fast_path_audit), and the dropped bits move the draw up by
// < 6.1e-8 / (1 - u) while 1 - u > 2^-12 (no upper bound is claimed beyond).  pd_fast_path_audit
// (pd_step_fast.cu) measures all of these against the exact path.
#pragma once

#include "pd_kmc.cuh"

namespace pd {

constexpr float kFastEps0 = 5e-5f;       // relative bound of the total rate
constexpr float kFastEps1 = 4e-6f;       // ... per unit of the prior's exponent
constexpr float kFastDrawAbs = 6e-7f;    // absolute bound of -log(1 - u24)
                                         // (+ 1e-6 relative, inside the slack)
constexpr float kFastMaxDwellS = 3000.f; // below graphene.py:668's 3600 s cap

enum : int { FAST_NO_HOP = 0, FAST_HOP = 1, FAST_UNSURE = 2 };

// Philox4x32-10 with the round keys precomputed (PhiloxKeys, pd_kmc.cuh) and
// the 32 x 32 -> 64 products as single wide multiplies.
__device__ __forceinline__ PhiloxKeys philox_keys(uint64_t seed) {
  PhiloxKeys k;
  uint32_t a = static_cast<uint32_t>(seed);
  uint32_t b = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    k.k0[r] = a;
    k.k1[r] = b;
    a += 0x9E3779B9u;
    b += 0xBB67AE85u;
  }
  return k;
}

__device__ __forceinline__ uint4 philox4x32_10k(uint32_t c0, uint32_t c1,
                                                uint32_t c2, uint32_t c3,
                                                const PhiloxKeys& k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0;
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
    c0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k.k0[r];
    c1 = static_cast<uint32_t>(p1);
    c2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k.k1[r];
    c3 = static_cast<uint32_t>(p0);
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------------------
// Neighbour geometry of the Si site in float32.
//   PD_RATE_PRIOR:  gx, gy = peak positions 0.85 (cos, -sin) of the three
//                   neighbour directions, bond units (graphene.py:210-229);
//   PD_RATE_SIMPLE: gx, gy = neighbour offsets, angstrom (graphene.py:151-166).
// Bulk sites come in two classes (pd_lattice.cu) whose geometries are
// g1[i] = -g0[2 - i], so a hop between bulk sites is a sign flip and a swap
// and no table is touched; everything is evaluated in float64 and rounded
// once.
// ---------------------------------------------------------------------------
// Scale of the beam offset that fast_event expects: bond units for the prior,
// angstrom for the simple rate.
template <int RATE>
__device__ __forceinline__ float fast_offset_scale() {
  return RATE == PD_RATE_PRIOR ? static_cast<float>(1.0 / kBond) : 1.0f;
}

struct FastGeo {
  float gx[3], gy[3];
};

template <int RATE>
__device__ __forceinline__ void fast_geo_one(double nx, double ny, double len,
                                             float* gx, float* gy) {
  if (RATE == PD_RATE_PRIOR) {
    const double inv = 0.85 / len;
    *gx = static_cast<float>(nx * inv);
    *gy = static_cast<float>(-ny * inv);  // mirror quirk, see rates_prior
  } else {
    *gx = static_cast<float>(nx);
    *gy = static_cast<float>(ny);
  }
}

// Bulk site of class `cls` (0 / 1) in the env's material frame
// (graphene.py:545-557: (x, y) -> (x c + y s, y c - x s)).
template <int RATE>
__device__ __forceinline__ void fast_geo_bulk(const Lattice4& lat, int cls,
                                              FastGeo* g) {
  const double kR = 0.8660254037844386;
  const double ux[3] = {-0.5, 1.0, -0.5};
  const double uy[3] = {-kR, 0.0, kR};
  const double sgn = cls == 0 ? kBond : -kBond;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int k = cls == 0 ? i : 2 - i;
    const double nx = sgn * (ux[k] * lat.c + uy[k] * lat.s);
    const double ny = sgn * (uy[k] * lat.c - ux[k] * lat.s);
    fast_geo_one<RATE>(nx, ny, kBond, &g->gx[i], &g->gy[i]);
  }
}

// Geometry of any site, out of line: bulk sites from the constants above,
// sheet-edge sites (class 2) from the tables.  Runs when an env is loaded,
// after an exact replay, and for hops that touch the sheet edge.
template <int RATE, class Tables>
__device__ __noinline__ FastGeo fast_geo_any(const Tables tab, int si, int n0,
                                             int n1, int n2, int cls, double c,
                                             double s) {
  FastGeo g;
  const Lattice4 lat{0.0, 0.0, c, s};
  if (cls < 2) {
    fast_geo_bulk<RATE>(lat, cls, &g);
    return g;
  }
  const int nb[3] = {n0, n1, n2};
  const double2 b0 = tab.position(si);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double2 bi = tab.position(nb[i]);
    const double dx = bi.x - b0.x, dy = bi.y - b0.y;
    const double nx = dx * c + dy * s;
    const double ny = dy * c - dx * s;
    fast_geo_one<RATE>(nx, ny, sqrt(nx * nx + ny * ny), &g.gx[i], &g.gy[i]);
  }
  return g;
}

// The Si site as the fast kernels carry it: the site, its three neighbours
// and geometry class (one table row) and the float32 geometry.
struct FastSite {
  int si, nb[3], cls;
  FastGeo geo;
};

template <int RATE, class Tables>
__device__ __forceinline__ FastSite fast_site(const Tables& tab, int si,
                                              double c, double s) {
  FastSite f;
  f.si = si;
  f.cls = tab.neighbors_class(si, f.nb);
  f.geo = fast_geo_any<RATE>(tab, si, f.nb[0], f.nb[1], f.nb[2], f.cls, c, s);
  return f;
}

// Offset (angstrom) of neighbour `slot` of a BULK site whose geometry is g.
template <int RATE>
__device__ __forceinline__ void bulk_offset(const FastGeo& g, int slot,
                                            float* ox, float* oy) {
  const float gx = slot == 0 ? g.gx[0] : (slot == 1 ? g.gx[1] : g.gx[2]);
  const float gy = slot == 0 ? g.gy[0] : (slot == 1 ? g.gy[1] : g.gy[2]);
  if (RATE == PD_RATE_PRIOR) {
    // a bond vector; peak = 0.85 (ux, -uy) of its unit vector u
    *ox = gx * static_cast<float>(kBond / 0.85);
    *oy = -gy * static_cast<float>(kBond / 0.85);
  } else {
    *ox = gx;
    *oy = gy;
  }
}

// bulk -> bulk always changes the sublattice: g'[i] = -g[2 - i]
__device__ __forceinline__ void flip_geo(FastGeo* g) {
  const float x0 = -g->gx[2], y0 = -g->gy[2];
  g->gx[2] = -g->gx[0];
  g->gy[2] = -g->gy[0];
  g->gx[1] = -g->gx[1];
  g->gy[1] = -g->gy[1];
  g->gx[0] = x0;
  g->gy[0] = y0;
}

// The hop to neighbour `slot`: new site and geometry, and the beam offset
// (the beam stays where its control put it; the Si moves by the neighbour
// offset, which the geometry holds in float32).  c, s: the env's lattice
// rotation; only read when the hop touches the sheet edge.
template <int RATE, class Tables, class LoadRotation>
__device__ __forceinline__ void fast_hop(const Tables& tab, int slot,
                                         LoadRotation rotation, FastSite* f,
                                         float* bx, float* by, float* ox,
                                         float* oy) {
  const int to = slot == 0 ? f->nb[0] : (slot == 1 ? f->nb[1] : f->nb[2]);
  const int cls_old = f->cls;
  // (*ox, *oy): the neighbour's offset in angstrom
  if (cls_old < 2) {
    bulk_offset<RATE>(f->geo, slot, ox, oy);
  } else {
    // sheet edge: the nearest sites need not be a bond away
    const double2 cs = rotation();
    const double2 b0 = tab.position(f->si), b1 = tab.position(to);
    const double dx = b1.x - b0.x, dy = b1.y - b0.y;
    *ox = static_cast<float>(dx * cs.x + dy * cs.y);
    *oy = static_cast<float>(dy * cs.x - dx * cs.y);
  }
  *bx -= *ox * fast_offset_scale<RATE>();
  *by -= *oy * fast_offset_scale<RATE>();
  f->si = to;
  f->cls = tab.neighbors_class(to, f->nb);
  if (cls_old < 2 && f->cls < 2) {
    flip_geo(&f->geo);
  } else {
    const double2 cs = rotation();
    f->geo = fast_geo_any<RATE>(tab, to, f->nb[0], f->nb[1], f->nb[2], f->cls,
                                cs.x, cs.y);
  }
}

// ---------------------------------------------------------------------------
// One iteration.  bx, by: beam - Si (scaled as above); wx, wz: Philox words x
// and z of the iteration; [e_lo, e_hi]: bounds of the control's clock in
// seconds; margin: FastTimes::margin.
//   FAST_NO_HOP  the waiting time certainly overshoots the dwell time;
//   FAST_HOP     it certainly does not and the loop certainly goes on after
//                it; *slot is the successor, [*t_lo, *t_hi] bounds the
//                waiting time;
//   FAST_UNSURE  anything else, NaNs included.
// ---------------------------------------------------------------------------
struct FastTimes {
  float dwell_s;  // dwell time
  float margin;   // microsecond rounding + float32 sums of the comparisons
};

__host__ __device__ __forceinline__ FastTimes fast_times(long long dwell_us) {
  FastTimes t;
  t.dwell_s = static_cast<float>(static_cast<double>(dwell_us) * 1e-6);
  t.margin = 1.5e-6f + 3e-7f * t.dwell_s;
  return t;
}
// The launch's FastTimes, computed once on the host (launch_fast) and read as
// constant-bank operands: the kernels would otherwise re-derive them from the
// dwell time wherever registers are short.
__device__ __forceinline__ FastTimes fast_times(const StepArgs& a) {
  FastTimes t;
  t.dwell_s = a.fast_dwell_s;
  t.margin = a.fast_margin;
  return t;
}

// The pieces of an iteration (fast_event and fast_quiet are the same
// arithmetic, instruction for instruction).
struct FastRates {
  float r0, r1, r2, sum, tot, eps;
};

template <int RATE>
__device__ __forceinline__ FastRates fast_rates(const FastGeo& g, float bx,
                                                float by) {
  FastRates f;
  const float dx0 = bx - g.gx[0], dy0 = by - g.gy[0];
  const float dx1 = bx - g.gx[1], dy1 = by - g.gy[1];
  const float dx2 = bx - g.gx[2], dy2 = by - g.gy[2];
  const float a0 = __fmaf_rn(dx0, dx0, dy0 * dy0);
  const float a1 = __fmaf_rn(dx1, dx1, dy1 * dy1);
  const float a2 = __fmaf_rn(dx2, dx2, dy2 * dy2);
  if (RATE == PD_RATE_PRIOR) {
    // exp(-5 a) = 2^(-5 log2(e) a); the factor ln 2 / 3 goes on the sum
    const float kC = -7.213475204444817f;
    f.r0 = ex2_approx(kC * a0);
    f.r1 = ex2_approx(kC * a1);
    f.r2 = ex2_approx(kC * a2);
    f.eps = __fmaf_rn(5.0f * kFastEps1, fminf(a0, fminf(a1, a2)), kFastEps0);
  } else {
    const float kS = static_cast<float>(16.0 / (kBond * kBond));
    f.r0 = rcp_approx(__fmaf_rn(a0, kS, 1.0f));
    f.r1 = rcp_approx(__fmaf_rn(a1, kS, 1.0f));
    f.r2 = rcp_approx(__fmaf_rn(a2, kS, 1.0f));
    f.eps = kFastEps0;
  }
  f.sum = (f.r0 + f.r1) + f.r2;
  f.tot = RATE == PD_RATE_PRIOR ? 0.23104906018664842f * f.sum : f.sum;
  return f;
}

// -log(1 - u): u53 lies in [u24, u24 + 2^-24); v = 1 - u24
__device__ __forceinline__ float fast_draw(uint32_t wx, float* v) {
  *v = 1.0f - static_cast<float>(wx >> 8) * (1.0f / 16777216.0f);
  return -0.6931471805599453f * lg2_approx(*v);
}

// Lower bound of the waiting time, and 1 / total.
__device__ __forceinline__ float fast_wait_lo(const FastRates& f, float draw,
                                              float* rc, float* t_mid,
                                              float* slack) {
  *rc = rcp_approx(f.tot);
  *t_mid = draw * *rc;
  *slack = __fmaf_rn(*t_mid, f.eps + 1e-6f, kFastDrawAbs * *rc);
  return *t_mid - *slack;
}

// A vanishing total rate (beam far away; float32 underflow of the prior is
// routine) means a waiting time of hours whatever the
// error, unless the draw is exactly zero.
__device__ __forceinline__ bool fast_no_rate(const FastRates& f) {
  return !(f.tot > 1e-30f) && f.tot == f.tot;
}

__device__ __forceinline__ bool fast_certain_no(bool no_rate, uint32_t wx,
                                                float e_lo, float t_lo,
                                                const FastTimes& tm) {
  return no_rate ? (wx >> 8) != 0u : (e_lo + t_lo) - tm.margin > tm.dwell_s;
}

template <int RATE>
__device__ __forceinline__ int fast_event(const FastGeo& g, float bx, float by,
                                          uint32_t wx, uint32_t wz, float e_lo,
                                          float e_hi, const FastTimes& tm,
                                          int* slot, float* t_lo,
                                          float* t_hi) {
  const FastRates f = fast_rates<RATE>(g, bx, by);
  float v, rc, t_mid, slack;
  const float draw = fast_draw(wx, &v);
  *t_lo = fast_wait_lo(f, draw, &rc, &t_mid, &slack);
  // the bits dropped from the uniform raise the draw by -log(1 - d / v),
  // d < 2^-24: below 6.1e-8 / v while v > 2^-12, unbounded as v -> 2^-24
  const float dropped =
      v > 2.5e-4f ? 6.1e-8f * rcp_approx(v) : __int_as_float(0x7f800000);
  *t_hi = (t_mid + slack) + dropped * rc;
  const bool no_rate = fast_no_rate(f);
  const bool certain_no = fast_certain_no(no_rate, wx, e_lo, *t_lo, tm);
  const bool certain_hop =
      !no_rate && (e_hi + *t_hi) + 2.0f * tm.margin < tm.dwell_s;
  if (certain_no) return FAST_NO_HOP;
  if (!certain_hop) return FAST_UNSURE;
  // rng.choice(3, p = rates / total): thresholds r0 / tot and (r0 + r1) / tot
  // against u53(z, w), which lies in [uc, uc + 2^-24)
  const float inv = rcp_approx(f.sum);
  const float p0 = f.r0 * inv, p01 = (f.r0 + f.r1) * inv;
  const float uc = static_cast<float>(wz >> 8) * (1.0f / 16777216.0f);
  const float eta = __fmaf_rn(3.0f, f.eps, 2e-6f);
  const bool le0 = p0 + eta < uc;             // c0 <= u for sure
  const bool gt0 = p0 - eta > uc + 6.0e-8f;   // c0 >  u for sure
  const bool le1 = p01 + eta < uc;
  const bool gt1 = p01 - eta > uc + 6.0e-8f;
  if (!((le0 || gt0) && (le1 || gt1))) return FAST_UNSURE;
  *slot = (le0 ? 1 : 0) + (le1 ? 1 : 0);
  return FAST_HOP;
}

// Iteration 0 of a control (clock at zero) for a Si on a bulk site of either
// class, when all that is asked is "does it certainly end the control without
// a hop": fast_event(...) == FAST_NO_HOP for the geometry g0 and for its flip,
// the draw shared (it depends on the Philox word only).
template <int RATE>
__device__ __forceinline__ void fast_quiet_both(const FastGeo& g0, float bx,
                                                float by, uint32_t wx,
                                                const FastTimes& tm,
                                                bool* quiet0, bool* quiet1) {
  float v, rc, t_mid, slack;
  const float draw = fast_draw(wx, &v);
  {
    const FastRates f = fast_rates<RATE>(g0, bx, by);
    const float t_lo = fast_wait_lo(f, draw, &rc, &t_mid, &slack);
    *quiet0 = fast_certain_no(fast_no_rate(f), wx, 0.f, t_lo, tm);
  }
  {
    FastGeo g1 = g0;
    flip_geo(&g1);
    const FastRates f = fast_rates<RATE>(g1, bx, by);
    const float t_lo = fast_wait_lo(f, draw, &rc, &t_mid, &slack);
    *quiet1 = fast_certain_no(fast_no_rate(f), wx, 0.f, t_lo, tm);
  }
}

// Clock bounds after a hop (directed rounding; 1 us for the rounding of the
// waiting time to whole microseconds, graphene.py:669).
__device__ __forceinline__ void fast_advance(float* e_lo, float* e_hi,
                                             float t_lo, float t_hi) {
  *e_lo = __fadd_rd(*e_lo, __fadd_rd(t_lo, -1e-6f));
  *e_hi = __fadd_ru(*e_hi, __fadd_ru(t_hi, 1e-6f));
}

}  // namespace pd
