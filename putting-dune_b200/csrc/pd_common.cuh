// Shared device helpers for the putting-dune B200 kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pdune_b200.h"

namespace pd {

constexpr double kBond = 1.42;                 // constants.py:23
constexpr double kMaxTransitionSeconds = 3600; // graphene.py:668
constexpr int kCarbon = 6;                     // constants.py:20
constexpr int kSilicon = 14;                   // constants.py:21

// Column 3 of pd_lattice.nbr: bits 0-23 an entry of the list of sites near the
// lattice centre (pd_reset), bits 24-25 the site's geometry class
// (pd_lattice.cu).
constexpr int kCentreListEnd = 0xFFFFFF;
constexpr int kSiteClassShift = 24;

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t err, const char* what);
int sm_count();

#define PD_CUDA_OK(expr)                                  \
  do {                                                    \
    int _rc = ::pd::check_cuda((expr), #expr);            \
    if (_rc != PD_OK) return _rc;                         \
  } while (0)

#define PD_REQUIRE(cond, msg)                             \
  do {                                                    \
    if (!(cond)) {                                        \
      ::pd::set_error("%s: %s", __func__, msg);           \
      return PD_ERR_INVALID_ARGUMENT;                     \
    }                                                     \
  } while (0)

// ---------------------------------------------------------------------------
// Philox4x32-10 (Random123 constants).  Counter = (env, seq, slot, stream).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1,
                                               uint32_t c2, uint32_t c3,
                                               uint64_t seed) {
  uint32_t k0 = static_cast<uint32_t>(seed);
  uint32_t k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0);
    const uint32_t lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2);
    const uint32_t lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// 53-bit uniform in [0, 1): ((hi >> 5) * 2^26 + (lo >> 6)) / 2^53.
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
  const uint64_t m = (static_cast<uint64_t>(hi >> 5) << 26) |
                     static_cast<uint64_t>(lo >> 6);
  // m < 2^53 converts exactly; the scaling is a power of two.
  return static_cast<double>(static_cast<long long>(m)) *
         (1.0 / 9007199254740992.0);
}

// k-th draw of a linear stream (slot k/2, half k%2).
__device__ __forceinline__ double draw_linear(uint64_t seed, uint32_t env,
                                              uint32_t seq, uint32_t stream,
                                              uint32_t k) {
  const uint4 w = philox4x32_10(env, seq, k >> 1, stream, seed);
  return (k & 1u) ? u53(w.z, w.w) : u53(w.x, w.y);
}

// ---------------------------------------------------------------------------
// Geometry: position of lattice site in the env's material frame.
// graphene.py:545-557: (base + off) @ [[c, -s], [s, c]]
//   x' = bx*c + by*s ; y' = by*c - bx*s  (each product rounded: no FMA).
// ---------------------------------------------------------------------------
// np.clip: a NaN stays a NaN (fmin / fmax would drop it), so that a NaN
// action reaches the rate function and is flagged there as in the reference
// (graphene.py:258).
__device__ __forceinline__ double clip_nan(double x, double lo, double hi) {
  const double c = fmin(fmax(x, lo), hi);
  return x == x ? c : x;
}
__device__ __forceinline__ float clip_nanf(float x, float lo, float hi) {
  const float c = fminf(fmaxf(x, lo), hi);
  return x == x ? c : x;
}

struct Lattice4 {
  double ox, oy, c, s;
};

__device__ __forceinline__ double2 site_position(const double2 base,
                                                 const Lattice4& t) {
  const double bx = __dadd_rn(base.x, t.ox);
  const double by = __dadd_rn(base.y, t.oy);
  double2 p;
  p.x = __dadd_rn(__dmul_rn(bx, t.c), __dmul_rn(by, t.s));
  p.y = __dsub_rn(__dmul_rn(by, t.c), __dmul_rn(bx, t.s));
  return p;
}

__device__ __forceinline__ Lattice4 load_lattice4(const double* lattice,
                                                  int64_t e) {
  const double2 a = reinterpret_cast<const double2*>(lattice)[2 * e];
  const double2 b = reinterpret_cast<const double2*>(lattice)[2 * e + 1];
  return Lattice4{a.x, a.y, b.x, b.y};
}

struct Fov4 {
  double llx, lly, urx, ury;
};

__device__ __forceinline__ Fov4 load_fov4(const double* fov, int64_t e) {
  const double2 a = reinterpret_cast<const double2*>(fov)[2 * e];
  const double2 b = reinterpret_cast<const double2*>(fov)[2 * e + 1];
  return Fov4{a.x, a.y, b.x, b.y};
}

__device__ __forceinline__ void store_fov4(double* fov, int64_t e,
                                           const Fov4& f) {
  reinterpret_cast<double2*>(fov)[2 * e] = make_double2(f.llx, f.lly);
  reinterpret_cast<double2*>(fov)[2 * e + 1] = make_double2(f.urx, f.ury);
}

}  // namespace pd
