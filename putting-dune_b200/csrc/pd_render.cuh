// Shared pieces of the STEM frame renderers (pd_render.cu: generic one-CTA
// kernel; pd_render_cluster.cu: the 8-CTA-cluster kernel that keeps the frame
// in shared memory).  Reference: putting_dune/imaging.py:117-265.
#pragma once

#include <math.h>

#include "pd_common.cuh"

namespace pd {

constexpr int kTiles = 8;         // CLAHE kernel = shape // 8
constexpr int kBins = 256;
constexpr int kGray = 16384;      // NR_OF_GRAY
constexpr int kBinSize = 1 + kGray / kBins;  // 65
constexpr int kInvTable = 256;
constexpr int kMaxFramesPerLaunch = 1 << 18;  // flag bytes in the workspace
constexpr int kGenericGrid = 16;  // CTAs (and scratch slots) of the generic kernel

struct RenderArgs {
  pd_lattice lat;
  pd_state st;
  const int32_t* env_ids;
  int32_t m;
  int32_t size;        // S
  int32_t log2_size;
  int32_t stop_stage;
  int32_t advance;
  float* out;          // [m][S][S]
  float* scratch;      // generic kernel: [grid][2][S][S]
  int32_t* n_generic;  // frames the cluster kernel handed to the generic one
  uint8_t* generic;    // [m] 1 = render with the generic kernel
  double buffer;       // imaging.py:123 buffer_size (0: atoms in view only)
  int32_t bw;          // int(buffer * S) pixels
};

// Pixel bin of an atom (imaging.py:129-143 np.histogram2d, then the flipud /
// transpose of :150 and the crop of :165-168), in coordinates of the cropped
// S x S image: row = S_ext - 1 - by - bw, col = bx - bw, both possibly
// outside [0, S) when a buffer is used.
struct AtomBin {
  bool keep;
  int row, col;
};

// np.linspace(lo, hi, n + 1)[i]: arange * step + start, last point = stop.
__device__ __forceinline__ double hist_edge(double lo, double hi, double step,
                                            int i, int n) {
  return i == n ? hi : __dadd_rn(__dmul_rn(static_cast<double>(i), step), lo);
}

// searchsorted(edges, q, 'right') - 1 with the rightmost edge closed
// (np.histogramdd); q must lie in [lo, hi].
__device__ __forceinline__ int hist_bin(double q, double lo, double hi,
                                        double step, int n) {
  int j = static_cast<int>(floor((q - lo) / step));
  j = j < 0 ? 0 : (j > n - 1 ? n - 1 : j);
  while (j < n - 1 && hist_edge(lo, hi, step, j + 1, n) <= q) ++j;
  while (j > 0 && hist_edge(lo, hi, step, j, n) > q) --j;
  return j;
}

__device__ __forceinline__ AtomBin atom_bin(const double2 p, const Fov4& fv,
                                            int S, double buffer, int bw) {
  AtomBin out{false, 0, 0};
  const double fw = __dsub_rn(fv.urx, fv.llx), fh = __dsub_rn(fv.ury, fv.lly);
  if (buffer <= 0.0) {
    // the caller passes get_atoms_in_bounds(fov): graphene.py:600-644
    if (!(fv.llx <= p.x && p.x <= fv.urx && fv.lly <= p.y && p.y <= fv.ury))
      return out;
    const double qx = __ddiv_rn(__dsub_rn(p.x, fv.llx), fw);
    const double qy = __ddiv_rn(__dsub_rn(p.y, fv.lly), fh);
    int bx = static_cast<int>(floor(__dmul_rn(qx, static_cast<double>(S))));
    int by = static_cast<int>(floor(__dmul_rn(qy, static_cast<double>(S))));
    if (bx > S - 1) bx = S - 1;  // q == 1 falls in the last bin
    if (by > S - 1) by = S - 1;
    out.keep = true;
    out.row = S - 1 - by;
    out.col = bx;
    return out;
  }
  // the caller passes the whole grid in the microscope frame
  const double qx = __ddiv_rn(__dsub_rn(p.x, fv.llx), fw);
  const double qy = __ddiv_rn(__dsub_rn(p.y, fv.lly), fh);
  const double lo = -buffer, hi = __dadd_rn(1.0, buffer);
  if (!(qx >= lo && qx <= hi && qy >= lo && qy <= hi)) return out;
  const int n = S + 2 * bw;
  const double step = __ddiv_rn(__dsub_rn(hi, lo), static_cast<double>(n));
  const int bx = hist_bin(qx, lo, hi, step, n);
  const int by = hist_bin(qy, lo, hi, step, n);
  out.keep = true;
  out.row = n - 1 - by - bw;
  out.col = bx - bw;
  return out;
}

__device__ __forceinline__ float u24(uint32_t w) {
  return static_cast<float>(w >> 8) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float u24_open(uint32_t w) {
  return (static_cast<float>(w >> 8) + 1.0f) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ uint32_t word_of(const uint4& w, int j) {
  return j == 0 ? w.x : j == 1 ? w.y : j == 2 ? w.z : w.w;
}

// Inverse-CDF Poisson: smallest k with CDF(k) > u (float64 recurrence; the
// injected-noise convention of include/pdune_b200.h).
__device__ __forceinline__ int poisson_icdf(double lam, double u) {
  double p = exp(-lam);
  double cdf = p;
  int k = 0;
  while (u >= cdf && k < 100000) {
    ++k;
    p = p * lam / static_cast<double>(k);
    cdf += p;
  }
  return k;
}

// Same recurrence with 1/k from a shared table (k < kInvTable).  p * lam / k
// and p * lam * (1/k) may differ in the last bit of p; the comparison u >= cdf
// only changes if u sits within ~1e-16 of a threshold (u has 24 bits).
__device__ __forceinline__ int poisson_icdf_tab(double lam, double u,
                                                const double* inv_k) {
  double p = exp(-lam);
  double cdf = p;
  int k = 0;
  while (u >= cdf && k < 100000) {
    ++k;
    p = p * lam * (k < kInvTable ? inv_k[k] : 1.0 / static_cast<double>(k));
    cdf += p;
  }
  return k;
}

// Four inverse-CDF searches in float32.  inv_k1[i] = 1 / (i + 1) (16-byte
// aligned).  The float32 CDF differs from the float64 one by
// < (4k + 3) * 2^-24; a search whose u lies within eps(k) = 1e-6 (k + 4) of
// the two thresholds that decide it, or with lam > 60 (exp(-lam) leaves the
// float32 range) or u > 0.9999 (the float32 sum may never reach it), is redone
// in float64, so the result always equals poisson_icdf_tab's.
__device__ __forceinline__ void poisson4(const float (&lam)[4],
                                         const float (&u)[4],
                                         const float* inv_k1,
                                         const double* inv_kd, int (&k)[4]) {
  unsigned redo = 0;
  const float4* tp = reinterpret_cast<const float4*>(inv_k1);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float l = lam[j], uu = u[j];
    k[j] = 0;
    if (!(l <= 60.f) || !(uu <= 0.9999f)) {
      redo |= 1u << j;
      continue;
    }
    // (p, cdf) = pmf and CDF at kk; below = CDF(kk - 1)
    float p = expf(-l);
    float cdf = p, below = 0.f;
    int kk = 0;
    if (uu >= cdf) {
      // four terms per trip: Poisson(60) passes 0.9999 before k = 100
      for (int i4 = 0;; ++i4) {
        const float4 ik = tp[i4];  // 1 / (4 i4 + 1) ... 1 / (4 i4 + 4)
        const float r1 = l * ik.x, r2 = l * ik.y, r3 = l * ik.z, r4 = l * ik.w;
        const float p1 = p * r1, p2 = p1 * r2, p3 = p2 * r3, p4 = p3 * r4;
        const float c1 = cdf + p1, c2 = c1 + p2, c3 = c2 + p3, c4 = c3 + p4;
        if (uu >= c4 && i4 < kInvTable / 4 - 1) {
          p = p4;
          cdf = c4;
          continue;
        }
        // the answer is among 4 i4 + 1 .. 4 i4 + 4 (or the table ended)
        const int base = 4 * i4;
        if (!(uu >= c1)) {
          kk = base + 1; p = p1; below = cdf; cdf = c1;
        } else if (!(uu >= c2)) {
          kk = base + 2; p = p2; below = c1; cdf = c2;
        } else if (!(uu >= c3)) {
          kk = base + 3; p = p3; below = c2; cdf = c3;
        } else {
          kk = base + 4; p = p4; below = c3; cdf = c4;
        }
        break;
      }
    }
    k[j] = kk;
    const float eps = 1e-6f * static_cast<float>(kk + 4);
    // the thresholds that decided: CDF(kk - 1) (<= u) and CDF(kk) (> u)
    if (!(cdf - uu >= eps) || (kk > 0 && !(uu - below >= eps)))
      redo |= 1u << j;
  }
  if (redo != 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (redo & (1u << j))
        k[j] = poisson_icdf_tab(static_cast<double>(lam[j]),
                                static_cast<double>(u[j]), inv_kd);
  }
}

// skimage exposure/_adapthist.py clip_histogram + map_histogram for one tile,
// one thread.
__device__ inline void clahe_tile_map(int* hist, unsigned short* map, int clim,
                                      int n_pixels) {
  int n_excess = 0;
  for (int i = 0; i < kBins; ++i)
    if (hist[i] > clim) {
      n_excess += hist[i] - clim;
      hist[i] = clim;
    }
  const int bin_incr = n_excess / kBins;
  const int upper = clim - bin_incr;
  for (int i = 0; i < kBins; ++i)
    if (hist[i] < upper) {
      n_excess -= bin_incr;
      hist[i] += bin_incr;
    }
  for (int i = 0; i < kBins; ++i)
    if (hist[i] >= upper && hist[i] < clim) {
      n_excess += hist[i] - clim;
      hist[i] = clim;
    }
  while (n_excess > 0) {
    const int prev = n_excess;
    for (int index = 0; index < kBins; ++index) {
      int n_under = 0;
      for (int i = 0; i < kBins; ++i) n_under += hist[i] < clim;
      int step = n_under / n_excess;
      if (step < 1) step = 1;
      int cnt = 0;
      for (int i = index; i < kBins; i += step)
        if (hist[i] < clim) {
          ++hist[i];
          ++cnt;
        }
      n_excess -= cnt;
      if (n_excess <= 0) break;
    }
    if (prev == n_excess) break;
  }
  long long cum = 0;
  for (int i = 0; i < kBins; ++i) {
    cum += hist[i];
    long long v = cum * (kGray - 1) / n_pixels;
    map[i] = static_cast<unsigned short>(v > kGray - 1 ? kGray - 1 : v);
  }
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// The same computation by one warp: lane l owns bins [8 l, 8 l + 8).
__device__ inline void clahe_tile_map_warp(const int* hist,
                                           unsigned short* map, int clim,
                                           int n_pixels, int lane) {
  int h[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) h[j] = hist[8 * lane + j];
  int part = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (h[j] > clim) {
      part += h[j] - clim;
      h[j] = clim;
    }
  int n_excess = warp_sum(part);
  const int bin_incr = n_excess / kBins;
  const int upper = clim - bin_incr;
  part = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (h[j] < upper) {
      h[j] += bin_incr;
      ++part;
    }
  n_excess -= warp_sum(part) * bin_incr;
  part = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (h[j] >= upper && h[j] < clim) {
      part += h[j] - clim;
      h[j] = clim;
    }
  n_excess += warp_sum(part);
  bool stuck = false;
  while (n_excess > 0 && !stuck) {
    const int prev = n_excess;
    for (int index = 0; index < kBins; ++index) {
      part = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) part += h[j] < clim;
      const int n_under = warp_sum(part);
      if (n_under == 0) {  // nothing can change any more
        stuck = true;
        break;
      }
      int step = n_under / n_excess;
      if (step < 1) step = 1;
      part = 0;
      // bins index, index + step, ...: offset of this lane's first bin
      int r = 8 * lane - index;
      r = r >= 0 ? r % step : (step - (-r) % step) % step;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (8 * lane + j >= index && r == 0 && h[j] < clim) {
          ++h[j];
          ++part;
        }
        r = r + 1 == step ? 0 : r + 1;
      }
      n_excess -= warp_sum(part);
      if (n_excess <= 0) break;
    }
    if (prev == n_excess) break;
  }
  int run = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) run += h[j];
  int incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  // n_pixels = ts^2 is a power of two <= 4096: cum * 16383 fits 32 bits
  const int sh_px = 31 - __clz(n_pixels);
  int cum = incl - run;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    cum += h[j];
    const int v = (cum * (kGray - 1)) >> sh_px;
    map[8 * lane + j] =
        static_cast<unsigned short>(v > kGray - 1 ? kGray - 1 : v);
  }
}

// Launches the cluster kernel over frames [0, a.m); frames it cannot take
// (more atoms in view / wider kernels than its shared-memory tables hold) are
// flagged in a.generic for the generic kernel.
int launch_render_cluster(const RenderArgs& a, cudaStream_t stream);
int render_cluster_count(int image_size, int* out_clusters);

}  // namespace pd
