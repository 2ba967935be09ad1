// Lattice generation and canonical neighbour table.
//
//   graphene.py:464-501  _generate_hexagonal_grid
//   graphene.py:537-543  scale by the bond length, centre on the mean
//   geometry.py:93-111   nearest_neighbors3 (k-NN semantics, canonical order)
//
// Compiled with -fmad=false: the centred coordinates must carry the same
// roundings as the reference's NumPy expression (one product, one subtract).
#include <limits.h>
#include <math.h>

#include "pd_common.cuh"

namespace pd {

struct LatticeShape {
  int n_cols, n_rows, cnt_even, cnt_odd, n_sites;
};

__host__ __device__ inline LatticeShape lattice_shape(int n_cols) {
  LatticeShape s;
  s.n_cols = n_cols;
  const double ratio = sqrt(3.0) / 2.0;
  s.n_rows = static_cast<int>(static_cast<double>(n_cols) / ratio);
  s.cnt_even = n_cols - (n_cols + 2) / 3;  // columns with i % 3 == 0 removed
  s.cnt_odd = n_cols - (n_cols + 1) / 3;   // columns with i % 3 == 1 removed
  const int pairs = s.n_rows / 2;
  s.n_sites = pairs * (s.cnt_even + s.cnt_odd) + (s.n_rows & 1) * s.cnt_even;
  return s;
}

// Site k -> (column i, row j), row-major over surviving sites.
__device__ __forceinline__ void site_ij(const LatticeShape& s, int k, int* i,
                                        int* j) {
  const int pair = s.cnt_even + s.cnt_odd;
  const int j2 = k / pair;
  const int rem = k - j2 * pair;
  if (rem < s.cnt_even) {
    *j = 2 * j2;
    *i = 1 + rem + rem / 2;  // 1,2,4,5,7,8,...
  } else {
    const int r = rem - s.cnt_even;
    *j = 2 * j2 + 1;
    *i = r + (r + 1) / 2;  // 0,2,3,5,6,8,...
  }
}

__global__ void __launch_bounds__(1024, 1)
    k_build_lattice(LatticeShape s, double* __restrict__ base_xy,
                    int32_t* __restrict__ nbr) {
  __shared__ double mean[2];
  const double ratio = sqrt(3.0) / 2.0;
  // (1) positions = grid * 1.42
  for (int k = threadIdx.x; k < s.n_sites; k += blockDim.x) {
    int i, j;
    site_ij(s, k, &i, &j);
    const double gx = static_cast<double>(i) + ((j & 1) ? 0.5 : 0.0);
    const double gy = __dmul_rn(static_cast<double>(j), ratio);
    base_xy[2 * k] = __dmul_rn(gx, kBond);
    base_xy[2 * k + 1] = __dmul_rn(gy, kBond);
  }
  __syncthreads();
  // (2) np.mean(axis=0): sequential per-column sum in site order, then / N.
  if (threadIdx.x < 2) {
    double acc = 0.0;
    for (int k = 0; k < s.n_sites; ++k)
      acc = __dadd_rn(acc, base_xy[2 * k + threadIdx.x]);
    mean[threadIdx.x] = __ddiv_rn(acc, static_cast<double>(s.n_sites));
  }
  __syncthreads();
  // (3) centre
  for (int k = threadIdx.x; k < s.n_sites; k += blockDim.x) {
    base_xy[2 * k] = __dsub_rn(base_xy[2 * k], mean[0]);
    base_xy[2 * k + 1] = __dsub_rn(base_xy[2 * k + 1], mean[1]);
  }
  // (4) 3 nearest neighbours by exact integer distance:
  //     4*d^2 = (2*dx)^2 + 3*dj^2 ; ties resolved by ascending site index.
  for (int k = threadIdx.x; k < s.n_sites; k += blockDim.x) {
    int ik, jk;
    site_ij(s, k, &ik, &jk);
    const int xk = 2 * ik + (jk & 1);
    long long best_d[3] = {LLONG_MAX, LLONG_MAX, LLONG_MAX};
    int best_m[3] = {-1, -1, -1};
    // Every site has three others within two bond lengths (4 d^2 <= 16) in
    // its own and the two rows on either side, and three rows away start at
    // 4 d^2 = 27: rows jk - 3 .. jk + 3 hold the three nearest (ids are row
    // major, so that is one contiguous id range).
    const int pair = s.cnt_even + s.cnt_odd;
    const int j_lo = jk - 3 < 0 ? 0 : jk - 3;
    const int j_hi = jk + 4 > s.n_rows ? s.n_rows : jk + 4;  // exclusive
    const int m_lo = (j_lo / 2) * pair + (j_lo & 1) * s.cnt_even;
    int m_hi = (j_hi / 2) * pair + (j_hi & 1) * s.cnt_even;
    if (m_hi > s.n_sites) m_hi = s.n_sites;
    for (int m = m_lo; m < m_hi; ++m) {
      if (m == k) continue;
      int im, jm;
      site_ij(s, m, &im, &jm);
      const long long dx = 2 * im + (jm & 1) - xk;
      const long long dj = jm - jk;
      const long long d2 = dx * dx + 3 * dj * dj;
      if (d2 < best_d[2]) {
        int pos = 2;
        if (d2 < best_d[1]) pos = 1;
        if (d2 < best_d[0]) pos = 0;
        for (int q = 2; q > pos; --q) {
          best_d[q] = best_d[q - 1];
          best_m[q] = best_m[q - 1];
        }
        best_d[pos] = d2;
        best_m[pos] = m;
      }
    }
    nbr[4 * k + 0] = best_m[0];
    nbr[4 * k + 1] = best_m[1];
    nbr[4 * k + 2] = best_m[2];
    // Geometry class of the site (bits 24-25 of column 3; pd_kmc.cuh
    // site_class): in the bulk the three neighbours are the bonded ones, in
    // the order (row below, same row, row above), and their offsets take one
    // of two forms -- class 0: (-1/2, -r), (1, 0), (-1/2, r) bond lengths
    // with r = sqrt(3)/2, class 1: the negatives in reverse order.  Sites at
    // the sheet edge whose nearest three are not of that form are class 2.
    int cls = 2;
    if (best_d[0] == 4 && best_d[1] == 4 && best_d[2] == 4) {
      int ox[3], oj[3];
      for (int q = 0; q < 3; ++q) {
        int im, jm;
        site_ij(s, best_m[q], &im, &jm);
        ox[q] = 2 * im + (jm & 1) - xk;
        oj[q] = jm - jk;
      }
      if (ox[0] == -1 && oj[0] == -1 && ox[1] == 2 && oj[1] == 0 &&
          ox[2] == -1 && oj[2] == 1)
        cls = 0;
      else if (ox[0] == 1 && oj[0] == -1 && ox[1] == -2 && oj[1] == 0 &&
               ox[2] == 1 && oj[2] == 1)
        cls = 1;
    }
    nbr[4 * k + 3] = (cls << kSiteClassShift) | kCentreListEnd;
  }
  __syncthreads();
  // (5) the low 24 bits of column 3 list, in ascending order and terminated
  //     by kCentreListEnd, the sites within 2.6 A of the lattice centre.  After the reset offset
  //     (|off| <= 0.71*sqrt(2) A) the site nearest the origin is always one
  //     of them (any point of a honeycomb is within one bond of a site), so
  //     pd_reset searches this list instead of all sites.
  if (threadIdx.x == 0) {
    int j = 0;
    for (int k = 0; k < s.n_sites && j < s.n_sites - 1; ++k) {
      const double x = base_xy[2 * k], y = base_xy[2 * k + 1];
      if (x * x + y * y <= 2.6 * 2.6) {
        nbr[4 * j + 3] = (nbr[4 * j + 3] & ~kCentreListEnd) | k;
        ++j;
      }
    }
  }
}

}  // namespace pd

extern "C" int pd_lattice_size(int32_t n_cols, int32_t* out_n_sites,
                               int32_t* out_n_rows) {
  PD_REQUIRE(n_cols >= 4, "n_cols must be >= 4");
  const pd::LatticeShape s = pd::lattice_shape(n_cols);
  if (out_n_sites) *out_n_sites = s.n_sites;
  if (out_n_rows) *out_n_rows = s.n_rows;
  return PD_OK;
}

extern "C" int pd_build_lattice(int32_t n_cols, double* base_xy, int32_t* nbr,
                                void* stream) {
  PD_REQUIRE(n_cols >= 4, "n_cols must be >= 4");
  PD_REQUIRE(base_xy != nullptr && nbr != nullptr, "null output");
  const pd::LatticeShape s = pd::lattice_shape(n_cols);
  PD_REQUIRE(s.n_sites < 65535, "lattice too large for 16-bit site ids");
  pd::k_build_lattice<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      s, base_xy, nbr);
  PD_CUDA_OK(cudaGetLastError());
  return PD_OK;
}
