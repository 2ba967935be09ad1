// Label masks for perception-model data: imaging.py:75-114 generate_grid_mask.
//
// One CTA per (frame, band of S/8 rows).  The CTA rebuilds the atoms in view
// (graphene.py:600-644) in the material frame exactly as the reference does
// -- normalise to the FOV, then microscope_utils.py:350-356 p * (ur - ll) + ll
// -- keeps those that can reach its rows, and every pixel takes the atomic
// number of the LAST atom (lattice order) whose squared distance in
// angstrom^2 is below radius = (Z / 6)^e * 0.1 (the reference compares the
// squared distance with the unsquared radius; kept).  Pixel centres are the
// midpoints of np.linspace(ll, ur, S + 1).  float64, every operation rounded
// like NumPy's (compiled with -fmad=false).
#include "pd_kmc.cuh"

namespace pd {

constexpr int kMaskThreads = 256;
constexpr int kMaskBands = 8;
constexpr int kMaskCap = 2048;

struct MaskShared {
  double2 pos[kMaskCap];
  uint8_t z[kMaskCap];
  int warp_count[kMaskThreads / 32];
  int n;
};

// np.linspace(lo, hi, S + 1)[i]: arange * step + start, last point = stop.
__device__ __forceinline__ double linspace_at(double lo, double hi, double step,
                                              int i, int s) {
  return i == s ? hi : __dadd_rn(__dmul_rn(static_cast<double>(i), step), lo);
}

__global__ void __launch_bounds__(kMaskThreads)
    k_grid_mask(const pd_lattice lat, const pd_state st,
                const int32_t* __restrict__ env_ids, int32_t image_size,
                double radius_c, double radius_si,
                uint8_t* __restrict__ mask_out) {
  __shared__ MaskShared sh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int f = blockIdx.x / kMaskBands, band = blockIdx.x % kMaskBands;
  const int e = env_ids ? env_ids[f] : f;
  const int S = image_size;
  const int rows = S / kMaskBands;
  const Lattice4 lt = load_lattice4(st.lattice, e);
  const Fov4 fv = load_fov4(st.fov, e);
  const int si = st.si_idx[e];
  const double w = __dsub_rn(fv.urx, fv.llx), h = __dsub_rn(fv.ury, fv.lly);
  const double stepx = __ddiv_rn(w, static_cast<double>(S));
  const double stepy = __ddiv_rn(h, static_cast<double>(S));
  // rows of the meshgrid (before flipud) this CTA owns: [iy0, iy0 + rows)
  const int iy0 = band * rows;
  const double y_lo = linspace_at(fv.lly, fv.ury, stepy, iy0, S);
  const double y_hi = linspace_at(fv.lly, fv.ury, stepy, iy0 + rows, S);
  const double reach = sqrt(radius_si > radius_c ? radius_si : radius_c) +
                       stepy;  // generous: culling only
  const double2* base = reinterpret_cast<const double2*>(lat.base_xy);
  if (tid == 0) sh.n = 0;
  __syncthreads();
  for (int k0 = 0; k0 < lat.n_sites; k0 += kMaskThreads) {
    const int k = k0 + tid;
    bool keep = false;
    double2 m = make_double2(0.0, 0.0);
    if (k < lat.n_sites) {
      const double2 p = site_position(__ldg(base + k), lt);
      if ((fv.llx <= p.x) && (p.x <= fv.urx) && (fv.lly <= p.y) &&
          (p.y <= fv.ury)) {
        const double qx = __ddiv_rn(__dsub_rn(p.x, fv.llx), w);
        const double qy = __ddiv_rn(__dsub_rn(p.y, fv.lly), h);
        m.x = __dadd_rn(__dmul_rn(qx, w), fv.llx);
        m.y = __dadd_rn(__dmul_rn(qy, h), fv.lly);
        keep = m.y >= y_lo - reach && m.y <= y_hi + reach;
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) sh.warp_count[warp] = __popc(bal);
    __syncthreads();
    int off = sh.n;
    for (int w2 = 0; w2 < warp; ++w2) off += sh.warp_count[w2];
    const int pos = off + __popc(bal & ((1u << lane) - 1u));
    if (keep && pos < kMaskCap) {
      sh.pos[pos] = m;
      sh.z[pos] = static_cast<uint8_t>(k == si ? kSilicon : kCarbon);
    }
    __syncthreads();
    if (tid == 0) {
      int total = sh.n;
      for (int w2 = 0; w2 < kMaskThreads / 32; ++w2) total += sh.warp_count[w2];
      sh.n = total < kMaskCap ? total : kMaskCap;
    }
    __syncthreads();
  }
  const int n = sh.n;
  uint8_t* out = mask_out + static_cast<size_t>(f) * S * S;
  for (int p = tid; p < rows * S; p += kMaskThreads) {
    const int iy = iy0 + p / S, ix = p % S;
    const double xa = linspace_at(fv.llx, fv.urx, stepx, ix, S);
    const double xb = linspace_at(fv.llx, fv.urx, stepx, ix + 1, S);
    const double ya = linspace_at(fv.lly, fv.ury, stepy, iy, S);
    const double yb = linspace_at(fv.lly, fv.ury, stepy, iy + 1, S);
    const double xx = __ddiv_rn(__dadd_rn(xa, xb), 2.0);
    const double yy = __ddiv_rn(__dadd_rn(ya, yb), 2.0);
    uint8_t v = 0;
    for (int i = 0; i < n; ++i) {
      const double2 a = sh.pos[i];
      const double dx = __dsub_rn(xx, a.x), dy = __dsub_rn(yy, a.y);
      const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
      const uint8_t z = sh.z[i];
      if (d2 < (z == kSilicon ? radius_si : radius_c)) v = z;
    }
    out[static_cast<size_t>(S - 1 - iy) * S + ix] = v;  // np.flipud
  }
}

int validate_common(const pd_lattice* lat, const pd_state* st,
                    const pd_rate_config* rc);

}  // namespace pd

extern "C" int pd_render_mask(const pd_lattice* lat, const pd_state* st,
                              const int32_t* env_ids, int32_t m,
                              int32_t image_size, double radius_carbon,
                              double radius_silicon, uint8_t* mask_out,
                              void* stream) {
  int rcode = pd::validate_common(lat, st, nullptr);
  if (rcode != PD_OK) return rcode;
  PD_REQUIRE(m >= 0, "negative frame count");
  PD_REQUIRE(env_ids != nullptr || m <= st->n_envs, "m exceeds n_envs");
  PD_REQUIRE(image_size >= 8 && image_size <= 4096 && image_size % 8 == 0,
             "image_size must be a multiple of 8 in [8, 4096]");
  PD_REQUIRE(radius_carbon >= 0.0 && radius_silicon >= 0.0, "negative radius");
  PD_REQUIRE(lat->n_sites <= pd::kMaskCap, "lattice exceeds 2048 sites");
  if (m == 0) return PD_OK;
  PD_REQUIRE(mask_out != nullptr, "null mask_out");
  pd::k_grid_mask<<<m * pd::kMaskBands, pd::kMaskThreads, 0,
                    static_cast<cudaStream_t>(stream)>>>(
      *lat, *st, env_ids, image_size, radius_carbon, radius_silicon, mask_out);
  PD_CUDA_OK(cudaGetLastError());
  return PD_OK;
}
