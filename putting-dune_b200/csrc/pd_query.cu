// Observation queries.
//
//   graphene.py:600-644  PristineSingleDopedGraphene.get_atoms_in_bounds
//   graphene.py:696-700  get_silicon_position
//   graphene.py:581      the `.grid` attribute (all atom positions)
//
// Compiled with -fmad=false.
#include "pd_kmc.cuh"

namespace pd {

constexpr int kQueryThreads = 128;

// One warp per env: ordered stream compaction of the sites inside the box.
__global__ void __launch_bounds__(kQueryThreads)
    k_atoms_in_bounds(const pd_lattice lat, const pd_state st,
                      const double* __restrict__ fov_override,
                      int32_t max_atoms, double* __restrict__ out_xy,
                      uint8_t* __restrict__ out_z,
                      int32_t* __restrict__ out_site,
                      int32_t* __restrict__ out_count) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_per_grid =
      static_cast<int64_t>(gridDim.x) * (kQueryThreads / 32);
  const double2* base = reinterpret_cast<const double2*>(lat.base_xy);
  for (int64_t e = blockIdx.x * (kQueryThreads / 32) + (threadIdx.x >> 5);
       e < st.n_envs; e += warps_per_grid) {
    const Lattice4 t = load_lattice4(st.lattice, e);
    const Fov4 f = load_fov4(fov_override ? fov_override : st.fov, e);
    const int si = st.si_idx[e];
    const double w = __dsub_rn(f.urx, f.llx);
    const double h = __dsub_rn(f.ury, f.lly);
    int count = 0;
    for (int k0 = 0; k0 < lat.n_sites; k0 += 32) {
      const int k = k0 + lane;
      bool keep = false;
      double2 p = make_double2(0.0, 0.0);
      if (k < lat.n_sites) {
        p = site_position(__ldg(base + k), t);
        keep = (f.llx <= p.x) && (p.x <= f.urx) && (f.lly <= p.y) &&
               (p.y <= f.ury);
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      const int pos = count + __popc(m & ((1u << lane) - 1u));
      if (keep && pos < max_atoms) {
        const int64_t o = e * max_atoms + pos;
        reinterpret_cast<double2*>(out_xy)[o] =
            make_double2(__ddiv_rn(__dsub_rn(p.x, f.llx), w),
                         __ddiv_rn(__dsub_rn(p.y, f.lly), h));
        out_z[o] = static_cast<uint8_t>(k == si ? kSilicon : kCarbon);
        if (out_site) out_site[o] = k;
      }
      count += __popc(m);
    }
    if (lane == 0) out_count[e] = count;
  }
}

__global__ void __launch_bounds__(kQueryThreads)
    k_silicon_position(const pd_lattice lat, const pd_state st,
                       double* __restrict__ out_xy) {
  const double2* base = reinterpret_cast<const double2*>(lat.base_xy);
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
       e < st.n_envs; e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const Lattice4 t = load_lattice4(st.lattice, e);
    reinterpret_cast<double2*>(out_xy)[e] =
        site_position(__ldg(base + st.si_idx[e]), t);
  }
}

__global__ void __launch_bounds__(kQueryThreads)
    k_grid(const pd_lattice lat, const pd_state st,
           const int32_t* __restrict__ env_ids, int32_t m,
           double* __restrict__ out_xy) {
  const double2* base = reinterpret_cast<const double2*>(lat.base_xy);
  const int64_t total = static_cast<int64_t>(m) * lat.n_sites;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
       i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t j = i / lat.n_sites;
    const int k = static_cast<int>(i - j * lat.n_sites);
    const Lattice4 t = load_lattice4(st.lattice, env_ids[j]);
    reinterpret_cast<double2*>(out_xy)[i] = site_position(__ldg(base + k), t);
  }
}

int validate_common(const pd_lattice* lat, const pd_state* st,
                    const pd_rate_config* rc);

static int blocks_for(int64_t items, int per_block) {
  const int64_t b = (items + per_block - 1) / per_block;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  return static_cast<int>(b < 1 ? 1 : (b < cap ? b : cap));
}

}  // namespace pd

extern "C" int pd_get_atoms_in_bounds(const pd_lattice* lat, const pd_state* st,
                                      const double* fov_override,
                                      int32_t max_atoms, double* out_xy,
                                      uint8_t* out_z, int32_t* out_site,
                                      int32_t* out_count, void* stream) {
  int rcode = pd::validate_common(lat, st, nullptr);
  if (rcode != PD_OK) return rcode;
  PD_REQUIRE(max_atoms >= 0 && out_count, "bad outputs");
  PD_REQUIRE(max_atoms == 0 || (out_xy && out_z), "null outputs");
  if (st->n_envs == 0) return PD_OK;
  pd::k_atoms_in_bounds<<<pd::blocks_for(st->n_envs, 4), pd::kQueryThreads, 0,
                          static_cast<cudaStream_t>(stream)>>>(
      *lat, *st, fov_override, max_atoms, out_xy, out_z, out_site, out_count);
  PD_CUDA_OK(cudaGetLastError());
  return PD_OK;
}

extern "C" int pd_get_silicon_position(const pd_lattice* lat,
                                       const pd_state* st, double* out_xy,
                                       void* stream) {
  int rcode = pd::validate_common(lat, st, nullptr);
  if (rcode != PD_OK) return rcode;
  PD_REQUIRE(out_xy != nullptr, "null output");
  if (st->n_envs == 0) return PD_OK;
  pd::k_silicon_position<<<pd::blocks_for(st->n_envs, pd::kQueryThreads),
                           pd::kQueryThreads, 0,
                           static_cast<cudaStream_t>(stream)>>>(*lat, *st,
                                                                out_xy);
  PD_CUDA_OK(cudaGetLastError());
  return PD_OK;
}

extern "C" int pd_get_grid(const pd_lattice* lat, const pd_state* st,
                           const int32_t* env_ids, int32_t m, double* out_xy,
                           void* stream) {
  int rcode = pd::validate_common(lat, st, nullptr);
  if (rcode != PD_OK) return rcode;
  PD_REQUIRE(m >= 0 && (m == 0 || (env_ids && out_xy)), "bad arguments");
  if (m == 0) return PD_OK;
  pd::k_grid<<<pd::blocks_for(static_cast<int64_t>(m) * lat->n_sites,
                              pd::kQueryThreads),
               pd::kQueryThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      *lat, *st, env_ids, m, out_xy);
  PD_CUDA_OK(cudaGetLastError());
  return PD_OK;
}
