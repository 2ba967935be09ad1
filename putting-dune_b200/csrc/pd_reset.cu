// K4: reset.  One thread per environment.
//
//   graphene.py:533-559  generate_pristine_graphene (offset + rotation draws)
//   graphene.py:584-598  PristineSingleDopedGraphene.reset (Si = site nearest
//                        the origin)
//   simulator.py:65-105  PuttingDuneSimulator.reset (FOV scale, FOV centred on
//                        the Si)
//   imaging.py:42-54     sample_image_parameters
//
// Draw order on the RESET stream (linear: draw k = slot k/2, half k%2):
//   0 offset x, 1 offset y, 2 angle, 3 fov scale, 4..12 image parameters.
// Compiled with -fmad=false.
#include <math.h>

#include "pd_kmc.cuh"

namespace pd {

void forget_plan_hint(const void* key);  // pd_step_fast.cu

constexpr int kResetThreads = 128;

// imaging.py:42-54 sample_image_parameters (mode 0) and :57-72
// sample_noisy_image_parameters (mode 1) from nine uniforms in dataclass
// order; rng.uniform(lo, hi) = lo + (hi - lo) * u, rng.exponential(15) =
// -log1p(-u) * 15.
__device__ __forceinline__ void write_image_params(double* ip, const double* u,
                                                   int mode) {
  const double var_hi = mode ? 0.3 : 5e-3, sp_hi = mode ? 1e-2 : 1e-3;
  const double blur_hi = mode ? 0.25 : 1.0;
  const double g_lo = mode ? 0.5 : 0.7, g_hi = mode ? 1.5 : 1.3;
  const double exp_hi = mode ? 0.25 : 0.2, uni_hi = mode ? 0.25 : 0.2;
  ip[0] = __dadd_rn(1.4, __dmul_rn(2.0 - 1.4, u[0]));
  ip[1] = __dmul_rn(var_hi, u[1]);
  ip[2] = __dmul_rn(5.0, u[2]);
  ip[3] = __dadd_rn(__dmul_rn(-log1p(-u[3]), 15.0), 1.0);
  ip[4] = __dmul_rn(sp_hi, u[4]);
  ip[5] = __dmul_rn(blur_hi, u[5]);
  ip[6] = __dadd_rn(g_lo, __dmul_rn(g_hi - g_lo, u[6]));
  ip[7] = __dmul_rn(exp_hi, u[7]);
  ip[8] = __dmul_rn(uni_hi, u[8]);
}

// Re-draws the image parameters of the env's current episode (the uniforms
// the last reset used: draws 4..12 of RESET sequence episode - 1).
__global__ void __launch_bounds__(kResetThreads)
    k_sample_image_params(const pd_state st, const uint8_t* __restrict__ mask,
                          int mode) {
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
       e < st.n_envs; e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (mask && !mask[e]) continue;
    const uint32_t env = st.env_offset + static_cast<uint32_t>(e);
    const uint32_t ep = st.episode[e] - 1u;
    double d[10];
#pragma unroll
    for (int k = 2; k < 7; ++k) {
      const uint4 w = philox4x32_10(env, ep, k, PD_STREAM_RESET, st.seed);
      d[2 * (k - 2)] = u53(w.x, w.y);
      d[2 * (k - 2) + 1] = u53(w.z, w.w);
    }
    write_image_params(st.image_params + 9 * e, d, mode);
  }
}

__global__ void __launch_bounds__(kResetThreads)
    k_reset(const pd_lattice lat, const pd_state st,
            const uint8_t* __restrict__ mask) {
  const double2* base = reinterpret_cast<const double2*>(lat.base_xy);
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
       e < st.n_envs; e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (mask && !mask[e]) continue;
    const uint32_t env = st.env_offset + static_cast<uint32_t>(e);
    const uint32_t ep = st.episode[e];
    // draws 2k and 2k+1 share one Philox call
    double d[14];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const uint4 w = philox4x32_10(env, ep, k, PD_STREAM_RESET, st.seed);
      d[2 * k] = u53(w.x, w.y);
      d[2 * k + 1] = u53(w.z, w.w);
    }
    // rng.uniform(-0.71, 0.71, size=(1, 2)) == low + (high - low) * u
    const double half = kBond / 2.0;
    Lattice4 t;
    t.ox = __dadd_rn(-half, __dmul_rn(__dsub_rn(half, -half), d[0]));
    t.oy = __dadd_rn(-half, __dmul_rn(__dsub_rn(half, -half), d[1]));
    const double angle = __dmul_rn(2.0 * 3.141592653589793, d[2]);
    sincos(angle, &t.s, &t.c);
    // argmin of the Euclidean norm over the central candidates (ascending
    // site order, strict <: first index on ties, like np.argmin).
    double best = 1e300;
    int best_k = 0;
    double2 psi = make_double2(0.0, 0.0);
    for (int j = 0; j < lat.n_sites; ++j) {
      const int k = __ldg(lat.nbr + 4 * j + 3) & kCentreListEnd;
      if (k == kCentreListEnd) break;
      const double2 p = site_position(__ldg(base + k), t);
      const double dist =
          __dsqrt_rn(__dadd_rn(__dmul_rn(p.x, p.x), __dmul_rn(p.y, p.y)));
      if (dist < best) {
        best = dist;
        best_k = k;
        psi = p;
      }
    }
    const double scale = __dadd_rn(15.0, __dmul_rn(30.0 - 15.0, d[3]));
    st.si_idx[e] = best_k;
    reinterpret_cast<double2*>(st.lattice)[2 * e] = make_double2(t.ox, t.oy);
    reinterpret_cast<double2*>(st.lattice)[2 * e + 1] = make_double2(t.c, t.s);
    st.fov_scale[e] = scale;
    store_fov4(st.fov, e, centred_fov(psi, scale));
    write_image_params(st.image_params + 9 * e, d + 4, 0);
    st.episode[e] = ep + 1;
    st.sim_time_us[e] = 0;
    st.n_events[e] = 0;
    st.n_transitions[e] = 0;
    st.status[e] = PD_ENV_OK;
  }
}

int validate_common(const pd_lattice* lat, const pd_state* st,
                    const pd_rate_config* rc);

}  // namespace pd

extern "C" int pd_sample_image_params(const pd_state* st, const uint8_t* mask,
                                      int32_t mode, void* stream) {
  PD_REQUIRE(st != nullptr && st->n_envs >= 0, "null state");
  PD_REQUIRE(mode == PD_IMAGE_PARAMS_DEFAULT || mode == PD_IMAGE_PARAMS_NOISY,
             "unknown mode");
  if (st->n_envs == 0) return PD_OK;
  PD_REQUIRE(st->image_params && st->episode, "state has null arrays");
  const int64_t blocks = (st->n_envs + pd::kResetThreads - 1) / pd::kResetThreads;
  const int64_t cap = static_cast<int64_t>(pd::sm_count()) * 16;
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  pd::k_sample_image_params<<<grid, pd::kResetThreads, 0,
                              static_cast<cudaStream_t>(stream)>>>(*st, mask,
                                                                   mode);
  PD_CUDA_OK(cudaGetLastError());
  return PD_OK;
}

extern "C" int pd_reset(const pd_lattice* lat, const pd_state* st,
                        const uint8_t* mask, void* stream) {
  int rcode = pd::validate_common(lat, st, nullptr);
  if (rcode != PD_OK) return rcode;
  if (st->n_envs == 0) return PD_OK;
  PD_REQUIRE(st->image_params && st->episode, "state has null arrays");
  const int64_t blocks = (st->n_envs + pd::kResetThreads - 1) / pd::kResetThreads;
  const int64_t cap = static_cast<int64_t>(pd::sm_count()) * 16;
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  pd::k_reset<<<grid, pd::kResetThreads, 0,
                static_cast<cudaStream_t>(stream)>>>(*lat, *st, mask);
  PD_CUDA_OK(cudaGetLastError());
  pd::forget_plan_hint(st->si_idx);  // pd_step_fast.cu: kernel choice hint
  return PD_OK;
}
