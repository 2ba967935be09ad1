// Batched RL environment layer.
//
//   putting_dune_environment.py:87-158  PuttingDuneEnvironment.reset / step
//   run_helpers.py:120-153              StepLimitWrapper
//   action_adapters.py:53-274           the four action adapters
//   feature_constructors.py:79-228      the two 10-float feature constructors
//   goals.py:143-181                    reward / terminal
//
// One pd_env_step = adapter kernel -> masked reset + goal selection for the
// envs whose episode ended -> the stepping kernel (K1) for the others ->
// feature / reward kernel.  Compiled with -fmad=false.
#include <math.h>

#include "pd_episode.cuh"

namespace pd {

constexpr int kEnvThreads = 128;

__global__ void __launch_bounds__(kEnvThreads)
    k_env_pre(const pd_lattice lat, const pd_state st, const pd_env_config cfg,
              const pd_env_buffers buf, const double* __restrict__ actions,
              long long fixed_dwell_us) {
  const double2* base = reinterpret_cast<const double2*>(lat.base_xy);
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
       e < st.n_envs; e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    // env.step on a fresh / finished episode resets instead
    // (putting_dune_environment.py:114-115, run_helpers.py:135-137)
    const bool resetting = buf.needs_reset[e] || buf.elapsed_steps[e] == -1;
    buf.resetting[e] = resetting ? 1 : 0;
    if (resetting) continue;
    const double* a = actions + e * cfg.action_dim;
    double2 ctl;
    long long dwell = fixed_dwell_us;
    if (cfg.adapter == PD_ADAPTER_DIRECT) {
      ctl = make_double2(clip_nan(a[0], 0.0, 1.0), clip_nan(a[1], 0.0, 1.0));
    } else if (cfg.adapter == PD_ADAPTER_DELTA) {
      double2 b = reinterpret_cast<double2*>(buf.beam_pos)[e];
      b.x = clip_nan(__dadd_rn(b.x, a[0]), 0.0, 1.0);
      b.y = clip_nan(__dadd_rn(b.y, a[1]), 0.0, 1.0);
      reinterpret_cast<double2*>(buf.beam_pos)[e] = b;
      ctl = b;
    } else {
      const Lattice4 lt = load_lattice4(st.lattice, e);
      const Fov4 fov = load_fov4(st.fov, e);
      const double2 psi = site_position(__ldg(base + st.si_idx[e]), lt);
      if (cfg.adapter == PD_ADAPTER_RELATIVE) {
        ctl = relative_to_silicon(fov, psi, make_double2(a[0], a[1]),
                                  cfg.max_distance_angstroms);
      } else {  // action_adapters.py:231-256
        const double2 si_m = round_trip(fov, psi);
        ctl = observe(fov, make_double2(__dadd_rn(si_m.x, a[0]),
                                        __dadd_rn(si_m.y, a[1])));
        ctl.x = clip_nan(ctl.x, 0.0, 1.0);
        ctl.y = clip_nan(ctl.y, 0.0, 1.0);
      }
      if (cfg.action_dim == 3) {  // action_adapters.py:193-199
        const double frac = fmin(fmax(a[2], 0.0), 1.0);
        const double secs = __dadd_rn(
            __dmul_rn(frac, __dsub_rn(cfg.max_dwell_s, cfg.min_dwell_s)),
            cfg.min_dwell_s);
        dwell = seconds_to_us(secs);
      }
    }
    reinterpret_cast<double2*>(buf.controls_xy)[e] = ctl;
    buf.dwell_us[e] = dwell;
  }
}

__global__ void __launch_bounds__(kEnvThreads)
    k_env_post(const pd_lattice lat, const pd_state st,
               const pd_env_config cfg, const pd_env_buffers buf,
               float* __restrict__ obs, float* __restrict__ reward,
               float* __restrict__ discount, int32_t* __restrict__ step_type) {
  const double2* base = reinterpret_cast<const double2*>(lat.base_xy);
  const int4* nbr = reinterpret_cast<const int4*>(lat.nbr);
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
       e < st.n_envs; e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const Lattice4 lt = load_lattice4(st.lattice, e);
    const Fov4 fov = load_fov4(st.fov, e);
    const int si = st.si_idx[e];
    const double2 psi = site_position(__ldg(base + si), lt);
    const double2 goal = reinterpret_cast<const double2*>(buf.goal_xy)[e];
    int type = PD_STEP_MID;
    float rew = 0.f, disc = 0.f;
    if (buf.resetting[e]) {
      type = PD_STEP_FIRST;
      disc = static_cast<float>(
          pow(kGamma, static_cast<double>(cfg.image_duration_us) / 1e6));
      buf.elapsed_steps[e] = 0;
      buf.needs_reset[e] = 0;
      if (cfg.adapter == PD_ADAPTER_DELTA) {
        // DeltaPositionActionAdapter.reset: rng.uniform(0, 1, size=2)
        const uint32_t env = st.env_offset + static_cast<uint32_t>(e);
        const uint32_t ep = st.episode[e] - 1u;
        reinterpret_cast<double2*>(buf.beam_pos)[e] = make_double2(
            draw_linear(st.seed, env, ep, PD_STREAM_RESET, 13),
            draw_linear(st.seed, env, ep, PD_STREAM_RESET, 14));
      }
    } else {
      const bool term = goal_reached(fov, psi, goal);
      const double g =
          pow(kGamma, static_cast<double>(buf.elapsed_us[e]) / 1e6);
      rew = term ? static_cast<float>(g) : 0.f;
      disc = term ? 0.f : static_cast<float>(g);
      if (term) {
        type = PD_STEP_LAST;
        buf.needs_reset[e] = 1;
      }
      int steps = buf.elapsed_steps[e] + 1;  // run_helpers.py:146-152
      if (steps >= cfg.step_limit) {
        steps = -1;
        type = PD_STEP_LAST;
      }
      buf.elapsed_steps[e] = steps;
    }
    // ---- features ----
    const int4 nb = __ldg(nbr + si);
    const int nbs[3] = {nb.x, nb.y, nb.z};
    const double2 q_si = observe(fov, psi);
    const double2 si_m = microscope_to_material(fov, q_si.x, q_si.y);
    float* o = obs + 10 * e;
    if (cfg.features == PD_FEATURES_MICROSCOPE) {
      o[0] = static_cast<float>(q_si.x);
      o[1] = static_cast<float>(q_si.y);
    } else {
      o[0] = static_cast<float>(si_m.x);
      o[1] = static_cast<float>(si_m.y);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double2 q_n = observe(fov, site_position(__ldg(base + nbs[i]), lt));
      double dx, dy;
      if (cfg.features == PD_FEATURES_MICROSCOPE) {
        dx = __dsub_rn(q_n.x, q_si.x);
        dy = __dsub_rn(q_n.y, q_si.y);
        const double dist =
            __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        dx = __ddiv_rn(dx, dist);
        dy = __ddiv_rn(dy, dist);
      } else {
        const double2 nm = microscope_to_material(fov, q_n.x, q_n.y);
        dx = __dsub_rn(nm.x, si_m.x);
        dy = __dsub_rn(nm.y, si_m.y);
      }
      o[2 + 2 * i] = static_cast<float>(dx);
      o[3 + 2 * i] = static_cast<float>(dy);
    }
    o[8] = static_cast<float>(__dsub_rn(goal.x, si_m.x));
    o[9] = static_cast<float>(__dsub_rn(goal.y, si_m.y));
    reward[e] = rew;
    discount[e] = disc;
    step_type[e] = type;
  }
}

int validate_common(const pd_lattice* lat, const pd_state* st,
                    const pd_rate_config* rc);
int step_and_image_masked(const pd_lattice* lat, const pd_state* st,
                          const pd_rate_config* rc, const double* controls_xy,
                          const int64_t* dwell_us, int64_t image_duration_us,
                          const uint8_t* skip, const pd_step_out* out,
                          void* stream);
int choose_goals_masked(const pd_lattice* lat, const pd_state* st,
                        double* goal_xy, const uint8_t* mask,
                        uint32_t draw_index, cudaStream_t s);

// dt.timedelta(seconds=x) on the host (same rounding as seconds_to_us).
static long long host_seconds_to_us(double t) {
  const double whole = trunc(t);
  return static_cast<long long>(whole) * 1000000LL +
         static_cast<long long>(nearbyint((t - whole) * 1e6));
}

}  // namespace pd

extern "C" int pd_env_step(const pd_lattice* lat, const pd_state* st,
                           const pd_rate_config* rc, const pd_env_config* cfg,
                           const pd_env_buffers* buf, const double* actions,
                           float* observation, float* reward, float* discount,
                           int32_t* step_type, void* stream) {
  int rcode = pd::validate_common(lat, st, rc);
  if (rcode != PD_OK) return rcode;
  PD_REQUIRE(rc && cfg && buf, "null config");
  PD_REQUIRE(cfg->adapter >= PD_ADAPTER_DIRECT &&
                 cfg->adapter <= PD_ADAPTER_RELATIVE_MATERIAL,
             "unknown adapter");
  PD_REQUIRE(cfg->features == PD_FEATURES_MICROSCOPE ||
                 cfg->features == PD_FEATURES_MATERIAL,
             "unknown feature constructor");
  PD_REQUIRE(cfg->action_dim == 2 ||
                 (cfg->action_dim == 3 && cfg->adapter >= PD_ADAPTER_RELATIVE),
             "action_dim must be 2 (or 3 for the relative adapters)");
  PD_REQUIRE((cfg->action_dim == 3) == (cfg->adapter >= PD_ADAPTER_RELATIVE &&
                                        cfg->min_dwell_s != cfg->max_dwell_s),
             "action_dim 3 <=> a dwell range on a relative adapter");
  PD_REQUIRE(cfg->step_limit > 0 && cfg->image_duration_us >= 0 &&
                 cfg->min_dwell_s >= 0 && cfg->max_dwell_s >= cfg->min_dwell_s,
             "bad limits");
  if (st->n_envs == 0) return PD_OK;
  PD_REQUIRE(buf->goal_xy && buf->beam_pos && buf->elapsed_steps &&
                 buf->needs_reset && buf->controls_xy && buf->dwell_us &&
                 buf->elapsed_us && buf->resetting,
             "null env buffers");
  PD_REQUIRE(actions && observation && reward && discount && step_type,
             "null actions / outputs");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t blocks = (st->n_envs + pd::kEnvThreads - 1) / pd::kEnvThreads;
  const int64_t cap = static_cast<int64_t>(pd::sm_count()) * 16;
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  // Direct / Delta adapters dwell 1.5 s (action_adapters.py:74,114)
  const long long fixed_dwell =
      cfg->adapter >= PD_ADAPTER_RELATIVE
          ? pd::host_seconds_to_us(cfg->min_dwell_s)
          : 1500000LL;
  pd::k_env_pre<<<grid, pd::kEnvThreads, 0, s>>>(*lat, *st, *cfg, *buf, actions,
                                                 fixed_dwell);
  PD_CUDA_OK(cudaGetLastError());
  rcode = pd_reset(lat, st, buf->resetting, stream);
  if (rcode != PD_OK) return rcode;
  rcode = pd::choose_goals_masked(
      lat, st, buf->goal_xy, buf->resetting,
      cfg->adapter == PD_ADAPTER_DELTA ? 15u : 13u, s);
  if (rcode != PD_OK) return rcode;
  pd_step_out out{};
  out.elapsed_us = buf->elapsed_us;
  rcode = pd::step_and_image_masked(lat, st, rc, buf->controls_xy,
                                    buf->dwell_us, cfg->image_duration_us,
                                    buf->resetting, &out, stream);
  if (rcode != PD_OK) return rcode;
  pd::k_env_post<<<grid, pd::kEnvThreads, 0, s>>>(
      *lat, *st, *cfg, *buf, observation, reward, discount, step_type);
  PD_CUDA_OK(cudaGetLastError());
  return PD_OK;
}
