// Whole goal-reaching episodes on the device (BASELINE configs[4]).
//
//   eval_lib.py:77-184                   evaluate (per-seed episode loop)
//   putting_dune_environment.py:87-158   reset / step
//   goals.py:70-185                      SingleSiliconGoalReaching
//   feature_constructors.py:157-228      material-frame features
//   agents/agent_lib.py:163-183          GreedyAgent.step
//   action_adapters.py:219-274           RelativeToSiliconMaterialFrame adapter
//   run_helpers.py:120-153               StepLimitWrapper
//
// Compiled with -fmad=false: the observation round trips
// (material -> microscope -> material) are reproduced operation by operation.
#include <math.h>

#include "pd_kmc.cuh"

namespace pd {

constexpr int kGoalThreads = 128;
constexpr double kGamma = 0.9967;  // constants.py:35

// graphene.py:623-638 then microscope_utils.py:362-369: the reference sees
// positions only through the normalised observed grid.
__device__ __forceinline__ double2 observe(const Fov4& f, const double2 p) {
  return make_double2(
      __ddiv_rn(__dsub_rn(p.x, f.llx), __dsub_rn(f.urx, f.llx)),
      __ddiv_rn(__dsub_rn(p.y, f.lly), __dsub_rn(f.ury, f.lly)));
}

__device__ __forceinline__ double2 round_trip(const Fov4& f, const double2 p) {
  const double2 q = observe(f, p);
  return microscope_to_material(f, q.x, q.y);
}

// goals.py:84-121: goal = the k-th (k = floor(u * n)) observed atom whose
// scaled distance from the Si lies in (0.1, 50) A.  One warp per env.
__global__ void __launch_bounds__(kGoalThreads)
    k_choose_goal(const pd_lattice lat, const pd_state st,
                  double* __restrict__ goal_xy,
                  int32_t* __restrict__ goal_site) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = static_cast<int64_t>(gridDim.x) * (kGoalThreads / 32);
  const double2* base = reinterpret_cast<const double2*>(lat.base_xy);
  for (int64_t e = blockIdx.x * (kGoalThreads / 32) + (threadIdx.x >> 5);
       e < st.n_envs; e += warps) {
    const Lattice4 t = load_lattice4(st.lattice, e);
    const Fov4 f = load_fov4(st.fov, e);
    const double w = __dsub_rn(f.urx, f.llx), h = __dsub_rn(f.ury, f.lly);
    const double2 q_si =
        observe(f, site_position(__ldg(base + st.si_idx[e]), t));
    const double u = draw_linear(st.seed, st.env_offset + static_cast<uint32_t>(e),
                                 st.episode[e] - 1u, PD_STREAM_RESET, 13);
    int target = -1;
    int found_site = -1;
    double2 found_q = make_double2(0.0, 0.0);
    for (int pass = 0; pass < 2; ++pass) {
      int count = 0;
      for (int k0 = 0; k0 < lat.n_sites; k0 += 32) {
        const int k = k0 + lane;
        bool valid = false;
        double2 q = make_double2(0.0, 0.0);
        if (k < lat.n_sites) {
          const double2 p = site_position(__ldg(base + k), t);
          if (f.llx <= p.x && p.x <= f.urx && f.lly <= p.y && p.y <= f.ury) {
            q = observe(f, p);
            const double dx = __dmul_rn(w, __dsub_rn(q.x, q_si.x));
            const double dy = __dmul_rn(h, __dsub_rn(q.y, q_si.y));
            const double dist =
                __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            valid = dist < 50.0 && dist > 0.1;
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, valid);
        if (pass == 1 && valid &&
            count + __popc(m & ((1u << lane) - 1u)) == target) {
          found_site = k;
          found_q = q;
        }
        count += __popc(m);
      }
      if (pass == 0) target = count > 0 ? static_cast<int>(floor(u * count)) : -1;
    }
    // exactly one lane found it; broadcast
    const unsigned who = __ballot_sync(0xffffffffu, found_site >= 0);
    if (who) {
      const int src = __ffs(who) - 1;
      found_site = __shfl_sync(0xffffffffu, found_site, src);
      found_q.x = __shfl_sync(0xffffffffu, found_q.x, src);
      found_q.y = __shfl_sync(0xffffffffu, found_q.y, src);
    }
    if (lane == 0) {
      double2 g = make_double2(nan(""), nan(""));
      if (found_site >= 0) g = microscope_to_material(f, found_q.x, found_q.y);
      reinterpret_cast<double2*>(goal_xy)[e] = g;
      if (goal_site) goal_site[e] = found_site;
      if (found_site < 0) st.status[e] |= PD_ENV_NOT_RESET;  // no valid goal
    }
  }
}

template <int RATE>
__device__ __forceinline__ void eval_rates_ep(const RateArgs& ra,
                                              const double2 beam,
                                              const double2 psi,
                                              const double2 pn[3],
                                              float r[3]) {
  if (RATE == PD_RATE_SIMPLE) {
    rates_simple(beam, psi, pn, r);
  } else if (RATE == PD_RATE_PRIOR) {
    rates_prior(beam, psi, pn, r);
  } else {
    r[0] = ra.constant_rates[0];
    r[1] = ra.constant_rates[1];
    r[2] = ra.constant_rates[2];
  }
}

template <int RATE, bool STAGE>
__global__ void __launch_bounds__(kStepThreads)
    k_episode(const pd_lattice lat, const pd_state st, const RateArgs ra,
              const pd_episode_config cfg, const double* __restrict__ goal_xy,
              pd_episode_stats* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smem[];
  typename std::conditional<STAGE, SharedTables, GlobalTables>::type tab;
  if constexpr (STAGE) {
    tab = stage_tables(lat, smem);
  } else {
    tab.base = reinterpret_cast<const double2*>(lat.base_xy);
    tab.nbr = reinterpret_cast<const int4*>(lat.nbr);
  }
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
       e < st.n_envs; e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const Lattice4 lt = load_lattice4(st.lattice, e);
    Fov4 fov = load_fov4(st.fov, e);
    const double scale = st.fov_scale[e];
    const double2 goal = reinterpret_cast<const double2*>(goal_xy)[e];
    const uint32_t env_id = st.env_offset + static_cast<uint32_t>(e);
    int si = st.si_idx[e];
    double2 psi = site_position(tab.position(si), lt);
    uint32_t ctrl_count = st.ctrl_count[e];
    uint8_t status = st.status[e];
    long long env_time = cfg.image_duration_us;  // eval_lib.py:121
    long long n_events = 0, n_transitions = 0;
    int actions = 0;
    bool reached = false;
    float reward = 0.f;
    const bool has_goal = goal.x == goal.x;
    while (has_goal && env_time < cfg.timeout_us) {  // eval_lib.py:128
      // ---- features (feature_constructors.py:190-221) ----
      int nb[3];
      tab.neighbors(si, nb);
      double2 pn[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) pn[i] = site_position(tab.position(nb[i]), lt);
      const double2 si_m = round_trip(fov, psi);
      const float gx = __double2float_rn(__dsub_rn(goal.x, si_m.x));
      const float gy = __double2float_rn(__dsub_rn(goal.y, si_m.y));
      // ---- GreedyAgent.step in float32 (agent_lib.py:163-183) ----
      int best = 0;
      float best_score = 0.f, bdx = 0.f, bdy = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double2 nm = round_trip(fov, pn[i]);
        const float dx = __double2float_rn(__dsub_rn(nm.x, si_m.x));
        const float dy = __double2float_rn(__dsub_rn(nm.y, si_m.y));
        const float ex = __fsub_rn(dx, gx), ey = __fsub_rn(dy, gy);
        const float score =
            __fsqrt_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)));
        if (i == 0 || score < best_score) {
          best = i;
          best_score = score;
          bdx = dx;
          bdy = dy;
        }
      }
      (void)best;
      const float angle = atan2f(bdy, bdx);
      const double c = static_cast<double>(cosf(angle));
      const double s = static_cast<double>(sinf(angle));
      // rotate_coordinates(argmax, angle): (x c - y s, x s + y c)
      const double ax = __dadd_rn(__dmul_rn(cfg.argmax_x, c),
                                  __dmul_rn(cfg.argmax_y, -s));
      const double ay = __dadd_rn(__dmul_rn(cfg.argmax_x, s),
                                  __dmul_rn(cfg.argmax_y, c));
      // ---- adapter (action_adapters.py:231-256) ----
      const double2 target =
          make_double2(__dadd_rn(si_m.x, ax), __dadd_rn(si_m.y, ay));
      double2 ctl = observe(fov, target);
      ctl.x = fmin(fmax(ctl.x, 0.0), 1.0);
      ctl.y = fmin(fmax(ctl.y, 0.0), 1.0);
      // ---- simulator.step_and_image (simulator.py:107-182) ----
      const double2 beam = microscope_to_material(fov, ctl.x, ctl.y);
      long long elapsed = 0;
      uint32_t it = 0;
      while (elapsed < cfg.dwell_us) {
        if (it > 0) {
          tab.neighbors(si, nb);
#pragma unroll
          for (int i = 0; i < 3; ++i)
            pn[i] = site_position(tab.position(nb[i]), lt);
        }
        float r[3];
        eval_rates_ep<RATE>(ra, beam, psi, pn, r);
        const uint4 w = philox4x32_10(env_id, ctrl_count, it, PD_STREAM_KMC,
                                      st.seed);
        int slot = 0;
        bool bad = false;
        const bool hit = kmc_event(r, u53(w.x, w.y), u53(w.z, w.w),
                                   cfg.dwell_us, &elapsed, &slot, &bad);
        if (bad) status |= PD_ENV_BAD_RATE;
        ++n_events;
        if (hit) {
          si = nb[slot];
          psi = slot == 0 ? pn[0] : (slot == 1 ? pn[1] : pn[2]);
          ++n_transitions;
        }
        ++it;
      }
      ++ctrl_count;
      long long step_us = cfg.dwell_us + cfg.image_duration_us;
      if (silicon_outside_safe_area(fov, psi)) {
        fov = centred_fov(psi, scale);
        step_us += cfg.image_duration_us;
      }
      env_time += step_us;
      ++actions;
      // ---- goals.py:143-181 on the new observation ----
      const double2 now_m = round_trip(fov, psi);
      const double ddx = __dsub_rn(now_m.x, goal.x);
      const double ddy = __dsub_rn(now_m.y, goal.y);
      const double dist =
          __dsqrt_rn(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)));
      if (dist < kBond * 0.5) {
        reached = true;
        reward = static_cast<float>(
            pow(kGamma, static_cast<double>(step_us) / 1e6));
        break;
      }
      if (actions >= cfg.step_limit) break;  // StepLimitWrapper truncation
    }
    st.si_idx[e] = si;
    st.ctrl_count[e] = ctrl_count;
    store_fov4(st.fov, e, fov);
    st.sim_time_us[e] = env_time;
    st.n_events[e] = n_events;
    st.n_transitions[e] = n_transitions;
    st.status[e] = status;
    pd_episode_stats out;
    out.num_actions = actions;
    out.env_seconds = reached ? static_cast<float>(
                                    static_cast<double>(env_time) / 1e6)
                              : nanf("");
    out.total_reward = reward;
    out.reached_goal = reached ? 1 : 0;
    out.pad_[0] = out.pad_[1] = out.pad_[2] = 0;
    stats[e] = out;
  }
}

int validate_common(const pd_lattice* lat, const pd_state* st,
                    const pd_rate_config* rc);

template <int RATE>
static int launch_episode(const pd_lattice* lat, const pd_state* st,
                          const RateArgs& ra, const pd_episode_config& cfg,
                          const double* goal_xy, pd_episode_stats* stats,
                          cudaStream_t s) {
  const int64_t blocks = (st->n_envs + kStepThreads - 1) / kStepThreads;
  const bool staged = st->n_envs >= 2LL * sm_count() * kStepThreads;
  const int64_t cap = static_cast<int64_t>(sm_count()) * (staged ? 4 : 16);
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  if (staged) {
    const size_t smem = static_cast<size_t>(lat->n_sites) *
                        (sizeof(double2) + sizeof(ushort4));
    PD_CUDA_OK(cudaFuncSetAttribute(
        k_episode<RATE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
        static_cast<int>(smem)));
    k_episode<RATE, true><<<grid, kStepThreads, smem, s>>>(*lat, *st, ra, cfg,
                                                           goal_xy, stats);
  } else {
    k_episode<RATE, false><<<grid, kStepThreads, 0, s>>>(*lat, *st, ra, cfg,
                                                         goal_xy, stats);
  }
  PD_CUDA_OK(cudaGetLastError());
  return PD_OK;
}

}  // namespace pd

extern "C" int pd_run_episodes(const pd_lattice* lat, const pd_state* st,
                               const pd_rate_config* rc,
                               const pd_episode_config* cfg, double* goal_xy,
                               int32_t* goal_site, pd_episode_stats* stats,
                               void* stream) {
  int rcode = pd::validate_common(lat, st, rc);
  if (rcode != PD_OK) return rcode;
  PD_REQUIRE(rc && cfg, "null rate / episode config");
  PD_REQUIRE(cfg->dwell_us >= 0 && cfg->image_duration_us >= 0 &&
                 cfg->timeout_us >= 0 && cfg->step_limit >= 0,
             "negative episode limits");
  if (st->n_envs == 0) return PD_OK;
  PD_REQUIRE(goal_xy && stats, "null goal workspace / stats");
  rcode = pd_reset(lat, st, nullptr, stream);
  if (rcode != PD_OK) return rcode;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  {
    const int64_t blocks = (st->n_envs + 3) / 4;
    const int64_t cap = static_cast<int64_t>(pd::sm_count()) * 16;
    pd::k_choose_goal<<<static_cast<int>(blocks < cap ? blocks : cap),
                        pd::kGoalThreads, 0, s>>>(*lat, *st, goal_xy,
                                                  goal_site);
    PD_CUDA_OK(cudaGetLastError());
  }
  pd::RateArgs ra{};
  for (int i = 0; i < 3; ++i) ra.constant_rates[i] = rc->constant_rates[i];
  switch (rc->rate_fn) {
    case PD_RATE_SIMPLE:
      return pd::launch_episode<PD_RATE_SIMPLE>(lat, st, ra, *cfg, goal_xy,
                                                stats, s);
    case PD_RATE_PRIOR:
      return pd::launch_episode<PD_RATE_PRIOR>(lat, st, ra, *cfg, goal_xy,
                                               stats, s);
    case PD_RATE_CONSTANT:
      return pd::launch_episode<PD_RATE_CONSTANT>(lat, st, ra, *cfg, goal_xy,
                                                  stats, s);
    default:
      pd::set_error("pd_run_episodes: rate_fn %d is not supported",
                    rc->rate_fn);
      return PD_ERR_UNSUPPORTED;
  }
}
