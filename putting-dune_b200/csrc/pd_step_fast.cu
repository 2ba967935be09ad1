// K1, fast form: rollouts on the prior / simple rates with the guarded
// float32 iteration of pd_fast.cuh.  Every decision an iteration takes (hop or
// not, which neighbour) is made in float32 when its error bound allows it and
// by replaying the control with the exact float64 code otherwise, so the
// results are those of k_rollout / k_walk (pd_step.cu) bit for bit
// (tests/test_gpu_fast.py) at ~1/4 of the instructions.
//
//   graphene.py:646-694   PristineSingleDopedGraphene.apply_control
//   simulator.py:107-182  PuttingDuneSimulator.step_and_image
//   action_adapters.py:163-188 RelativeToSiliconActionAdapter.get_action
//
// Compiled with -fmad=false (the exact replay shares this translation unit).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "pd_episode.cuh"
#include "pd_fast.cuh"

#ifndef PD_STEP_MIN_BLOCKS
#define PD_STEP_MIN_BLOCKS 4
#endif
#ifndef PD_FAST_MIN_BLOCKS  // resident CTAs per SM of the fast kernels
#define PD_FAST_MIN_BLOCKS 4
#endif

namespace pd {

// The exact control, out of line: it runs for ~2 controls in 10^4.  The env
// registers travel by value so that the callers' copies stay in registers.
template <int RATE, class Tables>
__device__ __noinline__ EnvRegs exact_control(const Tables tab,
                                              const RateArgs& ra, uint64_t seed,
                                              const double2 beam,
                                              long long dwell_us, EnvRegs r) {
  const LogSink log{0, nullptr, nullptr, nullptr};
  run_control<RATE>(tab, ra, seed, beam, dwell_us, 0, 0, log, &r);
  return r;
}

// The Si as the microscope sees it, in float32: its normalised position q in
// the FOV (graphene.py:623-638) and the reciprocal FOV extents.  Synchronised
// with the float64 state when an env is loaded and whenever the float64 test
// runs, advanced by the neighbour offset on a hop (1e-7 per hop; callers
// re-synchronise every kObservedSyncHops hops, the decisions below keep 1e-4
// of margin).  What it decides:
//   inside()    the Si is well inside the safe area (simulator.py:236-249): no
//               float64 test needed, no re-centre;
//   clip_free() the relative adapter's clip to [0, 1]
//               (action_adapters.py:186-188) cannot engage for any action, so
//               beam - Si = clip(action, -1, 1) * max_distance up to 1e-14 A.
constexpr int kObservedSyncHops = 128;

struct Observed {
  float qx, qy, iwx, iwy;
  __device__ __forceinline__ void sync(const Fov4& fov, const double2 psi) {
    iwx = __fdividef(1.0f, static_cast<float>(fov.urx - fov.llx));
    iwy = __fdividef(1.0f, static_cast<float>(fov.ury - fov.lly));
    qx = static_cast<float>(psi.x - fov.llx) * iwx;
    qy = static_cast<float>(psi.y - fov.lly) * iwy;
  }
  // the Si moved by (ox, oy) angstrom
  __device__ __forceinline__ void hop(float ox, float oy) {
    qx = __fmaf_rn(ox, iwx, qx);
    qy = __fmaf_rn(oy, iwy, qy);
  }
  __device__ __forceinline__ bool inside() const {
    return qx > 0.2501f && qx < 0.7499f && qy > 0.2501f && qy < 0.7499f;
  }
  __device__ __forceinline__ bool clip_free(float max_distance) const {
    const float rx = __fmaf_rn(max_distance, iwx, 1e-4f);
    const float ry = __fmaf_rn(max_distance, iwy, 1e-4f);
    return qx > rx && qx < 1.0f - rx && qy > ry && qy < 1.0f - ry;
  }
};

// Host-format policy of the fast kernels.
//   IO == 0  pd_rollout_actions: float64 actions [T][n][2] in, int32 Si site
//            and int64 elapsed microseconds [T][n] out (either may be null);
//   IO == 1  pd_rollout_actions_host_packed: float32 actions in, one uint16
//            per env-step out: Si site | re-centred << 15 (the elapsed time of
//            a step is dwell + image duration * (1 + re-centred),
//            simulator.py:131-169);
//   IO == 2  pd_rollout_actions_host_f32: float32 actions in, int32 Si site
//            and int32 elapsed microseconds out.
template <int IO>
struct ActionStream {
  const void* base;
  int64_t n;
  __device__ __forceinline__ ActionStream(const StepArgs& a)
      : base(IO != 0 ? static_cast<const void*>(a.actions_f32)
                     : static_cast<const void*>(a.controls_xy)),
        n(a.st.n_envs) {}
  // i: linear element index step * n + env
  __device__ __forceinline__ const void* at(int64_t i) const {
    return IO != 0 ? static_cast<const void*>(
                         static_cast<const float2*>(base) + i)
                   : static_cast<const void*>(
                         static_cast<const double2*>(base) + i);
  }
  __device__ __forceinline__ double2 load(int64_t i) const {
    if (IO != 0) {
      const float2 v = *static_cast<const float2*>(at(i));
      return make_double2(static_cast<double>(v.x), static_cast<double>(v.y));
    }
    return *static_cast<const double2*>(at(i));
  }
  __device__ __forceinline__ const void* at(int64_t step, int64_t env) const {
    return at(step * n + env);
  }
  __device__ __forceinline__ double2 load(int64_t step, int64_t env) const {
    return load(step * n + env);
  }
};

// i: linear element index step * n + env
template <int IO>
__device__ __forceinline__ void store_step(const StepArgs& a, int64_t i,
                                           int si, bool recentred,
                                           long long step_us) {
  if (IO == 1) {
    a.packed_out[i] =
        static_cast<uint16_t>(si | (recentred ? 0x8000 : 0));
  } else if (IO == 2) {
    if (a.si_idx_out) a.si_idx_out[i] = si;
    if (a.elapsed32_out)
      a.elapsed32_out[i] = static_cast<int32_t>(
          step_us + (recentred ? a.image_duration_us : 0));
  } else {
    if (a.si_idx_out) a.si_idx_out[i] = si;
    if (a.elapsed_us_out)
      a.elapsed_us_out[i] = step_us + (recentred ? a.image_duration_us : 0);
  }
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
constexpr int kActionsAhead = 8;  // steps of the action stream requested ahead

// ---------------------------------------------------------------------------
// Rare paths of the fast kernels, out of line so that the hot loops hold only
// their own registers.  They read the launch arguments through the kernel's
// (__grid_constant__) parameter.
// ---------------------------------------------------------------------------
struct ReplayResult {
  FastSite s;
  Observed obs;
  uint32_t ctrl_count;
  int transitions, events;
  uint8_t status;
  bool hopped;
};

// The control of step t of env `env`, which began with the Si at si0 and has
// made `it` float32 hops so far, again and exactly (run_control).
template <int RATE, int IO>
__device__ __noinline__ void replay_control(const StepArgs& a, int64_t env,
                                            int t, int si0, uint32_t it,
                                            uint32_t ctrl_count,
                                            int transitions, int events,
                                            uint8_t status, const Fov4 fov,
                                            ReplayResult* out) {
  const GlobalTables tab{reinterpret_cast<const double2*>(a.lat.base_xy),
                         reinterpret_cast<const int4*>(a.lat.nbr)};
  const ActionStream<IO> ctl(a);
  const Lattice4 lat = load_lattice4(a.st.lattice, env);
  EnvRegs r;
  r.si = si0;
  r.psi = site_position(tab.position(si0), lat);
  r.lat = lat;
  r.env_id = a.st.env_offset + static_cast<uint32_t>(env);
  r.ctrl_count = ctrl_count;
  r.transitions = transitions - static_cast<int>(it);
  r.events = events - static_cast<int>(it);
  r.log_n = 0;
  r.status = status;
  const int tr0 = r.transitions;
  double2 pos = ctl.load(t, env);
  if (a.action_mode == PD_ACTION_RELATIVE_TO_SILICON)
    pos = relative_to_silicon(fov, r.psi, pos, a.max_distance);
  const double2 beam = microscope_to_material(fov, pos.x, pos.y);
  const LogSink log{0, nullptr, nullptr, nullptr};
  run_control<RATE>(tab, a.ra, a.st.seed, beam, a.dwell_us_scalar, 0, 0, log,
                    &r);
  out->s = fast_site<RATE>(tab, r.si, lat.c, lat.s);
  out->obs.sync(fov, r.psi);
  out->ctrl_count = r.ctrl_count;
  out->transitions = r.transitions;
  out->events = r.events;
  out->status = r.status;
  out->hopped = r.transitions != tr0;
}

// simulator.py:156-169 in float64 for the Si at site `si`: re-centres the FOV
// (in memory) if the Si has left the safe area; returns whether it did and the
// re-synchronised float32 view.
__device__ __noinline__ bool exact_area_check(const StepArgs& a, int64_t env,
                                              int si, Observed* obs) {
  const double2 base =
      __ldg(reinterpret_cast<const double2*>(a.lat.base_xy) + si);
  Fov4 fov = load_fov4(a.st.fov, env);
  const double2 psi = site_position(base, load_lattice4(a.st.lattice, env));
  bool recentred = false;
  if (silicon_outside_safe_area(fov, psi)) {
    fov = centred_fov(psi, a.st.fov_scale[env]);
    store_fov4(a.st.fov, env, fov);
    recentred = true;
  }
  obs->sync(fov, psi);
  return recentred;
}

// Beam offset from the Si (fast_event's units) of a control given in the
// microscope frame (PD_ACTION_DIRECT; simulator.py:137 in float64).
template <int RATE>
__device__ __noinline__ float2 direct_offset(const StepArgs& a, int64_t env,
                                             int si, const double2 act) {
  const double2 base =
      __ldg(reinterpret_cast<const double2*>(a.lat.base_xy) + si);
  const Fov4 fov = load_fov4(a.st.fov, env);
  const double2 psi = site_position(base, load_lattice4(a.st.lattice, env));
  const double2 beam = microscope_to_material(fov, act.x, act.y);
  const float k = fast_offset_scale<RATE>();
  return make_float2(static_cast<float>(beam.x - psi.x) * k,
                     static_cast<float>(beam.y - psi.y) * k);
}

// ---------------------------------------------------------------------------
// k_walk_fast: large batches.  A lane owns one environment and walks it
// through its n_steps controls, one iteration per trip of the loop; the 32
// environments of a warp start together and the warp moves on when all of
// them are done.  A trip is the float32 iteration for every lane, then one of
// two short blocks: the hop (~10 % of the lanes) or the end of the control
// and the start of the next (the others).  Between trips a lane holds ~35
// registers of env state: the lattice transform, the FOV and the float64 Si
// position are re-read / re-derived in the (out-of-line) places that need
// them.  REL: the relative adapter (action_adapters.py:163-188).
// ---------------------------------------------------------------------------
template <int RATE, int IO, bool REL>
__global__ void __launch_bounds__(kStepThreads, PD_FAST_MIN_BLOCKS)
    k_walk_fast(const __grid_constant__ StepArgs a) {
  // The tables are read through L1: a hop touches one 16-byte row, in one
  // iteration of ten, and staging them would take the shared memory that the
  // action stream's L1 lines need.
  const GlobalTables tab{reinterpret_cast<const double2*>(a.lat.base_xy),
                         reinterpret_cast<const int4*>(a.lat.nbr)};
  const FastTimes tm = fast_times(a.dwell_us_scalar);
  const float md_f = static_cast<float>(a.max_distance);
  // action -> beam offset in the units fast_event expects
  const float md_s = static_cast<float>(
      a.max_distance * (RATE == PD_RATE_PRIOR ? 1.0 / kBond : 1.0));
  const long long step_us = a.dwell_us_scalar + a.image_duration_us;
  const int64_t n = a.st.n_envs;
  const int n_steps = a.n_steps;
  const ActionStream<IO> ctl(a);
  const int lane = threadIdx.x & 31;
  const int64_t n_batches = (n + 31) / 32;
  const int64_t warps_total =
      static_cast<int64_t>(gridDim.x) * (kStepThreads / 32);
  const int64_t wid = static_cast<int64_t>(blockIdx.x) * (kStepThreads / 32) +
                      (threadIdx.x >> 5);

  for (int64_t b = wid; b < n_batches; b += warps_total) {
    const int64_t env = b * 32 + lane;
    bool active = env < n;
    // ---- the env's registers ----
    FastSite s;
    s.si = 0;
    s.nb[0] = s.nb[1] = s.nb[2] = 0;
    s.cls = 2;
#pragma unroll
    for (int i = 0; i < 3; ++i) s.geo.gx[i] = s.geo.gy[i] = 0.f;
    uint32_t env_id = 0, ctrl_count = 0;
    int events = 0, transitions = 0, recentres = 0;
    uint8_t status = 0;
    double2 act_next = make_double2(0.0, 0.0);
    int t = 0;
    int64_t row = env;  // t * n + env: this step's element of every [T][n] array
    uint32_t it = 0;    // iteration of the current control = its hops so far
    int si0 = 0;        // Si site when the control began
    float e_lo = 0.f, e_hi = 0.f, bx = 0.f, by = 0.f;
    bool check_area = true;  // simulator.py:156 can only change its answer
                             // after a hop (and is unknown at call start)
    bool usable = false;
    Observed obs{0.5f, 0.5f, 0.f, 0.f};
    auto rotation = [&]() {
      return reinterpret_cast<const double2*>(a.st.lattice)[2 * env + 1];
    };
    // Starts the control `act` for the Si at s.si.
    auto begin_control = [&](const double2 act) {
      it = 0;
      e_lo = e_hi = 0.f;
      si0 = s.si;
      if (REL) {
        // action_adapters.py:163-188 without the clip to the frame
        const float ax = fminf(fmaxf(static_cast<float>(act.x), -1.f), 1.f);
        const float ay = fminf(fmaxf(static_cast<float>(act.y), -1.f), 1.f);
        bx = ax * md_s;
        by = ay * md_s;
        usable = obs.clip_free(md_f);
      } else {
        const float2 o = direct_offset<RATE>(a, env, s.si, act);
        bx = o.x;
        by = o.y;
        usable = true;
      }
    };
    if (active) {
      const double2 act = ctl.load(row);
      if (n_steps > 1) act_next = ctl.load(row + n);
      // a step takes a few hundred cycles, DRAM a thousand: the action
      // stream is requested several steps ahead (L2 now, L1 two steps ahead
      // in the loop), the next batch's state a whole batch ahead
#pragma unroll 1
      for (int k = 2; k < n_steps && k < 2 + kActionsAhead; ++k)
        prefetch_l2(ctl.at(row + k * n));
      if (env + warps_total * 32 < n) {
        prefetch_env(a, env + warps_total * 32);
        prefetch_l1(ctl.at(row + warps_total * 32));
        if (n_steps > 1) prefetch_l1(ctl.at(row + n + warps_total * 32));
      }
      prefetch_l1(a.st.fov_scale + env);
      const Lattice4 lat = load_lattice4(a.st.lattice, env);
      const Fov4 fov = load_fov4(a.st.fov, env);
      env_id = a.st.env_offset + static_cast<uint32_t>(env);
      ctrl_count = a.st.ctrl_count[env];
      status = a.st.status[env];
      s = fast_site<RATE>(tab, a.st.si_idx[env], lat.c, lat.s);
      obs.sync(fov, site_position(tab.position(s.si), lat));
      begin_control(act);
    }

    while (__any_sync(0xffffffffu, active)) {
      if (!active) continue;
      int kind = FAST_UNSURE, slot = 0;
      float t_lo = 0.f, t_hi = 0.f;
      if (usable) {
        const uint4 w =
            philox4x32_10k(env_id, ctrl_count, it, PD_STREAM_KMC, a.keys);
        kind = fast_event<RATE>(s.geo, bx, by, w.x, w.z, e_lo, e_hi, tm, &slot,
                                &t_lo, &t_hi);
      }
      if (kind == FAST_HOP) {
        float ox, oy;
        fast_hop<RATE>(tab, slot, rotation, &s, &bx, &by, &ox, &oy);
        obs.hop(ox, oy);
        transitions += 1;
        events += 1;
        ++it;
        fast_advance(&e_lo, &e_hi, t_lo, t_hi);
        continue;
      }
      // ---- the control has ended ----
      bool hopped = it > 0;
      if (kind == FAST_UNSURE) {
        // replay it from its start with the exact code
        ReplayResult rr;
        replay_control<RATE, IO>(a, env, t, si0, it, ctrl_count, transitions,
                                 events, status, load_fov4(a.st.fov, env), &rr);
        if (rr.hopped || it > 0) {
          s = rr.s;
          obs = rr.obs;
        }
        hopped = rr.hopped;
        ctrl_count = rr.ctrl_count;
        transitions = rr.transitions;
        events = rr.events;
        status = rr.status;
      } else {
        events += 1;
        ctrl_count += 1;
      }
      // image, safe area (simulator.py:152-169)
      bool recentred = false;
      if (hopped || check_area) {
        check_area = false;
        // the float64 test (and a re-synchronised float32 view) only when
        // float32 cannot rule the re-centre out, or is due for a refresh
        const bool due = (transitions / kObservedSyncHops) !=
                         ((transitions - static_cast<int>(it)) /
                          kObservedSyncHops);
        if (!obs.inside() || due) {
          recentred = exact_area_check(a, env, s.si, &obs);
          recentres += recentred ? 1 : 0;
        }
      }
      store_step<IO>(a, row, s.si, recentred, step_us);
      ++t;
      row += n;
      if (t < n_steps) {
        const double2 act = act_next;
        if (t + 1 < n_steps) act_next = ctl.load(row + n);
        if (t + 3 < n_steps) prefetch_l1(ctl.at(row + 3 * n));
        if (t + 2 + kActionsAhead < n_steps)
          prefetch_l2(ctl.at(row + (2 + kActionsAhead) * n));
        begin_control(act);
      } else {
        const long long total =
            static_cast<long long>(n_steps) * step_us +
            static_cast<long long>(recentres) * a.image_duration_us;
        atomicAdd(reinterpret_cast<unsigned long long*>(a.st.sim_time_us + env),
                  static_cast<unsigned long long>(total));
        a.st.si_idx[env] = s.si;
        a.st.ctrl_count[env] = ctrl_count;
        atomicAdd(reinterpret_cast<unsigned long long*>(a.st.n_events + env),
                  static_cast<unsigned long long>(events));
        atomicAdd(
            reinterpret_cast<unsigned long long*>(a.st.n_transitions + env),
            static_cast<unsigned long long>(transitions));
        a.st.status[env] = status;
        active = false;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// k_rollout_fast: small batches (BASELINE configs[1]: 4096 envs).  A rollout
// of one env is a dependent chain, and a small batch leaves most lanes of the
// machine idle, so a group of G lanes owns one env (its state replicated in
// the group) and looks ahead: lane 0 evaluates the env's true next iteration
// (step t, iteration `it`), lane j > 0 iteration 0 of step t + j under the
// assumption that nothing before it hops -- a control that does not hop
// changes nothing but counters, and ~89 % of the relative_random controls do
// not.  Philox is counter based, so lane j simply uses control counter + j.
// The group commits the prefix of steps that certainly end without a hop and
// then applies what the first other lane found: a hop (float32, certain) or
// an iteration float32 cannot settle, whose control is replayed exactly.
// ---------------------------------------------------------------------------
template <int RATE, int IO>
__global__ void __launch_bounds__(kStepThreads, PD_STEP_MIN_BLOCKS)
    k_rollout_fast(const __grid_constant__ StepArgs a) {
  // The tables are read through L1: a hop touches one 16-byte row, in one
  // iteration of ten, and staging them would take the shared memory that the
  // action stream's L1 lines need.
  const GlobalTables tab{reinterpret_cast<const double2*>(a.lat.base_xy),
                         reinterpret_cast<const int4*>(a.lat.nbr)};
  const FastTimes tm = fast_times(a.dwell_us_scalar);
  const bool relative = a.action_mode == PD_ACTION_RELATIVE_TO_SILICON;
  const float md_f = static_cast<float>(a.max_distance);
  const float md_s = static_cast<float>(
      a.max_distance * (RATE == PD_RATE_PRIOR ? 1.0 / kBond : 1.0));
  const float off_s = fast_offset_scale<RATE>();
  const long long dwell = a.dwell_us_scalar;
  const long long step_us = dwell + a.image_duration_us;
  const int64_t n = a.st.n_envs;
  const int n_steps = a.n_steps;
  const ActionStream<IO> ctl(a);
  const int G = a.lane_stride;  // power of two, 2..32
  const int lane = threadIdx.x & 31;
  const int j = lane & (G - 1);
  const int gbase = lane - j;
  const unsigned gfull = G >= 32 ? 0xffffffffu : ((1u << G) - 1u);
  const unsigned gmask = gfull << gbase;
  const int64_t gtid =
      blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t n_groups = static_cast<int64_t>(gridDim.x) * blockDim.x / G;

  for (int64_t e = gtid / G; e < n; e += n_groups) {
    // the first windows of the env's action column into L1
    // (requesting the env's whole action column into L2 here was measured:
    // 3 % slower than the rolling L1 prefetch alone)
    if (j < n_steps) prefetch_l1(ctl.at(j, e));
    if (G + j < n_steps) prefetch_l1(ctl.at(G + j, e));
    // ---- state of the env, replicated in the G lanes of its group ----
    const Lattice4 lat = load_lattice4(a.st.lattice, e);
    Fov4 fov = load_fov4(a.st.fov, e);
    const double scale = a.st.fov_scale[e];
    const uint32_t env_id = a.st.env_offset + static_cast<uint32_t>(e);
    uint32_t ctrl_count = a.st.ctrl_count[e];
    uint8_t status = a.st.status[e];
    int events = 0, transitions = 0, recentres = 0;
    FastSite s = fast_site<RATE>(tab, a.st.si_idx[e], lat.c, lat.s);
    // float64 Si position: kept current in the direct mode (the beam offset
    // needs it), derived on demand otherwise
    double2 psi = site_position(tab.position(s.si), lat);
    bool psi_ok = true;
    auto get_psi = [&]() {
      if (!psi_ok) {
        psi = site_position(tab.position(s.si), lat);
        psi_ok = true;
      }
      return psi;
    };
    Observed obs;
    obs.sync(fov, psi);
    int hops_synced = 0;    // transitions at the last obs.sync
    int t = 0;              // current step
    uint32_t it = 0;        // next iteration of the current step's control
                            // = its hops so far
    float e_lo = 0.f, e_hi = 0.f;  // clock bounds of the current control
    float bx0 = 0.f, by0 = 0.f;    // its beam offset (it > 0)
    int si0 = s.si;                // Si site at its start (it > 0)
    bool first = true;       // first round: lane 0 only, un-re-centred FOV
    bool need_check = true;  // simulator.py:156 runs at t = 0 and after a hop
    bool stale = true;       // Si or FOV changed since the values below
    bool pending_rec = false, clip_free = false, clip_free_n = false;
    auto rotation = [&]() { return make_double2(lat.c, lat.s); };

    while (t < n_steps) {
      if (stale) {
        // Will the step that ends the current control re-centre the FOV
        // (simulator.py:156-169)?  The steps after it see the new FOV.
        pending_rec = false;
        if (need_check && (!obs.inside() ||
                           transitions - hops_synced >= kObservedSyncHops)) {
          pending_rec = silicon_outside_safe_area(fov, get_psi());
          obs.sync(fov, psi);
          hops_synced = transitions;
        }
        clip_free = obs.clip_free(md_f);
        clip_free_n = clip_free;
        if (pending_rec) {
          Observed next;
          next.sync(centred_fov(psi, scale), psi);
          clip_free_n = next.clip_free(md_f);
        }
        stale = false;
      }
      const bool cont = it > 0;  // lane 0 continues a control already begun
      const int step = t + j;
      const bool valid = (first ? j == 0 : true) && step < n_steps;
      int kind = FAST_UNSURE, slot = 0;
      float t_lo = 0.f, t_hi = 0.f, bx = bx0, by = by0;
      if (valid) {
        bool usable = true;
        if (!(j == 0 && cont)) {
          const double2 c = ctl.load(step, e);
          if (relative) {
            const float ax = fminf(fmaxf(static_cast<float>(c.x), -1.f), 1.f);
            const float ay = fminf(fmaxf(static_cast<float>(c.y), -1.f), 1.f);
            bx = ax * md_s;
            by = ay * md_s;
            usable = j == 0 ? clip_free : clip_free_n;
          } else {
            const double2 p = get_psi();
            const Fov4 f =
                (j > 0 && pending_rec) ? centred_fov(p, scale) : fov;
            const double2 beam = microscope_to_material(f, c.x, c.y);
            bx = static_cast<float>(beam.x - p.x) * off_s;
            by = static_cast<float>(beam.y - p.y) * off_s;
          }
        }
        if (usable) {
          const uint4 w = philox4x32_10k(
              env_id, ctrl_count + static_cast<uint32_t>(j), j == 0 ? it : 0u,
              PD_STREAM_KMC, a.keys);
          kind = fast_event<RATE>(s.geo, bx, by, w.x, w.z, j == 0 ? e_lo : 0.f,
                                  j == 0 ? e_hi : 0.f, tm, &slot, &t_lo, &t_hi);
        }
      }
      if (step + 2 * G < n_steps)
        prefetch_l1(ctl.at(step + 2 * G, e));
      const unsigned valids = (__ballot_sync(gmask, valid) & gmask) >> gbase;
      const unsigned quiet =
          (__ballot_sync(gmask, valid && kind == FAST_NO_HOP) & gmask) >> gbase;
      const unsigned stop = valids & ~quiet;
      const int n_done = stop ? __ffs(stop) - 1 : __popc(valids);
      if (n_done > 0) {
        // the current control and the n_done - 1 after it end without a hop
        const bool rec = pending_rec;  // simulator.py:156-169
        if (j < n_done)
          store_step<IO>(a, static_cast<int64_t>(step) * n + e, s.si,
                         j == 0 && rec, step_us);
        if (rec) {
          fov = centred_fov(get_psi(), scale);
          obs.sync(fov, psi);
          hops_synced = transitions;
          recentres += 1;
          pending_rec = false;
          clip_free = clip_free_n;
        }
        need_check = false;
        events += n_done;
        ctrl_count += static_cast<uint32_t>(n_done);
        t += n_done;
        it = 0;
        e_lo = e_hi = 0.f;
      }
      first = false;
      if (!stop) continue;
      // ---- the first lane that did not stay quiet: step t, its iteration ----
      const int src = gbase + __ffs(stop) - 1;
      const int kind_s = __shfl_sync(gmask, kind, src);
      if (kind_s == FAST_HOP) {
        const int slot_s = __shfl_sync(gmask, slot, src);
        const float tl = __shfl_sync(gmask, t_lo, src);
        const float th = __shfl_sync(gmask, t_hi, src);
        bx0 = __shfl_sync(gmask, bx, src);
        by0 = __shfl_sync(gmask, by, src);
        if (it == 0) si0 = s.si;
        float ox, oy;
        fast_hop<RATE>(tab, slot_s, rotation, &s, &bx0, &by0, &ox, &oy);
        obs.hop(ox, oy);
        psi_ok = false;
        transitions += 1;
        events += 1;
        ++it;
        fast_advance(&e_lo, &e_hi, tl, th);
        need_check = true;
        stale = true;
        continue;
      }
      // ---- float32 cannot settle it: the control of step t, exactly ----
      {
        ReplayResult rr;
        replay_control<RATE, IO>(a, e, t, it > 0 ? si0 : s.si, it, ctrl_count,
                                 transitions, events, status, fov, &rr);
        if (rr.hopped || it > 0) {
          s = rr.s;
          psi_ok = false;
        }
        if (rr.hopped) need_check = true;
        ctrl_count = rr.ctrl_count;
        transitions = rr.transitions;
        events = rr.events;
        status = rr.status;
      }
      const double2 p_now = get_psi();
      const bool rec = need_check && silicon_outside_safe_area(fov, p_now);
      if (j == 0)
        store_step<IO>(a, static_cast<int64_t>(t) * n + e, s.si, rec, step_us);
      if (rec) {
        fov = centred_fov(p_now, scale);
        recentres += 1;
      }
      obs.sync(fov, p_now);
      hops_synced = transitions;
      need_check = false;
      stale = true;
      t += 1;
      it = 0;
      e_lo = e_hi = 0.f;
    }
    if (j == 0) {
      if (recentres > 0) store_fov4(a.st.fov, e, fov);
      const long long total =
          static_cast<long long>(n_steps) * step_us +
          static_cast<long long>(recentres) * a.image_duration_us;
      atomicAdd(reinterpret_cast<unsigned long long*>(a.st.sim_time_us + e),
                static_cast<unsigned long long>(total));
      a.st.si_idx[e] = s.si;
      a.st.ctrl_count[e] = ctrl_count;
      atomicAdd(reinterpret_cast<unsigned long long*>(a.st.n_events + e),
                static_cast<unsigned long long>(events));
      atomicAdd(reinterpret_cast<unsigned long long*>(a.st.n_transitions + e),
                static_cast<unsigned long long>(transitions));
      a.st.status[e] = status;
    }
  }
}

// ---------------------------------------------------------------------------
// Launch (called from launch_step, pd_step.cu).
// ---------------------------------------------------------------------------
template <int RATE>
int launch_fast(const StepArgs& a, bool walk, int grid, cudaStream_t stream) {
  const int io = a.packed_out ? 1 : (a.actions_f32 ? 2 : 0);
  const bool rel = a.action_mode == PD_ACTION_RELATIVE_TO_SILICON;
  void (*kern)(const StepArgs) = nullptr;
  if (walk) {
    kern = io == 1   ? (rel ? k_walk_fast<RATE, 1, true>
                            : k_walk_fast<RATE, 1, false>)
           : io == 2 ? (rel ? k_walk_fast<RATE, 2, true>
                            : k_walk_fast<RATE, 2, false>)
                     : (rel ? k_walk_fast<RATE, 0, true>
                            : k_walk_fast<RATE, 0, false>);
  } else {
    kern = io == 1   ? k_rollout_fast<RATE, 1>
           : io == 2 ? k_rollout_fast<RATE, 2>
                     : k_rollout_fast<RATE, 0>;
  }
  if (walk) {
    // persistent warps, one batch of 32 envs at a time: a whole number of
    // CTAs per SM
    const int64_t want =
        (a.st.n_envs + kStepThreads - 1) / kStepThreads;
    const int64_t cap = static_cast<int64_t>(sm_count()) * PD_FAST_MIN_BLOCKS;
    grid = static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
  }
  kern<<<grid, kStepThreads, 0, stream>>>(a);
  PD_CUDA_OK(cudaGetLastError());
  return PD_OK;
}

template int launch_fast<PD_RATE_SIMPLE>(const StepArgs&, bool, int,
                                         cudaStream_t);
template int launch_fast<PD_RATE_PRIOR>(const StepArgs&, bool, int,
                                        cudaStream_t);

// ---------------------------------------------------------------------------
// pd_fast_path_audit: how far the float32 quantities of fast_event are from
// the exact ones, against the bounds fast_event assumes for them.
// One sample = one iteration at a random bulk or edge site, lattice angle,
// beam offset (a disc of `max_distance` around the Si), Philox draw and
// clock; the exact side is rates_* / kmc_event_drawn of pd_kmc.cuh.
// ---------------------------------------------------------------------------
struct AuditStats {
  unsigned long long samples, no_hop, hop, unsure, wrong_decision, wrong_slot,
      t_outside, reserved;
  // largest observed error / assumed bound (float bits, all non-negative)
  unsigned int tot_ratio, t_ratio, choice_ratio, draw_abs;
};

template <int RATE>
__global__ void __launch_bounds__(256)
    k_fast_audit(const pd_lattice lat, uint64_t seed, int64_t n_samples,
                 long long dwell_us, double max_distance, AuditStats* out) {
  GlobalTables tab{reinterpret_cast<const double2*>(lat.base_xy),
                   reinterpret_cast<const int4*>(lat.nbr)};
  const PhiloxKeys keys = philox_keys(seed);
  const FastTimes tm = fast_times(dwell_us);
  const double off_s = RATE == PD_RATE_PRIOR ? 1.0 / kBond : 1.0;
  unsigned long long c_no = 0, c_hop = 0, c_un = 0, c_wd = 0, c_ws = 0, c_to = 0,
                     c_n = 0;
  float m_tot = 0.f, m_t = 0.f, m_ch = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
       i < n_samples; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint32_t lo = static_cast<uint32_t>(i), hi = static_cast<uint32_t>(i >> 32);
    const uint4 s0 = philox4x32_10k(lo, hi, 0u, 100u, keys);
    const uint4 s1 = philox4x32_10k(lo, hi, 1u, 100u, keys);
    const uint4 w = philox4x32_10k(lo, hi, 2u, 100u, keys);
    const uint4 s2 = philox4x32_10k(lo, hi, 3u, 100u, keys);
    const int si = static_cast<int>(s0.x % static_cast<uint32_t>(lat.n_sites));
    double sn, cs;
    sincos(6.283185307179586 * u53(s0.y, s0.z), &sn, &cs);
    const Lattice4 lt{1.42 * (u53(s0.w, s1.x) - 0.5),
                      1.42 * (u53(s1.y, s1.z) - 0.5), cs, sn};
    // beam offset: uniform in the square the relative adapter reaches
    const double ax = 2.0 * u53(s1.w, s2.x) - 1.0;
    const double ay = 2.0 * u53(s2.y, s2.z) - 1.0;
    const double2 psi = site_position(tab.position(si), lt);
    const double2 beam = make_double2(psi.x + ax * max_distance,
                                      psi.y + ay * max_distance);
    // a clock somewhere in the control, exactly representable
    const long long el_us =
        (s2.w & 1u) ? 0 : static_cast<long long>((s2.w >> 1) % 1000u) *
                              (dwell_us / 1000);
    const float e_s = static_cast<float>(static_cast<double>(el_us) * 1e-6);
    const FastGeo geo = fast_site<RATE>(tab, si, lt.c, lt.s).geo;
    const float bx = static_cast<float>(ax) *
                     static_cast<float>(max_distance * off_s);
    const float by = static_cast<float>(ay) *
                     static_cast<float>(max_distance * off_s);
    int slot = 0;
    float t_lo = 0.f, t_hi = 0.f;
    const int kind =
        fast_event<RATE>(geo, bx, by, w.x, w.z, __fadd_rd(e_s, -1e-6f),
                         __fadd_ru(e_s, 1e-6f), tm, &slot, &t_lo, &t_hi);
    // ---- exact ----
    int nb[3];
    tab.neighbors(si, nb);
    double2 pn[3];
    for (int k = 0; k < 3; ++k) pn[k] = site_position(tab.position(nb[k]), lt);
    float r[3];
    if (RATE == PD_RATE_PRIOR)
      rates_prior(beam, psi, pn, r);
    else
      rates_simple(beam, psi, pn, r);
    long long el = el_us;
    int slot_x = 0;
    bool bad = false;
    const bool hop_x = kmc_event(r, u53(w.x, w.y), u53(w.z, w.w), dwell_us, &el,
                                 &slot_x, &bad);
    const bool goes_on = el < dwell_us;
    ++c_n;
    if (kind == FAST_NO_HOP) {
      ++c_no;
      if (hop_x) ++c_wd;
    } else if (kind == FAST_HOP) {
      ++c_hop;
      if (!hop_x || !goes_on) ++c_wd;
      else if (slot != slot_x) ++c_ws;
    } else {
      ++c_un;
    }
    // ---- measured error / assumed bound ----
    const float tot32 = __fadd_rn(__fadd_rn(r[0], r[1]), r[2]);
    if (tot32 > 1e-30f) {
      // the float32 side again, piece by piece (same expressions)
      float a[3], rr[3];
      for (int k = 0; k < 3; ++k) {
        const float dx = bx - geo.gx[k], dy = by - geo.gy[k];
        a[k] = __fmaf_rn(dx, dx, dy * dy);
        rr[k] = RATE == PD_RATE_PRIOR
                    ? ex2_approx(-7.213475204444817f * a[k])
                    : rcp_approx(__fmaf_rn(
                          a[k], static_cast<float>(16.0 / (kBond * kBond)),
                          1.0f));
      }
      const float eps =
          RATE == PD_RATE_PRIOR
              ? __fmaf_rn(5.0f * kFastEps1, fminf(a[0], fminf(a[1], a[2])),
                          kFastEps0)
              : kFastEps0;
      const float sum = (rr[0] + rr[1]) + rr[2];
      const double tot_f =
          (RATE == PD_RATE_PRIOR ? 0.23104906018664842 : 1.0) * sum;
      m_tot = fmaxf(m_tot, static_cast<float>(
                               fabs(tot_f / static_cast<double>(tot32) - 1.0) /
                               eps));
      const double t_x =
          -log1p(-u53(w.x, w.y)) *
          static_cast<double>(__fdiv_rn(1.0f, tot32));
      const double mid = 0.5 * (static_cast<double>(t_lo) + t_hi);
      const double half = 0.5 * (static_cast<double>(t_hi) - t_lo);
      if (half > 0.0 && t_x < 3600.0) {
        m_t = fmaxf(m_t, static_cast<float>(fabs(t_x - mid) / half));
        if (t_x < t_lo || t_x > t_hi) ++c_to;
      }
      const double c0 = static_cast<double>(__fdiv_rn(r[0], tot32));
      const double c1 = c0 + static_cast<double>(__fdiv_rn(r[1], tot32));
      const double c2 = c1 + static_cast<double>(__fdiv_rn(r[2], tot32));
      const float eta = __fmaf_rn(3.0f, eps, 2e-6f);
      const double inv = 1.0 / sum;
      m_ch = fmaxf(m_ch, static_cast<float>(
                             fmax(fabs(rr[0] * inv - c0 / c2),
                                  fabs((rr[0] + rr[1]) * inv - c1 / c2)) /
                             eta));
    }
  }
  atomicAdd(&out->samples, c_n);
  atomicAdd(&out->no_hop, c_no);
  atomicAdd(&out->hop, c_hop);
  atomicAdd(&out->unsure, c_un);
  atomicAdd(&out->wrong_decision, c_wd);
  atomicAdd(&out->wrong_slot, c_ws);
  atomicAdd(&out->t_outside, c_to);
  atomicMax(&out->tot_ratio, __float_as_uint(m_tot));
  atomicMax(&out->t_ratio, __float_as_uint(m_t));
  atomicMax(&out->choice_ratio, __float_as_uint(m_ch));
}

// All 2^24 values of the top-24-bit uniform: the float32 draw
// -ln 2 * lg2.approx(1 - u24) against -log1p(-u24) in float64, as a fraction
// of what fast_event allows for it (kFastDrawAbs absolute + 1e-6 relative).
__global__ void __launch_bounds__(256) k_draw_audit(unsigned int* max_ratio_bits) {
  float m = 0.f;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < (1u << 24);
       k += gridDim.x * blockDim.x) {
    const float v = 1.0f - static_cast<float>(k) * (1.0f / 16777216.0f);
    const float draw = -0.6931471805599453f * lg2_approx(v);
    const double exact = -log1p(-static_cast<double>(k) / 16777216.0);
    const double allowed = static_cast<double>(kFastDrawAbs) + 1e-6 * exact;
    m = fmaxf(m, static_cast<float>(
                     fabs(static_cast<double>(draw) - exact) / allowed));
  }
  atomicMax(max_ratio_bits, __float_as_uint(m));
}

// The short float64 rate forms against the reference's operation sequence
// (pd_kmc.cuh) at random sites / lattice angles / beam offsets.
struct RateOpsStats {
  unsigned long long samples, simple_cast_differs_unguarded,
      simple_cast_differs, simple_guard_taken, prior_cast_differs;
  unsigned int simple_max_ulps, prior_max_ulps;
};

__global__ void __launch_bounds__(256)
    k_rate_ops_audit(const pd_lattice lat, uint64_t seed, int64_t n_samples,
                     double max_distance, RateOpsStats* out) {
  GlobalTables tab{reinterpret_cast<const double2*>(lat.base_xy),
                   reinterpret_cast<const int4*>(lat.nbr)};
  const PhiloxKeys keys = philox_keys(seed);
  unsigned long long c_n = 0, c_su = 0, c_s = 0, c_g = 0, c_p = 0;
  unsigned int m_s = 0, m_p = 0;
  auto ulps = [](double a, double b) {
    const long long d = __double_as_longlong(a) - __double_as_longlong(b);
    const unsigned long long u = d < 0 ? -d : d;
    return u > 0xFFFFFFFFull ? 0xFFFFFFFFu : static_cast<unsigned int>(u);
  };
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
       i < n_samples; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint32_t lo = static_cast<uint32_t>(i), hi = static_cast<uint32_t>(i >> 32);
    const uint4 s0 = philox4x32_10k(lo, hi, 0u, 101u, keys);
    const uint4 s1 = philox4x32_10k(lo, hi, 1u, 101u, keys);
    const uint4 s2 = philox4x32_10k(lo, hi, 2u, 101u, keys);
    const int si = static_cast<int>(s0.x % static_cast<uint32_t>(lat.n_sites));
    double sn, cs;
    sincos(6.283185307179586 * u53(s0.y, s0.z), &sn, &cs);
    const Lattice4 lt{1.42 * (u53(s0.w, s1.x) - 0.5),
                      1.42 * (u53(s1.y, s1.z) - 0.5), cs, sn};
    const double2 psi = site_position(tab.position(si), lt);
    const double2 beam =
        make_double2(psi.x + (2.0 * u53(s1.w, s2.x) - 1.0) * max_distance,
                     psi.y + (2.0 * u53(s2.y, s2.z) - 1.0) * max_distance);
    int nb[3];
    tab.neighbors(si, nb);
    for (int k = 0; k < 3; ++k) {
      const double2 p = site_position(tab.position(nb[k]), lt);
      const double a = rate_simple_ops(beam, psi, p);
      const double b = rate_simple_short(beam, p);
      const bool guard = cast_margin_ulps(b) < kCastGuardUlps || !(b > 1e-37);
      ++c_n;
      m_s = max(m_s, ulps(a, b));
      if (__double2float_rn(a) != __double2float_rn(b)) {
        ++c_su;
        if (!guard) ++c_s;
      }
      if (guard) ++c_g;
      const double pa = rate_prior_ops(beam, psi, p);
      const double pb = rate_prior_short(beam, psi, p);
      if (pa > 1e-300) m_p = max(m_p, ulps(pa, pb));
      if (__double2float_rn(pa) != __double2float_rn(pb)) ++c_p;
    }
  }
  atomicAdd(&out->samples, c_n);
  atomicAdd(&out->simple_cast_differs_unguarded, c_su);
  atomicAdd(&out->simple_cast_differs, c_s);
  atomicAdd(&out->simple_guard_taken, c_g);
  atomicAdd(&out->prior_cast_differs, c_p);
  atomicMax(&out->simple_max_ulps, m_s);
  atomicMax(&out->prior_max_ulps, m_p);
}

}  // namespace pd

extern "C" int pd_rate_ops_audit(const pd_lattice* lat, uint64_t seed,
                                 int64_t n_samples,
                                 double max_distance_angstroms,
                                 pd_rate_ops_stats* out, void* stream) {
  PD_REQUIRE(lat && lat->base_xy && lat->nbr && out, "null lattice / output");
  PD_REQUIRE(n_samples > 0, "nothing to audit");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  pd::RateOpsStats* d = nullptr;
  PD_CUDA_OK(cudaMalloc(&d, sizeof(pd::RateOpsStats)));
  PD_CUDA_OK(cudaMemsetAsync(d, 0, sizeof(pd::RateOpsStats), s));
  pd::k_rate_ops_audit<<<pd::sm_count() * 8, 256, 0, s>>>(
      *lat, seed, n_samples, max_distance_angstroms, d);
  pd::RateOpsStats h{};
  cudaError_t err = cudaGetLastError();
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, s);
  if (err == cudaSuccess) err = cudaStreamSynchronize(s);
  cudaFree(d);
  PD_CUDA_OK(err);
  out->evaluations = static_cast<int64_t>(h.samples);
  out->simple_cast_differs_unguarded =
      static_cast<int64_t>(h.simple_cast_differs_unguarded);
  out->simple_cast_differs = static_cast<int64_t>(h.simple_cast_differs);
  out->simple_guard_taken = static_cast<int64_t>(h.simple_guard_taken);
  out->prior_cast_differs = static_cast<int64_t>(h.prior_cast_differs);
  out->simple_max_ulps = h.simple_max_ulps;
  out->prior_max_ulps = h.prior_max_ulps;
  out->guard_ulps = pd::kCastGuardUlps;
  return PD_OK;
}

extern "C" int pd_fast_path_audit(const pd_lattice* lat, int32_t rate_fn,
                                  uint64_t seed, int64_t n_samples,
                                  int64_t dwell_us,
                                  double max_distance_angstroms,
                                  pd_fast_audit* out, void* stream) {
  PD_REQUIRE(lat && lat->base_xy && lat->nbr && out, "null lattice / output");
  PD_REQUIRE(rate_fn == PD_RATE_PRIOR || rate_fn == PD_RATE_SIMPLE,
             "the fast path covers the prior and simple rates");
  PD_REQUIRE(n_samples > 0 && dwell_us > 0, "nothing to audit");
  static_assert(sizeof(pd::AuditStats) == 8 * 8 + 4 * 4, "layout");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  pd::AuditStats* d = nullptr;
  PD_CUDA_OK(cudaMalloc(&d, sizeof(pd::AuditStats)));
  PD_CUDA_OK(cudaMemsetAsync(d, 0, sizeof(pd::AuditStats), s));
  const int grid = pd::sm_count() * 8;
  if (rate_fn == PD_RATE_PRIOR)
    pd::k_fast_audit<PD_RATE_PRIOR><<<grid, 256, 0, s>>>(
        *lat, seed, n_samples, dwell_us, max_distance_angstroms, d);
  else
    pd::k_fast_audit<PD_RATE_SIMPLE><<<grid, 256, 0, s>>>(
        *lat, seed, n_samples, dwell_us, max_distance_angstroms, d);
  pd::k_draw_audit<<<grid, 256, 0, s>>>(&d->draw_abs);
  pd::AuditStats h{};
  cudaError_t err = cudaGetLastError();
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, s);
  if (err == cudaSuccess) err = cudaStreamSynchronize(s);
  cudaFree(d);
  PD_CUDA_OK(err);
  out->samples = static_cast<int64_t>(h.samples);
  out->no_hop = static_cast<int64_t>(h.no_hop);
  out->hop = static_cast<int64_t>(h.hop);
  out->unsure = static_cast<int64_t>(h.unsure);
  out->wrong_decision = static_cast<int64_t>(h.wrong_decision);
  out->wrong_slot = static_cast<int64_t>(h.wrong_slot);
  out->waiting_time_outside_bounds = static_cast<int64_t>(h.t_outside);
  auto f = [](unsigned int bits) {
    float v;
    memcpy(&v, &bits, sizeof(v));
    return static_cast<double>(v);
  };
  out->total_rate_error_over_bound = f(h.tot_ratio);
  out->waiting_time_error_over_bound = f(h.t_ratio);
  out->choice_error_over_bound = f(h.choice_ratio);
  out->draw_error_over_bound = f(h.draw_abs);
  return PD_OK;
}
