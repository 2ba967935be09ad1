// K1, fast form: rollouts on the prior / simple rates with the guarded
// float32 iteration of pd_fast.cuh.  Every decision an iteration takes (hop or
// not, which neighbour) is made in float32 when its error bound allows it and
// by replaying the control with the exact float64 code otherwise, so the
// results are those of k_rollout / k_walk (pd_step.cu) bit for bit
// (tests/test_gpu_fast.py) at ~1/4 of the instructions.
//
//   k_rollout_fast   small batches: 16 lanes per env look ahead over the steps
//   k_walk_fast      large batches: a lane per env, one iteration per trip;
//                    <LIST>: the envs (and first steps) k_walk_plan hands over
//   k_walk_plan      large batches under the relative adapter: every (env,
//                    step, bulk site class) evaluated densely first, then a
//                    walk through tables (the default within an episode's
//                    length of a reset, see launch_fast)
//   k_rollout_plan   the same plan for small batches (opt-in: slower there)
//
//   graphene.py:646-694   PristineSingleDopedGraphene.apply_control
//   simulator.py:107-182  PuttingDuneSimulator.step_and_image
//   action_adapters.py:163-188 RelativeToSiliconActionAdapter.get_action
//
// Compiled with -fmad=false (the exact replay shares this translation unit).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

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
// LIST: the envs and their first steps come from a.defer_list (the envs that
// k_walk_plan handed over); otherwise every env of the batch from step 0.
template <int RATE, int IO, bool REL, bool LIST = false>
__global__ void __launch_bounds__(kStepThreads, PD_FAST_MIN_BLOCKS)
    k_walk_fast(const __grid_constant__ StepArgs a) {
  // The tables are read through L1: a hop touches one 16-byte row, in one
  // iteration of ten, and staging them would take the shared memory that the
  // action stream's L1 lines need.
  const GlobalTables tab{reinterpret_cast<const double2*>(a.lat.base_xy),
                         reinterpret_cast<const int4*>(a.lat.nbr)};
  const FastTimes tm = fast_times(a);
  const float md_f = static_cast<float>(a.max_distance);
  // action -> beam offset in the units fast_event expects
  const float md_s = static_cast<float>(
      a.max_distance * (RATE == PD_RATE_PRIOR ? 1.0 / kBond : 1.0));
  const long long step_us = a.dwell_us_scalar + a.image_duration_us;
  const int64_t n = a.st.n_envs;
  const int n_steps = a.n_steps;
  const ActionStream<IO> ctl(a);
  const int lane = threadIdx.x & 31;
  const int64_t n_items = LIST ? static_cast<int64_t>(*a.defer_count) : n;
  // (the host reads the list's length before its next launch on this batch:
  // a mapped host word, written once)
  if (LIST && a.list_hint && blockIdx.x == 0 && threadIdx.x == 0)
    *a.list_hint = static_cast<uint32_t>(n_items);
  const int64_t n_batches = (n_items + 31) / 32;
  const int64_t warps_total =
      static_cast<int64_t>(gridDim.x) * (kStepThreads / 32);
  const int64_t wid = static_cast<int64_t>(blockIdx.x) * (kStepThreads / 32) +
                      (threadIdx.x >> 5);

  for (int64_t b = wid; b < n_batches; b += warps_total) {
    bool active = b * 32 + lane < n_items;
    int t_first = 0;
    int64_t env = b * 32 + lane;
    if (LIST) {
      const int2 item = active ? a.defer_list[b * 32 + lane] : make_int2(0, 0);
      env = item.x;
      t_first = item.y;
    }
    // ---- the env's registers ----
    FastSite s;
    s.si = 0;
    s.nb[0] = s.nb[1] = s.nb[2] = 0;
    s.cls = 2;
#pragma unroll
    for (int i = 0; i < 3; ++i) s.geo.gx[i] = s.geo.gy[i] = 0.f;
    uint32_t env_id = 0, ctrl_count = 0;
    int events = 0, transitions = 0, recentres = 0;
    uint8_t status = 0, status_in = 0;
    double2 act_next = make_double2(0.0, 0.0);
    int t = t_first;
    // t * n + env: this step's element of every [T][n] array
    int64_t row = static_cast<int64_t>(t_first) * n + env;
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
        const float ax = clip_nanf(static_cast<float>(act.x), -1.f, 1.f);
        const float ay = clip_nanf(static_cast<float>(act.y), -1.f, 1.f);
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
      if (t + 1 < n_steps) act_next = ctl.load(row + n);
      // a step takes a few hundred cycles, DRAM a thousand: the action
      // stream is requested several steps ahead (L2 now, L1 two steps ahead
      // in the loop), the next batch's state a whole batch ahead
#pragma unroll 1
      for (int k = 2; t + k < n_steps && k < 2 + kActionsAhead; ++k)
        prefetch_l2(ctl.at(row + k * n));
      if (!LIST && env + warps_total * 32 < n) {
        prefetch_env(a, env + warps_total * 32);
        prefetch_l1(ctl.at(row + warps_total * 32));
        if (n_steps > 1) prefetch_l1(ctl.at(row + n + warps_total * 32));
      }
      prefetch_l1(a.st.fov_scale + env);
      const Lattice4 lat = load_lattice4(a.st.lattice, env);
      const Fov4 fov = load_fov4(a.st.fov, env);
      env_id = a.st.env_offset + static_cast<uint32_t>(env);
      ctrl_count = a.st.ctrl_count[env];
      status = status_in = a.st.status[env];
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
            static_cast<long long>(n_steps - t_first) * step_us +
            static_cast<long long>(recentres) * a.image_duration_us;
        atomicAdd(reinterpret_cast<unsigned long long*>(a.st.sim_time_us + env),
                  static_cast<unsigned long long>(total));
        a.st.si_idx[env] = s.si;
        a.st.ctrl_count[env] = ctrl_count;
        atomicAdd(reinterpret_cast<unsigned long long*>(a.st.n_events + env),
                  static_cast<unsigned long long>(events));
        // (one-step launches: 9 envs in 10 do not hop, and the status byte
        // changes for a handful of envs per 10^9 -- skip those lines)
        if (transitions > 0)
          atomicAdd(
              reinterpret_cast<unsigned long long*>(a.st.n_transitions + env),
              static_cast<unsigned long long>(transitions));
        if (status != status_in) a.st.status[env] = status;
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
  const FastTimes tm = fast_times(a);
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
            const float ax = clip_nanf(static_cast<float>(c.x), -1.f, 1.f);
            const float ay = clip_nanf(static_cast<float>(c.y), -1.f, 1.f);
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
// k_rollout_plan: small-batch rollouts under the relative adapter, with the
// time dimension taken out of the dependent chain.
//
// Under the relative adapter the beam offset of a control is clip(action, -1,
// 1) * max_distance whatever the state (while the adapter's clip to the frame
// cannot engage), the Philox draws of a control depend on (env, control
// counter, iteration) only, and the neighbour geometry of a bulk site depends
// on the env's lattice angle and the site's class (0 / 1) only.  So what a
// control does to an env on a bulk site -- how many hops, to which neighbour
// slots -- is a function of (env, step, class) alone.  A CTA owns up to
// kPlanEnvs envs and walks the call in chunks of kPlanChunk steps:
//   plan    every (env, step) of the chunk at full density, all threads, no
//           dependence between items: iteration 0 of the control for both
//           classes with fast_event (one Philox call serves both); the ~11 %
//           that do not end there go to a queue and are run to their end by
//           the next free lane.  One uint16 per class in shared memory:
//           hops | UNSURE << 3 | slots << 4;
//   commit  one warp per env walks the plan: ballots give the non-quiet
//           steps of a 32-step window for either class, the warp jumps from
//           one to the next, follows the slots through the neighbour table,
//           keeps the float32 view of the Si in the FOV, and hands everything
//           the plan does not cover (UNSURE controls, sheet-edge sites, a clip
//           that may engage, the safe-area test when float32 cannot rule a
//           re-centre out) to the exact float64 code;
//   store   all threads write the chunk's per-step results, env-contiguous.
// Results are those of k_rollout_fast / the float64 kernels bit for bit
// (tests/test_gpu_fast.py runs every case through all three).
// ---------------------------------------------------------------------------
constexpr int kPlanThreads = 512;
constexpr int kPlanEnvs = kPlanThreads / 32;  // one warp per env in `commit`
constexpr int kPlanChunk = 256;               // steps per chunk
constexpr int kPlanMaxHops = 5;               // hops one plan entry describes
constexpr int kPlanQueue = 2048;              // queued (env, step, class)
constexpr uint32_t kPlanUnsure = 8u;

// The control of (env_id, ctrl) for a Si on a bulk site with geometry g and
// beam offset (bx, by), run to its end in float32.
template <int RATE>
__device__ __forceinline__ uint32_t plan_control(FastGeo g, float bx, float by,
                                                 uint32_t env_id, uint32_t ctrl,
                                                 const FastTimes& tm,
                                                 const PhiloxKeys& keys) {
  const float off_s = fast_offset_scale<RATE>();
  float e_lo = 0.f, e_hi = 0.f;
  uint32_t hops = 0, slots = 0;
  for (uint32_t it = 0;; ++it) {
    const uint4 w = philox4x32_10k(env_id, ctrl, it, PD_STREAM_KMC, keys);
    int slot = 0;
    float t_lo, t_hi;
    const int kind = fast_event<RATE>(g, bx, by, w.x, w.z, e_lo, e_hi, tm,
                                      &slot, &t_lo, &t_hi);
    if (kind == FAST_NO_HOP) break;
    if (kind == FAST_UNSURE || hops == kPlanMaxHops) return kPlanUnsure;
    float ox, oy;
    bulk_offset<RATE>(g, slot, &ox, &oy);
    bx -= ox * off_s;
    by -= oy * off_s;
    flip_geo(&g);
    slots |= static_cast<uint32_t>(slot) << (2 * hops);
    ++hops;
    fast_advance(&e_lo, &e_hi, t_lo, t_hi);
  }
  return hops | (slots << 4);
}

// Rare paths of `commit`, out of line.  A Si on a sheet-edge site (class 2:
// its neighbour geometry is the site's own, pd_lattice.cu) is not covered by
// the plan; its controls still run in float32:
//   site_busy_mask  iteration 0 of the steps [t_first, t_first + len) for a
//                   Si parked on `si`, one step per lane: bit j = step
//                   t_first + j does not certainly end without a hop;
//   serial_control  one control from its start, iteration by iteration
//                   (what a lane of k_walk_fast does).
template <int RATE, int IO, class Tables>
__device__ __noinline__ unsigned site_busy_mask(const StepArgs& a,
                                                const Tables tab,
                                                const double2 cs, int64_t env,
                                                int t_first, int len, int si,
                                                uint32_t ctrl_first) {
  const FastSite s = fast_site<RATE>(tab, si, cs.x, cs.y);
  const int lane = threadIdx.x & 31;
  bool busy = false;
  if (lane < len) {
    const FastTimes tm = fast_times(a);
    const float md_s = static_cast<float>(
        a.max_distance * (RATE == PD_RATE_PRIOR ? 1.0 / kBond : 1.0));
    const double2 act = ActionStream<IO>(a).load(t_first + lane, env);
    const float bx = clip_nanf(static_cast<float>(act.x), -1.f, 1.f) * md_s;
    const float by = clip_nanf(static_cast<float>(act.y), -1.f, 1.f) * md_s;
    const uint4 w = philox4x32_10k(
        a.st.env_offset + static_cast<uint32_t>(env),
        ctrl_first + static_cast<uint32_t>(lane), 0u, PD_STREAM_KMC, a.keys);
    int slot;
    float t_lo, t_hi;
    busy = fast_event<RATE>(s.geo, bx, by, w.x, w.z, 0.f, 0.f, tm, &slot, &t_lo,
                            &t_hi) != FAST_NO_HOP;
  }
  return __ballot_sync(0xffffffffu, busy);
}

struct SerialResult {
  int si, nb[3], cls;  // the Si after the control
  float ox, oy;        // how far it moved, angstrom
  int hops;
  bool unsure;         // float32 could not settle an iteration: replay
};

template <int RATE, int IO, class Tables>
__device__ __noinline__ void serial_control(const StepArgs& a, const Tables tab,
                                            const double2 cs, int64_t env,
                                            int t, int si, uint32_t ctrl,
                                            SerialResult* out) {
  auto rotation = [&]() { return cs; };
  FastSite s = fast_site<RATE>(tab, si, cs.x, cs.y);
  const FastTimes tm = fast_times(a);
  const float md_s = static_cast<float>(
      a.max_distance * (RATE == PD_RATE_PRIOR ? 1.0 / kBond : 1.0));
  const double2 act = ActionStream<IO>(a).load(t, env);
  float bx = clip_nanf(static_cast<float>(act.x), -1.f, 1.f) * md_s;
  float by = clip_nanf(static_cast<float>(act.y), -1.f, 1.f) * md_s;
  const uint32_t env_id = a.st.env_offset + static_cast<uint32_t>(env);
  float e_lo = 0.f, e_hi = 0.f, oxs = 0.f, oys = 0.f;
  int hops = 0;
  out->unsure = false;
  for (uint32_t it = 0;; ++it) {
    const uint4 w = philox4x32_10k(env_id, ctrl, it, PD_STREAM_KMC, a.keys);
    int slot = 0;
    float t_lo, t_hi;
    const int kind = fast_event<RATE>(s.geo, bx, by, w.x, w.z, e_lo, e_hi, tm,
                                      &slot, &t_lo, &t_hi);
    if (kind == FAST_NO_HOP) break;
    if (kind == FAST_UNSURE) {
      out->unsure = true;
      return;
    }
    float ox, oy;
    fast_hop<RATE>(tab, slot, rotation, &s, &bx, &by, &ox, &oy);
    oxs += ox;
    oys += oy;
    ++hops;
    fast_advance(&e_lo, &e_hi, t_lo, t_hi);
  }
  out->si = s.si;
  out->nb[0] = s.nb[0];
  out->nb[1] = s.nb[1];
  out->nb[2] = s.nb[2];
  out->cls = s.cls;
  out->ox = oxs;
  out->oy = oys;
  out->hops = hops;
}

// State of an env between the phases of a chunk (shared memory).
struct WalkState {
  int si;
  float qx, qy, iwx, iwy;  // Observed
  int events, transitions, recentres, hops_synced;
  int fov_site;    // >= 0: the FOV in memory is stale, the current one is the
                   // FOV centred on this site (simulator.py:161-165)
  int pos;         // next step of the chunk to commit
  int status;
  int check_area;  // simulator.py:156 at the first step of the call
};

// FOV centred on `site` (what exact_area_check stores when it re-centres).
__device__ __noinline__ void store_centred_fov(const StepArgs& a, int64_t env,
                                               int site) {
  const double2 base =
      __ldg(reinterpret_cast<const double2*>(a.lat.base_xy) + site);
  const double2 psi = site_position(base, load_lattice4(a.st.lattice, env));
  store_fov4(a.st.fov, env, centred_fov(psi, a.st.fov_scale[env]));
}

__device__ __forceinline__ bool q_inside(float qx, float qy) {
  return qx > 0.2501f && qx < 0.7499f && qy > 0.2501f && qy < 0.7499f;
}
// certainly outside the safe area (simulator.py:236-249): a re-centre
__device__ __forceinline__ bool q_outside(float qx, float qy) {
  return qx < 0.2499f || qx > 0.7501f || qy < 0.2499f || qy > 0.7501f;
}
__device__ __forceinline__ bool q_clip_free(float qx, float qy, float iwx,
                                            float iwy, float max_distance) {
  const float rx = __fmaf_rn(max_distance, iwx, 1e-4f);
  const float ry = __fmaf_rn(max_distance, iwy, 1e-4f);
  return qx > rx && qx < 1.0f - rx && qy > ry && qy < 1.0f - ry;
}

#ifdef PD_PLAN_CLOCKS
// phase clocks of k_rollout_plan (profiles/prof_walk.py PLAN_CLOCKS=1)
__device__ unsigned long long g_plan_clocks[1024 * 8];
#define PLAN_CLOCK(i)                                              \
  do {                                                             \
    if (threadIdx.x == 0 && blockIdx.x < 1024) {                   \
      unsigned long long t_;                                       \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));       \
      g_plan_clocks[blockIdx.x * 8 + (i)] = t_;                    \
    }                                                              \
  } while (0)
// event counts of the warp walker: 16 bits each of replays, area checks,
// serial controls, busy masks
#define PLAN_COUNT(shift)                                                   \
  do {                                                                      \
    if ((threadIdx.x & 31) == 0 && blockIdx.x < 1024)                       \
      atomicAdd(&g_plan_clocks[blockIdx.x * 8 + 4], 1ull << (shift));       \
  } while (0)
#else
#define PLAN_CLOCK(i)
#define PLAN_COUNT(shift)
#endif

constexpr uint32_t kOutMark = 1u << 30;  // tile entry: a step's result
constexpr uint32_t kOutRec = 1u << 31;   //   ... which re-centred the FOV

template <int RATE, int IO>
__global__ void __launch_bounds__(kPlanThreads, 2)
    k_rollout_plan(const __grid_constant__ StepArgs a) {
  extern __shared__ __align__(16) unsigned char plan_smem[];
  // plan records (uint16 per class: hops | UNSURE << 3 | slots << 4, bits 14
  // and 15 unused), overwritten step by step with the results
  // (site | kOutMark | kOutRec)
  __shared__ uint32_t tile[kPlanEnvs][kPlanChunk + 1];
  __shared__ uint32_t queue[kPlanQueue];
  __shared__ uint32_t s_mask[kPlanEnvs][2][kPlanChunk / 32];  // busy steps
  __shared__ float s_geo[kPlanEnvs][6];  // class-0 geometry (fast_event)
  __shared__ float s_off[kPlanEnvs][6];  // class-0 neighbour offsets, angstrom
  __shared__ float s_iws[kPlanEnvs];     // 1 / fov_scale
  __shared__ double2 s_rot[kPlanEnvs];   // lattice rotation (cos, sin)
  __shared__ uint32_t s_ctrl0[kPlanEnvs];
  __shared__ int s_si_start[kPlanEnvs];
  __shared__ WalkState s_ws[kPlanEnvs];
  __shared__ uint32_t q_count;
  const FastTimes tm = fast_times(a);
  const float md_f = static_cast<float>(a.max_distance);
  const float md_s = static_cast<float>(
      a.max_distance * (RATE == PD_RATE_PRIOR ? 1.0 / kBond : 1.0));
  const long long step_us = a.dwell_us_scalar + a.image_duration_us;
  const int64_t n = a.st.n_envs;
  const int n_steps = a.n_steps;
  const ActionStream<IO> ctl(a);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int epc = a.plan_envs_per_cta;
  const int64_t env0 = static_cast<int64_t>(blockIdx.x) * epc;
  const int n_env = n - env0 < epc ? static_cast<int>(n - env0) : epc;
  const uint32_t env_id0 = a.st.env_offset + static_cast<uint32_t>(env0);
  PLAN_CLOCK(0);
#ifdef PD_PLAN_CLOCKS
  if (tid == 0 && blockIdx.x < 1024) g_plan_clocks[blockIdx.x * 8 + 4] = 0;
#endif

  // the neighbour table as 8-byte rows in shared memory (a planned hop is a
  // dependent lookup) and the per-env constants
  double2* s_base = reinterpret_cast<double2*>(plan_smem);
  ushort4* s_nbr = reinterpret_cast<ushort4*>(s_base + a.lat.n_sites);
  {
    const double2* gbase = reinterpret_cast<const double2*>(a.lat.base_xy);
    const int4* gnbr = reinterpret_cast<const int4*>(a.lat.nbr);
    for (int k = tid; k < a.lat.n_sites; k += kPlanThreads) {
      s_base[k] = __ldg(gbase + k);
      const int4 v = __ldg(gnbr + k);
      s_nbr[k] = make_ushort4(
          static_cast<unsigned short>(v.x), static_cast<unsigned short>(v.y),
          static_cast<unsigned short>(v.z),
          static_cast<unsigned short>((v.w >> kSiteClassShift) & 3));
    }
  }
  const SharedTables stab{s_base, s_nbr};
  if (tid < n_env) {
    const int64_t e = env0 + tid;
    const Lattice4 lat = load_lattice4(a.st.lattice, e);
    s_rot[tid] = make_double2(lat.c, lat.s);
    FastGeo g;
    fast_geo_bulk<RATE>(Lattice4{0.0, 0.0, lat.c, lat.s}, 0, &g);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      s_geo[tid][i] = g.gx[i];
      s_geo[tid][3 + i] = g.gy[i];
      bulk_offset<RATE>(g, i, &s_off[tid][i], &s_off[tid][3 + i]);
    }
    s_iws[tid] = __fdividef(1.0f, static_cast<float>(a.st.fov_scale[e]));
    s_ctrl0[tid] = a.st.ctrl_count[e];
    WalkState ws;
    ws.si = a.st.si_idx[e];
    Observed obs;
    obs.sync(load_fov4(a.st.fov, e),
             site_position(
                 __ldg(reinterpret_cast<const double2*>(a.lat.base_xy) + ws.si),
                 lat));
    ws.qx = obs.qx;
    ws.qy = obs.qy;
    ws.iwx = obs.iwx;
    ws.iwy = obs.iwy;
    ws.events = ws.transitions = ws.recentres = ws.hops_synced = 0;
    ws.fov_site = -1;
    ws.pos = 0;
    ws.status = a.st.status[e];
    ws.check_area = 1;
    s_ws[tid] = ws;
  }
  __syncthreads();

  for (int t0 = 0; t0 < n_steps; t0 += kPlanChunk) {
    const int len_c = n_steps - t0 < kPlanChunk ? n_steps - t0 : kPlanChunk;
    const int n_win = (len_c + 31) >> 5;
    PLAN_CLOCK(1);
    if (tid == 0) q_count = 0;
    if (tid < kPlanEnvs * 2 * (kPlanChunk / 32))
      (&s_mask[0][0][0])[tid] = 0u;
    __syncthreads();
    // ---- plan, dense pass: iteration 0 of every control, both classes ----
    const int items = n_env * len_c;
    for (int idx = tid; idx < items; idx += kPlanThreads) {
      const int k = idx / n_env, el = idx - k * n_env;
      const int t = t0 + k;
      const double2 act = ctl.load(static_cast<int64_t>(t) * n + env0 + el);
      const float bx =
          clip_nanf(static_cast<float>(act.x), -1.f, 1.f) * md_s;
      const float by =
          clip_nanf(static_cast<float>(act.y), -1.f, 1.f) * md_s;
      FastGeo g;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        g.gx[i] = s_geo[el][i];
        g.gy[i] = s_geo[el][3 + i];
      }
      const uint4 w = philox4x32_10k(env_id0 + static_cast<uint32_t>(el),
                                     s_ctrl0[el] + static_cast<uint32_t>(t), 0u,
                                     PD_STREAM_KMC, a.keys);
      bool quiet0, quiet1;
      fast_quiet_both<RATE>(g, bx, by, w.x, tm, &quiet0, &quiet1);
      const bool busy0 = !quiet0, busy1 = !quiet1;
      // (a queue entry that does not fit stays UNSURE: the exact code runs it)
      tile[el][k] =
          (busy0 ? kPlanUnsure : 0u) | (busy1 ? kPlanUnsure << 16 : 0u);
      const int want = (busy0 ? 1 : 0) + (busy1 ? 1 : 0);
      if (want) {
        if (busy0) atomicOr(&s_mask[el][0][k >> 5], 1u << (k & 31));
        if (busy1) atomicOr(&s_mask[el][1][k >> 5], 1u << (k & 31));
        uint32_t q = atomicAdd(&q_count, static_cast<uint32_t>(want));
        if (busy0 && q < kPlanQueue) queue[q] = static_cast<uint32_t>(idx) << 1;
        q += busy0 ? 1u : 0u;
        if (busy1 && q < kPlanQueue)
          queue[q] = (static_cast<uint32_t>(idx) << 1) | 1u;
      }
    }
    __syncthreads();
    PLAN_CLOCK(2);
    // ---- plan, queue pass: the controls that went on, to their end ----
    {
      const int n_q = q_count < kPlanQueue ? static_cast<int>(q_count)
                                           : kPlanQueue;
      for (int q = tid; q < n_q; q += kPlanThreads) {
        const uint32_t entry = queue[q];
        const int idx = static_cast<int>(entry >> 1), c = entry & 1u;
        const int k = idx / n_env, el = idx - k * n_env;
        const int t = t0 + k;
        const double2 act = ctl.load(static_cast<int64_t>(t) * n + env0 + el);
        const float bx =
            clip_nanf(static_cast<float>(act.x), -1.f, 1.f) * md_s;
        const float by =
            clip_nanf(static_cast<float>(act.y), -1.f, 1.f) * md_s;
        FastGeo g;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          g.gx[i] = s_geo[el][i];
          g.gy[i] = s_geo[el][3 + i];
        }
        if (c) flip_geo(&g);
        const uint32_t r16 = plan_control<RATE>(
            g, bx, by, env_id0 + static_cast<uint32_t>(el),
            s_ctrl0[el] + static_cast<uint32_t>(t), tm, a.keys);
        reinterpret_cast<unsigned short*>(&tile[el][k])[c] =
            static_cast<unsigned short>(r16);
      }
    }
    __syncthreads();
    PLAN_CLOCK(3);
    // ---- commit: a warp per env.  Lane 0 follows the plan from one busy
    // step to the next while the Si stays on bulk sites, re-centres included
    // (the FOV itself is formed when something needs it); it stops where the
    // plan ends: UNSURE, a sheet-edge site, a clip that may engage, a
    // safe-area test float32 cannot settle.  The whole warp then takes the
    // env through those steps (float32 look-ahead over the window on edge
    // sites, the exact code where float32 cannot speak) and lane 0 goes on.
    // ----
    if (tid < n_env) {
      s_si_start[tid] = s_ws[tid].si;
      s_ws[tid].pos = 0;
    }
    __syncthreads();
    if (warp < n_env) for (;;) {
      if (lane == 0 && s_ws[warp].pos < len_c) {
        WalkState ws = s_ws[warp];
        int si = ws.si;
        ushort4 row = s_nbr[si];
        int cls = row.w;
        float qx = ws.qx, qy = ws.qy;
        int stop_at = len_c, hops = 0;
        bool ok = cls < 2 && q_clip_free(qx, qy, ws.iwx, ws.iwy, md_f);
        if (ws.check_area) {
          if (q_inside(qx, qy)) ws.check_area = 0;
          else ok = false;
        }
        if (!ok) stop_at = ws.pos;
        int w = ws.pos >> 5;
        unsigned from = ~0u << (ws.pos & 31);
        while (ok) {
          unsigned m = 0;
          while (w < n_win) {
            m = s_mask[warp][cls][w] & from;
            if (m) break;
            ++w;
            from = ~0u;
          }
          if (!m) break;  // the chunk is done
          const int b = __ffs(m) - 1;
          const int k = w * 32 + b;
          from = b == 31 ? 0u : (~0u << (b + 1));
          const uint32_t r16 =
              reinterpret_cast<const unsigned short*>(&tile[warp][k])[cls];
          if (r16 & kPlanUnsure) {
            stop_at = k;
            break;
          }
          const int n_hops = static_cast<int>(r16 & 7u);
          int si_t = si, cls_t = cls;
          ushort4 row_t = row;
          float qx_t = qx, qy_t = qy;
          for (int h = 0; h < n_hops; ++h) {
            const int slot = static_cast<int>((r16 >> (4 + 2 * h)) & 3u);
            // class 1: g1[i] = -g0[2 - i]
            const int j = cls_t == 0 ? slot : 2 - slot;
            const float sg = cls_t == 0 ? 1.f : -1.f;
            qx_t = __fmaf_rn(sg * s_off[warp][j], ws.iwx, qx_t);
            qy_t = __fmaf_rn(sg * s_off[warp][3 + j], ws.iwy, qy_t);
            si_t = slot == 0 ? row_t.x : (slot == 1 ? row_t.y : row_t.z);
            row_t = s_nbr[si_t];
            cls_t = row_t.w;
            if (cls_t == 2) break;  // onto the sheet's edge: the plan's next
                                    // iteration assumed a bulk site's geometry
          }
          const bool in = q_inside(qx_t, qy_t), out = q_outside(qx_t, qy_t);
          const bool due = ws.transitions + hops + n_hops - ws.hops_synced >=
                           kObservedSyncHops;
          if (cls_t == 2 || (!in && !out) || (due && !out)) {
            stop_at = k;
            break;
          }
          si = si_t;
          row = row_t;
          cls = cls_t;
          qx = qx_t;
          qy = qy_t;
          hops += n_hops;
          if (out) {
            // simulator.py:156-169: the FOV is centred on the Si again
            qx = qy = 0.5f;
            ws.iwx = ws.iwy = s_iws[warp];
            ws.recentres += 1;
            ws.fov_site = si;
            ws.hops_synced = ws.transitions + hops;
          }
          tile[warp][k] =
              static_cast<uint32_t>(si) | kOutMark | (out ? kOutRec : 0u);
          if (!q_clip_free(qx, qy, ws.iwx, ws.iwy, md_f)) {
            stop_at = k + 1;
            break;
          }
        }
        const int reached = stop_at < len_c ? stop_at : len_c;
        ws.si = si;
        ws.qx = qx;
        ws.qy = qy;
        ws.transitions += hops;
        ws.events += (reached - ws.pos) + hops;
        ws.pos = reached;
        s_ws[warp] = ws;
      }
      __syncwarp();
      if (s_ws[warp].pos >= len_c) break;
      {
        const int64_t e = env0 + warp;
        WalkState ws = s_ws[warp];
        const double2 cs = s_rot[warp];
        __syncwarp();
        if (ws.fov_site >= 0) {
          store_centred_fov(a, e, ws.fov_site);
          ws.fov_site = -1;
        }
        int si = ws.si;
        int nb0, nb1, nb2, cls;
        {
          const ushort4 row = s_nbr[si];
          nb0 = row.x;
          nb1 = row.y;
          nb2 = row.z;
          cls = row.w;
        }
        Observed obs{ws.qx, ws.qy, ws.iwx, ws.iwy};
        uint32_t ctrl_count =
            s_ctrl0[warp] + static_cast<uint32_t>(t0 + ws.pos);
        uint8_t status = static_cast<uint8_t>(ws.status);
        int events = ws.events, transitions = ws.transitions,
            recentres = ws.recentres, hops_synced = ws.hops_synced;
        bool check_area = ws.check_area != 0;
        int next = len_c;     // where the lane walker takes over again
        bool handed = false;
        for (int w0 = ws.pos & ~31; w0 < len_c && !handed; w0 += 32) {
          const int len = len_c - w0 < 32 ? len_c - w0 : 32;
          int pos = w0 < ws.pos ? ws.pos - w0 : 0;
          const uint32_t rec =
              lane >= pos && lane < len ? tile[warp][w0 + lane] : 0u;
          bool edge_ok = false;  // m_e describes the Si's current (edge) site
          unsigned m_e = 0;
          while (pos < len) {
            // the plan speaks for a Si on a bulk site, the look-ahead below
            // for one on a sheet-edge site, both while the adapter's clip
            // cannot engage; anything else goes step by step through the
            // exact code
            const bool clip_free = obs.clip_free(md_f);
            const bool planned = cls < 2 && clip_free;
            const bool edge = cls == 2 && clip_free;
            if (planned && !check_area && w0 + pos > ws.pos) {
              // back on the plan: the lane walker goes on from here
              next = w0 + pos;
              handed = true;
              break;
            }
            if (edge && !edge_ok && !check_area) {
              PLAN_COUNT(48);
              m_e = site_busy_mask<RATE, IO>(
                  a, stab, cs, e, t0 + w0, len, si,
                  s_ctrl0[warp] + static_cast<uint32_t>(t0 + w0));
              edge_ok = true;
            }
            int stop;
            if (!clip_free || check_area || planned) {
              stop = pos;
            } else {
              const unsigned m = m_e & (~0u << pos);
              stop = m ? __ffs(m) - 1 : len;
            }
            // steps [pos, stop) end without a hop
            events += stop - pos;
            ctrl_count += static_cast<uint32_t>(stop - pos);
            if (stop >= len) break;
            bool hopped = false;
            bool exact = !clip_free;
            bool serial = edge;
            if (planned) {
              const uint32_t r32 = __shfl_sync(0xffffffffu, rec, stop);
              const uint32_t r16 = (cls == 1 ? r32 >> 16 : r32) & 0xFFFFu;
              exact = (r16 & kPlanUnsure) != 0u;
              if (!exact) {
                // follow the planned hops through the neighbour table
                const int n_hops = static_cast<int>(r16 & 7u);
                int si_t = si, n0 = nb0, n1 = nb1, n2 = nb2, cls_t = cls;
                float qx = obs.qx, qy = obs.qy;
                for (int h = 0; h < n_hops; ++h) {
                  const int slot =
                      static_cast<int>((r16 >> (4 + 2 * h)) & 3u);
                  const int j = cls_t == 0 ? slot : 2 - slot;
                  const float sg = cls_t == 0 ? 1.f : -1.f;
                  qx = __fmaf_rn(sg * s_off[warp][j], obs.iwx, qx);
                  qy = __fmaf_rn(sg * s_off[warp][3 + j], obs.iwy, qy);
                  si_t = slot == 0 ? n0 : (slot == 1 ? n1 : n2);
                  const ushort4 row = s_nbr[si_t];
                  n0 = row.x;
                  n1 = row.y;
                  n2 = row.z;
                  cls_t = row.w;
                  if (cls_t == 2) {
                    // onto the sheet's edge: the plan's next iteration
                    // assumed a bulk site's geometry
                    serial = true;
                    break;
                  }
                }
                if (!serial) {
                  si = si_t;
                  nb0 = n0;
                  nb1 = n1;
                  nb2 = n2;
                  cls = cls_t;
                  obs.qx = qx;
                  obs.qy = qy;
                  hopped = n_hops > 0;
                  transitions += n_hops;
                  events += n_hops + 1;
                  ctrl_count += 1;
                }
              }
            }
            if (serial) {
              SerialResult sr;
              PLAN_COUNT(32);
              serial_control<RATE, IO>(a, stab, cs, e, t0 + w0 + stop, si,
                                       ctrl_count, &sr);
              if (sr.unsure) {
                exact = true;
              } else {
                hopped = sr.hops > 0;
                if (hopped) {
                  si = sr.si;
                  nb0 = sr.nb[0];
                  nb1 = sr.nb[1];
                  nb2 = sr.nb[2];
                  cls = sr.cls;
                  obs.hop(sr.ox, sr.oy);
                }
                transitions += sr.hops;
                events += sr.hops + 1;
                ctrl_count += 1;
              }
            }
            if (exact) {
              ReplayResult rr;
              PLAN_COUNT(0);
              replay_control<RATE, IO>(a, e, t0 + w0 + stop, si, 0u,
                                       ctrl_count, transitions, events, status,
                                       load_fov4(a.st.fov, e), &rr);
              hopped = rr.hopped;
              if (hopped) {
                si = rr.s.si;
                nb0 = rr.s.nb[0];
                nb1 = rr.s.nb[1];
                nb2 = rr.s.nb[2];
                cls = rr.s.cls;
                obs = rr.obs;
              }
              ctrl_count = rr.ctrl_count;
              transitions = rr.transitions;
              events = rr.events;
              status = rr.status;
            }
            if (hopped) edge_ok = false;
            // image, safe area (simulator.py:152-169)
            bool recentred = false;
            if (hopped || check_area) {
              check_area = false;
              if (!obs.inside() ||
                  transitions - hops_synced >= kObservedSyncHops) {
                PLAN_COUNT(16);
                recentred = exact_area_check(a, e, si, &obs);
                recentres += recentred ? 1 : 0;
                hops_synced = transitions;
              }
            }
            if (lane == stop)
              tile[warp][w0 + stop] = static_cast<uint32_t>(si) | kOutMark |
                                      (recentred ? kOutRec : 0u);
            pos = stop + 1;
          }
        }
        if (lane == 0) {
          ws.si = si;
          ws.qx = obs.qx;
          ws.qy = obs.qy;
          ws.iwx = obs.iwx;
          ws.iwy = obs.iwy;
          ws.events = events;
          ws.transitions = transitions;
          ws.recentres = recentres;
          ws.hops_synced = hops_synced;
          ws.status = status;
          ws.check_area = check_area ? 1 : 0;
          ws.pos = next;
          s_ws[warp] = ws;
        }
      }
      __syncwarp();
      if (s_ws[warp].pos >= len_c) break;
    }
    __syncthreads();
    PLAN_CLOCK(5);
    // ---- fill: every step's result = that of the last step before it that
    // changed something ----
    if (warp < n_env) {
      uint32_t carry = static_cast<uint32_t>(s_si_start[warp]);
      for (int w0 = 0; w0 < len_c; w0 += 32) {
        const uint32_t v = w0 + lane < len_c ? tile[warp][w0 + lane] : 0u;
        const unsigned marks = __ballot_sync(0xffffffffu, (v & kOutMark) != 0u);
        const unsigned upto = marks & (0xffffffffu >> (31 - lane));
        const int src = upto ? 31 - __clz(upto) : 0;
        const uint32_t got = __shfl_sync(0xffffffffu, v, src);
        const uint32_t out = upto ? ((got & (kOutMark - 1u)) |
                                     (src == lane ? (v & kOutRec) : 0u))
                                  : carry;
        if (w0 + lane < len_c) tile[warp][w0 + lane] = out;
        carry = __shfl_sync(0xffffffffu, out, 31) & (kOutMark - 1u);
      }
    }
    __syncthreads();
    PLAN_CLOCK(6);
    // ---- store: the chunk's results, env-contiguous ----
    for (int idx = tid; idx < items; idx += kPlanThreads) {
      const int k = idx / n_env, el = idx - k * n_env;
      const uint32_t v = tile[el][k];
      store_step<IO>(a, static_cast<int64_t>(t0 + k) * n + env0 + el,
                     static_cast<int>(v & (kOutMark - 1u)),
                     (v & kOutRec) != 0u, step_us);
    }
    // (the next chunk's first barrier orders these reads before its writes)
  }
  PLAN_CLOCK(7);
  if (tid < n_env) {
    const int64_t e = env0 + tid;
    const WalkState ws = s_ws[tid];
    if (ws.fov_site >= 0) store_centred_fov(a, e, ws.fov_site);
    const long long total =
        static_cast<long long>(n_steps) * step_us +
        static_cast<long long>(ws.recentres) * a.image_duration_us;
    atomicAdd(reinterpret_cast<unsigned long long*>(a.st.sim_time_us + e),
              static_cast<unsigned long long>(total));
    a.st.si_idx[e] = ws.si;
    a.st.ctrl_count[e] = s_ctrl0[tid] + static_cast<uint32_t>(n_steps);
    atomicAdd(reinterpret_cast<unsigned long long*>(a.st.n_events + e),
              static_cast<unsigned long long>(ws.events));
    atomicAdd(reinterpret_cast<unsigned long long*>(a.st.n_transitions + e),
              static_cast<unsigned long long>(ws.transitions));
    a.st.status[e] = static_cast<uint8_t>(ws.status);
  }
}

// ---------------------------------------------------------------------------
// k_walk_plan: the same plan for large batches (n_envs >= a few waves of
// lanes), where there are enough envs to give each its own lane and a launch
// covers few steps.  A CTA of 256 threads owns 256 envs and works through the
// call in chunks of 16 steps; thread = env in every phase, so the env's
// geometry, counters and busy masks stay in registers:
//   plan    for k in the chunk: action (coalesced over the envs), Philox,
//           iteration 0 for both bulk classes -- uniform over the lanes, no
//           bookkeeping between the evaluations -- one busy bit per class and
//           step in a register; the controls that go on (~11 % per class) are
//           queued and run to their end by the next free thread
//           (plan_control), one uint16 per class in shared memory;
//   commit  the thread jumps from one busy step of its env's current class to
//           the next and follows the planned slots through the neighbour
//           table; a certain re-centre is recorded as "FOV centred on site s";
//   store   the thread writes its env's results step by step (coalesced over
//           the envs).
// An env whose next control the plan does not cover (UNSURE, a sheet-edge
// site, a clip that may engage, a safe-area test within 1e-4 of its
// threshold) is written back as it stands and handed, with the step it
// stopped at, to k_walk_fast<LIST>, which runs such envs 32 to a warp: what
// is a serial tail in k_rollout_plan is dense work here.
// ---------------------------------------------------------------------------
constexpr int kWalkPlanThreads = 256;
constexpr int kWalkPlanChunk = 16;
constexpr int kWalkPlanQueue = 4096;  // (simple rate, 5 s: ~2300 per chunk)

#ifndef PD_WALK_PLAN_BLOCKS  // resident CTAs per SM (registers: 65536 / 256 / this)
#define PD_WALK_PLAN_BLOCKS 4
#endif

template <int RATE, int IO>
__global__ void __launch_bounds__(kWalkPlanThreads, PD_WALK_PLAN_BLOCKS)
    k_walk_plan(const __grid_constant__ StepArgs a) {
  __shared__ uint32_t tile[kWalkPlanChunk][kWalkPlanThreads];
  __shared__ uint32_t queue[kWalkPlanQueue];
  __shared__ float s_geo[6][kWalkPlanThreads];
  __shared__ uint32_t s_ctrl0[kWalkPlanThreads];
  __shared__ uint32_t q_count;
  const GlobalTables tab{reinterpret_cast<const double2*>(a.lat.base_xy),
                         reinterpret_cast<const int4*>(a.lat.nbr)};
  const FastTimes tm = fast_times(a);
  const float md_f = static_cast<float>(a.max_distance);
  const float md_s = static_cast<float>(
      a.max_distance * (RATE == PD_RATE_PRIOR ? 1.0 / kBond : 1.0));
  const long long step_us = a.dwell_us_scalar + a.image_duration_us;
  const int64_t n = a.st.n_envs;
  const int n_steps = a.n_steps;
  const ActionStream<IO> ctl(a);
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t env0 = static_cast<int64_t>(blockIdx.x) * kWalkPlanThreads;
  const int64_t env = env0 + tid;
  const uint32_t env_id0 = a.st.env_offset + static_cast<uint32_t>(env0);
  const bool mine = env < n;
  PLAN_CLOCK(0);
  // the first chunk's actions are DRAM-cold: ask for their lines now, the
  // prologue's loads and float64 arithmetic cover the wait
  if (mine && (lane & (IO != 0 ? 15 : 7)) == 0) {
    const int first = n_steps < kWalkPlanChunk ? n_steps : kWalkPlanChunk;
    for (int k = 0; k < first; ++k)
      prefetch_l2(ctl.at(static_cast<int64_t>(k) * n + env));
  }

  // ---- the env of this thread ----
  FastGeo g0;
  float o0x[3], o0y[3];
  uint32_t ctrl0 = 0;
  int si = 0, cls = 2;
  int4 row = make_int4(0, 0, 0, 0);
  float qx = 0.5f, qy = 0.5f, iwx = 0.f, iwy = 0.f, iws = 0.f;
  int hops = 0, recentres = 0, hops_synced = 0, fov_site = -1;
  bool check_area = true;   // simulator.py:156 at the first step of the call
  int stopped_at = -1;      // >= 0: handed over at this step of the call
#pragma unroll
  for (int i = 0; i < 3; ++i) g0.gx[i] = g0.gy[i] = o0x[i] = o0y[i] = 0.f;
  if (mine) {
    const Lattice4 lat = load_lattice4(a.st.lattice, env);
    fast_geo_bulk<RATE>(Lattice4{0.0, 0.0, lat.c, lat.s}, 0, &g0);
#pragma unroll
    for (int i = 0; i < 3; ++i) bulk_offset<RATE>(g0, i, &o0x[i], &o0y[i]);
    ctrl0 = a.st.ctrl_count[env];
    si = a.st.si_idx[env];
    row = __ldg(tab.nbr + si);
    cls = (row.w >> kSiteClassShift) & 3;
    Observed obs;
    obs.sync(load_fov4(a.st.fov, env), site_position(tab.position(si), lat));
    qx = obs.qx;
    qy = obs.qy;
    iwx = obs.iwx;
    iwy = obs.iwy;
    iws = __fdividef(1.0f, static_cast<float>(a.st.fov_scale[env]));
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    s_geo[i][tid] = g0.gx[i];
    s_geo[3 + i][tid] = g0.gy[i];
  }
  s_ctrl0[tid] = ctrl0;
  int steps_done = 0;  // steps of the call committed by this kernel
  PLAN_CLOCK(1);

  for (int t0 = 0; t0 < n_steps; t0 += kWalkPlanChunk) {
    const int len_c =
        n_steps - t0 < kWalkPlanChunk ? n_steps - t0 : kWalkPlanChunk;
    __syncthreads();  // the previous chunk's tile and queue are done with
    if (tid == 0) q_count = 0;
    __syncthreads();
    const bool live = mine && stopped_at < 0;
    // ---- plan, dense pass ----
    uint32_t busy = 0;  // bit k: class 0, bit 16 + k: class 1
    if (live) {
      int64_t at = static_cast<int64_t>(t0) * n + env;
      uint32_t ctrl = ctrl0 + static_cast<uint32_t>(t0);
      for (int k = 0; k < len_c; ++k, at += n, ++ctrl) {
        const double2 act = ctl.load(at);
        const float bx =
            clip_nanf(static_cast<float>(act.x), -1.f, 1.f) * md_s;
        const float by =
            clip_nanf(static_cast<float>(act.y), -1.f, 1.f) * md_s;
        const uint4 w = philox4x32_10k(env_id0 + static_cast<uint32_t>(tid),
                                       ctrl, 0u, PD_STREAM_KMC, a.keys);
        bool quiet0, quiet1;
        fast_quiet_both<RATE>(g0, bx, by, w.x, tm, &quiet0, &quiet1);
        busy |= (quiet0 ? 0u : 1u) << k | (quiet1 ? 0u : 1u) << (16 + k);
      }
    }
    // queue the controls that go on: a warp scan of the counts, one
    // shared-memory atomic per warp
    {
      const int cnt = __popc(busy);
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      uint32_t base = 0;
      if (total > 0) {
        if (lane == 31)
          base = atomicAdd(&q_count, static_cast<uint32_t>(total));
        base = __shfl_sync(0xffffffffu, base, 31);
      }
      uint32_t q = base + static_cast<uint32_t>(incl - cnt);
      for (uint32_t bits = busy; bits; bits &= bits - 1u, ++q) {
        const int bit = __ffs(bits) - 1;
        const int k = bit & 15, c = bit >> 4;
        if (q < kWalkPlanQueue)
          queue[q] =
              (static_cast<uint32_t>(k * kWalkPlanThreads + tid) << 1) | c;
        else  // no room: the exact code runs it (the env is handed over)
          reinterpret_cast<unsigned short*>(&tile[k][tid])[c] =
              static_cast<unsigned short>(kPlanUnsure);
      }
    }
    __syncthreads();
    if (t0 == 0) PLAN_CLOCK(2);
    // ---- plan, queue pass ----
    {
      const int n_q = q_count < kWalkPlanQueue ? static_cast<int>(q_count)
                                               : kWalkPlanQueue;
      for (int q = tid; q < n_q; q += kWalkPlanThreads) {
        const uint32_t entry = queue[q];
        const int c = entry & 1u;
        const int k = static_cast<int>(entry >> 1) / kWalkPlanThreads;
        const int el = static_cast<int>(entry >> 1) % kWalkPlanThreads;
        const int t = t0 + k;
        const double2 act = ctl.load(static_cast<int64_t>(t) * n + env0 + el);
        const float bx =
            clip_nanf(static_cast<float>(act.x), -1.f, 1.f) * md_s;
        const float by =
            clip_nanf(static_cast<float>(act.y), -1.f, 1.f) * md_s;
        FastGeo g;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          g.gx[i] = s_geo[i][el];
          g.gy[i] = s_geo[3 + i][el];
        }
        if (c) flip_geo(&g);
        const uint32_t r16 = plan_control<RATE>(
            g, bx, by, env_id0 + static_cast<uint32_t>(el),
            s_ctrl0[el] + static_cast<uint32_t>(t), tm, a.keys);
        reinterpret_cast<unsigned short*>(&tile[k][el])[c] =
            static_cast<unsigned short>(r16);
      }
    }
    __syncthreads();
    if (t0 == 0) PLAN_CLOCK(3);
    // ---- commit ----
    if (mine && stopped_at < 0 && (lane & (IO != 0 ? 15 : 7)) == 0)
      for (int k = t0 + kWalkPlanChunk;
           k < n_steps && k < t0 + 2 * kWalkPlanChunk; ++k)
        prefetch_l2(ctl.at(static_cast<int64_t>(k) * n + env));
    int committed = 0;  // steps of this chunk committed
    uint32_t changed = 0, recd = 0;  // bit k: step k moved the Si / re-centred
    const int si_start = si;
    if (live) {
      committed = len_c;
      bool ok = cls < 2 && q_clip_free(qx, qy, iwx, iwy, md_f);
      if (check_area) {
        if (q_inside(qx, qy)) check_area = false;
        else ok = false;
      }
      if (!ok) committed = 0;
      unsigned from = ~0u;
      while (ok) {
        const unsigned m = (cls == 0 ? busy : busy >> 16) & 0xFFFFu & from;
        if (!m) break;  // the chunk is done
        const int k = __ffs(m) - 1;
        from = ~0u << (k + 1);
        const uint32_t r16 =
            reinterpret_cast<const unsigned short*>(&tile[k][tid])[cls];
        if (r16 & kPlanUnsure) {
          committed = k;
          break;
        }
        const int n_hops = static_cast<int>(r16 & 7u);
        int si_t = si, cls_t = cls;
        int4 row_t = row;
        float qx_t = qx, qy_t = qy;
        for (int h = 0; h < n_hops; ++h) {
          const int slot = static_cast<int>((r16 >> (4 + 2 * h)) & 3u);
          // class 1: g1[i] = -g0[2 - i]
          const int j = cls_t == 0 ? slot : 2 - slot;
          const float sg = cls_t == 0 ? 1.f : -1.f;
          const float ox = j == 0 ? o0x[0] : (j == 1 ? o0x[1] : o0x[2]);
          const float oy = j == 0 ? o0y[0] : (j == 1 ? o0y[1] : o0y[2]);
          qx_t = __fmaf_rn(sg * ox, iwx, qx_t);
          qy_t = __fmaf_rn(sg * oy, iwy, qy_t);
          si_t = slot == 0 ? row_t.x : (slot == 1 ? row_t.y : row_t.z);
          row_t = __ldg(tab.nbr + si_t);
          cls_t = (row_t.w >> kSiteClassShift) & 3;
          if (cls_t == 2) break;  // onto the sheet's edge: the plan's next
                                  // iteration assumed a bulk site's geometry
        }
        const bool in = q_inside(qx_t, qy_t), out = q_outside(qx_t, qy_t);
        const bool due = hops + n_hops - hops_synced >= kObservedSyncHops;
        if (cls_t == 2 || (!in && !out) || (due && !out)) {
          committed = k;
          break;
        }
        si = si_t;
        row = row_t;
        cls = cls_t;
        qx = qx_t;
        qy = qy_t;
        hops += n_hops;
        if (out) {
          // simulator.py:156-169: the FOV is centred on the Si again
          qx = qy = 0.5f;
          iwx = iwy = iws;
          recentres += 1;
          fov_site = si;
          hops_synced = hops;
        }
        tile[k][tid] = static_cast<uint32_t>(si);
        changed |= 1u << k;
        if (out) recd |= 1u << k;
        if (!q_clip_free(qx, qy, iwx, iwy, md_f)) {
          committed = k + 1;
          break;
        }
      }
      steps_done = t0 + committed;
      if (committed < len_c) stopped_at = steps_done;
    }
    if (t0 == 0) PLAN_CLOCK(4);
    // ---- store ----
    {
      int cur = si_start;
      int64_t at = static_cast<int64_t>(t0) * n + env;
      for (int k = 0; k < committed; ++k, at += n) {
        if ((changed >> k) & 1u) cur = static_cast<int>(tile[k][tid]);
        store_step<IO>(a, at, cur, ((recd >> k) & 1u) != 0u, step_us);
      }
    }
  }
  PLAN_CLOCK(5);
  if (mine) {
    // the env as it stands after steps_done steps
    if (fov_site >= 0) store_centred_fov(a, env, fov_site);
    if (steps_done > 0) {
      const long long total =
          static_cast<long long>(steps_done) * step_us +
          static_cast<long long>(recentres) * a.image_duration_us;
      atomicAdd(reinterpret_cast<unsigned long long*>(a.st.sim_time_us + env),
                static_cast<unsigned long long>(total));
      a.st.si_idx[env] = si;
      a.st.ctrl_count[env] = ctrl0 + static_cast<uint32_t>(steps_done);
      atomicAdd(reinterpret_cast<unsigned long long*>(a.st.n_events + env),
                static_cast<unsigned long long>(steps_done + hops));
      if (hops > 0)
        atomicAdd(
            reinterpret_cast<unsigned long long*>(a.st.n_transitions + env),
            static_cast<unsigned long long>(hops));
    }
    if (stopped_at >= 0) {
      const uint32_t at = atomicAdd(a.defer_count, 1u);
      a.defer_list[at] = make_int2(static_cast<int>(env), stopped_at);
    }
  }
  PLAN_CLOCK(6);
}

// ---------------------------------------------------------------------------
// Launch (called from launch_step, pd_step.cu).
// ---------------------------------------------------------------------------
// k_rollout_plan keeps the lattice tables (float64 positions, ushort4
// neighbour rows) in shared memory: 24 bytes per site, two CTAs per SM.
constexpr int kPlanMaxSites = 3072;
static size_t shared_tables_bytes_host(int n_sites) {
  return static_cast<size_t>(n_sites) * (sizeof(double2) + sizeof(ushort4));
}

template <int RATE, int IO>
static int launch_plan(const StepArgs& a_in, cudaStream_t stream) {
  StepArgs a = a_in;
  // two CTAs per SM, one wave when the batch allows it
  const int64_t slots = 2LL * sm_count();
  int64_t epc = (a.st.n_envs + slots - 1) / slots;
  if (epc > kPlanEnvs) epc = kPlanEnvs;
  if (epc < 1) epc = 1;
  a.plan_envs_per_cta = static_cast<int32_t>(epc);
  const int64_t grid = (a.st.n_envs + epc - 1) / epc;
  const size_t smem = shared_tables_bytes_host(a.lat.n_sites);
  auto kern = k_rollout_plan<RATE, IO>;
  PD_CUDA_OK(cudaFuncSetAttribute(
      kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
      static_cast<int>(smem)));
  kern<<<static_cast<unsigned>(grid), kPlanThreads, smem, stream>>>(a);
  PD_CUDA_OK(cudaGetLastError());
  return PD_OK;
}

#ifdef PD_PLAN_CLOCKS
extern "C" int pd_debug_plan_clocks(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_plan_clocks, sizeof(g_plan_clocks)) ==
                 cudaSuccess
             ? 0
             : 1;
}
#endif

// How many envs the last k_walk_plan launch on a batch handed over.  A batch
// that has run for thousands of steps without a reset has 30-50 % of its Si
// atoms on the sheet's edge sites (the boundary is sticky under the
// relative_random workload), half of the batch goes through the list, and
// k_walk_fast alone is the faster kernel (measured at 1 Mi envs: 0.32 against
// 0.37 ms per 8 steps, 2.1 against 2.8 ms per 64; within an episode's length
// of a reset it is 0.23 against 0.20 and 1.8 against 1.1).  So the list kernel
// leaves the length of its list in a mapped host word, keyed by the batch's
// si_idx array, and launch_fast reads it before the next launch: above
// kListFraction of the batch it takes k_walk_fast, and tries the plan again
// every kProbeEvery-th launch; pd_reset forgets the hint.
struct PlanHint {
  const void* key;
  volatile uint32_t* count;  // cudaHostAlloc, mapped
  uint32_t launches;
};
static std::mutex& plan_hint_mutex() {
  static std::mutex m;
  return m;
}
static std::vector<PlanHint>& plan_hints() {
  static std::vector<PlanHint> v;
  return v;
}
constexpr double kListFraction = 0.15;
constexpr uint32_t kProbeEvery = 16;

static PlanHint* plan_hint_for(const void* key, bool create) {
  auto& v = plan_hints();
  for (auto& h : v)
    if (h.key == key) return &h;
  if (!create) return nullptr;
  if (v.size() >= 64) v.erase(v.begin());  // (the words are never freed)
  void* word = nullptr;
  if (cudaHostAlloc(&word, sizeof(uint32_t),
                    cudaHostAllocMapped | cudaHostAllocPortable) !=
      cudaSuccess) {
    (void)cudaGetLastError();
    return nullptr;
  }
  *static_cast<volatile uint32_t*>(word) = 0u;
  v.push_back(PlanHint{key, static_cast<volatile uint32_t*>(word), 0u});
  return &v.back();
}

// pd_reset: the batch starts afresh.
void forget_plan_hint(const void* key) {
  std::lock_guard<std::mutex> lock(plan_hint_mutex());
  if (PlanHint* h = plan_hint_for(key, false)) {
    *h->count = 0u;
    h->launches = 0u;
  }
}

// Large batches under the relative adapter: k_walk_plan, then k_walk_fast over
// the list of envs (and first steps) it handed over.
template <int RATE, int IO>
static int launch_walk_plan(const StepArgs& a_in, cudaStream_t stream) {
  StepArgs a = a_in;
  const int64_t n = a.st.n_envs;
  a.list_hint = nullptr;
  // (no page-locked allocation while the stream is being captured into a
  // graph: the launch then goes without a hint)
  cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &capturing) != cudaSuccess) {
    (void)cudaGetLastError();
    capturing = cudaStreamCaptureStatusNone;
  }
  if (capturing == cudaStreamCaptureStatusNone) {
    std::lock_guard<std::mutex> lock(plan_hint_mutex());
    if (PlanHint* h = plan_hint_for(a.st.si_idx, true)) {
      void* dev_word = nullptr;
      if (cudaHostGetDevicePointer(&dev_word, const_cast<uint32_t*>(h->count),
                                   0) == cudaSuccess)
        a.list_hint = static_cast<uint32_t*>(dev_word);
      else
        (void)cudaGetLastError();
    }
  }
  // stream-ordered scratch from the device's default pool, which is told once
  // to keep what it has been given (no trip to the driver per call)
  static thread_local int pool_ready_for = -1;
  int dev = 0;
  PD_CUDA_OK(cudaGetDevice(&dev));
  if (pool_ready_for != dev) {
    cudaMemPool_t pool;
    PD_CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, dev));
    unsigned long long keep = ~0ull;
    PD_CUDA_OK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold,
                                       &keep));
    pool_ready_for = dev;
  }
  void* scratch = nullptr;
  const size_t bytes = 16 + static_cast<size_t>(n) * sizeof(int2);
  PD_CUDA_OK(cudaMallocAsync(&scratch, bytes, stream));
  a.defer_count = static_cast<uint32_t*>(scratch);
  a.defer_list =
      reinterpret_cast<int2*>(static_cast<unsigned char*>(scratch) + 16);
  cudaError_t err = cudaMemsetAsync(scratch, 0, 16, stream);
  if (err == cudaSuccess) {
    const int64_t grid_p = (n + kWalkPlanThreads - 1) / kWalkPlanThreads;
    k_walk_plan<RATE, IO>
        <<<static_cast<unsigned>(grid_p), kWalkPlanThreads, 0, stream>>>(a);
    // the list: a few envs per 10^4 in a fresh batch (UNSURE controls,
    // ambiguous area tests), the envs in the sheet's edge region later on; a
    // full wave, warps without work exit
    k_walk_fast<RATE, IO, true, true>
        <<<sm_count() * PD_FAST_MIN_BLOCKS, kStepThreads, 0, stream>>>(a);
    err = cudaGetLastError();
  }
  cudaFreeAsync(scratch, stream);
  PD_CUDA_OK(err);
  return PD_OK;
}

// plan_mode: bit 0 = k_rollout_plan where it applies (small batches, relative
// adapter) instead of k_rollout_fast; bit 1 = k_walk_plan (large batches,
// relative adapter, eight steps or more) instead of k_walk_fast; bit 2 =
// k_walk_plan for every batch the walk kernels take.
template <int RATE>
int launch_fast(const StepArgs& a_in, bool walk, int grid, cudaStream_t stream,
                int plan_mode) {
  StepArgs a = a_in;
  {
    const FastTimes tm = fast_times(static_cast<long long>(a.dwell_us_scalar));
    a.fast_dwell_s = tm.dwell_s;
    a.fast_margin = tm.margin;
  }
  const int io = a.packed_out ? 1 : (a.actions_f32 ? 2 : 0);
  const bool rel = a.action_mode == PD_ACTION_RELATIVE_TO_SILICON;
  if (!walk && rel && (plan_mode & 1) && a.n_steps >= 8 &&
      a.lat.n_sites <= kPlanMaxSites)
    return io == 1   ? launch_plan<RATE, 1>(a, stream)
           : io == 2 ? launch_plan<RATE, 2>(a, stream)
                     : launch_plan<RATE, 0>(a, stream);
  // (measured: k_walk_plan pays from three waves of its CTAs and eight steps
  // per launch on; below that the fixed cost of the second launch and the
  // last, partly filled wave eat the gain)
  // plan_mode bit 2: whatever the sizes and the hint (the parity tests)
  bool walk_plan = walk && rel && (plan_mode & 4);
  if (walk && rel && !walk_plan && (plan_mode & 2) && a.n_steps >= 8 &&
      a.st.n_envs >=
          3LL * sm_count() * PD_WALK_PLAN_BLOCKS * kWalkPlanThreads) {
    walk_plan = true;
    std::lock_guard<std::mutex> lock(plan_hint_mutex());
    if (PlanHint* h = plan_hint_for(a.st.si_idx, false)) {
      h->launches += 1;
      if (static_cast<double>(*h->count) >
              kListFraction * static_cast<double>(a.st.n_envs) &&
          h->launches % kProbeEvery != 0)
        walk_plan = false;
    }
  }
  if (walk_plan)
    return io == 1   ? launch_walk_plan<RATE, 1>(a, stream)
           : io == 2 ? launch_walk_plan<RATE, 2>(a, stream)
                     : launch_walk_plan<RATE, 0>(a, stream);
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
                                         cudaStream_t, int);
template int launch_fast<PD_RATE_PRIOR>(const StepArgs&, bool, int,
                                        cudaStream_t, int);

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
