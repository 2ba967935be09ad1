// K1: the event step.  One thread owns one environment for the duration of a
// call: neighbour geometry -> rate function -> direct-method event loop with a
// microsecond clock -> Si update -> FOV safe-area check / re-centre.
//
//   graphene.py:238-276   PristineSingleSiGrRatePredictor.__call__
//   graphene.py:646-694   PristineSingleDopedGraphene.apply_control
//   simulator.py:107-182  PuttingDuneSimulator.step_and_image
//
// Compiled with -fmad=false (see pd_kmc.cuh).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <vector>

#include "pd_episode.cuh"
#include "pd_fast.cuh"

#ifndef PD_STEP_MIN_BLOCKS
#define PD_STEP_MIN_BLOCKS 4
#endif

namespace pd {

// One step_and_image (or apply_control) call for every env.
template <int RATE, bool STAGE>
__global__ void __launch_bounds__(kStepThreads)
    k_step(const StepArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  typename std::conditional<STAGE, SharedTables, GlobalTables>::type tab;
  if constexpr (STAGE) {
    tab = stage_tables(a.lat, smem);
  } else {
    tab.base = reinterpret_cast<const double2*>(a.lat.base_xy);
    tab.nbr = reinterpret_cast<const int4*>(a.lat.nbr);
  }
  const LogSink log{a.out.log_count ? a.out.log_capacity : 0,
                    a.out.log_elapsed_us, a.out.log_site, a.out.log_ctrl};
  const int64_t n = a.st.n_envs;
  const int64_t gtid =
      blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (gtid % a.lane_stride) return;  // after the cooperative table staging
  for (int64_t e = gtid / a.lane_stride; e < n;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x / a.lane_stride) {
    if (a.skip && a.skip[e]) continue;
    EnvRegs r = load_env(tab, a, e);
    Fov4 fov = load_fov4(a.st.fov, e);
    long long elapsed = 0;
    for (int c = 0; c < a.n_controls; ++c) {
      const double2 ctl = reinterpret_cast<const double2*>(
          a.controls_xy)[e * a.n_controls + c];
      const long long dwell =
          a.dwell_us ? a.dwell_us[e * a.n_controls + c] : a.dwell_us_scalar;
      // simulator.py:137 microscope frame -> material frame
      const double2 beam = a.material_frame
                               ? ctl
                               : microscope_to_material(fov, ctl.x, ctl.y);
      run_control<RATE>(tab, a.ra, a.st.seed, beam, dwell, c, e, log, &r);
      elapsed += dwell;  // simulator.py:149
    }
    uint8_t recentred = 0;
    if (!a.material_frame) {
      elapsed += a.image_duration_us;  // simulator.py:152-153
      if (silicon_outside_safe_area(fov, r.psi)) {  // simulator.py:156
        fov = centred_fov(r.psi, a.st.fov_scale[e]);
        store_fov4(a.st.fov, e, fov);
        elapsed += a.image_duration_us;  // simulator.py:168-169
        recentred = 1;
      }
      atomicAdd(reinterpret_cast<unsigned long long*>(a.st.sim_time_us + e),
                static_cast<unsigned long long>(elapsed));
    }
    a.st.si_idx[e] = r.si;
    a.st.ctrl_count[e] = r.ctrl_count;
    atomicAdd(reinterpret_cast<unsigned long long*>(a.st.n_events + e),
              static_cast<unsigned long long>(r.events));
    atomicAdd(reinterpret_cast<unsigned long long*>(a.st.n_transitions + e),
              static_cast<unsigned long long>(r.transitions));
    a.st.status[e] = r.status;
    if (a.out.elapsed_us) a.out.elapsed_us[e] = elapsed;
    if (a.out.transitions) a.out.transitions[e] = r.transitions;
    if (a.out.events) a.out.events[e] = r.events;
    if (a.out.recentred) a.out.recentred[e] = recentred;
    if (a.out.si_xy)
      reinterpret_cast<double2*>(a.out.si_xy)[e] = r.psi;
    if (a.out.log_count) a.out.log_count[e] = r.log_n;
  }
}

// n_steps consecutive single-control step_and_image calls in one launch.
template <int RATE, bool STAGE>
__global__ void __launch_bounds__(kStepThreads)
    k_rollout(const StepArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  typename std::conditional<STAGE, SharedTables, GlobalTables>::type tab;
  if constexpr (STAGE) {
    tab = stage_tables(a.lat, smem);
  } else {
    tab.base = reinterpret_cast<const double2*>(a.lat.base_xy);
    tab.nbr = reinterpret_cast<const int4*>(a.lat.nbr);
  }
  const LogSink log{0, nullptr, nullptr, nullptr};
  const int64_t n = a.st.n_envs;
  const int64_t gtid =
      blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (gtid % a.lane_stride) return;  // after the cooperative table staging
  for (int64_t e = gtid / a.lane_stride; e < n;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x / a.lane_stride) {
    EnvRegs r = load_env(tab, a, e);
    Fov4 fov = load_fov4(a.st.fov, e);
    const double scale = a.st.fov_scale[e];
    long long total = 0;
    bool fov_dirty = false;
    // Software prefetch of the next control hides the HBM latency of the
    // action stream behind the current step's arithmetic.
    double2 next = reinterpret_cast<const double2*>(a.controls_xy)[e];
    // The observed Si position q = (P_si - ll) / (ur - ll) feeds both the
    // RelativeToSilicon adapter and the safe-area test; it only changes when
    // the Si hops or the FOV is re-centred, so its four divisions (and the
    // adapter's two) are redone only then.  The cached values are the exact
    // same expressions, so results are unchanged.
    const bool relative = a.action_mode == PD_ACTION_RELATIVE_TO_SILICON;
    bool q_stale = true, check_area = true;
    double2 q = make_double2(0.0, 0.0);
    double rx = 0.0, ry = 0.0;
    for (int t = 0; t < a.n_steps; ++t) {
      const double2 ctl = next;
      if (t + 1 < a.n_steps)
        next = reinterpret_cast<const double2*>(
            a.controls_xy)[static_cast<int64_t>(t + 1) * n + e];
      double2 pos = ctl;
      if (relative) {
        if (q_stale) {
          q = observe(fov, r.psi);
          rx = __ddiv_rn(a.max_distance, __dsub_rn(fov.urx, fov.llx));
          ry = __ddiv_rn(a.max_distance, __dsub_rn(fov.ury, fov.lly));
          q_stale = false;
        }
        // action_adapters.py:163-188
        const double ax = clip_nan(ctl.x, -1.0, 1.0);
        const double ay = clip_nan(ctl.y, -1.0, 1.0);
        pos.x = clip_nan(__dadd_rn(q.x, __dmul_rn(ax, rx)), 0.0, 1.0);
        pos.y = clip_nan(__dadd_rn(q.y, __dmul_rn(ay, ry)), 0.0, 1.0);
      }
      const double2 beam = microscope_to_material(fov, pos.x, pos.y);
      const int hops_before = r.transitions;
      run_control<RATE>(tab, a.ra, a.st.seed, beam, a.dwell_us_scalar, 0, e,
                        log, &r);
      long long elapsed = a.dwell_us_scalar + a.image_duration_us;
      if (r.transitions != hops_before) {
        q_stale = true;
        check_area = true;
      }
      if (check_area) {
        check_area = false;
        if (silicon_outside_safe_area(fov, r.psi)) {
          fov = centred_fov(r.psi, scale);
          elapsed += a.image_duration_us;
          fov_dirty = true;
          q_stale = true;
        }
      }
      total += elapsed;
      if (a.si_idx_out) a.si_idx_out[static_cast<int64_t>(t) * n + e] = r.si;
      if (a.elapsed_us_out)
        a.elapsed_us_out[static_cast<int64_t>(t) * n + e] = elapsed;
    }
    if (fov_dirty) store_fov4(a.st.fov, e, fov);
    atomicAdd(reinterpret_cast<unsigned long long*>(a.st.sim_time_us + e),
              static_cast<unsigned long long>(total));
    a.st.si_idx[e] = r.si;
    a.st.ctrl_count[e] = r.ctrl_count;
    atomicAdd(reinterpret_cast<unsigned long long*>(a.st.n_events + e),
              static_cast<unsigned long long>(r.events));
    atomicAdd(reinterpret_cast<unsigned long long*>(a.st.n_transitions + e),
              static_cast<unsigned long long>(r.transitions));
    a.st.status[e] = r.status;
  }
}

// ---------------------------------------------------------------------------
// k_rollout_spec: the small-batch rollout with exact speculation.
//
// A small batch (BASELINE configs[1]: 4096 envs) leaves most lanes of the
// machine idle and the rollout is one long dependent chain per env: one KMC
// iteration (~1000 cycles of float64 latency) after the other.  But a control
// whose first iteration draws a waiting time beyond the dwell time changes
// nothing except the control counter, and that is the common case (~70 % of
// the controls of the relative_random workload).  So a group of G lanes owns
// one env and every round evaluates G iterations at once:
//   lane 0      the true next iteration: (step t, iteration `it`);
//   lane j > 0  iteration 0 of step t + j, assuming the control of step t ends
//               without a further hop and steps t+1 .. t+j-1 do not hop.
// Philox is counter based, so lane j simply uses (ctrl_count + j, 0).  After
// the round the group commits the longest prefix whose assumptions held:
// lane 0 alone if it hopped, otherwise lane 0 and the following no-hop steps
// up to and including the first lane that hopped (its control then continues
// in lane 0 of the next round).  Discarded lanes cost nothing the batch could
// have used.  Every committed value is computed by the same expressions as
// k_rollout from the same inputs, so results are bit-identical
// (test_rollout_speculative_equals_serial).
//
// The FOV re-centre of simulator.py:156-169 happens at the end of a step that
// hopped; whether it will happen is known as soon as the hop is (it depends on
// the Si position only), so the speculative lanes already use the re-centred
// FOV for the steps that follow.
// ---------------------------------------------------------------------------
template <int RATE, bool STAGE>
__global__ void __launch_bounds__(kStepThreads)
    k_rollout_spec(const StepArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  typename std::conditional<STAGE, SharedTables, GlobalTables>::type tab;
  if constexpr (STAGE) {
    tab = stage_tables(a.lat, smem);
  } else {
    tab.base = reinterpret_cast<const double2*>(a.lat.base_xy);
    tab.nbr = reinterpret_cast<const int4*>(a.lat.nbr);
  }
  const int G = a.lane_stride;  // power of two, 2..32
  const int lane = threadIdx.x & 31;
  const int j = lane & (G - 1);
  const int gbase = lane - j;
  const unsigned gmask = (G >= 32 ? 0xffffffffu : ((1u << G) - 1u)) << gbase;
  const int64_t n = a.st.n_envs;
  const int64_t gtid =
      blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t n_groups = static_cast<int64_t>(gridDim.x) * blockDim.x / G;
  const bool relative = a.action_mode == PD_ACTION_RELATIVE_TO_SILICON;
  const long long dwell = a.dwell_us_scalar;
  const long long step_us = dwell + a.image_duration_us;
  const double2* ctl = reinterpret_cast<const double2*>(a.controls_xy);
  const int n_steps = a.n_steps;

  for (int64_t e = gtid / G; e < n; e += n_groups) {
    // ---- state of the env, replicated in the G lanes of its group ----
    int si = a.st.si_idx[e];
    const Lattice4 lat = load_lattice4(a.st.lattice, e);
    double2 psi = site_position(tab.position(si), lat);
    Fov4 fov = load_fov4(a.st.fov, e);
    const double scale = a.st.fov_scale[e];
    const uint32_t env_id = a.st.env_offset + static_cast<uint32_t>(e);
    uint32_t ctrl_count = a.st.ctrl_count[e];
    uint8_t status = a.st.status[e];
    int transitions = 0, events = 0;
    long long total = 0;
    bool fov_dirty = false;
    int t = 0;              // current step
    uint32_t it = 0;        // next iteration of the current step's control
    long long elapsed = 0;  // clock of the current control
    double2 beam0 = make_double2(0.0, 0.0);  // beam of the current control
    bool first = true;       // first round: lane 0 only, un-re-centred FOV
    bool need_check = true;  // simulator.py:156 runs at t = 0 and after a hop
    bool geo_stale = true, obs_stale = true;
    bool pending_rec = false;
    int nb[3] = {0, 0, 0};
    double2 pn[3];
    pn[0] = pn[1] = pn[2] = make_double2(0.0, 0.0);
    Fov4 fov_n = fov;
    double2 q_n = make_double2(0.0, 0.0);
    double rx_n = 0.0, ry_n = 0.0;

    if (j < n_steps) prefetch_l1(ctl + static_cast<int64_t>(j) * n + e);
    if (G + j < n_steps) prefetch_l1(ctl + static_cast<int64_t>(G + j) * n + e);

    while (t < n_steps) {
      if (geo_stale) {
        tab.neighbors(si, nb);
#pragma unroll
        for (int i = 0; i < 3; ++i)
          pn[i] = site_position(tab.position(nb[i]), lat);
        geo_stale = false;
      }
      if (obs_stale) {
        // What the steps after the current control will see if it ends
        // without another hop: the FOV after the pending re-centre and the
        // observed Si position in it (action_adapters.py:163-188).
        pending_rec = need_check && silicon_outside_safe_area(fov, psi);
        fov_n = (pending_rec && !first) ? centred_fov(psi, scale) : fov;
        if (relative) {
          q_n = observe(fov_n, psi);
          rx_n = __ddiv_rn(a.max_distance, __dsub_rn(fov_n.urx, fov_n.llx));
          ry_n = __ddiv_rn(a.max_distance, __dsub_rn(fov_n.ury, fov_n.lly));
        }
        obs_stale = false;
      }
      const bool cont = it > 0;  // lane 0 continues a control already begun
      const int step = t + j;
      const bool valid = (first ? j == 0 : true) && step < n_steps;
      double2 beam = beam0;
      if (valid && !(j == 0 && cont)) {
        const double2 c = ctl[static_cast<int64_t>(step) * n + e];
        double2 pos = c;
        if (relative) {
          const double ax = clip_nan(c.x, -1.0, 1.0);
          const double ay = clip_nan(c.y, -1.0, 1.0);
          pos.x = clip_nan(__dadd_rn(q_n.x, __dmul_rn(ax, rx_n)), 0.0, 1.0);
          pos.y = clip_nan(__dadd_rn(q_n.y, __dmul_rn(ay, ry_n)), 0.0, 1.0);
        }
        beam = microscope_to_material(fov_n, pos.x, pos.y);
      }
      if (step + 2 * G < n_steps)
        prefetch_l1(ctl + static_cast<int64_t>(step + 2 * G) * n + e);

      // ---- one KMC iteration per lane (graphene.py:658-694) ----
      // A control whose clock landed exactly on the dwell time has ended
      // (graphene.py:658): lane 0 then has nothing to evaluate.
      const bool c_done = cont && elapsed >= dwell;
      long long el = j == 0 ? elapsed : 0;
      const uint4 w =
          philox4x32_10(env_id, ctrl_count + static_cast<uint32_t>(j),
                        j == 0 ? it : 0u, PD_STREAM_KMC, a.st.seed);
      int slot = 0;
      bool bad = false;
      bool hop = rate_event<RATE>(a.ra, beam, psi, pn, u53(w.x, w.y),
                                  u53(w.z, w.w), dwell, &el, &slot, &bad);
      if (!valid || (j == 0 && c_done)) hop = bad = false;
      const unsigned hops = (__ballot_sync(gmask, hop) & gmask) >> gbase;
      const unsigned bads = (__ballot_sync(gmask, bad) & gmask) >> gbase;
      const unsigned valids = (__ballot_sync(gmask, valid) & gmask) >> gbase;

      // ---- commit the prefix whose assumptions held ----
      int n_done = 0;  // steps completed by this round
      int jh = -1;     // lane whose hop is applied
      if (hops & 1u) {
        jh = 0;
        events += 1;
        if (bads & 1u) status |= PD_ENV_BAD_RATE;
      } else {
        // the control of step t has ended: the step completes
        const unsigned later = hops & ~1u;
        if (later) {
          jh = __ffs(later) - 1;
          n_done = jh;
        } else {
          n_done = __popc(valids);
        }
        const int n_iter = n_done + (jh > 0 ? 1 : 0);  // iterations committed
        events += n_iter - (c_done ? 1 : 0);
        if (bads & ((n_iter >= 32 ? 0u : (1u << n_iter)) - 1u))
          status |= PD_ENV_BAD_RATE;
        const bool rec = pending_rec;  // simulator.py:156-169
        if (j < n_done) {
          if (a.si_idx_out)
            a.si_idx_out[static_cast<int64_t>(step) * n + e] = si;
          if (a.elapsed_us_out)
            a.elapsed_us_out[static_cast<int64_t>(step) * n + e] =
                step_us + ((j == 0 && rec) ? a.image_duration_us : 0);
        }
        total += static_cast<long long>(n_done) * step_us +
                 (rec ? a.image_duration_us : 0);
        if (rec) {
          fov = centred_fov(psi, scale);
          fov_dirty = true;
          obs_stale = true;
        }
        need_check = false;
        ctrl_count += static_cast<uint32_t>(n_done);
        t += n_done;
        it = 0;
        elapsed = 0;
      }
      if (jh >= 0) {
        const int src = gbase + jh;
        const int slot_h = __shfl_sync(gmask, slot, src);
        const long long el_h = __shfl_sync(gmask, el, src);
        beam0.x = shfl_double(gmask, beam.x, src);
        beam0.y = shfl_double(gmask, beam.y, src);
        si = slot_h == 0 ? nb[0] : (slot_h == 1 ? nb[1] : nb[2]);
        psi = slot_h == 0 ? pn[0] : (slot_h == 1 ? pn[1] : pn[2]);
        transitions += 1;
        elapsed = el_h;
        it += 1;  // it was reset to 0 if the hop came from a later lane
        need_check = true;
        geo_stale = obs_stale = true;
      }
      if (first) {
        first = false;
        obs_stale = true;
      }
    }
    if (j == 0) {
      if (fov_dirty) store_fov4(a.st.fov, e, fov);
      atomicAdd(reinterpret_cast<unsigned long long*>(a.st.sim_time_us + e),
                static_cast<unsigned long long>(total));
      a.st.si_idx[e] = si;
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
// k_rollout_pre: the small-batch rollout for the prior / simple rates.
//
// Same ownership as k_rollout_spec (a group of G lanes per env, state
// replicated in the group), but the look-ahead runs on the float32 pre-pass
// of pd_kmc.cuh instead of the float64 chain:
//   phase A  lane 0 tests the true next iteration (step t, iteration `it`),
//            lane j > 0 iteration 0 of step t + j, with certainly_no_hop.
//            The prefix of lanes that are certain is committed: those steps
//            end without a hop, whatever came before them in the prefix.
//   phase B  the first iteration the pre-pass could not settle is now the
//            env's current iteration; every lane of the group evaluates it
//            exactly (float64, same expressions as k_rollout) and applies the
//            outcome.  No shuffles: the lanes stay bit-identical replicas.
// ~89 % of the controls of the relative_random workload never reach phase B.
//
// STREAM (1: pd_rollout_actions_host_f32, float32 actions / int32 elapsed;
// 2: pd_rollout_actions_host, float64 actions / int64 elapsed, the "not yet"
// pattern being two all-ones 64-bit words): the launch runs WHILE the actions
// arrive and the results leave over PCIe, instead of between copy-engine
// chunks.  In: one copy-engine H2D copy of the whole float32 action stream
// into a staging the call pre-filled with 0xFF bytes; a stepping lane reads
// its action straight from L2 and treats an element whose two words are both
// 0xFFFFFFFF (a NaN no adapter produces) as not there yet -- the groups follow
// the copy front element by element, no flags, no chunks.  (`copy_done`, a
// word copied behind the stream, settles the pathological case of real data
// with that bit pattern.)  Out: the int32 results go to HBM stagings that are
// pre-filled with 0xFF too; -1 is neither a site nor an elapsed time and every
// word is stored exactly once, so a word that is not -1 is final.  The CTAs
// of a few SMs are writers: they walk the result rows in order, re-read (from
// L2) the 16-byte units that still hold a -1 and stream complete 16 KB blocks
// to the caller's pinned buffers with full-warp stores -- the stepping groups
// pay no fence and no counter for that.  Measured on this box
// (profiles/pcie_sm_copy.cu): copy engine in + SM stores out sustain 43.5 GB/s
// per direction, SM loads in + SM stores out only 35 (the SMs' host reads
// slow down under write traffic), the copy engines both ways 49.
// ---------------------------------------------------------------------------
// Loads that go to L2 (no L1 allocation: a line that was read before its data
// arrived must not be served from L1 later).
__device__ __forceinline__ uint4 ld_relaxed_v4(const uint32_t* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ float2 ld_relaxed_f2(const float2* p) {
  float2 v;
  asm volatile("ld.relaxed.gpu.global.v2.f32 {%0, %1}, [%2];"
               : "=f"(v.x), "=f"(v.y)
               : "l"(p)
               : "memory");
  return v;
}
// `copy_done` is written by the copy engine behind the action copy: acquire at
// system scope, so that the re-read of the element that follows it observes
// everything the copy wrote before the flag.
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double2 ld_relaxed_d2(const double2* p) {
  double2 v;
  asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];"
               : "=d"(v.x), "=d"(v.y)
               : "l"(p)
               : "memory");
  return v;
}
// Bound on the polls of the streamed rollout (~15 s of 100 ns sleeps; the
// writers sleep 300 ns and use a quarter of it): beyond it the launch traps.
constexpr unsigned kStreamPollBound = 1u << 27;

// An element has arrived only when NEITHER of its words holds the fill pattern
// any more (a half-landed element must not be consumed); real data with that
// pattern in one word is settled by `copy_done`.
__device__ __forceinline__ bool action_missing(const double2 v) {
  return __double_as_longlong(v.x) == -1LL || __double_as_longlong(v.y) == -1LL;
}
__device__ __forceinline__ bool action_missing(const float2 v) {
  return __float_as_uint(v.x) == 0xFFFFFFFFu ||
         __float_as_uint(v.y) == 0xFFFFFFFFu;
}
// A 16-byte unit of results is final when none of its elements (int32 words,
// or int64 as word pairs) is -1 any more.
__device__ __forceinline__ bool unit_final(const uint4 v, int words) {
  if (words == 2)
    return !(v.x == 0xFFFFFFFFu && v.y == 0xFFFFFFFFu) &&
           !(v.z == 0xFFFFFFFFu && v.w == 0xFFFFFFFFu);
  return v.x != 0xFFFFFFFFu && v.y != 0xFFFFFFFFu && v.z != 0xFFFFFFFFu &&
         v.w != 0xFFFFFFFFu;
}

// sm_ctl words: election and tickets of the streamed rollout (zeroed by the
// call).  Roles per SM id follow at kCtlRoles.
enum : int {
  kCtlCopySms = 0,   // SMs that have been given the copy role so far
  kCtlCopyIdx = 1,   // copy CTAs so far
  kCtlStepTicket = 2,
  kCtlWriteTicket = 3,
  kCtlCopyDone = 4,  // written by the H2D stream behind the action copy
  kCtlRoles = 8,
  kCtlSmSlots = 1024,
  kCtlWords = kCtlRoles + kCtlSmSlots
};

struct StreamCopyArgs {
  const int32_t* si_idx_out;
  const void* elapsed_out;  // int32 (el_words 1) or int64 (el_words 2)
  int32_t* h_si_idx_out;
  void* h_elapsed_out;
  uint32_t* sm_ctl;
  int64_t n;
  int T;
  int el_words;
  unsigned long long* trace;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// trace[0] = first CTA start (min), [2] = last stepping CTA end, [3] = last
// writer end, [4] = first stepping block past its first step
__device__ __forceinline__ void trace_mark(unsigned long long* trace, int slot,
                                           bool is_min) {
  if (trace && threadIdx.x == 0) {
    if (is_min)
      atomicMin(trace + slot, globaltimer_ns());
    else
      atomicMax(trace + slot, globaltimer_ns());
  }
}

// Next ticket of a counter for the whole CTA.
__device__ __forceinline__ uint32_t cta_ticket(uint32_t* counter,
                                               uint32_t* s_slot) {
  __syncthreads();
  if (threadIdx.x == 0) *s_slot = atomicAdd(counter, 1u);
  __syncthreads();
  return *s_slot;
}

// Writer work: blocks of kStepThreads * 8 units (16 KB) of the result
// stagings in row order.  The CTA re-reads (L2) the units of its block that
// still hold a -1 until the whole block is final, then stores it to the
// caller's buffer with full-warp 16-byte stores.
__device__ __forceinline__ void stream_write(const StreamCopyArgs& a,
                                             uint32_t* s_slot) {
  constexpr int kInFlight = 8;
  const int64_t block = static_cast<int64_t>(kStepThreads) * kInFlight;
  const int64_t units = static_cast<int64_t>(a.T) * a.n * 4 / 16;
  const uint32_t blocks = static_cast<uint32_t>((units + block - 1) / block);
  for (;;) {
    const uint32_t tk = cta_ticket(a.sm_ctl + kCtlWriteTicket, s_slot);
    if (tk >= blocks) break;
    // a ticket = the same result elements of both arrays: one block of si
    // words, el_words blocks of elapsed words
#pragma unroll 1
    for (int part = 0; part < 1 + a.el_words; ++part) {
      const bool el = part > 0;
      const uint4* src = reinterpret_cast<const uint4*>(
          el ? a.elapsed_out : static_cast<const void*>(a.si_idx_out));
      uint4* dst = reinterpret_cast<uint4*>(
          el ? a.h_elapsed_out : static_cast<void*>(a.h_si_idx_out));
      if (!dst) continue;
      const int64_t base =
          el ? block * (static_cast<int64_t>(tk) * a.el_words + (part - 1))
             : block * tk;
      const int64_t lim = el ? units * a.el_words : units;
      uint4 v[kInFlight];
      unsigned pending = 0u;
#pragma unroll
      for (int k = 0; k < kInFlight; ++k)
        if (base + k * kStepThreads + threadIdx.x < lim) pending |= 1u << k;
      const unsigned mine = pending;
      for (unsigned polls = 0u;; ++polls) {
        // (results come from the stepping CTAs of this launch: if none can
        // become resident -- an SM limit at or below copy_sms -- this ends
        // the launch with an error instead of spinning for ever)
        if (polls > kStreamPollBound / 4u) __trap();
#pragma unroll
        for (int k = 0; k < kInFlight; ++k)
          if (pending >> k & 1u) {
            v[k] = ld_relaxed_v4(reinterpret_cast<const uint32_t*>(
                src + base + k * kStepThreads + threadIdx.x));
            if (unit_final(v[k], el ? a.el_words : 1)) pending &= ~(1u << k);
          }
        if (__syncthreads_and(pending == 0u)) break;
        __nanosleep(300);
      }
#pragma unroll
      for (int k = 0; k < kInFlight; ++k)
        if (mine >> k & 1u) dst[base + k * kStepThreads + threadIdx.x] = v[k];
    }
  }
  trace_mark(a.trace, 3, false);
}

// A copy CTA of the streamed rollout: a writer.
__device__ __noinline__ void stream_copy_role(const StreamCopyArgs a,
                                              uint32_t* s_slot) {
  trace_mark(a.trace, 0, true);
  if (a.h_si_idx_out || a.h_elapsed_out) stream_write(a, s_slot);
}

template <int RATE, bool STAGE, int STREAM = 0>
__global__ void __launch_bounds__(kStepThreads)
    k_rollout_pre(const StepArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint32_t s_slot[2];
  uint32_t block_id = blockIdx.x;
  if constexpr (STREAM != 0) {
    // Role by SM: the first `copy_sms` SMs on which a CTA of this launch
    // starts run only writer CTAs, so that no stepping CTA shares its SM's
    // load/store path with the PCIe traffic; every other CTA takes
    // stepping blocks from a ticket counter.  All work is handed out by
    // tickets and every CTA ends by draining the stepping tickets, so the
    // launch completes whatever subset of its CTAs is resident.
    if (threadIdx.x == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      uint32_t* role_p = a.sm_ctl + kCtlRoles + (smid & (kCtlSmSlots - 1));
      uint32_t r = atomicCAS(role_p, 0u, 1u);  // 0 unset, 1 being decided,
                                               // 2 copy, 3 stepping
      if (r == 0u) {
        const uint32_t k = atomicAdd(a.sm_ctl + kCtlCopySms, 1u);
        r = k < static_cast<uint32_t>(a.copy_sms) ? 2u : 3u;
        atomicExch(role_p, r);
      } else {
        for (unsigned polls = 0u; r == 1u; ++polls) {
          if (polls > kStreamPollBound) __trap();
          r = *reinterpret_cast<volatile uint32_t*>(role_p);
        }
      }
      s_slot[1] = r == 2u ? atomicAdd(a.sm_ctl + kCtlCopyIdx, 1u) : ~0u;
    }
    __syncthreads();
    const uint32_t copy_idx = s_slot[1];
    if (copy_idx != ~0u)
      stream_copy_role(
          StreamCopyArgs{
              a.si_idx_out,
              STREAM == 2 ? static_cast<const void*>(a.elapsed_us_out)
                          : static_cast<const void*>(a.elapsed32_out),
              a.h_si_idx_out,
              STREAM == 2 ? static_cast<void*>(a.h_elapsed64_out)
                          : static_cast<void*>(a.h_elapsed32_out),
              a.sm_ctl, a.st.n_envs, a.n_steps, STREAM == 2 ? 2 : 1, a.trace},
          s_slot);
    block_id = cta_ticket(a.sm_ctl + kCtlStepTicket, s_slot);
    if (block_id >= static_cast<uint32_t>(a.step_ctas)) return;
  }
  typename std::conditional<STAGE, SharedTables, GlobalTables>::type tab;
  if constexpr (STAGE) {
    tab = stage_tables(a.lat, smem);
  } else {
    tab.base = reinterpret_cast<const double2*>(a.lat.base_xy);
    tab.nbr = reinterpret_cast<const int4*>(a.lat.nbr);
  }
  const int G = a.lane_stride;  // power of two, 2..32
  const int lane = threadIdx.x & 31;
  const int j = lane & (G - 1);
  const int gbase = lane - j;
  const unsigned gfull = G >= 32 ? 0xffffffffu : ((1u << G) - 1u);
  const unsigned gmask = gfull << gbase;
  const int64_t n = a.st.n_envs;
  const int64_t n_groups =
      static_cast<int64_t>(STREAM != 0 ? a.step_ctas : gridDim.x) * blockDim.x / G;
  const bool relative = a.action_mode == PD_ACTION_RELATIVE_TO_SILICON;
  const long long dwell = a.dwell_us_scalar;
  const long long step_us = dwell + a.image_duration_us;
  const double2* ctl = reinterpret_cast<const double2*>(a.controls_xy);
  const int n_steps = a.n_steps;
  const float md = static_cast<float>(a.max_distance);

  for (;;) {  // STREAM: one pass per stepping ticket; otherwise one pass
  const int64_t gtid =
      block_id * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  for (int64_t e = gtid / G; e < n; e += n_groups) {
    // ---- state of the env, replicated in the G lanes of its group ----
    int si = a.st.si_idx[e];
    const Lattice4 lat = load_lattice4(a.st.lattice, e);
    double2 psi = site_position(tab.position(si), lat);
    Fov4 fov = load_fov4(a.st.fov, e);
    const double scale = a.st.fov_scale[e];
    const uint32_t env_id = a.st.env_offset + static_cast<uint32_t>(e);
    uint32_t ctrl_count = a.st.ctrl_count[e];
    uint8_t status = a.st.status[e];
    int transitions = 0, events = 0;
    long long total = 0;
    bool fov_dirty = false;
    int t = 0;              // current step
    uint32_t it = 0;        // next iteration of the current step's control
    long long elapsed = 0;  // clock of the current control
    double2 beam0 = make_double2(0.0, 0.0);  // beam of the current control
    bool first = true;       // first round: lane 0 only, un-re-centred FOV
    bool need_check = true;  // simulator.py:156 runs at t = 0 and after a hop
    bool stale = true;       // Si or FOV changed since the constants below
    bool pending_rec = false;
    PrepassGeo<RATE> geo;
#pragma unroll
    for (int i = 0; i < 3; ++i) geo.cx[i] = geo.cy[i] = 0.f;
    float qfx = 0.f, qfy = 0.f, wfx = 1.f, wfy = 1.f;
    unsigned stalled = 0u;  // STREAM: consecutive polls without an action

    if constexpr (STREAM == 0) {
      if (j < n_steps) prefetch_l1(ctl + static_cast<int64_t>(j) * n + e);
      if (G + j < n_steps)
        prefetch_l1(ctl + static_cast<int64_t>(G + j) * n + e);
    }

    while (t < n_steps) {
      if (stale) {
        prepass_geometry<RATE>(tab, si, lat, &geo);
        // Will the step that ends the current control re-centre the FOV
        // (simulator.py:156-169)?  The steps after it see the new FOV.
        wfx = static_cast<float>(fov.urx - fov.llx);
        wfy = static_cast<float>(fov.ury - fov.lly);
        qfx = __fdividef(static_cast<float>(psi.x - fov.llx), wfx);
        qfy = __fdividef(static_cast<float>(psi.y - fov.lly), wfy);
        pending_rec = false;
        if (need_check) {
          // well inside the safe area in float32: no float64 test needed
          const bool inside = qfx > 0.2501f && qfx < 0.7499f &&
                              qfy > 0.2501f && qfy < 0.7499f;
          if (!inside) pending_rec = silicon_outside_safe_area(fov, psi);
        }
        if (pending_rec && !first) {
          const Fov4 fn = centred_fov(psi, scale);
          wfx = static_cast<float>(fn.urx - fn.llx);
          wfy = static_cast<float>(fn.ury - fn.lly);
          qfx = __fdividef(static_cast<float>(psi.x - fn.llx), wfx);
          qfy = __fdividef(static_cast<float>(psi.y - fn.lly), wfy);
        }
        stale = false;
      }
      // ---- phase A: float32 look-ahead over the next G iterations ----
      const bool cont = it > 0;  // lane 0 continues a control already begun
      // a control whose clock landed exactly on the dwell time has ended
      const bool c_done = cont && elapsed >= dwell;
      const int step = t + j;
      bool valid = (first ? j == 0 : true) && step < n_steps;
      float2 act = make_float2(0.f, 0.f);
      double2 actd = make_double2(0.0, 0.0);  // STREAM == 2: the float64 action
      if constexpr (STREAM != 0) {
        // this lane's action, if it has arrived; the lanes up to the first
        // one whose action has not are this round's look-ahead
        if (valid && !(j == 0 && cont)) {
          if constexpr (STREAM == 2) {
            const double2* src = ctl + static_cast<int64_t>(step) * n + e;
            actd = ld_relaxed_d2(src);
            if (action_missing(actd)) {
              if (ld_relaxed_u32(a.copy_done))
                actd = ld_relaxed_d2(src);  // the copy has ended: it is data
              else
                valid = false;
            }
            act = make_float2(static_cast<float>(actd.x),
                              static_cast<float>(actd.y));
          } else {
            const float2* src =
                a.actions_f32 + static_cast<int64_t>(step) * n + e;
            act = ld_relaxed_f2(src);
            if (action_missing(act)) {
              if (ld_relaxed_u32(a.copy_done))
                act = ld_relaxed_f2(src);  // the copy has ended: it is data
              else
                valid = false;
            }
          }
        }
        const unsigned vm = (__ballot_sync(gmask, valid) & gmask) >> gbase;
        const unsigned inv = ~vm & gfull;
        const unsigned keep = inv ? (inv & (0u - inv)) - 1u : gfull;
        valid = valid && ((keep >> j) & 1u);
        if (!(vm & keep & 1u)) {
          // not even the current step is there: the copy front is behind us
          // (bounded: a copy that never arrives ends the launch with an
          // error the host sees, not with a hung device)
          if (++stalled > kStreamPollBound) __trap();
          __nanosleep(100);
          continue;
        }
        stalled = 0u;
      }
      bool certain = false;
      uint4 w = make_uint4(0u, 0u, 0u, 0u);  // Philox words of this lane's
                                             // iteration (re-used by phase B)
      if (valid) {
        if (j == 0 && c_done) {
          certain = true;
        } else {
          float bx, by;
          if (j == 0 && cont) {
            bx = static_cast<float>(beam0.x - psi.x);
            by = static_cast<float>(beam0.y - psi.y);
          } else {
            float px, py;
            if constexpr (STREAM != 0) {
              px = act.x;
              py = act.y;
            } else {
              const double2 c = ctl[static_cast<int64_t>(step) * n + e];
              px = static_cast<float>(c.x);
              py = static_cast<float>(c.y);
            }
            if (relative) {
              px = clip_nanf(px, -1.f, 1.f);
              py = clip_nanf(py, -1.f, 1.f);
              px = clip_nanf(qfx + px * __fdividef(md, wfx), 0.f, 1.f);
              py = clip_nanf(qfy + py * __fdividef(md, wfy), 0.f, 1.f);
            }
            bx = (px - qfx) * wfx;
            by = (py - qfy) * wfy;
          }
          w = philox4x32_10(env_id, ctrl_count + static_cast<uint32_t>(j),
                            j == 0 ? it : 0u, PD_STREAM_KMC, a.st.seed);
          certain = certainly_no_hop<RATE>(geo, bx, by, w.x,
                                           dwell - (j == 0 ? elapsed : 0));
        }
      }
      if constexpr (STREAM == 0) {
        if (step + 2 * G < n_steps)
          prefetch_l1(ctl + static_cast<int64_t>(step + 2 * G) * n + e);
      }
      const unsigned valids = (__ballot_sync(gmask, valid) & gmask) >> gbase;
      const unsigned certs = (__ballot_sync(gmask, certain) & gmask) >> gbase;
      const unsigned unsure = valids & ~certs;
      const int n_done = unsure ? __ffs(unsure) - 1 : __popc(valids);
      if (n_done > 0) {
        // the current control and the n_done - 1 after it end without a hop
        events += n_done - (c_done ? 1 : 0);
        const bool rec = pending_rec;  // simulator.py:156-169
        if (j < n_done) {
          if (a.si_idx_out)
            a.si_idx_out[static_cast<int64_t>(step) * n + e] = si;
          const long long el_out =
              step_us + ((j == 0 && rec) ? a.image_duration_us : 0);
          if constexpr (STREAM == 1) {
            if (a.elapsed32_out)
              a.elapsed32_out[static_cast<int64_t>(step) * n + e] =
                  static_cast<int32_t>(el_out);
          } else {
            if (a.elapsed_us_out)
              a.elapsed_us_out[static_cast<int64_t>(step) * n + e] = el_out;
          }
        }
        total += static_cast<long long>(n_done) * step_us +
                 (rec ? a.image_duration_us : 0);
        if (rec) {
          fov = centred_fov(psi, scale);
          fov_dirty = true;
          pending_rec = false;
          if (first) stale = true;  // otherwise the constants are the new FOV's
        }
        need_check = false;
        ctrl_count += static_cast<uint32_t>(n_done);
        t += n_done;
        it = 0;
        elapsed = 0;
      }
      first = false;
      if (!unsure) continue;

      // ---- phase B: the current iteration (t, it), exactly ----
      // It is the iteration lane `ju` looked at in phase A, so its Philox
      // words come from there; the three neighbour rates are evaluated by
      // three lanes (lane j takes neighbour j % 3) and gathered, instead of
      // three times by every lane.
      const int ju = __ffs(unsure) - 1;
      const uint4 wb = make_uint4(__shfl_sync(gmask, w.x, gbase + ju),
                                  __shfl_sync(gmask, w.y, gbase + ju),
                                  __shfl_sync(gmask, w.z, gbase + ju),
                                  __shfl_sync(gmask, w.w, gbase + ju));
      if (it == 0) {
        double2 c;
        if constexpr (STREAM != 0) {
          // lane ju loaded it in phase A
          if constexpr (STREAM == 2)
            c = make_double2(shfl_double(gmask, actd.x, gbase + ju),
                             shfl_double(gmask, actd.y, gbase + ju));
          else
            c = make_double2(
                static_cast<double>(__shfl_sync(gmask, act.x, gbase + ju)),
                static_cast<double>(__shfl_sync(gmask, act.y, gbase + ju)));
        } else {
          c = ctl[static_cast<int64_t>(t) * n + e];
        }
        double2 pos = c;
        if (relative) {
          if (G >= 4) {
            // relative_to_silicon(fov, psi, c, max_distance) with its four
            // divisions (observed Si x, y; max_distance / FOV width, height)
            // taken by four lanes
            const int part = j & 3;
            const double den = (part & 1) ? __dsub_rn(fov.ury, fov.lly)
                                          : __dsub_rn(fov.urx, fov.llx);
            const double num = part == 0   ? __dsub_rn(psi.x, fov.llx)
                               : part == 1 ? __dsub_rn(psi.y, fov.lly)
                                           : a.max_distance;
            const double val = __ddiv_rn(num, den);
            const double qx = shfl_double(gmask, val, gbase + 0);
            const double qy = shfl_double(gmask, val, gbase + 1);
            const double rx = shfl_double(gmask, val, gbase + 2);
            const double ry = shfl_double(gmask, val, gbase + 3);
            const double ax = clip_nan(c.x, -1.0, 1.0);
            const double ay = clip_nan(c.y, -1.0, 1.0);
            pos.x = clip_nan(__dadd_rn(qx, __dmul_rn(ax, rx)), 0.0, 1.0);
            pos.y = clip_nan(__dadd_rn(qy, __dmul_rn(ay, ry)), 0.0, 1.0);
          } else {
            pos = relative_to_silicon(fov, psi, c, a.max_distance);
          }
        }
        beam0 = microscope_to_material(fov, pos.x, pos.y);
      }
      int nb[3];
      tab.neighbors(si, nb);
      float r[3];
      double2 pm;  // position of this lane's neighbour
      if (G >= 4) {
        const int mine = j % 3;
        pm = site_position(
            tab.position(mine == 0 ? nb[0] : (mine == 1 ? nb[1] : nb[2])), lat);
        const float rm = RATE == PD_RATE_PRIOR ? rate_prior_one(beam0, psi, pm)
                                               : rate_simple_one(beam0, psi, pm);
#pragma unroll
        for (int i = 0; i < 3; ++i) r[i] = __shfl_sync(gmask, rm, gbase + i);
      } else {
        double2 pn[3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
          pn[i] = site_position(tab.position(nb[i]), lat);
        eval_rates<RATE>(a.ra, beam0, psi, pn, r);
        pm = make_double2(0.0, 0.0);
      }
      int slot = 0;
      bool bad = false;
      long long el = elapsed;
      const bool hop = kmc_event(r, u53(wb.x, wb.y), u53(wb.z, wb.w), dwell,
                                 &el, &slot, &bad);
      double2 p_new;
      if (G >= 4) {
        p_new.x = shfl_double(gmask, pm.x, gbase + slot);
        p_new.y = shfl_double(gmask, pm.y, gbase + slot);
      } else {
        p_new = site_position(
            tab.position(slot == 0 ? nb[0] : (slot == 1 ? nb[1] : nb[2])), lat);
      }
      if (bad) status |= PD_ENV_BAD_RATE;
      events += 1;
      if (hop) {
        si = slot == 0 ? nb[0] : (slot == 1 ? nb[1] : nb[2]);
        psi = p_new;
        transitions += 1;
        elapsed = el;
        it += 1;
        need_check = true;
        stale = true;
      } else {
        // the pre-pass was unsure but the control ends here: step t completes
        const bool rec = need_check && silicon_outside_safe_area(fov, psi);
        if (j == 0) {
          if (a.si_idx_out)
            a.si_idx_out[static_cast<int64_t>(t) * n + e] = si;
          const long long el_out = step_us + (rec ? a.image_duration_us : 0);
          if constexpr (STREAM == 1) {
            if (a.elapsed32_out)
              a.elapsed32_out[static_cast<int64_t>(t) * n + e] =
                  static_cast<int32_t>(el_out);
          } else {
            if (a.elapsed_us_out)
              a.elapsed_us_out[static_cast<int64_t>(t) * n + e] = el_out;
          }
        }
        total += step_us + (rec ? a.image_duration_us : 0);
        if (rec) {
          fov = centred_fov(psi, scale);
          fov_dirty = true;
        }
        if (rec || need_check) stale = true;
        need_check = false;
        ctrl_count += 1;
        t += 1;
        it = 0;
        elapsed = 0;
      }
    }
    if (j == 0) {
      if (fov_dirty) store_fov4(a.st.fov, e, fov);
      atomicAdd(reinterpret_cast<unsigned long long*>(a.st.sim_time_us + e),
                static_cast<unsigned long long>(total));
      a.st.si_idx[e] = si;
      a.st.ctrl_count[e] = ctrl_count;
      atomicAdd(reinterpret_cast<unsigned long long*>(a.st.n_events + e),
                static_cast<unsigned long long>(events));
      atomicAdd(reinterpret_cast<unsigned long long*>(a.st.n_transitions + e),
                static_cast<unsigned long long>(transitions));
      a.st.status[e] = status;
    }
  }
  if constexpr (STREAM != 0) {
    trace_mark(a.trace, 2, false);
    block_id = cta_ticket(a.sm_ctl + kCtlStepTicket, s_slot);
    if (block_id >= static_cast<uint32_t>(a.step_ctas)) break;
  } else {
    break;
  }
  }
}

// ---------------------------------------------------------------------------
// k_walk: the stepping kernel.  A *lane* owns one environment at a time and
// walks it through its work (n_steps x n_controls controls); a *warp* owns a
// contiguous range of environments.  Every trip of the main loop executes
// exactly one KMC iteration for every lane that has one pending; lanes whose
// control, step or environment just finished do their (short) bookkeeping and
// pull the next control / step / environment before the next trip.  The
// expensive part -- neighbour geometry, rate function, Philox, waiting time --
// therefore always runs with all lanes active instead of the whole warp
// waiting for its slowest environment (ncu on the one-thread-per-step kernel
// this replaces: 17.7 of 32 lanes active, profiles/r01_*baseline*).
//
// One launch covers pd_apply_control (material frame, no observation),
// pd_step_and_image (n_steps = 1) and pd_rollout (n_controls = 1).
// ---------------------------------------------------------------------------
template <int RATE, bool STAGE, bool EPISODE>
__global__ void __launch_bounds__(kStepThreads, PD_STEP_MIN_BLOCKS)
    k_walk(const StepArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  typename std::conditional<STAGE, SharedTables, GlobalTables>::type tab;
  if constexpr (STAGE) {
    tab = stage_tables(a.lat, smem);
  } else {
    tab.base = reinterpret_cast<const double2*>(a.lat.base_xy);
    tab.nbr = reinterpret_cast<const int4*>(a.lat.nbr);
  }
  const LogSink log{a.out.log_count ? a.out.log_capacity : 0,
                    a.out.log_elapsed_us, a.out.log_site, a.out.log_ctrl};
  const bool rollout = a.n_steps > 0;
  const int n_steps = rollout ? a.n_steps : 1;
  const int n_controls = a.n_controls;
  const int64_t n = a.st.n_envs;
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int64_t warps_total =
      static_cast<int64_t>(gridDim.x) * (kStepThreads / 32);
  const int64_t wid =
      static_cast<int64_t>(blockIdx.x) * (kStepThreads / 32) +
      (threadIdx.x >> 5);
  const int64_t per = (n + warps_total - 1) / warps_total;
  int64_t cursor = wid * per;
  const int64_t hi = cursor + per < n ? cursor + per : n;

  // lane state
  int64_t env = -1;
  EnvRegs r;
  Fov4 fov;
  double2 beam, next_ctl = make_double2(0.0, 0.0);
  long long dwell = 0, elapsed = 0, step_elapsed = 0, total = 0;
  uint32_t it = 0;
  int t = 0, c = 0;
  bool fov_dirty = false, any_recentre = false;
  bool ready = false;  // an iteration of the current control is pending
  // float32 pre-pass (pd_kmc.cuh certainly_no_hop): `checked` = the pending
  // iteration went through it and needs the float64 chain; `beam` holds the
  // raw control until the float64 beam is needed (beam_ready).
  constexpr int kCtlPrefetch = 5;
  constexpr bool kPrepass =
      !EPISODE && (RATE == PD_RATE_SIMPLE || RATE == PD_RATE_PRIOR);
  bool checked = false, beam_ready = false, geo_ok = false, obs_ok = false;
  bool need_check = true;
  PrepassGeo<RATE> geo;
  float qfx = 0.f, qfy = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) geo.cx[i] = geo.cy[i] = 0.f;
  r.si = 0;
  // episode mode: goal, simulated clock, actions taken (eval_lib.py:110-150)
  double2 goal = make_double2(0.0, 0.0);
  long long env_time = 0;
  int actions = 0;
  // episode mode on the prior / simple rates: the guarded float32 iteration
  // of pd_fast.cuh.  A control runs in float32 (site, geometry and beam
  // offset in `fs`, `bxf`, `byf`, its clock as the interval [e_lo, e_hi])
  // until it ends or float32 cannot settle an iteration; then the control is
  // taken again from its start (`si0` ...) by the exact code below.
  constexpr bool kFastRates = RATE == PD_RATE_SIMPLE || RATE == PD_RATE_PRIOR;
  const bool fast_on = EPISODE && kFastRates && a.fast_episode != 0 &&
                       a.ep.dwell_us > 0 &&
                       a.ep.dwell_us < 3000LL * 1000000LL && log.capacity == 0;
  const FastTimes tm = fast_times(static_cast<long long>(a.ep.dwell_us));
  FastSite fs;
  fs.si = -1;
  fs.cls = 2;
  fs.nb[0] = fs.nb[1] = fs.nb[2] = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) fs.geo.gx[i] = fs.geo.gy[i] = 0.f;
  float bxf = 0.f, byf = 0.f, e_lo = 0.f, e_hi = 0.f;
  bool exact_ctl = true;
  int si0 = 0, tr0 = 0, ev0 = 0;
  double2 psi0 = make_double2(0.0, 0.0);

  while (true) {
    // Bookkeeping repeats until enough lanes hold an iteration that needs the
    // float64 chain (a lane whose environment just ended finalises it, pulls
    // the next one and sets up its first control; with the pre-pass most
    // controls are consumed here as well), or until nothing can change.
#pragma unroll 1
    for (int rep = 0;; ++rep) {
      // Episodes: the controller (float64 round trips, atan2 / cos / sin) is
      // by far the longest block of a lane's cycle, so the lanes of a warp
      // take it together: no bookkeeping while any lane is inside a control.
      if constexpr (EPISODE) {
        if (a.walk_lockstep != 0 &&
            __any_sync(0xffffffffu, env >= 0 && ready))
          break;
      }
      // ---- next control / end of step / end of environment ----
      // A lane leaves this block with an iteration that needs the float64
      // chain pending (ready && checked) or without an env.  Controls whose
      // iteration the float32 pre-pass can settle (certainly_no_hop) are
      // consumed right here.
      if (env >= 0 && !(ready && checked)) {
        // A lane settles at most walk_controls_per_pass controls per pass
        // (it still finishes its step / env): without a cap the warp waits
        // for its luckiest lane's run of settled controls, with a cap of one
        // the per-pass overhead (ballots, env pull) is paid per control.
        int settled = 0;
        while (true) {
          if (!ready) {
            if constexpr (EPISODE) {
              bool done = false, reached = false;
              float reward = 0.f;
              if (c > 0) {
                // step end: image, re-centre, goal test (simulator.py:152-169,
                // goals.py:143-181)
                long long step_us = a.ep.dwell_us + a.ep.image_duration_us;
                if (silicon_outside_safe_area(fov, r.psi)) {
                  fov = centred_fov(r.psi, a.st.fov_scale[env]);
                  step_us += a.ep.image_duration_us;
                }
                env_time += step_us;
                ++actions;
                c = 0;
                if (goal_reached(fov, r.psi, goal)) {
                  done = reached = true;
                  reward = static_cast<float>(
                      pow(kGamma, static_cast<double>(step_us) / 1e6));
                } else if (actions >= a.ep.step_limit) {
                  done = true;  // StepLimitWrapper truncation
                }
              }
              // eval_lib.py:128: simulated-time limit; no goal: nothing to do
              if (!done && (!(goal.x == goal.x) ||
                            !(env_time < a.ep.timeout_us)))
                done = true;
              if (!done) {
                int nb[3];
                tab.neighbors(r.si, nb);
                double2 pn[3];
#pragma unroll
                for (int i = 0; i < 3; ++i)
                  pn[i] = site_position(tab.position(nb[i]), r.lat);
                const double2 ctl = greedy_control(
                    fov, r.psi, pn, goal, a.ep.argmax_x, a.ep.argmax_y);
                beam = microscope_to_material(fov, ctl.x, ctl.y);
                beam_ready = true;
                dwell = a.ep.dwell_us;
                elapsed = 0;
                it = 0;
                if constexpr (kFastRates) {
                  exact_ctl = !fast_on;
                  e_lo = e_hi = 0.f;
                  si0 = r.si;
                  psi0 = r.psi;
                  tr0 = r.transitions;
                  ev0 = r.events;
                  const float k = fast_offset_scale<kFastRates ? RATE
                                                               : PD_RATE_SIMPLE>();
                  bxf = static_cast<float>(beam.x - r.psi.x) * k;
                  byf = static_cast<float>(beam.y - r.psi.y) * k;
                }
                if (dwell > 0) {
                  ready = true;
                  checked = true;  // the greedy beam sits on a neighbour: hops
                  break;
                }
                r.ctrl_count += 1;  // zero dwell: the control is a no-op
                c = 1;
                continue;
              }
              // ---- episode finished: EvalResult + state write-back ----
              store_fov4(a.st.fov, env, fov);
              a.st.sim_time_us[env] = env_time;
              a.st.si_idx[env] = r.si;
              a.st.ctrl_count[env] = r.ctrl_count;
              a.st.n_events[env] = r.events;
              a.st.n_transitions[env] = r.transitions;
              a.st.status[env] = r.status;
              pd_episode_stats out;
              out.num_actions = actions;
              out.env_seconds =
                  reached ? static_cast<float>(
                                static_cast<double>(env_time) / 1e6)
                          : nanf("");
              out.total_reward = reward;
              out.reached_goal = reached ? 1 : 0;
              out.pad_[0] = out.pad_[1] = out.pad_[2] = 0;
              a.stats[env] = out;
              env = -1;
              break;
            } else if (c < n_controls) {
              if (settled >= a.walk_controls_per_pass) break;
              const int64_t ci = rollout ? static_cast<int64_t>(t) * n + env
                                         : env * n_controls + c;
              dwell = a.dwell_us ? a.dwell_us[ci] : a.dwell_us_scalar;
              if (dwell <= 0) {
                r.ctrl_count += 1;  // zero dwell: no rate evaluation
                ++c;
                continue;
              }
              // controls are fetched one step ahead (the first one when the
              // env is pulled) so their HBM latency hides behind arithmetic
              beam = next_ctl;  // raw control until beam_ready
              if (rollout) {
                if (t + 1 < n_steps)
                  next_ctl = reinterpret_cast<const double2*>(
                      a.controls_xy)[static_cast<int64_t>(t + 1) * n + env];
                // a control the pre-pass settles takes a few hundred cycles:
                // pull the action stream into L1 several steps ahead
                if (t + kCtlPrefetch < n_steps)
                  prefetch_l1(reinterpret_cast<const double2*>(a.controls_xy) +
                              static_cast<int64_t>(t + kCtlPrefetch) * n + env);
              } else if (c > 0) {
                beam = reinterpret_cast<const double2*>(a.controls_xy)[ci];
              }
              beam_ready = false;
              elapsed = 0;
              it = 0;
              ready = true;
              checked = false;
            } else {
              // all controls applied: take the image (simulator.py:152)
              if (!a.material_frame) {
                step_elapsed += a.image_duration_us;
                // simulator.py:156; the answer only changes when the Si hops
                // (need_check: set when the env is pulled and by every hop)
                if (need_check) {
                  need_check = false;
                  if (silicon_outside_safe_area(fov, r.psi)) {
                    fov = centred_fov(r.psi, a.st.fov_scale[env]);
                    step_elapsed += a.image_duration_us;  // simulator.py:168-169
                    fov_dirty = true;
                    any_recentre = true;
                    obs_ok = false;
                  }
                }
              }
              if (rollout) {
                if (a.si_idx_out)
                  a.si_idx_out[static_cast<int64_t>(t) * n + env] = r.si;
                if (a.elapsed_us_out)
                  a.elapsed_us_out[static_cast<int64_t>(t) * n + env] =
                      step_elapsed;
              }
              total += step_elapsed;
              ++t;
              if (t < n_steps) {
                c = 0;
                step_elapsed = 0;
                continue;
              }
              // ---- environment finished: write back ----
              if (fov_dirty) store_fov4(a.st.fov, env, fov);
              if (!a.material_frame)
                atomicAdd(reinterpret_cast<unsigned long long*>(
                              a.st.sim_time_us + env),
                          static_cast<unsigned long long>(total));
              a.st.si_idx[env] = r.si;
              a.st.ctrl_count[env] = r.ctrl_count;
              atomicAdd(
                  reinterpret_cast<unsigned long long*>(a.st.n_events + env),
                  static_cast<unsigned long long>(r.events));
              atomicAdd(reinterpret_cast<unsigned long long*>(
                            a.st.n_transitions + env),
                        static_cast<unsigned long long>(r.transitions));
              a.st.status[env] = r.status;
              if (a.out.elapsed_us) a.out.elapsed_us[env] = total;
              if (a.out.transitions) a.out.transitions[env] = r.transitions;
              if (a.out.events) a.out.events[env] = r.events;
              if (a.out.recentred) a.out.recentred[env] = any_recentre ? 1 : 0;
              if (a.out.si_xy)
                reinterpret_cast<double2*>(a.out.si_xy)[env] = r.psi;
              if (a.out.log_count) a.out.log_count[env] = r.log_n;
              env = -1;
              break;
            }
          }
          // ---- an iteration is pending: can float32 settle it? ----
          if (checked) break;
          checked = true;
          if constexpr (kPrepass) {
            if (a.prepass) {
              if (!geo_ok) {
                prepass_geometry<RATE>(tab, r.si, r.lat, &geo);
                geo_ok = true;
              }
              float bx, by;
              if (beam_ready || a.material_frame) {
                bx = static_cast<float>(beam.x - r.psi.x);
                by = static_cast<float>(beam.y - r.psi.y);
              } else {
                // beam - Si through the microscope frame, as the exact path
                // forms it (action_adapters.py:163-188, simulator.py:137)
                const float wx = static_cast<float>(fov.urx - fov.llx);
                const float wy = static_cast<float>(fov.ury - fov.lly);
                if (!obs_ok) {
                  qfx = static_cast<float>(r.psi.x - fov.llx) / wx;
                  qfy = static_cast<float>(r.psi.y - fov.lly) / wy;
                  obs_ok = true;
                }
                float px = static_cast<float>(beam.x);
                float py = static_cast<float>(beam.y);
                if (rollout &&
                    a.action_mode == PD_ACTION_RELATIVE_TO_SILICON) {
                  const float md = static_cast<float>(a.max_distance);
                  px = clip_nanf(px, -1.f, 1.f);
                  py = clip_nanf(py, -1.f, 1.f);
                  px = clip_nanf(qfx + px * __fdividef(md, wx), 0.f, 1.f);
                  py = clip_nanf(qfy + py * __fdividef(md, wy), 0.f, 1.f);
                }
                bx = (px - qfx) * wx;
                by = (py - qfy) * wy;
              }
              const uint4 w = philox4x32_10(r.env_id, r.ctrl_count, it,
                                            PD_STREAM_KMC, a.st.seed);
              if (certainly_no_hop<RATE>(geo, bx, by, w.x, dwell - elapsed)) {
                r.events += 1;
                step_elapsed += dwell;  // simulator.py:149
                r.ctrl_count += 1;
                ++c;
                ready = false;
                ++settled;
                continue;
              }
            }
          }
          if (!beam_ready) {
            double2 ctl = beam;
            if (rollout && a.action_mode == PD_ACTION_RELATIVE_TO_SILICON)
              ctl = relative_to_silicon(fov, r.psi, ctl, a.max_distance);
            // simulator.py:137 microscope frame -> material frame
            beam = a.material_frame
                       ? ctl
                       : microscope_to_material(fov, ctl.x, ctl.y);
            beam_ready = true;
          }
          break;
        }
      }
      // ---- idle lanes pull the next environments, in order ----
      const bool idle = env < 0;
      const unsigned im = __ballot_sync(0xffffffffu, idle);
      if (im && cursor < hi) {
        const int64_t cand = cursor + __popc(im & lt_mask);
        if (idle && cand < hi && !(a.skip && a.skip[cand])) {
          env = cand;
          // first control of this env, and the lines of the env this lane is
          // likely to pull next, are requested before anything waits
          if (!EPISODE && n_controls > 0) {
            next_ctl = reinterpret_cast<const double2*>(
                a.controls_xy)[rollout ? env : env * n_controls];
            if (rollout) {
#pragma unroll
              for (int k = 2; k < kCtlPrefetch; ++k)
                if (k < n_steps)
                  prefetch_l1(reinterpret_cast<const double2*>(a.controls_xy) +
                              static_cast<int64_t>(k) * n + env);
            }
          }
          if (cand + 32 < hi) {
            prefetch_env(a, cand + 32);
            if (!EPISODE && n_controls > 0)
              prefetch_l1(reinterpret_cast<const double2*>(a.controls_xy) +
                          (rollout ? cand + 32 : (cand + 32) * n_controls));
          }
          r = load_env(tab, a, env);
          fov = load_fov4(a.st.fov, env);
          if constexpr (EPISODE) {
            goal = reinterpret_cast<const double2*>(a.goal_xy)[env];
            env_time = a.ep.image_duration_us;  // eval_lib.py:121
            actions = 0;
            fs.si = -1;  // (its geometry is the previous env's)
          }
          t = 0;
          c = 0;
          total = 0;
          step_elapsed = 0;
          fov_dirty = false;
          any_recentre = false;
          ready = false;
          checked = false;
          geo_ok = obs_ok = false;
          need_check = true;
        }
        cursor += __popc(im);
      }
      const unsigned busy = __ballot_sync(0xffffffffu, env >= 0);
      const unsigned pending =
          __ballot_sync(0xffffffffu, env >= 0 && ready && checked);
      const bool can_change =
          (busy & ~pending) != 0u || (cursor < hi && busy != 0xffffffffu);
      if (!can_change || __popc(pending) >= a.walk_min_ready ||
          rep + 1 >= a.walk_max_reps)
        break;
    }
    if (!__any_sync(0xffffffffu, env >= 0)) break;

    bool settled_fast = false;
    if constexpr (EPISODE && kFastRates) {
      if (ready && !exact_ctl) {
        // ---- one KMC iteration in float32 (pd_fast.cuh) ----
        settled_fast = true;
        if (fs.si != r.si) fs = fast_site<RATE>(tab, r.si, r.lat.c, r.lat.s);
        const uint4 w = philox4x32_10k(r.env_id, r.ctrl_count, it,
                                       PD_STREAM_KMC, a.keys);
        int slot = 0;
        float t_lo = 0.f, t_hi = 0.f;
        const int kind = fast_event<RATE>(fs.geo, bxf, byf, w.x, w.z, e_lo,
                                          e_hi, tm, &slot, &t_lo, &t_hi);
        if (kind == FAST_NO_HOP) {
          r.events += 1;
          r.ctrl_count += 1;
          ++c;
          ready = false;
        } else if (kind == FAST_HOP) {
          float ox, oy;
          auto rotation = [&]() { return make_double2(r.lat.c, r.lat.s); };
          fast_hop<RATE>(tab, slot, rotation, &fs, &bxf, &byf, &ox, &oy);
          r.si = fs.si;
          r.psi = site_position(tab.position(r.si), r.lat);
          r.transitions += 1;
          r.events += 1;
          ++it;
          fast_advance(&e_lo, &e_hi, t_lo, t_hi);
        } else {
          // float32 cannot settle it: the control again, from its start, by
          // the exact code (this trip)
          r.si = si0;
          r.psi = psi0;
          r.transitions = tr0;
          r.events = ev0;
          it = 0;
          elapsed = 0;
          exact_ctl = true;
          settled_fast = false;
        }
      }
    }
    if (ready && !settled_fast) {
      // ---- one KMC iteration (graphene.py:658-694) ----
      int nb[3];
      tab.neighbors(r.si, nb);
      double2 pn[3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
        pn[i] = site_position(tab.position(nb[i]), r.lat);
      const uint4 w = philox4x32_10(r.env_id, r.ctrl_count, it, PD_STREAM_KMC,
                                    a.st.seed);
      int slot = 0;
      bool bad = false;
      const bool hit =
          rate_event<RATE>(a.ra, beam, r.psi, pn, u53(w.x, w.y), u53(w.z, w.w),
                           dwell, &elapsed, &slot, &bad);
      if (bad) r.status |= PD_ENV_BAD_RATE;
      r.events += 1;
      ++it;
      if (hit) {
        r.si = slot == 0 ? nb[0] : (slot == 1 ? nb[1] : nb[2]);
        r.psi = slot == 0 ? pn[0] : (slot == 1 ? pn[1] : pn[2]);
        r.transitions += 1;
        if (log.capacity > 0) {
          if (r.log_n < log.capacity) {
            const int64_t o = env * log.capacity + r.log_n;
            log.elapsed_us[o] = elapsed;
            log.site[o] = r.si;
            if (log.ctrl) log.ctrl[o] = c;
          } else {
            r.status |= PD_ENV_LOG_OVERFLOW;
          }
          r.log_n += 1;
        }
      }
      if (hit) {
        geo_ok = obs_ok = false;
        need_check = true;
      }
      if (elapsed >= dwell) {  // control finished
        step_elapsed += dwell;  // simulator.py:149
        r.ctrl_count += 1;
        ++c;
        ready = false;
      } else {
        checked = false;  // the control continues: pre-pass its next iteration
      }
    }
  }
}

// RateFunction seam: rates + successor sites, no state change.
template <int RATE>
__global__ void __launch_bounds__(kStepThreads)
    k_rates(const StepArgs a, float* __restrict__ rates_out,
            int32_t* __restrict__ nbr_out) {
  GlobalTables tab{reinterpret_cast<const double2*>(a.lat.base_xy),
                   reinterpret_cast<const int4*>(a.lat.nbr)};
  const int64_t n = a.st.n_envs;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
       e < n; e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int si = a.st.si_idx[e];
    const Lattice4 lat = load_lattice4(a.st.lattice, e);
    const double2 psi = site_position(tab.position(si), lat);
    int nb[3];
    tab.neighbors(si, nb);
    double2 pn[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) pn[i] = site_position(tab.position(nb[i]), lat);
    const double2 beam = reinterpret_cast<const double2*>(a.controls_xy)[e];
    float r[3];
    eval_rates<RATE>(a.ra, beam, psi, pn, r);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (rates_out) rates_out[3 * e + i] = r[i];
      if (nbr_out) nbr_out[3 * e + i] = nb[i];
    }
  }
}

// ---------------------------------------------------------------------------
// Launch helpers
// ---------------------------------------------------------------------------
// Staging the tables costs ~45 KB of L2 reads per CTA; it pays once every SM
// holds at least a few full CTAs of envs.
static bool use_staging(int64_t n_envs, int64_t work_per_env) {
  static const int forced = [] {  // PD_STAGE=0|1 overrides (A/B timing)
    const char* v = getenv("PD_STAGE");
    return v ? (v[0] != '0' ? 1 : 0) : -1;
  }();
  if (forced >= 0) return forced == 1;
  return n_envs * work_per_env >= 4LL * sm_count() * kStepThreads;
}

// Kernel choice.  Large (staged) batches use k_walk, whose converged KMC
// iterations are 10-20 % faster once every scheduler has several warps
// (profiles/r01_kernel_choice.md); small batches are latency-bound with one
// warp per scheduler, where the shorter instruction stream of the
// one-thread-per-step kernels wins.  PD_STEP_KERNEL=walk|simple overrides.
static bool walk_kernel(bool staged) {
  static const int choice = [] {
    const char* v = getenv("PD_STEP_KERNEL");
    return !v ? -1 : (v[0] == 'w' ? 1 : 0);
  }();
  return choice < 0 ? staged : choice == 1;
}

static int grid_for(int64_t n_envs, bool staged) {
  const int64_t blocks = (n_envs + kStepThreads - 1) / kStepThreads;
  // Persistent grid-stride loop: a whole number of CTAs per SM.
  const int64_t cap =
      static_cast<int64_t>(sm_count()) * (staged ? PD_STEP_MIN_BLOCKS : 16);
  return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

// Small batches cannot fill the machine with one env per lane: every
// scheduler would hold at most one warp and each warp would pay for its
// slowest lane.  Giving an env to only one lane in `stride` multiplies the
// number of warps (latency hiding) and divides the chance that a warp needs a
// second KMC iteration; the idle lanes cost nothing the batch could have used.
// Target: ~2.5 warps per scheduler (profiles/r01_kernel_choice.md).
static int lane_stride_forced() {
  static const int forced = [] {
    const char* v = getenv("PD_LANE_STRIDE");
    return v ? atoi(v) : 0;
  }();
  return forced;
}

static int lane_stride_for(int64_t n_envs) {
  const int forced = lane_stride_forced();
  if (forced > 0) return forced;
  const int64_t want_threads = static_cast<int64_t>(sm_count()) * 4 * 32 * 5 / 2;
  int stride = 1;
  while (stride < 32 && n_envs * stride * 2 <= want_threads) stride *= 2;
  return stride;
}

static int env_int(const char* name, int fallback) {
  const char* v = getenv(name);
  return v ? atoi(v) : fallback;
}

// PD_PREPASS=0 sends every iteration through the float64 chain (A/B timing,
// parity tests).
// Process-wide options: read from the environment once, changed at run time
// through pd_set_option (tests, A/B timing).
static int& option_prepass() {
  static int on = [] {
    const char* v = getenv("PD_PREPASS");
    return (!v || v[0] != '0') ? 1 : 0;
  }();
  return on;
}
static int& option_rollout_spec() {
  static int on = [] {
    const char* v = getenv("PD_ROLLOUT_SPEC");
    return (!v || v[0] != '0') ? 1 : 0;
  }();
  return on;
}
static int& option_race_sampling() {
  static int on = [] {
    const char* v = getenv("PD_SAMPLING_RACE");
    return (v && v[0] == '1') ? 1 : 0;
  }();
  return on;
}
// k_rollout_plan (pd_step_fast.cu) instead of k_rollout_fast: opt-in, it is
// the slower of the two on the benchmarked workload.
static int& option_plan() {
  static int on = [] {
    const char* v = getenv("PD_PLAN");
    return (v && v[0] == '1') ? 1 : 0;
  }();
  return on;
}
// k_walk_plan + k_walk_fast<LIST> instead of k_walk_fast alone for large
// batches under the relative adapter.
static int& option_walk_plan() {
  static int on = [] {
    const char* v = getenv("PD_WALK_PLAN");
    return !v ? 1 : (v[0] == '2' ? 2 : (v[0] != '0' ? 1 : 0));
  }();
  return on;
}
// (the float32 pre-pass and the fast kernels reason about the direct method)
static bool prepass_enabled() {
  return option_prepass() != 0 && option_race_sampling() == 0;
}

// PD_ROLLOUT_SPEC=0 keeps the serial k_rollout (A/B timing, parity tests).
static bool speculation_enabled() { return option_rollout_spec() != 0; }

template <int RATE, bool STAGE, int STREAM = 0>
static auto rollout_pre_kernel() -> void (*)(const StepArgs) {
  if constexpr (RATE == PD_RATE_SIMPLE || RATE == PD_RATE_PRIOR)
    return k_rollout_pre<RATE, STAGE, STREAM>;
  else
    return nullptr;
}

// What launch_step will run for a call (shared with the streamed host rollout,
// which needs to know beforehand that k_rollout_pre covers the batch in one
// wave).
struct StepPlan {
  bool staged, walk, spec, pre;
  int lane_stride, grid;
};

static StepPlan plan_step(const StepArgs& a, bool rollout, bool has_prepass) {
  StepPlan p;
  p.staged = use_staging(a.st.n_envs, rollout ? a.n_steps : 1);
  p.walk = walk_kernel(a.st.n_envs >= 4LL * sm_count() * kStepThreads);
  p.lane_stride = p.walk ? 1 : lane_stride_for(a.st.n_envs);
  // Rollouts of small batches speculate over the idle lanes (k_rollout_spec);
  // every lane then does useful work, so the group is twice as wide as the
  // idle-lane stride (measured at 4096 envs: G = 8: 5.1e9, 16: 6.0e9, 32:
  // 4.1e9 env-steps/s).
  p.spec = rollout && !p.walk && p.lane_stride >= 2 && a.n_steps >= 2 &&
           a.dwell_us_scalar > 0 && speculation_enabled();
  // ... and on the float32 pre-pass where the rate function has one
  // (k_rollout_pre; G = 8: 5.3e9, 16: 6.0e9, 32: 4.0e9).
  p.pre = p.spec && has_prepass && prepass_enabled();
  if (p.spec && p.lane_stride < 32 && !lane_stride_forced()) p.lane_stride *= 2;
  p.grid = grid_for(a.st.n_envs * p.lane_stride, p.staged);
  return p;
}

// The streamed host rollout: k_rollout_pre<.., STREAM> launched with one full
// wave of CTAs (resident CTAs per SM x SMs).  `copy_sms` SMs are set aside
// for the reader / writer CTAs; the others must hold every stepping block at
// once for the pipeline to flow (the launch still completes if they do not:
// all work goes by tickets).  Needs action rows that are whole 128-byte
// lines.  wave = 0 if the call does not qualify.
struct StreamPlan {
  int wave, step_ctas, copy_sms, per_sm;
};

template <int RATE>
static StreamPlan stream_plan_for(const StepArgs& a, int want_copy_sms,
                                  int mode) {
  StreamPlan sp{0, 0, 0, 0};
  const StepPlan p = plan_step(a, true, true);
  if (!p.pre || !p.staged || a.st.n_envs % 16 != 0) return sp;
  const int64_t want =
      (a.st.n_envs * p.lane_stride + kStepThreads - 1) / kStepThreads;
  auto kern = mode == 2 ? rollout_pre_kernel<RATE, true, 2>()
                        : rollout_pre_kernel<RATE, true, 1>();
  const size_t smem = static_cast<size_t>(a.lat.n_sites) *
                      (sizeof(double2) + sizeof(ushort4));
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kStepThreads,
                                                    smem) != cudaSuccess ||
      per_sm < 1) {
    (void)cudaGetLastError();
    return sp;
  }
  const int64_t step_sms = (want + per_sm - 1) / per_sm;
  int copy_sms = sm_count() - static_cast<int>(step_sms);
  if (copy_sms > want_copy_sms) copy_sms = want_copy_sms;
  if (copy_sms < 1) return sp;
  sp.wave = per_sm * sm_count();
  sp.step_ctas = static_cast<int>(want);
  sp.copy_sms = copy_sms;
  sp.per_sm = per_sm;
  return sp;
}

static StreamPlan stream_plan(const pd_rate_config* rc, const StepArgs& a,
                              int want_copy_sms, int mode) {
  if (rc->rate_fn == PD_RATE_SIMPLE)
    return stream_plan_for<PD_RATE_SIMPLE>(a, want_copy_sms, mode);
  if (rc->rate_fn == PD_RATE_PRIOR && !rc->prior)
    return stream_plan_for<PD_RATE_PRIOR>(a, want_copy_sms, mode);
  return StreamPlan{0, 0, 0, 0};
}

// pd_step_fast.cu: the guarded float32 kernels (prior / simple rates).
template <int RATE>
int launch_fast(const StepArgs& a, bool walk, int grid, cudaStream_t stream,
                int plan_mode);

// PD_FAST=0 keeps every iteration on the float64 chain (A/B timing; the
// parity tests compare the two).
static int& fast_flag() {
  static int on = [] {
    const char* v = getenv("PD_FAST");
    return (!v || v[0] != '0') ? 1 : 0;
  }();
  return on;
}
static int& option_race_sampling();
static bool fast_enabled() {
  return fast_flag() != 0 && option_race_sampling() == 0;
}

template <int RATE>
static int launch_step(const StepArgs& a_in, bool rollout,
                       cudaStream_t stream) {
  StepArgs a = a_in;
  constexpr bool kHasPrepass =
      RATE == PD_RATE_SIMPLE || RATE == PD_RATE_PRIOR;
  const StepPlan plan = plan_step(a, rollout, kHasPrepass);
  const bool staged = plan.staged, walk = plan.walk, spec = plan.spec,
             pre = plan.pre;
  const int grid = plan.grid;
  a.lane_stride = plan.lane_stride;
  a.keys = philox_keys_host(a.st.seed);
  a.prepass = prepass_enabled() ? 1 : 0;
  // PD_* knobs are read once per process, not per launch
  static const int walk_min_ready = env_int("PD_WALK_MIN_READY", 12);
  static const int walk_max_reps = env_int("PD_WALK_MAX_REPS", 4);
  static const int walk_controls = env_int("PD_WALK_CONTROLS", 4);
  a.walk_min_ready = walk_min_ready;
  a.walk_max_reps = walk_max_reps;
  a.walk_controls_per_pass = walk_controls;
  if constexpr (kHasPrepass) {
    // Rollouts with one positive dwell time below the reference's waiting
    // time cap (graphene.py:668; pd_fast.cuh kFastMaxDwellS): every decision
    // in float32 with an error bound, exact replay of what it cannot settle.
    if (rollout && !a.stream_mode && fast_enabled() && a.dwell_us_scalar > 0 &&
        a.dwell_us_scalar < 3000LL * 1000000LL && !a.skip) {
      const int plan_mode = (option_plan() ? 1 : 0) |
                            (option_walk_plan() ? 2 : 0) |
                            (option_walk_plan() == 2 ? 4 : 0);
      if (walk || !spec)
        return launch_fast<RATE>(a, true, grid_for(a.st.n_envs, true), stream,
                                 plan_mode);
      return launch_fast<RATE>(a, false, grid, stream, plan_mode);
    }
  }
  if (a.packed_out || (a.actions_f32 && !a.stream_mode)) {
    set_error("the float32 / packed rollout formats need the prior / simple "
              "rates, one positive dwell time below 3000 s and "
              "pd_set_fast_path(1)");
    return PD_ERR_UNSUPPORTED;
  }
  if (a.stream_mode) {
    // streamed host rollout: the caller went through stream_plan
    PD_REQUIRE(pre && staged && kHasPrepass,
               "streamed rollout needs the staged k_rollout_pre");
    const size_t smem = static_cast<size_t>(a.lat.n_sites) *
                        (sizeof(double2) + sizeof(ushort4));
    auto kern = a.stream_mode == 2 ? rollout_pre_kernel<RATE, true, 2>()
                                   : rollout_pre_kernel<RATE, true, 1>();
    PD_CUDA_OK(cudaFuncSetAttribute(
        kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
        static_cast<int>(smem)));
    kern<<<a.stream_wave, kStepThreads, smem, stream>>>(a);
    PD_CUDA_OK(cudaGetLastError());
    return PD_OK;
  }
  if (staged) {
    const size_t smem = static_cast<size_t>(a.lat.n_sites) *
                        (sizeof(double2) + sizeof(ushort4));
    auto kern = walk ? k_walk<RATE, true, false>
                : pre         ? rollout_pre_kernel<RATE, true>()
                : spec        ? k_rollout_spec<RATE, true>
                : rollout     ? k_rollout<RATE, true>
                              : k_step<RATE, true>;
    PD_CUDA_OK(cudaFuncSetAttribute(
        kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
        static_cast<int>(smem)));
    kern<<<grid, kStepThreads, smem, stream>>>(a);
  } else {
    auto kern = walk ? k_walk<RATE, false, false>
                : pre         ? rollout_pre_kernel<RATE, false>()
                : spec        ? k_rollout_spec<RATE, false>
                : rollout     ? k_rollout<RATE, false>
                              : k_step<RATE, false>;
    kern<<<grid, kStepThreads, 0, stream>>>(a);
  }
  PD_CUDA_OK(cudaGetLastError());
  return PD_OK;
}

// Copies the rate-function parameters that travel by value to the kernels.
int fill_rate_args(const pd_rate_config* rc, RateArgs* ra) {
  for (int i = 0; i < 3; ++i) ra->constant_rates[i] = rc->constant_rates[i];
  ra->gmm_n = 0;
  ra->prior_general = 0;
  ra->race_sampling = option_race_sampling();
  if (rc->rate_fn == PD_RATE_PRIOR && rc->prior) {
    const pd_prior* p = rc->prior;
    const bool defaults = p->mean[0] == 0.85 && p->mean[1] == 0.0 &&
                          p->cov[0][0] == 0.1 && p->cov[1][1] == 0.1 &&
                          p->cov[0][1] == 0.0 && p->cov[1][0] == 0.0 &&
                          p->max_rate == 0.23104906018664842;
    if (!defaults) {
      const double det = p->cov[0][0] * p->cov[1][1] -
                         p->cov[0][1] * p->cov[1][0];
      PD_REQUIRE(det > 0 && p->cov[0][0] > 0 && p->cov[1][1] > 0 &&
                     p->max_rate >= 0,
                 "pd_prior: covariance must be positive definite, max_rate "
                 ">= 0");
      ra->prior_general = 1;
      ra->prior_mean[0] = p->mean[0];
      ra->prior_mean[1] = p->mean[1];
      ra->prior_prec[0] = p->cov[1][1] / det;
      ra->prior_prec[1] = -(p->cov[0][1] + p->cov[1][0]) / det;
      ra->prior_prec[2] = p->cov[0][0] / det;
      ra->prior_max_rate = p->max_rate;
    }
  }
  if (rc->rate_fn != PD_RATE_GMM) return PD_OK;
  const pd_gmm* g = rc->gmm;
  PD_REQUIRE(g != nullptr, "PD_RATE_GMM needs rc->gmm");
  PD_REQUIRE(g->n_mixtures >= 1 && g->n_mixtures <= PD_GMM_MAX_MIXTURES,
             "n_mixtures out of range");
  const double kTwoPi = 6.283185307179586;
  // graphene.py:289-301 _normalizing_factor
  double max_mode = 0.0;
  for (int m = 0; m < g->n_mixtures; ++m) {
    PD_REQUIRE(g->variances[m][0] > 0 && g->variances[m][1] > 0,
               "variances must be positive");
    const double mode = g->mixture_weights[m] /
                        (kTwoPi * sqrt(g->variances[m][0] * g->variances[m][1]));
    if (mode > max_mode) max_mode = mode;
  }
  PD_REQUIRE(max_mode > 0, "mixture weights must be positive");
  const double norm = g->max_rate / max_mode;
  ra->gmm_n = g->n_mixtures;
  for (int m = 0; m < g->n_mixtures; ++m) {
    ra->gmm_coef[m] =
        norm * g->mixture_weights[m] /
        (kTwoPi * sqrt(g->variances[m][0] * g->variances[m][1]));
    ra->gmm_loc[m] = g->loc_distances[m];
    ra->gmm_nh_inv_v[m][0] = -0.5 / g->variances[m][0];
    ra->gmm_nh_inv_v[m][1] = -0.5 / g->variances[m][1];
  }
  return PD_OK;
}

static int dispatch_step(const pd_rate_config* rc, StepArgs& a, bool rollout,
                         cudaStream_t stream) {
  int frc = fill_rate_args(rc, &a.ra);
  if (frc != PD_OK) return frc;
  switch (rc->rate_fn) {
    case PD_RATE_GMM:
      return launch_step<PD_RATE_GMM>(a, rollout, stream);
    case PD_RATE_SIMPLE:
      return launch_step<PD_RATE_SIMPLE>(a, rollout, stream);
    case PD_RATE_PRIOR:
      if (a.ra.prior_general)
        return launch_step<kRatePriorGeneral>(a, rollout, stream);
      return launch_step<PD_RATE_PRIOR>(a, rollout, stream);
    case PD_RATE_CONSTANT:
      for (int i = 0; i < 3; ++i) a.ra.constant_rates[i] = rc->constant_rates[i];
      return launch_step<PD_RATE_CONSTANT>(a, rollout, stream);
    default:
      set_error("rate_fn %d is not handled by the scalar event kernel",
                rc->rate_fn);
      return PD_ERR_UNSUPPORTED;
  }
}

template <int RATE>
static int launch_episode_walk(const StepArgs& a_in, cudaStream_t stream) {
  StepArgs a = a_in;
  a.keys = philox_keys_host(a.st.seed);
  a.fast_episode = fast_enabled() ? 1 : 0;
  static const int lockstep = env_int("PD_EPISODE_LOCKSTEP", 1);
  a.walk_lockstep = lockstep;
  a.walk_min_ready = 33;  // episodes: two bookkeeping passes per trip
  a.walk_max_reps = 2;
  a.walk_controls_per_pass = 1 << 30;
  const bool staged = a.st.n_envs >= 2LL * sm_count() * kStepThreads;
  const int grid = grid_for(a.st.n_envs, staged);
  if (staged) {
    const size_t smem = static_cast<size_t>(a.lat.n_sites) *
                        (sizeof(double2) + sizeof(ushort4));
    PD_CUDA_OK(cudaFuncSetAttribute(
        k_walk<RATE, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
        static_cast<int>(smem)));
    k_walk<RATE, true, true><<<grid, kStepThreads, smem, stream>>>(a);
  } else {
    k_walk<RATE, false, true><<<grid, kStepThreads, 0, stream>>>(a);
  }
  PD_CUDA_OK(cudaGetLastError());
  return PD_OK;
}

// pd_run_episodes (after reset and goal selection, pd_episode.cu).
int launch_episodes(const pd_lattice* lat, const pd_state* st,
                    const pd_rate_config* rc, const pd_episode_config* cfg,
                    const double* goal_xy, pd_episode_stats* stats,
                    cudaStream_t stream) {
  StepArgs a{};
  a.lat = *lat;
  a.st = *st;
  a.n_controls = 1;
  a.n_steps = 1;
  a.ep = *cfg;
  a.goal_xy = goal_xy;
  a.stats = stats;
  a.dwell_us_scalar = cfg->dwell_us;
  a.image_duration_us = cfg->image_duration_us;
  int frc = fill_rate_args(rc, &a.ra);
  if (frc != PD_OK) return frc;
  switch (rc->rate_fn) {
    case PD_RATE_GMM:
      return launch_episode_walk<PD_RATE_GMM>(a, stream);
    case PD_RATE_SIMPLE:
      return launch_episode_walk<PD_RATE_SIMPLE>(a, stream);
    case PD_RATE_PRIOR:
      if (a.ra.prior_general)
        return launch_episode_walk<kRatePriorGeneral>(a, stream);
      return launch_episode_walk<PD_RATE_PRIOR>(a, stream);
    case PD_RATE_CONSTANT:
      return launch_episode_walk<PD_RATE_CONSTANT>(a, stream);
    default:
      set_error("pd_run_episodes: rate_fn %d is not supported", rc->rate_fn);
      return PD_ERR_UNSUPPORTED;
  }
}

int validate_common(const pd_lattice* lat, const pd_state* st,
                    const pd_rate_config* rc) {
  PD_REQUIRE(lat && st, "null lattice/state");
  PD_REQUIRE(lat->base_xy && lat->nbr && lat->n_sites > 0, "lattice not built");
  PD_REQUIRE(st->n_envs >= 0, "negative n_envs");
  PD_REQUIRE(st->n_envs == 0 || (st->si_idx && st->lattice && st->fov && st->fov_scale &&
                 st->ctrl_count && st->sim_time_us && st->n_events &&
                 st->n_transitions && st->status),
             "state has null arrays");
  if (rc) {
    PD_REQUIRE(rc->rate_fn >= PD_RATE_SIMPLE && rc->rate_fn <= PD_RATE_GMM,
               "unknown rate_fn");
    if (rc->rate_fn == PD_RATE_LEARNED)
      PD_REQUIRE(rc->mlp != nullptr, "PD_RATE_LEARNED needs rc->mlp");
  }
  return PD_OK;
}

// Implemented in pd_mlp.cu.
int& option_mlp_slim();
int learned_step(const pd_lattice* lat, const pd_state* st, const pd_mlp* mlp,
                 const StepArgs& a, bool rollout, cudaStream_t stream);
int learned_rates(const pd_lattice* lat, const pd_state* st, const pd_mlp* mlp,
                  const double* beam_xy, float* rates_out, int32_t* nbr_out,
                  cudaStream_t stream);

}  // namespace pd

using pd::StepArgs;

extern "C" int pd_set_fast_path(int enabled) {
  const int before = pd::fast_flag();
  pd::fast_flag() = enabled ? 1 : 0;
  return before;
}

extern "C" int pd_set_option(const char* name, int value) {
  PD_REQUIRE(name != nullptr, "null option name");
  int* slot = nullptr;
  if (!strcmp(name, "fast_path")) slot = &pd::fast_flag();
  if (!strcmp(name, "prepass")) slot = &pd::option_prepass();
  if (!strcmp(name, "rollout_spec")) slot = &pd::option_rollout_spec();
  if (!strcmp(name, "race_sampling")) slot = &pd::option_race_sampling();
  if (!strcmp(name, "plan")) slot = &pd::option_plan();
  if (!strcmp(name, "walk_plan")) slot = &pd::option_walk_plan();
  if (!strcmp(name, "mlp_slim")) slot = &pd::option_mlp_slim();
  if (!slot) {
    pd::set_error("pd_set_option: unknown option '%s'", name);
    return PD_ERR_INVALID_ARGUMENT;
  }
  *slot = slot == &pd::option_walk_plan() && value == 2 ? 2 : (value ? 1 : 0);
  return PD_OK;
}

extern "C" int pd_rates(const pd_lattice* lat, const pd_state* st,
                        const pd_rate_config* rc, const double* beam_xy,
                        float* rates_out, int32_t* nbr_out, void* stream) {
  int rcode = pd::validate_common(lat, st, rc);
  if (rcode != PD_OK) return rcode;
  PD_REQUIRE(rc && beam_xy, "null rate config / beam");
  if (st->n_envs == 0) return PD_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (rc->rate_fn == PD_RATE_LEARNED)
    return pd::learned_rates(lat, st, rc->mlp, beam_xy, rates_out, nbr_out, s);
  StepArgs a{};
  a.lat = *lat;
  a.st = *st;
  a.controls_xy = beam_xy;
  rcode = pd::fill_rate_args(rc, &a.ra);
  if (rcode != PD_OK) return rcode;
  const int grid = pd::grid_for(st->n_envs, false);
  switch (rc->rate_fn) {
    case PD_RATE_GMM:
      pd::k_rates<PD_RATE_GMM><<<grid, pd::kStepThreads, 0, s>>>(
          a, rates_out, nbr_out);
      break;
    case PD_RATE_SIMPLE:
      pd::k_rates<PD_RATE_SIMPLE><<<grid, pd::kStepThreads, 0, s>>>(
          a, rates_out, nbr_out);
      break;
    case PD_RATE_PRIOR:
      if (a.ra.prior_general)
        pd::k_rates<pd::kRatePriorGeneral><<<grid, pd::kStepThreads, 0, s>>>(
            a, rates_out, nbr_out);
      else
        pd::k_rates<PD_RATE_PRIOR><<<grid, pd::kStepThreads, 0, s>>>(
            a, rates_out, nbr_out);
      break;
    default:
      pd::k_rates<PD_RATE_CONSTANT><<<grid, pd::kStepThreads, 0, s>>>(
          a, rates_out, nbr_out);
  }
  PD_CUDA_OK(cudaGetLastError());
  return PD_OK;
}

static int step_common(const pd_lattice* lat, const pd_state* st,
                       const pd_rate_config* rc, const double* controls_xy,
                       const int64_t* dwell_us, int64_t dwell_us_scalar,
                       int32_t n_controls, int64_t image_duration_us,
                       int material_frame, const pd_step_out* out,
                       void* stream, const uint8_t* skip = nullptr) {
  int rcode = pd::validate_common(lat, st, rc);
  if (rcode != PD_OK) return rcode;
  PD_REQUIRE(rc != nullptr, "null rate config");
  PD_REQUIRE(n_controls >= 0, "negative n_controls");
  if (st->n_envs == 0) return PD_OK;
  PD_REQUIRE(n_controls == 0 || controls_xy != nullptr, "null controls");
  PD_REQUIRE(dwell_us != nullptr || dwell_us_scalar >= 0, "negative dwell");
  PD_REQUIRE(image_duration_us >= 0, "negative image duration");
  if (out && out->log_count) {
    PD_REQUIRE(out->log_capacity > 0 && out->log_elapsed_us && out->log_site,
               "event log requested without buffers");
  }
  if (st->n_envs == 0) return PD_OK;
  StepArgs a{};
  a.lat = *lat;
  a.st = *st;
  a.controls_xy = controls_xy;
  a.dwell_us = dwell_us;
  a.dwell_us_scalar = dwell_us_scalar;
  a.n_controls = n_controls;
  a.image_duration_us = image_duration_us;
  a.material_frame = material_frame;
  a.skip = skip;
  if (out) a.out = *out;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (rc->rate_fn == PD_RATE_LEARNED)
    return pd::learned_step(lat, st, rc->mlp, a, false, s);
  return pd::dispatch_step(rc, a, false, s);
}

namespace pd {
// pd_env.cu: step_and_image for the envs whose skip[e] == 0.
int step_and_image_masked(const pd_lattice* lat, const pd_state* st,
                          const pd_rate_config* rc, const double* controls_xy,
                          const int64_t* dwell_us, int64_t image_duration_us,
                          const uint8_t* skip, const pd_step_out* out,
                          void* stream) {
  return step_common(lat, st, rc, controls_xy, dwell_us, 0, 1,
                     image_duration_us, 0, out, stream, skip);
}
}  // namespace pd

extern "C" int pd_apply_control(const pd_lattice* lat, const pd_state* st,
                                const pd_rate_config* rc, const double* beam_xy,
                                const int64_t* dwell_us,
                                int64_t dwell_us_scalar, const pd_step_out* out,
                                void* stream) {
  return step_common(lat, st, rc, beam_xy, dwell_us, dwell_us_scalar, 1, 0, 1,
                     out, stream);
}

extern "C" int pd_step_and_image(const pd_lattice* lat, const pd_state* st,
                                 const pd_rate_config* rc,
                                 const double* controls_xy,
                                 const int64_t* dwell_us,
                                 int64_t dwell_us_scalar, int32_t n_controls,
                                 int64_t image_duration_us,
                                 const pd_step_out* out, void* stream) {
  return step_common(lat, st, rc, controls_xy, dwell_us, dwell_us_scalar,
                     n_controls, image_duration_us, 0, out, stream);
}

extern "C" int pd_step_and_image_host(
    const pd_lattice* lat, const pd_state* st, const pd_rate_config* rc,
    const double* h_controls_xy, const int64_t* h_dwell_us,
    int64_t dwell_us_scalar, int32_t n_controls, int64_t image_duration_us,
    double* d_controls_xy, int64_t* d_dwell_us, const pd_step_out* d_out,
    int64_t* h_elapsed_us, double* h_si_xy, double* h_fov, void* stream) {
  PD_REQUIRE(st != nullptr, "null state");
  PD_REQUIRE(n_controls == 0 || (h_controls_xy && d_controls_xy),
             "null controls / staging");
  PD_REQUIRE(h_dwell_us == nullptr || d_dwell_us != nullptr,
             "per-env dwell needs device staging");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = static_cast<size_t>(st->n_envs);
  if (n_controls > 0)
    PD_CUDA_OK(cudaMemcpyAsync(d_controls_xy, h_controls_xy,
                               n * n_controls * 2 * sizeof(double),
                               cudaMemcpyHostToDevice, s));
  if (h_dwell_us)
    PD_CUDA_OK(cudaMemcpyAsync(d_dwell_us, h_dwell_us,
                               n * n_controls * sizeof(int64_t),
                               cudaMemcpyHostToDevice, s));
  int rcode = pd_step_and_image(lat, st, rc, d_controls_xy,
                                h_dwell_us ? d_dwell_us : nullptr,
                                dwell_us_scalar, n_controls, image_duration_us,
                                d_out, stream);
  if (rcode != PD_OK) return rcode;
  if (h_elapsed_us) {
    PD_REQUIRE(d_out && d_out->elapsed_us, "elapsed_us needs device staging");
    PD_CUDA_OK(cudaMemcpyAsync(h_elapsed_us, d_out->elapsed_us,
                               n * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  }
  if (h_si_xy) {
    PD_REQUIRE(d_out && d_out->si_xy, "si_xy needs device staging");
    PD_CUDA_OK(cudaMemcpyAsync(h_si_xy, d_out->si_xy, n * 2 * sizeof(double),
                               cudaMemcpyDeviceToHost, s));
  }
  if (h_fov)
    PD_CUDA_OK(cudaMemcpyAsync(h_fov, st->fov, n * 4 * sizeof(double),
                               cudaMemcpyDeviceToHost, s));
  PD_CUDA_OK(cudaStreamSynchronize(s));
  return PD_OK;
}

extern "C" int pd_rollout(const pd_lattice* lat, const pd_state* st,
                          const pd_rate_config* rc, const double* controls_xy,
                          int64_t dwell_us_scalar, int32_t n_steps,
                          int64_t image_duration_us, int32_t* si_idx_out,
                          int64_t* elapsed_us_out, void* stream) {
  return pd_rollout_actions(lat, st, rc, controls_xy, PD_ACTION_DIRECT, 0.0,
                            dwell_us_scalar, n_steps, image_duration_us,
                            si_idx_out, elapsed_us_out, stream);
}

extern "C" int pd_rollout_actions(const pd_lattice* lat, const pd_state* st,
                                  const pd_rate_config* rc,
                                  const double* controls_xy,
                                  int32_t action_mode,
                                  double max_distance_angstroms,
                                  int64_t dwell_us_scalar, int32_t n_steps,
                                  int64_t image_duration_us,
                                  int32_t* si_idx_out, int64_t* elapsed_us_out,
                                  void* stream) {
  PD_REQUIRE(action_mode == PD_ACTION_DIRECT ||
                 action_mode == PD_ACTION_RELATIVE_TO_SILICON,
             "unknown action_mode");
  int rcode = pd::validate_common(lat, st, rc);
  if (rcode != PD_OK) return rcode;
  PD_REQUIRE(rc != nullptr, "null rate config");
  PD_REQUIRE(n_steps >= 0 && (n_steps == 0 || controls_xy), "bad action stream");
  PD_REQUIRE(dwell_us_scalar >= 0 && image_duration_us >= 0, "negative time");
  if (st->n_envs == 0 || n_steps == 0) return PD_OK;
  StepArgs a{};
  a.lat = *lat;
  a.st = *st;
  a.controls_xy = controls_xy;
  a.dwell_us_scalar = dwell_us_scalar;
  a.n_controls = 1;
  a.n_steps = n_steps;
  a.action_mode = action_mode;
  a.max_distance = max_distance_angstroms;
  a.image_duration_us = image_duration_us;
  a.si_idx_out = si_idx_out;
  a.elapsed_us_out = elapsed_us_out;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (rc->rate_fn == PD_RATE_LEARNED)
    return pd::learned_step(lat, st, rc->mlp, a, true, s);
  return pd::dispatch_step(rc, a, true, s);
}

extern "C" int pd_rollout_host(const pd_lattice* lat, const pd_state* st,
                               const pd_rate_config* rc,
                               const double* h_controls_xy,
                               int64_t dwell_us_scalar, int32_t n_steps,
                               int64_t image_duration_us,
                               double* d_controls_xy, int32_t* d_si_idx,
                               int64_t* d_elapsed_us, int32_t* h_si_idx,
                               int64_t* h_elapsed_us, void* stream) {
  return pd_rollout_actions_host(lat, st, rc, h_controls_xy, PD_ACTION_DIRECT,
                                 0.0, dwell_us_scalar, n_steps,
                                 image_duration_us, d_controls_xy, d_si_idx,
                                 d_elapsed_us, h_si_idx, h_elapsed_us, stream);
}

namespace pd {
// Side streams for the host-buffer entry points: the action stream is split
// into chunks so that the H2D copy of chunk i+1, the kernel of chunk i and
// the D2H copy of chunk i-1 overlap (PCIe is full duplex).
struct HostPipeline {
  cudaStream_t h2d = nullptr, d2h = nullptr;
  cudaEvent_t start = nullptr, copied[16] = {}, stepped[16] = {};
  int device = -1;
  // streamed rollout: sm_ctl[kCtlWords] + trace marks (device), a pinned 1
  uint32_t* flags = nullptr;
  uint32_t* h_one = nullptr;
  // Stagings the library keeps when the caller passes none (grow-only):
  // [0] float32 actions, [1] float64 controls, [2] si, [3] int64 elapsed,
  // [4] int32 elapsed.  The streamed form re-fills [0], [2], [4] and the
  // control words behind each call on the d2h stream, so that the next call
  // finds them ready: `clean_*` = leading 16-byte units known to hold 0xFF.
  void* own[5] = {};
  size_t own_bytes[5] = {};
  int64_t clean_in = 0, clean_out = 0;
  cudaEvent_t done = nullptr, copied_all = nullptr, cleaned = nullptr;
};

// Leaves no copy in flight on the caller's buffers when a host-buffer call
// returns early (an error code from PD_CUDA_OK or a failed launch): the
// caller may free or reuse its host and device buffers as soon as the call
// is back.  The streamed form's "ready" state of the stagings is dropped too.
struct DrainOnError {
  HostPipeline* p;
  cudaStream_t s;
  bool armed = true;
  ~DrainOnError() {
    if (!armed) return;
    cudaStreamSynchronize(p->h2d);
    cudaStreamSynchronize(p->d2h);
    cudaStreamSynchronize(s);
    p->clean_in = p->clean_out = 0;
  }
};

// Makes sure the library-owned staging `k` holds `bytes` bytes.
static int own_staging(HostPipeline* p, int k, size_t bytes) {
  if (p->own_bytes[k] >= bytes) return PD_OK;
  PD_CUDA_OK(cudaDeviceSynchronize());
  if (p->own[k]) PD_CUDA_OK(cudaFree(p->own[k]));
  p->own[k] = nullptr;
  p->own_bytes[k] = 0;
  p->clean_in = p->clean_out = 0;
  PD_CUDA_OK(cudaMalloc(&p->own[k], bytes));
  p->own_bytes[k] = bytes;
  return PD_OK;
}

static int host_pipeline(HostPipeline** out) {
  static thread_local HostPipeline p;
  int dev = 0;
  PD_CUDA_OK(cudaGetDevice(&dev));
  if (p.device != dev) {
    PD_CUDA_OK(cudaStreamCreateWithFlags(&p.h2d, cudaStreamNonBlocking));
    PD_CUDA_OK(cudaStreamCreateWithFlags(&p.d2h, cudaStreamNonBlocking));
    PD_CUDA_OK(cudaEventCreateWithFlags(&p.start, cudaEventDisableTiming));
    for (int i = 0; i < 16; ++i) {
      PD_CUDA_OK(cudaEventCreateWithFlags(&p.copied[i], cudaEventDisableTiming));
      PD_CUDA_OK(
          cudaEventCreateWithFlags(&p.stepped[i], cudaEventDisableTiming));
    }
    PD_CUDA_OK(cudaEventCreateWithFlags(&p.done, cudaEventDisableTiming));
    PD_CUDA_OK(cudaEventCreateWithFlags(&p.copied_all, cudaEventDisableTiming));
    PD_CUDA_OK(cudaEventCreateWithFlags(&p.cleaned, cudaEventDisableTiming));
    for (int k = 0; k < 5; ++k) {
      p.own[k] = nullptr;  // (a previous device's buffers stay with it)
      p.own_bytes[k] = 0;
    }
    p.clean_in = p.clean_out = 0;
    p.flags = nullptr;
    p.h_one = nullptr;
    if (cudaMalloc(&p.flags, (kCtlWords + 16) * sizeof(uint32_t)) !=
            cudaSuccess ||
        cudaHostAlloc(&p.h_one, sizeof(uint32_t), cudaHostAllocDefault) !=
            cudaSuccess) {
      (void)cudaGetLastError();
      p.flags = nullptr;  // the chunked pipeline is used instead
    } else {
      *p.h_one = 1u;
    }
    p.device = dev;
  }
  *out = &p;
  return PD_OK;
}
}  // namespace pd

namespace pd {
// One launch for the fills of the streamed rollout: zeroes the control words
// and sets the action / result stagings to 0xFF bytes (null or empty ranges
// are skipped).  16-byte stores.
struct FillRange {
  uint4* p;
  int64_t units;
  uint32_t word;
};
__global__ void __launch_bounds__(256) k_stream_fill(FillRange r0, FillRange r1,
                                                     FillRange r2, FillRange r3) {
  const FillRange rs[4] = {r0, r1, r2, r3};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint4 v = make_uint4(rs[k].word, rs[k].word, rs[k].word, rs[k].word);
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
         i < rs[k].units; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
      rs[k].p[i] = v;
  }
}

// ---- streamed form of the host-buffer rollouts: one launch that overlaps
// both PCIe copies (k_rollout_pre<.., STREAM>; see the comment above that
// kernel).  mode 1: float32 actions / int32 elapsed; mode 2: float64 actions
// / int64 elapsed.  Small batches on the prior / simple rates with
// page-locked host buffers (device-visible result buffers); sets *handled =
// false when the call does not qualify, and the caller takes the chunked
// copy-engine pipeline.  PD_HOST_STREAMED=0 forces the latter;
// PD_HOST_COPY_SMS sets the SMs given to the writer CTAs.
static int streamed_rollout(HostPipeline* pipe, const pd_lattice* lat,
                            const pd_state* st, const pd_rate_config* rc,
                            const void* h_actions, int mode,
                            int32_t action_mode, double max_distance_angstroms,
                            int64_t dwell_us_scalar, int32_t n_steps,
                            int64_t image_duration_us, void* d_actions,
                            int32_t* d_si_idx, void* d_elapsed,
                            int32_t* h_si_idx, void* h_elapsed, bool owned,
                            cudaStream_t s, bool* handled) {
  *handled = false;
  const int64_t n = st->n_envs;
  const int in_bytes = mode == 2 ? 16 : 8;   // per env-step
  const int el_bytes = mode == 2 ? 8 : 4;
  int rcode = PD_OK;
  static const int copy_sms = [] {
    // Opt-in (PD_HOST_STREAMED=1): the launch reads a staging that a
    // copy-engine copy is still writing, which CUDA's memory model leaves
    // undefined; it is correct on this hardware as long as the copy engine's
    // writes land in units of >= 8 bytes (the checks below; tests/
    // test_gpu_events.py stress test), not by specification.  The default is
    // the chunked pipeline, where no kernel touches a buffer in flight.
    const char* on = getenv("PD_HOST_STREAMED");
    if (!on || on[0] != '1') return 0;
    const char* v = getenv("PD_HOST_COPY_SMS");
    const int c = v ? atoi(v) : 4;
    return c < 1 ? 1 : (c > 64 ? 64 : c);
  }();
  auto device_visible = [](const void* h) -> void* {
    if (!h) return nullptr;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, h) != cudaSuccess) {
      (void)cudaGetLastError();
      return nullptr;
    }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
  };
  auto page_locked = [](const void* h) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, h) != cudaSuccess) {
      (void)cudaGetLastError();
      return false;
    }
    return at.type == cudaMemoryTypeHost;
  };
  auto now_us = [] {
    return std::chrono::duration<double, std::micro>(
               std::chrono::steady_clock::now().time_since_epoch())
        .count();
  };
  const double cpu_in = now_us();
  if (copy_sms > 0 && pipe->flags && rc && lat &&
      static_cast<int64_t>(n_steps) * n >= (1 << 18) && n_steps >= 32 &&
      n < (1LL << 31) &&
      (reinterpret_cast<uintptr_t>(d_actions) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(d_si_idx) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(d_elapsed) & 15) == 0 &&
      (action_mode == PD_ACTION_DIRECT ||
       action_mode == PD_ACTION_RELATIVE_TO_SILICON) &&
      validate_common(lat, st, rc) == PD_OK) {
    void* hv_si = device_visible(h_si_idx);
    void* hv_el = device_visible(h_elapsed);
    StepArgs a{};
    a.lat = *lat;
    a.st = *st;
    a.dwell_us_scalar = dwell_us_scalar;
    a.n_controls = 1;
    a.n_steps = n_steps;
    a.action_mode = action_mode;
    a.max_distance = max_distance_angstroms;
    a.image_duration_us = image_duration_us;
    const bool ptrs_ok =
        page_locked(h_actions) &&
        (!h_si_idx || (hv_si && (reinterpret_cast<uintptr_t>(hv_si) & 15) == 0)) &&
        (!h_elapsed ||
         (hv_el && (reinterpret_cast<uintptr_t>(hv_el) & 15) == 0));
    const StreamPlan sp =
        ptrs_ok ? stream_plan(rc, a, copy_sms, mode) : StreamPlan{0, 0, 0, 0};
    if (sp.wave > 0) {
      a.stream_mode = mode;
      a.si_idx_out = h_si_idx ? d_si_idx : nullptr;
      a.h_si_idx_out = static_cast<int32_t*>(hv_si);
      if (mode == 2) {
        a.controls_xy = static_cast<const double*>(d_actions);
        a.elapsed_us_out = h_elapsed ? static_cast<int64_t*>(d_elapsed) : nullptr;
        a.h_elapsed64_out = static_cast<int64_t*>(hv_el);
      } else {
        a.actions_f32 = static_cast<const float2*>(d_actions);
        a.elapsed32_out = h_elapsed ? static_cast<int32_t*>(d_elapsed) : nullptr;
        a.h_elapsed32_out = static_cast<int32_t*>(hv_el);
      }
      a.sm_ctl = pipe->flags;
      a.copy_done = pipe->flags + kCtlCopyDone;
      a.copy_sms = sp.copy_sms;
      a.step_ctas = sp.step_ctas;
      a.stream_wave = sp.wave;
      static const bool trace = getenv("PD_HOST_TRACE") != nullptr;
      unsigned long long* d_trace = reinterpret_cast<unsigned long long*>(
          pipe->flags + kCtlWords);
      if (trace) {
        const unsigned long long init[5] = {~0ull, 0, 0, 0, 0};
        PD_CUDA_OK(cudaMemcpyAsync(d_trace, init, sizeof(init),
                                   cudaMemcpyHostToDevice, s));
        a.trace = d_trace;
      }
      const double cpu0 = now_us();
      // fills -> (H2D stream) the action copy and the word behind it
      //       -> (s) the launch, which follows the copy front.
      // With library-owned stagings the fills were done behind the previous
      // call (d2h stream), and this call leaves the same behind itself.
      const int64_t in_units = static_cast<int64_t>(n_steps) * n * in_bytes / 16;
      const int64_t out_units = static_cast<int64_t>(n_steps) * n * 4 / 16;
      const int64_t el_units = static_cast<int64_t>(n_steps) * n * el_bytes / 16;
      auto fill = [&](cudaStream_t fs) {
        k_stream_fill<<<sm_count() * 4, 256, 0, fs>>>(
            FillRange{reinterpret_cast<uint4*>(pipe->flags),
                          kCtlWords / 4, 0u},
            FillRange{reinterpret_cast<uint4*>(d_actions), in_units,
                      0xFFFFFFFFu},
            FillRange{reinterpret_cast<uint4*>(d_si_idx),
                      h_si_idx ? out_units : 0, 0xFFFFFFFFu},
            FillRange{reinterpret_cast<uint4*>(d_elapsed),
                      h_elapsed ? el_units : 0, 0xFFFFFFFFu});
        return cudaGetLastError();
      };
      const bool ready = owned && !trace && pipe->clean_in >= in_units &&
                         pipe->clean_out >= out_units;
      if (ready) {
        PD_CUDA_OK(cudaStreamWaitEvent(s, pipe->cleaned, 0));
        PD_CUDA_OK(cudaStreamWaitEvent(pipe->h2d, pipe->cleaned, 0));
      } else {
        // A re-fill behind an earlier call (library-owned stagings) may still
        // be running, and it writes the control words this call is about to
        // use; after this call they are used, whoever owns the stagings.
        PD_CUDA_OK(cudaStreamSynchronize(pipe->d2h));
        pipe->clean_in = pipe->clean_out = 0;
        PD_CUDA_OK(fill(s));
        PD_CUDA_OK(cudaEventRecord(pipe->start, s));
        PD_CUDA_OK(cudaStreamWaitEvent(pipe->h2d, pipe->start, 0));
      }
      PD_CUDA_OK(cudaMemcpyAsync(d_actions, h_actions,
                                 static_cast<size_t>(in_units) * 16,
                                 cudaMemcpyHostToDevice, pipe->h2d));
      PD_CUDA_OK(cudaMemcpyAsync(pipe->flags + kCtlCopyDone, pipe->h_one,
                                 sizeof(uint32_t), cudaMemcpyHostToDevice,
                                 pipe->h2d));
      const double cpu_copy = now_us();
      rcode = dispatch_step(rc, a, true, s);
      if (rcode != PD_OK) {
        cudaStreamSynchronize(pipe->h2d);
        pipe->clean_in = pipe->clean_out = 0;
        return rcode;
      }
      const double cpu_launch = now_us();
      if (owned && !trace) {
        PD_CUDA_OK(cudaEventRecord(pipe->done, s));
        PD_CUDA_OK(cudaEventRecord(pipe->copied_all, pipe->h2d));
        PD_CUDA_OK(cudaStreamWaitEvent(pipe->d2h, pipe->done, 0));
        PD_CUDA_OK(cudaStreamWaitEvent(pipe->d2h, pipe->copied_all, 0));
        PD_CUDA_OK(fill(pipe->d2h));
        PD_CUDA_OK(cudaEventRecord(pipe->cleaned, pipe->d2h));
        pipe->clean_in = in_units;
        pipe->clean_out = (h_si_idx && h_elapsed) ? out_units : 0;
        // the results are in the caller's buffers once the launch has ended
        PD_CUDA_OK(cudaEventSynchronize(pipe->done));
        *handled = true;
        return PD_OK;
      }
      PD_CUDA_OK(cudaStreamSynchronize(s));
      PD_CUDA_OK(cudaStreamSynchronize(pipe->h2d));
      if (trace) {
        const double cpu1 = now_us();
        fprintf(stderr,
                "pd host trace (us): cpu checks %.1f, fill+copy enqueued %.1f, "
                "launched %.1f, synced %.1f\n",
                cpu0 - cpu_in, cpu_copy - cpu_in, cpu_launch - cpu_in,
                cpu1 - cpu_in);
        unsigned long long h[5];
        PD_CUDA_OK(cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr,
                "pd host trace (us): call %.1f | kernel: stepping done %.1f, "
                "writers done %.1f\n",
                cpu1 - cpu0, (h[2] - h[0]) * 1e-3, (h[3] - h[0]) * 1e-3);
      }
      *handled = true;
      return PD_OK;
    }
  }

  return PD_OK;
}
}  // namespace pd

extern "C" int pd_rollout_actions_host(
    const pd_lattice* lat, const pd_state* st, const pd_rate_config* rc,
    const double* h_controls_xy, int32_t action_mode,
    double max_distance_angstroms, int64_t dwell_us_scalar, int32_t n_steps,
    int64_t image_duration_us, double* d_controls_xy, int32_t* d_si_idx,
    int64_t* d_elapsed_us, int32_t* h_si_idx, int64_t* h_elapsed_us,
    void* stream) {
  PD_REQUIRE(st != nullptr, "null state");
  PD_REQUIRE(n_steps >= 0, "negative n_steps");
  PD_REQUIRE(n_steps == 0 || (h_controls_xy && d_controls_xy),
             "null controls / staging");
  PD_REQUIRE(!h_si_idx || d_si_idx, "si_idx needs device staging");
  PD_REQUIRE(!h_elapsed_us || d_elapsed_us, "elapsed needs device staging");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t n = st->n_envs;
  if (n == 0 || n_steps == 0) return PD_OK;
  pd::HostPipeline* pipe = nullptr;
  int rcode = pd::host_pipeline(&pipe);
  if (rcode != PD_OK) return rcode;
  {
    // small batches: one launch that follows the H2D copy and writes the
    // results back itself (float64 actions, int64 elapsed)
    bool handled = false;
    rcode = pd::streamed_rollout(
        pipe, lat, st, rc, h_controls_xy, 2, action_mode,
        max_distance_angstroms, dwell_us_scalar, n_steps, image_duration_us,
        d_controls_xy, d_si_idx, d_elapsed_us, h_si_idx, h_elapsed_us, false, s,
        &handled);
    if (rcode != PD_OK || handled) return rcode;
  }
  // Otherwise: copy-engine chunks.
  // The call is bound by the H2D copy of the actions (16 B per env-step over
  // PCIe, 47 GB/s with both directions busy: profiles/prof_pcie.py); the
  // kernel and the D2H copy of chunk i hide behind the H2D copy of chunk i+1,
  // what is left over is the last chunk's kernel + D2H copy.  Each chunk costs
  // ~10 us of stream hand-offs, so chunks are 4 MiB (measured at 16.8 MB in:
  // 4 or 8 chunks 0.51 ms, 12: 0.56, 16: 0.60).  PD_HOST_CHUNKS overrides.
  static const int forced_chunks = [] {
    const char* v = getenv("PD_HOST_CHUNKS");
    return v ? atoi(v) : 0;
  }();
  int n_chunks = static_cast<int>(
      static_cast<int64_t>(n_steps) * n * 2 * sizeof(double) >> 22);
  if (forced_chunks > 0) n_chunks = forced_chunks;
  if (n_chunks > 16) n_chunks = 16;
  if (n_chunks > n_steps) n_chunks = n_steps;
  if (n_chunks < 1) n_chunks = 1;
  // staging may still be read by earlier work queued on `s`
  pd::DrainOnError guard{pipe, s};
  PD_CUDA_OK(cudaEventRecord(pipe->start, s));
  PD_CUDA_OK(cudaStreamWaitEvent(pipe->h2d, pipe->start, 0));
  PD_CUDA_OK(cudaStreamWaitEvent(pipe->d2h, pipe->start, 0));
  int t0[17];
  for (int c = 0; c <= n_chunks; ++c)
    t0[c] = static_cast<int>(static_cast<int64_t>(n_steps) * c / n_chunks);
  for (int c = 0; c < n_chunks; ++c) {
    const size_t off = static_cast<size_t>(t0[c]) * n * 2;
    const size_t cnt = static_cast<size_t>(t0[c + 1] - t0[c]) * n * 2;
    PD_CUDA_OK(cudaMemcpyAsync(d_controls_xy + off, h_controls_xy + off,
                               cnt * sizeof(double), cudaMemcpyHostToDevice,
                               pipe->h2d));
    PD_CUDA_OK(cudaEventRecord(pipe->copied[c], pipe->h2d));
  }
  for (int c = 0; c < n_chunks; ++c) {
    const size_t off = static_cast<size_t>(t0[c]) * n;
    const int steps = t0[c + 1] - t0[c];
    PD_CUDA_OK(cudaStreamWaitEvent(s, pipe->copied[c], 0));
    rcode = pd_rollout_actions(
        lat, st, rc, d_controls_xy + off * 2, action_mode,
        max_distance_angstroms, dwell_us_scalar, steps, image_duration_us,
        h_si_idx ? d_si_idx + off : nullptr,
        h_elapsed_us ? d_elapsed_us + off : nullptr, stream);
    if (rcode != PD_OK) return rcode;  // (the guard drains the streams)
    PD_CUDA_OK(cudaEventRecord(pipe->stepped[c], s));
    PD_CUDA_OK(cudaStreamWaitEvent(pipe->d2h, pipe->stepped[c], 0));
    if (h_si_idx)
      PD_CUDA_OK(cudaMemcpyAsync(h_si_idx + off, d_si_idx + off,
                                 static_cast<size_t>(steps) * n * sizeof(int32_t),
                                 cudaMemcpyDeviceToHost, pipe->d2h));
    if (h_elapsed_us)
      PD_CUDA_OK(cudaMemcpyAsync(h_elapsed_us + off, d_elapsed_us + off,
                                 static_cast<size_t>(steps) * n * sizeof(int64_t),
                                 cudaMemcpyDeviceToHost, pipe->d2h));
  }
  PD_CUDA_OK(cudaStreamSynchronize(pipe->d2h));
  PD_CUDA_OK(cudaStreamSynchronize(s));
  guard.armed = false;
  return PD_OK;
}


// ---------------------------------------------------------------------------
// Compact host formats: float32 actions in (the dtype every action adapter
// declares, action_adapters.py:80-84,124-128,202-216), int32 elapsed
// microseconds out.  Both conversions are exact (float32 -> float64 widening;
// dwell + 2 x image duration is checked to fit int32), so results equal
// pd_rollout_actions_host fed the widened actions.  8 + 8 instead of 16 + 12
// bytes per env-step cross PCIe, which is what bounds the call.
// ---------------------------------------------------------------------------
namespace pd {
__global__ void __launch_bounds__(256)
    k_widen_actions(const float2* __restrict__ in, double2* __restrict__ out,
                    int64_t count) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
       i < count; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float2 v = in[i];
    out[i] = make_double2(static_cast<double>(v.x), static_cast<double>(v.y));
  }
}

__global__ void __launch_bounds__(256)
    k_narrow_elapsed(const int64_t* __restrict__ in, int32_t* __restrict__ out,
                     int64_t count) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
       i < count; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = static_cast<int32_t>(in[i]);
}

static int convert_grid(int64_t count) {
  const int64_t want = (count + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  return static_cast<int>(want < 1 ? 1 : (want < cap ? want : cap));
}
}  // namespace pd

extern "C" int pd_rollout_actions_host_f32(
    const pd_lattice* lat, const pd_state* st, const pd_rate_config* rc,
    const float* h_actions_xy, int32_t action_mode,
    double max_distance_angstroms, int64_t dwell_us_scalar, int32_t n_steps,
    int64_t image_duration_us, float* d_actions_f32, double* d_controls_xy,
    int32_t* d_si_idx, int64_t* d_elapsed_us, int32_t* d_elapsed_us32,
    int32_t* h_si_idx, int32_t* h_elapsed_us32, void* stream) {
  PD_REQUIRE(st != nullptr, "null state");
  PD_REQUIRE(n_steps >= 0, "negative n_steps");
  // all five stagings NULL: the library keeps its own
  const bool owned = !d_actions_f32 && !d_controls_xy && !d_si_idx &&
                     !d_elapsed_us && !d_elapsed_us32;
  PD_REQUIRE(n_steps == 0 || h_actions_xy, "null actions");
  PD_REQUIRE(owned || n_steps == 0 || (d_actions_f32 && d_controls_xy),
             "null staging (pass all five stagings or none)");
  PD_REQUIRE(owned || !h_si_idx || d_si_idx, "si_idx needs device staging");
  PD_REQUIRE(owned || !h_elapsed_us32 || (d_elapsed_us && d_elapsed_us32),
             "elapsed needs device staging");
  PD_REQUIRE(dwell_us_scalar >= 0 && image_duration_us >= 0 &&
                 dwell_us_scalar + 2 * image_duration_us < (1LL << 31),
             "per-step elapsed time does not fit int32 microseconds");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t n = st->n_envs;
  if (n == 0 || n_steps == 0) return PD_OK;
  pd::HostPipeline* pipe = nullptr;
  int rcode = pd::host_pipeline(&pipe);
  if (rcode != PD_OK) return rcode;
  if (owned) {
    const size_t items = static_cast<size_t>(n_steps) * n;
    if ((rcode = pd::own_staging(pipe, 0, items * 8)) != PD_OK) return rcode;
    if ((rcode = pd::own_staging(pipe, 2, items * 4)) != PD_OK) return rcode;
    if ((rcode = pd::own_staging(pipe, 4, items * 4)) != PD_OK) return rcode;
    d_actions_f32 = static_cast<float*>(pipe->own[0]);
    d_si_idx = static_cast<int32_t*>(pipe->own[2]);
    d_elapsed_us32 = static_cast<int32_t*>(pipe->own[4]);
  }
  // With 8 + 8 bytes per env-step the copies no longer bound the call: the
  // kernels of the chunks run back to back and what is left over is the fill
  // (H2D copy of the first chunk) and the drain (kernel + D2H copy of the
  // last one) plus ~20 us of stream hand-offs per chunk, so the schedule has
  // few chunks and a short last one (measured at 4096 envs x 256 steps, in
  // sixteenths: "2,5,5,3,1" 0.38 ms, "1,3,4,4,3,1" 0.41, "1,2,3,4,3,2,1" 0.44,
  // one chunk 0.54).  PD_HOST_SCHEDULE overrides.
  static const std::vector<int> schedule = [] {
    std::vector<int> w;
    const char* v = getenv("PD_HOST_SCHEDULE");
    const char* p = v ? v : "2,5,5,3,1";
    while (*p) {
      w.push_back(atoi(p));
      while (*p && *p != ',') ++p;
      if (*p == ',') ++p;
    }
    if (w.empty() || w.size() > 16) w.assign(1, 1);
    return w;
  }();
  {
    bool handled = false;
    rcode = pd::streamed_rollout(
        pipe, lat, st, rc, h_actions_xy, 1, action_mode, max_distance_angstroms,
        dwell_us_scalar, n_steps, image_duration_us, d_actions_f32, d_si_idx,
        d_elapsed_us32, h_si_idx, h_elapsed_us32, owned, s, &handled);
    if (rcode != PD_OK || handled) return rcode;
  }

  // Prior / simple rates: the fast kernels read the float32 actions and write
  // the int32 results themselves; other rate functions go through a widening
  // and a narrowing pass around the float64 rollout.
  const bool direct32 =
      rc && ((rc->rate_fn == PD_RATE_PRIOR && !rc->prior) ||
             rc->rate_fn == PD_RATE_SIMPLE) &&
      pd::fast_enabled() && dwell_us_scalar > 0 &&
      dwell_us_scalar < 3000LL * 1000000LL;
  if (owned) {
    const size_t items = static_cast<size_t>(n_steps) * n;
    PD_CUDA_OK(cudaStreamSynchronize(pipe->d2h));  // a re-fill may be running
    pipe->clean_in = pipe->clean_out = 0;
    if (!direct32) {
      if ((rcode = pd::own_staging(pipe, 1, items * 16)) != PD_OK) return rcode;
      if ((rcode = pd::own_staging(pipe, 3, items * 8)) != PD_OK) return rcode;
      d_controls_xy = static_cast<double*>(pipe->own[1]);
      d_elapsed_us = static_cast<int64_t*>(pipe->own[3]);
    }
  }
  // On an error nothing may stay in flight on the caller's buffers: the
  // guard drains the three streams on every early return (PD_CUDA_OK too).
  pd::DrainOnError guard{pipe, s};
  auto fail = [&](int code) { return code; };
  int total_w = 0;
  for (int wgt : schedule) total_w += wgt > 0 ? wgt : 1;
  int n_chunks = static_cast<int>(schedule.size());
  const bool small = static_cast<int64_t>(n_steps) * n < (1 << 18) ||
                     n_steps < 2 * n_chunks;
  if (small) n_chunks = 1;
  PD_CUDA_OK(cudaEventRecord(pipe->start, s));
  PD_CUDA_OK(cudaStreamWaitEvent(pipe->h2d, pipe->start, 0));
  PD_CUDA_OK(cudaStreamWaitEvent(pipe->d2h, pipe->start, 0));
  int t0[17];
  t0[0] = 0;
  for (int c = 0, acc = 0; c < n_chunks; ++c) {
    acc += schedule[c] > 0 ? schedule[c] : 1;
    t0[c + 1] = small ? n_steps
                      : static_cast<int>(static_cast<int64_t>(n_steps) * acc /
                                         total_w);
  }
  for (int c = 0; c < n_chunks; ++c) {
    const size_t off = static_cast<size_t>(t0[c]) * n * 2;
    const size_t cnt = static_cast<size_t>(t0[c + 1] - t0[c]) * n * 2;
    PD_CUDA_OK(cudaMemcpyAsync(d_actions_f32 + off, h_actions_xy + off,
                               cnt * sizeof(float), cudaMemcpyHostToDevice,
                               pipe->h2d));
    PD_CUDA_OK(cudaEventRecord(pipe->copied[c], pipe->h2d));
  }
  for (int c = 0; c < n_chunks; ++c) {
    const size_t off = static_cast<size_t>(t0[c]) * n;
    const int steps = t0[c + 1] - t0[c];
    const int64_t items = static_cast<int64_t>(steps) * n;
    PD_CUDA_OK(cudaStreamWaitEvent(s, pipe->copied[c], 0));
    if (direct32) {
      rcode = pd::validate_common(lat, st, rc);
      if (rcode != PD_OK) return fail(rcode);
      StepArgs a{};
      a.lat = *lat;
      a.st = *st;
      a.actions_f32 = reinterpret_cast<const float2*>(d_actions_f32) + off;
      a.si_idx_out = h_si_idx ? d_si_idx + off : nullptr;
      a.elapsed32_out = h_elapsed_us32 ? d_elapsed_us32 + off : nullptr;
      a.dwell_us_scalar = dwell_us_scalar;
      a.n_controls = 1;
      a.n_steps = steps;
      a.action_mode = action_mode;
      a.max_distance = max_distance_angstroms;
      a.image_duration_us = image_duration_us;
      rcode = pd::dispatch_step(rc, a, true, s);
      if (rcode != PD_OK) return fail(rcode);
    } else {
      pd::k_widen_actions<<<pd::convert_grid(items), 256, 0, s>>>(
          reinterpret_cast<const float2*>(d_actions_f32) + off,
          reinterpret_cast<double2*>(d_controls_xy) + off, items);
      rcode = pd_rollout_actions(
          lat, st, rc, d_controls_xy + off * 2, action_mode,
          max_distance_angstroms, dwell_us_scalar, steps, image_duration_us,
          h_si_idx ? d_si_idx + off : nullptr,
          h_elapsed_us32 ? d_elapsed_us + off : nullptr, stream);
      if (rcode != PD_OK) return fail(rcode);
      if (h_elapsed_us32)
        pd::k_narrow_elapsed<<<pd::convert_grid(items), 256, 0, s>>>(
            d_elapsed_us + off, d_elapsed_us32 + off, items);
      if (cudaGetLastError() != cudaSuccess)
        return fail(pd::check_cuda(cudaErrorLaunchFailure, "convert kernels"));
    }
    PD_CUDA_OK(cudaEventRecord(pipe->stepped[c], s));
    PD_CUDA_OK(cudaStreamWaitEvent(pipe->d2h, pipe->stepped[c], 0));
    if (h_si_idx)
      PD_CUDA_OK(cudaMemcpyAsync(h_si_idx + off, d_si_idx + off,
                                 static_cast<size_t>(items) * sizeof(int32_t),
                                 cudaMemcpyDeviceToHost, pipe->d2h));
    if (h_elapsed_us32)
      PD_CUDA_OK(cudaMemcpyAsync(h_elapsed_us32 + off, d_elapsed_us32 + off,
                                 static_cast<size_t>(items) * sizeof(int32_t),
                                 cudaMemcpyDeviceToHost, pipe->d2h));
  }
  PD_CUDA_OK(cudaStreamSynchronize(pipe->d2h));
  PD_CUDA_OK(cudaStreamSynchronize(s));
  guard.armed = false;
  return PD_OK;
}

// ---------------------------------------------------------------------------
// Packed host format: float32 actions in, one uint16 per env-step out (Si site
// | re-centred << 15; the elapsed time of a step is dwell + image duration *
// (1 + re-centred), simulator.py:131-169).  8 + 2 instead of 8 + 8 bytes per
// env-step cross PCIe.  Copy-engine pipeline: the action stream goes in as a
// few chunks of whole steps, the fast kernels (pd_step_fast.cu) run chunk c
// while chunk c + 1 is on the wire and the results of chunk c - 1 leave; what
// is left over at the end is the last (short) chunk's launch and its D2H copy.
// ---------------------------------------------------------------------------
extern "C" int pd_rollout_actions_host_packed(
    const pd_lattice* lat, const pd_state* st, const pd_rate_config* rc,
    const float* h_actions_xy, int32_t action_mode,
    double max_distance_angstroms, int64_t dwell_us_scalar, int32_t n_steps,
    int64_t image_duration_us, uint16_t* h_packed, void* stream) {
  PD_REQUIRE(action_mode == PD_ACTION_DIRECT ||
                 action_mode == PD_ACTION_RELATIVE_TO_SILICON,
             "unknown action_mode");
  int rcode = pd::validate_common(lat, st, rc);
  if (rcode != PD_OK) return rcode;
  PD_REQUIRE(rc != nullptr, "null rate config");
  PD_REQUIRE(n_steps >= 0, "negative n_steps");
  PD_REQUIRE(dwell_us_scalar >= 0 && image_duration_us >= 0, "negative time");
  PD_REQUIRE(lat->n_sites <= 32768, "site ids do not fit 15 bits");
  const int64_t n = st->n_envs;
  if (n == 0 || n_steps == 0) return PD_OK;
  PD_REQUIRE(h_actions_xy && h_packed, "null host buffer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  pd::HostPipeline* pipe = nullptr;
  rcode = pd::host_pipeline(&pipe);
  if (rcode != PD_OK) return rcode;
  const size_t items = static_cast<size_t>(n_steps) * n;
  // a re-fill of the streamed form may still be using stagings 0 / 2
  PD_CUDA_OK(cudaStreamSynchronize(pipe->d2h));
  pipe->clean_in = pipe->clean_out = 0;
  if ((rcode = pd::own_staging(pipe, 0, items * 8)) != PD_OK) return rcode;
  if ((rcode = pd::own_staging(pipe, 2, items * 4)) != PD_OK) return rcode;
  float* d_act = static_cast<float*>(pipe->own[0]);
  uint16_t* d_packed = static_cast<uint16_t*>(pipe->own[2]);
  // Chunks in sixteenths of the steps; short last chunk (it is the one whose
  // launch and D2H copy nothing hides).  PD_PACKED_SCHEDULE overrides.
  static const std::vector<int> schedule = [] {
    std::vector<int> w;
    const char* v = getenv("PD_PACKED_SCHEDULE");
    const char* p = v ? v : "5,5,4,2";
    while (*p) {
      w.push_back(atoi(p));
      while (*p && *p != ',') ++p;
      if (*p == ',') ++p;
    }
    if (w.empty() || w.size() > 16) w.assign(1, 1);
    return w;
  }();
  int total_w = 0;
  for (int wgt : schedule) total_w += wgt > 0 ? wgt : 1;
  int n_chunks = static_cast<int>(schedule.size());
  if (items * 8 < (1u << 20) || n_steps < 2 * n_chunks) n_chunks = 1;
  int t0[17];
  t0[0] = 0;
  for (int c = 0, acc = 0; c < n_chunks; ++c) {
    acc += schedule[c] > 0 ? schedule[c] : 1;
    t0[c + 1] = n_chunks == 1
                    ? n_steps
                    : static_cast<int>(static_cast<int64_t>(n_steps) * acc /
                                       total_w);
  }
  t0[n_chunks] = n_steps;
  // (The stagings are the library's and the previous call has synchronised:
  // the copies need not wait for anything queued on `s`.)
  // PD_PACKED_TRACE=1: device-side timeline of the call on stderr
  static const bool trace = getenv("PD_PACKED_TRACE") != nullptr;
  cudaEvent_t tev[40];
  int n_tev = 0;
  const char* tev_name[40];
  auto mark = [&](cudaStream_t q, const char* name) {
    if (!trace || n_tev >= 40) return;
    cudaEventCreate(&tev[n_tev]);
    cudaEventRecord(tev[n_tev], q);
    tev_name[n_tev++] = name;
  };
  const auto cpu_t0 = std::chrono::steady_clock::now();
  mark(pipe->h2d, "h2d begin");
  auto drain = [&] {  // nothing may be in flight on the caller's buffers
    cudaStreamSynchronize(pipe->h2d);
    cudaStreamSynchronize(pipe->d2h);
    cudaStreamSynchronize(s);
  };
  for (int c = 0; c < n_chunks; ++c) {
    const size_t off = static_cast<size_t>(t0[c]) * n;
    const size_t cnt = static_cast<size_t>(t0[c + 1] - t0[c]) * n;
    if (cnt == 0) continue;
    cudaError_t e = cudaMemcpyAsync(d_act + 2 * off, h_actions_xy + 2 * off,
                                    cnt * 2 * sizeof(float),
                                    cudaMemcpyHostToDevice, pipe->h2d);
    if (e == cudaSuccess) e = cudaEventRecord(pipe->copied[c], pipe->h2d);
    mark(pipe->h2d, "h2d chunk done");
    if (e != cudaSuccess) {
      drain();
      PD_CUDA_OK(e);
    }
  }
  for (int c = 0; c < n_chunks; ++c) {
    const size_t off = static_cast<size_t>(t0[c]) * n;
    const int steps = t0[c + 1] - t0[c];
    if (steps == 0) continue;
    StepArgs a{};
    a.lat = *lat;
    a.st = *st;
    a.actions_f32 = reinterpret_cast<const float2*>(d_act) + off;
    a.packed_out = d_packed + off;
    a.dwell_us_scalar = dwell_us_scalar;
    a.n_controls = 1;
    a.n_steps = steps;
    a.action_mode = action_mode;
    a.max_distance = max_distance_angstroms;
    a.image_duration_us = image_duration_us;
    cudaError_t e = cudaStreamWaitEvent(s, pipe->copied[c], 0);
    if (e == cudaSuccess) {
      mark(s, "kernel may start");
      rcode = pd::dispatch_step(rc, a, true, s);
      if (rcode != PD_OK) {
        drain();
        return rcode;
      }
      mark(s, "kernel done");
      // the last chunk's results leave on `s` itself (no stream hand-off
      // on the path nothing hides), the others on the D2H stream
      if (c + 1 < n_chunks) {
        e = cudaEventRecord(pipe->stepped[c], s);
        if (e == cudaSuccess)
          e = cudaStreamWaitEvent(pipe->d2h, pipe->stepped[c], 0);
      }
    }
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(h_packed + off, d_packed + off,
                          static_cast<size_t>(steps) * n * sizeof(uint16_t),
                          cudaMemcpyDeviceToHost,
                          c + 1 < n_chunks ? pipe->d2h : s);
    mark(c + 1 < n_chunks ? pipe->d2h : s, "d2h chunk done");
    if (e != cudaSuccess) {
      drain();
      PD_CUDA_OK(e);
    }
  }
  const auto cpu_t1 = std::chrono::steady_clock::now();
  PD_CUDA_OK(cudaStreamSynchronize(pipe->d2h));
  PD_CUDA_OK(cudaStreamSynchronize(s));
  if (trace) {
    const auto cpu_t2 = std::chrono::steady_clock::now();
    fprintf(stderr, "pd packed trace: enqueue %.1f us, call %.1f us |",
            std::chrono::duration<double, std::micro>(cpu_t1 - cpu_t0).count(),
            std::chrono::duration<double, std::micro>(cpu_t2 - cpu_t0).count());
    for (int k = 0; k < n_tev; ++k) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, tev[0], tev[k]);
      fprintf(stderr, " %s %.1f;", tev_name[k], ms * 1e3);
      if (k > 0) cudaEventDestroy(tev[k]);
    }
    if (n_tev > 0) cudaEventDestroy(tev[0]);
    fprintf(stderr, "\n");
  }
  return PD_OK;
}
