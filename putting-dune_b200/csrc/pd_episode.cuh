// Device-side pieces of the goal-reaching episode loop (BASELINE configs[4]),
// shared by pd_episode.cu (goal selection) and the stepping kernel k_walk.
//
//   feature_constructors.py:157-228  material-frame features
//   agents/agent_lib.py:163-183      GreedyAgent.step
//   action_adapters.py:219-274       RelativeToSiliconMaterialFrame adapter
//   goals.py:143-181                 SingleSiliconGoalReaching terminal test
#pragma once

#include "pd_kmc.cuh"

namespace pd {

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


// action_adapters.py:163-188 RelativeToSiliconActionAdapter.get_action:
// control = clip(si_observed + clip(a, -1, 1) * max_distance / fov_extent, 0, 1)
__device__ __forceinline__ double2 relative_to_silicon(const Fov4& fov,
                                                       const double2 psi,
                                                       const double2 action,
                                                       double max_distance) {
  const double2 q = observe(fov, psi);
  const double ax = clip_nan(action.x, -1.0, 1.0);
  const double ay = clip_nan(action.y, -1.0, 1.0);
  const double rx = __ddiv_rn(max_distance, __dsub_rn(fov.urx, fov.llx));
  const double ry = __ddiv_rn(max_distance, __dsub_rn(fov.ury, fov.lly));
  return make_double2(
      clip_nan(__dadd_rn(q.x, __dmul_rn(ax, rx)), 0.0, 1.0),
      clip_nan(__dadd_rn(q.y, __dmul_rn(ay, ry)), 0.0, 1.0));
}

// Features -> GreedyAgent.step (float32) -> adapter: the control position in
// the microscope frame for an env whose Si sits at `psi` with neighbours
// `pn`, aiming at `goal` (material frame).
__device__ __forceinline__ double2 greedy_control(const Fov4& fov,
                                                  const double2 psi,
                                                  const double2 pn[3],
                                                  const double2 goal,
                                                  double argmax_x,
                                                  double argmax_y) {
  const double2 si_m = round_trip(fov, psi);
  const float gx = __double2float_rn(__dsub_rn(goal.x, si_m.x));
  const float gy = __double2float_rn(__dsub_rn(goal.y, si_m.y));
  float best_score = 0.f, bdx = 0.f, bdy = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double2 nm = round_trip(fov, pn[i]);
    const float dx = __double2float_rn(__dsub_rn(nm.x, si_m.x));
    const float dy = __double2float_rn(__dsub_rn(nm.y, si_m.y));
    const float ex = __fsub_rn(dx, gx), ey = __fsub_rn(dy, gy);
    const float score =
        __fsqrt_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)));
    if (i == 0 || score < best_score) {  // np.argmin: first minimum
      best_score = score;
      bdx = dx;
      bdy = dy;
    }
  }
  // float32 trigonometry (agent_lib.py:172-181 on float32 features) evaluated
  // in float64 and rounded: float32 library routines differ between libm,
  // NumPy and CUDA by an ulp now and then, the rounded float64 ones do not
  // (the parity tests do the same on the NumPy side)
  const float angle = static_cast<float>(
      atan2(static_cast<double>(bdy), static_cast<double>(bdx)));
  const double c = static_cast<double>(
      static_cast<float>(cos(static_cast<double>(angle))));
  const double s = static_cast<double>(
      static_cast<float>(sin(static_cast<double>(angle))));
  // rotate_coordinates(argmax, angle): (x c - y s, x s + y c)
  const double ax =
      __dadd_rn(__dmul_rn(argmax_x, c), __dmul_rn(argmax_y, -s));
  const double ay = __dadd_rn(__dmul_rn(argmax_x, s), __dmul_rn(argmax_y, c));
  // action_adapters.py:231-256
  const double2 target =
      make_double2(__dadd_rn(si_m.x, ax), __dadd_rn(si_m.y, ay));
  double2 ctl = observe(fov, target);
  ctl.x = fmin(fmax(ctl.x, 0.0), 1.0);
  ctl.y = fmin(fmax(ctl.y, 0.0), 1.0);
  return ctl;
}

// goals.py:160-168: terminal when the observed Si is within half a bond of
// the goal.
__device__ __forceinline__ bool goal_reached(const Fov4& fov, const double2 psi,
                                             const double2 goal) {
  const double2 now_m = round_trip(fov, psi);
  const double dx = __dsub_rn(now_m.x, goal.x);
  const double dy = __dsub_rn(now_m.y, goal.y);
  return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))) <
         kBond * 0.5;
}

}  // namespace pd
