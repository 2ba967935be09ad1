// K2: learned rate model (rate_learning/learn_rates.py:80-99, :925-972).
#include "pd_kmc.cuh"

namespace pd {

int learned_step(const pd_lattice*, const pd_state*, const pd_mlp*,
                 const StepArgs&, bool, cudaStream_t) {
  set_error("PD_RATE_LEARNED stepping is not built yet");
  return PD_ERR_UNSUPPORTED;
}

int learned_rates(const pd_lattice*, const pd_state*, const pd_mlp*,
                  const double*, float*, int32_t*, cudaStream_t) {
  set_error("PD_RATE_LEARNED rates are not built yet");
  return PD_ERR_UNSUPPORTED;
}

}  // namespace pd
