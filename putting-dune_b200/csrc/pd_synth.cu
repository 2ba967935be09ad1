// Synthetic rate-learning datasets on the device.
//
//   rate_learning/data_utils.py:158-303  generate_synthetic_data (PRIOR mode:
//       sample_from_prior :237-283, sample_dataset :288-295)
//   rate_learning/data_utils.py:49-72    get_all_position_rotations,
//       rotate_attributes, rotate_index
//   graphene.py:121-130                  single_silicon_prior_rates
//   constants.py:26-28                   SIGR_PRIOR_RATE_MEAN / COV / MAX_RATE
//
// One thread per sample, float32 like the reference's JAX code.  The
// reference keys its draws by jax.random (threefry), which is not
// reproducible here; this kernel keys them by Philox (stream
// PD_STREAM_SYNTH, counter = (sample, split, slot)):
//   slot 0   position: two standard normals (Box-Muller), scaled by
//            sqrt(1.5 * 0.1) around the prior mean (0.85, 0)
//   slot 1   words x,y -> next-state uniform; words z,w -> rotation uniform
//   slot 2   words x,y -> waiting-time uniform; z,w -> window uniform
//   slot 3+k context normals 2k, 2k+1
// Given the draws, the arithmetic is the reference's (the oracle runs the
// reference's own helper functions on the same draws).
#include <math.h>

#include "pd_common.cuh"

namespace pd {

constexpr int kSynthMaxStates = 8;

__device__ __forceinline__ float2 box_muller(uint4 w) {
  // u1 in (0, 1], u2 in [0, 1)
  const float u1 = static_cast<float>(1.0 - u53(w.x, w.y));
  const float u2 = static_cast<float>(u53(w.z, w.w));
  const float r = sqrtf(-2.0f * logf(fmaxf(u1, 1e-37f)));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

__global__ void __launch_bounds__(128)
    k_synthetic_prior(uint64_t seed, uint32_t split, int64_t n,
                      int32_t num_states, int32_t context_dim, float time_lo,
                      float time_hi, int32_t* __restrict__ next_state,
                      float* __restrict__ dt, float* __restrict__ rates_out,
                      float* __restrict__ context,
                      float* __restrict__ position) {
  const float kMeanX = 0.85f, kMeanY = 0.0f;  // constants.py:26
  const float kMaxRate = 0.23104906f;         // np.log(2) / 3
  const float kPi = 3.14159265358979f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
       i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint32_t id = static_cast<uint32_t>(i);
    // position ~ N(mean, 1.5 * cov), cov = 0.1 I  (data_utils.py:246-250)
    const float2 z = box_muller(philox4x32_10(id, split, 0u, PD_STREAM_SYNTH, seed));
    const float sd = sqrtf(0.15f);
    float px = kMeanX + sd * z.x, py = kMeanY + sd * z.y;
    // rates of the num_states rotations of the position (:252-257):
    // max_rate * pdf(x) / pdf(mean) = max_rate * exp(-|x - mean|^2 / 0.2)
    float r[kSynthMaxStates];
    float total = 0.f;
    for (int k = 0; k < num_states; ++k) {
      float s, c;
      sincosf(2.0f * static_cast<float>(k) * kPi / static_cast<float>(num_states), &s, &c);
      // coord @ [[c, s], [-s, c]]  (geometry.py:81-84)
      const float rx = px * c - py * s, ry = px * s + py * c;
      const float dx = rx - kMeanX, dy = ry - kMeanY;
      r[k] = kMaxRate * expf(-0.5f * (dx * dx + dy * dy) / 0.1f);
      total += r[k];
    }
    const uint4 w1 = philox4x32_10(id, split, 1u, PD_STREAM_SYNTH, seed);
    const float u_state = static_cast<float>(u53(w1.x, w1.y));
    const float u_rot = static_cast<float>(u53(w1.z, w1.w));
    // next_state ~ categorical(rates / total)  (:258-260)
    int state = num_states - 1;
    float cdf = 0.f;
    for (int k = 0; k < num_states; ++k) {
      cdf += r[k] / total;
      if (u_state < cdf) {
        state = k;
        break;
      }
    }
    // random rotation of the whole sample (:262-269)
    int rf = static_cast<int>(u_rot * static_cast<float>(num_states));
    if (rf >= num_states) rf = num_states - 1;
    {
      float s, c;
      sincosf(2.0f * static_cast<float>(rf) * kPi / static_cast<float>(num_states), &s, &c);
      const float nx = px * c - py * s, ny = px * s + py * c;
      px = nx;
      py = ny;
    }
    state = (state + rf) % num_states;
    const uint4 w2 = philox4x32_10(id, split, 2u, PD_STREAM_SYNTH, seed);
    const float u_time = static_cast<float>(1.0 - u53(w2.x, w2.y));  // (0, 1]
    const float u_win = static_cast<float>(u53(w2.z, w2.w));
    const float next_time = -logf(fmaxf(u_time, 1e-37f)) / total;  // :270
    const float actual = time_lo + u_win * (time_hi - time_lo);     // :271-276
    const bool transitioned = next_time < actual;                   // :277
    next_state[i] = transitioned ? state + 1 : 0;                   // :278
    dt[i] = actual;
    for (int k = 0; k < num_states; ++k)  // jnp.roll(rates, rf)
      rates_out[i * num_states + (k + rf) % num_states] = r[k];
    position[2 * i] = px;
    position[2 * i + 1] = py;
    for (int k = 0; k < context_dim; k += 2) {
      const float2 g = box_muller(philox4x32_10(
          id, split, 3u + static_cast<uint32_t>(k / 2), PD_STREAM_SYNTH, seed));
      context[i * context_dim + k] = g.x;
      if (k + 1 < context_dim) context[i * context_dim + k + 1] = g.y;
    }
  }
}

}  // namespace pd

extern "C" int pd_generate_synthetic_data(
    uint64_t seed, int32_t split, int64_t n, int32_t num_states,
    int32_t context_dim, float time_lo, float time_hi, int32_t* next_state,
    float* dt, float* rates, float* context, float* position, void* stream) {
  PD_REQUIRE(n >= 0 && split >= 0, "bad sizes");
  PD_REQUIRE(num_states >= 1 && num_states <= pd::kSynthMaxStates,
             "num_states out of range");
  PD_REQUIRE(context_dim >= 0, "negative context_dim");
  PD_REQUIRE(time_hi >= time_lo, "empty time range");
  if (n == 0) return PD_OK;
  PD_REQUIRE(next_state && dt && rates && position &&
                 (context || context_dim == 0),
             "null outputs");
  const int64_t want = (n + 127) / 128;
  const int64_t cap = static_cast<int64_t>(pd::sm_count()) * 16;
  pd::k_synthetic_prior<<<static_cast<int>(want < cap ? want : cap), 128, 0,
                          static_cast<cudaStream_t>(stream)>>>(
      seed, static_cast<uint32_t>(split), n, num_states, context_dim, time_lo,
      time_hi, next_state, dt, rates, context, position);
  PD_CUDA_OK(cudaGetLastError());
  return PD_OK;
}
