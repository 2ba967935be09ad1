// Synthetic rate-learning datasets on the device.
//
//   rate_learning/data_utils.py:158-303  generate_synthetic_data (PRIOR mode:
//       sample_from_prior :237-283, sample_dataset :288-295; NETWORK mode:
//       sample_network_rates :201-234 with the MLP of learn_rates.py:80-99)
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

// NETWORK mode (data_utils.py:201-234): x ~ N(0, I) of context_dim +
// position_dim entries, rates = softplus(MLP(x))[:num_states] (hk.nets.MLP,
// swish between the layers, no batch norm: learn_rates.py:80-99 with
// batchnorm=False), next state ~ categorical(rates / total), waiting time ~
// Exp(total) against a uniform window.  Weights are the caller's ([in][out]
// row-major like Haiku's).  Philox slots: 1 (x,y) next-state uniform, 2
// waiting-time / window uniforms, 3 + k the normals 2k, 2k + 1 of x.
constexpr int kSynthMaxHidden0 = 64;
constexpr int kSynthMaxInputs = 16;

__device__ __forceinline__ float swishf(float z) { return z / (1.0f + expf(-z)); }
__device__ __forceinline__ float softplusf(float z) {
  // logaddexp(z, 0)
  return fmaxf(z, 0.0f) + log1pf(expf(-fabsf(z)));
}

__global__ void __launch_bounds__(128)
    k_synthetic_network(uint64_t seed, uint32_t split, int64_t n,
                        int32_t num_states, int32_t context_dim,
                        int32_t position_dim, float time_lo, float time_hi,
                        const float* __restrict__ w0, const float* __restrict__ b0,
                        const float* __restrict__ w1, const float* __restrict__ b1,
                        const float* __restrict__ w2, const float* __restrict__ b2,
                        int32_t h0n, int32_t h1n, int32_t* __restrict__ next_state,
                        float* __restrict__ dt, float* __restrict__ rates_out,
                        float* __restrict__ context,
                        float* __restrict__ position) {
  const int d_in = context_dim + position_dim;
  const int n_out = num_states + 1;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
       i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint32_t id = static_cast<uint32_t>(i);
    float x[kSynthMaxInputs];
    for (int k = 0; k < d_in; k += 2) {
      const float2 g = box_muller(philox4x32_10(
          id, split, 3u + static_cast<uint32_t>(k / 2), PD_STREAM_SYNTH, seed));
      x[k] = g.x;
      if (k + 1 < d_in) x[k + 1] = g.y;
    }
    float h0[kSynthMaxHidden0];
    for (int j = 0; j < h0n; ++j) {
      float acc = 0.f;
      for (int k = 0; k < d_in; ++k) acc = fmaf(x[k], __ldg(w0 + k * h0n + j), acc);
      h0[j] = swishf(acc + __ldg(b0 + j));
    }
    float o[kSynthMaxStates + 1];
    for (int k = 0; k < n_out; ++k) o[k] = 0.f;
    for (int j = 0; j < h1n; ++j) {
      float acc = 0.f;
      for (int k = 0; k < h0n; ++k) acc = fmaf(h0[k], __ldg(w1 + k * h1n + j), acc);
      const float hj = swishf(acc + __ldg(b1 + j));
      for (int k = 0; k < n_out; ++k) o[k] = fmaf(hj, __ldg(w2 + j * n_out + k), o[k]);
    }
    float total = 0.f;
    for (int k = 0; k < num_states; ++k) {  // rates[0, :-1]  (:213)
      o[k] = softplusf(o[k] + __ldg(b2 + k));
      total += o[k];
    }
    const uint4 w1d = philox4x32_10(id, split, 1u, PD_STREAM_SYNTH, seed);
    const float u_state = static_cast<float>(u53(w1d.x, w1d.y));
    int state = num_states - 1;
    float cdf = 0.f;
    for (int k = 0; k < num_states; ++k) {
      cdf += o[k] / total;
      if (u_state < cdf) {
        state = k;
        break;
      }
    }
    const uint4 w2d = philox4x32_10(id, split, 2u, PD_STREAM_SYNTH, seed);
    const float u_time = static_cast<float>(1.0 - u53(w2d.x, w2d.y));  // (0, 1]
    const float u_win = static_cast<float>(u53(w2d.z, w2d.w));
    const float next_time = -logf(fmaxf(u_time, 1e-37f)) / total;  // :218
    const float actual = time_lo + u_win * (time_hi - time_lo);     // :219-224
    next_state[i] = next_time < actual ? state + 1 : 0;             // :226-227
    dt[i] = actual;
    for (int k = 0; k < num_states; ++k) rates_out[i * num_states + k] = o[k];
    for (int k = 0; k < context_dim; ++k) context[i * context_dim + k] = x[k];
    for (int k = 0; k < position_dim; ++k)
      position[i * position_dim + k] = x[context_dim + k];
  }
}

}  // namespace pd

extern "C" int pd_generate_synthetic_data_network(
    uint64_t seed, int32_t split, int64_t n, int32_t num_states,
    int32_t context_dim, int32_t position_dim, float time_lo, float time_hi,
    const float* w0, const float* b0, const float* w1, const float* b1,
    const float* w2, const float* b2, int32_t hidden0, int32_t hidden1,
    int32_t* next_state, float* dt, float* rates, float* context,
    float* position, void* stream) {
  PD_REQUIRE(n >= 0 && split >= 0, "bad sizes");
  PD_REQUIRE(num_states >= 1 && num_states <= pd::kSynthMaxStates,
             "num_states out of range");
  PD_REQUIRE(context_dim >= 0 && position_dim >= 0 &&
                 context_dim + position_dim >= 1 &&
                 context_dim + position_dim <= pd::kSynthMaxInputs,
             "context_dim + position_dim out of range");
  PD_REQUIRE(hidden0 >= 1 && hidden0 <= pd::kSynthMaxHidden0 && hidden1 >= 1,
             "hidden sizes out of range");
  PD_REQUIRE(time_hi >= time_lo, "empty time range");
  PD_REQUIRE(w0 && b0 && w1 && b1 && w2 && b2, "null weights");
  if (n == 0) return PD_OK;
  PD_REQUIRE(next_state && dt && rates && (context || context_dim == 0) &&
                 (position || position_dim == 0),
             "null outputs");
  const int64_t want = (n + 127) / 128;
  const int64_t cap = static_cast<int64_t>(pd::sm_count()) * 16;
  pd::k_synthetic_network<<<static_cast<int>(want < cap ? want : cap), 128, 0,
                            static_cast<cudaStream_t>(stream)>>>(
      seed, static_cast<uint32_t>(split), n, num_states, context_dim,
      position_dim, time_lo, time_hi, w0, b0, w1, b1, w2, b2, hidden0, hidden1,
      next_state, dt, rates, context, position);
  PD_CUDA_OK(cudaGetLastError());
  return PD_OK;
}

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
