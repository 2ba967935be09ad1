// K2: learned rate model as the rate function of the event step.
//
//   rate_learning/learn_rates.py:80-99    get_mlp_fn (eval): BatchNorm(EMA) ->
//                                         Linear(D,H1) swish Linear(H1,H2)
//                                         swish Linear(H2,4) -> softplus
//   rate_learning/learn_rates.py:925-972  LearnedTransitionRatePredictor.predict
//   rate_learning/data_utils.py:389-432   standardize_beam_and_neighbors
//   rate_learning/learn_rates.py:704-732  apply_model (softmax(o[:3]) * o[3])
//
// The KMC loop is data dependent (1..7+ rate evaluations per control), and one
// evaluation is 2*(D*H1 + H1*H2 + 4*H2) flops (134k at H=256), so the step is
// organised around the contraction: a persistent CTA keeps a work queue of
// (env, control, iteration) items and evaluates the network for 128 items at
// a time as a register-tiled FP32 GEMM (the hidden layer-1 activations are a
// function of only two inputs and are recomputed per k-chunk instead of being
// stored).  Survivors of an iteration go back into the queue and are topped
// up with fresh environments, so every GEMM wave is full until the tail.
//
// FP32 FMA throughout (the parity path of BASELINE.json's north_star).
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "pd_kmc.cuh"

namespace pd {

// -DPD_MLP_PHASE_CLOCKS: thread 0 of every CTA accumulates the cycles between
// phase marks of k_step_learned; pd_debug_mlp_phases() reads and clears them.
#ifdef PD_MLP_PHASE_CLOCKS
__device__ unsigned long long g_mlp_phase[16];
#define PD_MLP_PHASE(i)                                         \
  do {                                                          \
    if (threadIdx.x == 0) {                                     \
      const long long now_ = clock64();                         \
      atomicAdd(&g_mlp_phase[i],                                \
                static_cast<unsigned long long>(now_ - ph_t0)); \
      ph_t0 = now_;                                             \
    }                                                           \
  } while (0)
// ... and inside the tensor-core wave (slots 8..11: operand generation, MMA
// issue + wait, epilogue, head reduction)
#define PD_MLP_SUB(i)                                           \
  do {                                                          \
    if (threadIdx.x == 0) {                                     \
      const long long now_ = clock64();                         \
      atomicAdd(&g_mlp_phase[i],                                \
                static_cast<unsigned long long>(now_ - sub_t0)); \
      sub_t0 = now_;                                            \
    }                                                           \
  } while (0)
// ... and inside the event phase (slots 12..15: rates + Philox, event draw,
// hop bookkeeping, finalise), on warp 0's own timeline
#define PD_MLP_EV(i)                                            \
  do {                                                          \
    if (threadIdx.x == 0) {                                     \
      const long long now_ = clock64();                         \
      atomicAdd(&g_mlp_phase[i],                                \
                static_cast<unsigned long long>(now_ - ev_t0)); \
      ev_t0 = now_;                                             \
    }                                                           \
  } while (0)
// ... and inside the item build (slots 4..7: state loads + dwell loop, key
// hand-over barrier, positions of the Si and its neighbours, canonicalise)
#define PD_MLP_BLD(i)                                            \
  do {                                                           \
    if (threadIdx.x == 0) {                                      \
      const long long now_ = clock64();                          \
      atomicAdd(&g_mlp_phase[i],                                 \
                static_cast<unsigned long long>(now_ - bld_t0)); \
      bld_t0 = now_;                                             \
    }                                                            \
  } while (0)
#else
#define PD_MLP_BLD(i) do { } while (0)
#define PD_MLP_PHASE(i) do { } while (0)
#define PD_MLP_SUB(i) do { } while (0)
#define PD_MLP_EV(i) do { } while (0)
#endif

constexpr int kMlpThreads = 512;
constexpr int kMlpBatch = 128;   // GEMM M tile = queue batch
constexpr int kChunk = 16;       // k-chunk of the hidden contraction
constexpr int kEnvPerThread = 4;  // 32 row groups x 16 column groups

struct MlpView {
  int d, h1, h2, batchnorm;
  const float *bn_scale, *bn_offset, *bn_mean, *bn_var;
  const float *w0, *b0, *w1, *b1, *w2, *b2;
  int tensor_core;
  const void* w1_umma;
  int tc_k_phases;  // tensor path: K handled in this many parts per wave (1 =
                    // W1 resident in shared memory; 2 = its tiles streamed
                    // from L2, half of K at a time)
};

__device__ __forceinline__ float swishf(float z) {
  return z / (1.0f + expf(-z));
}

// Tensor-core paths only: the activations use the SFU approximations (ex2 /
// rcp, ~2^-21 relative) as z / (1 + 2^(-z log2 e)); the library forms of
// __expf / __fdividef carry range fix-ups worth another six instructions.
// Two at a time with one reciprocal: 1/a = b / (a b), 1/b = a / (a b).  The
// wave is bound by the MUFU unit (ex2 + rcp per activation, a quarter-rate
// pipe); this trades one MUFU per pair for three multiplies.  The exponent is
// capped at 60 so that the product stays finite: below z = -41.6 the result
// is z 2^-60 instead of z 2^(z log2 e), both zero at float32 resolution of
// anything they are added to.
__device__ __forceinline__ void swish_fast2(float z0, float z1, float* s0,
                                            float* s1) {
  float e0, e1, r;
  asm("ex2.approx.ftz.f32 %0, %1;"
      : "=f"(e0) : "f"(fminf(z0 * -1.4426950408889634f, 60.f)));
  asm("ex2.approx.ftz.f32 %0, %1;"
      : "=f"(e1) : "f"(fminf(z1 * -1.4426950408889634f, 60.f)));
  const float a = 1.0f + e0, b = 1.0f + e1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a * b));
  *s0 = z0 * (r * b);
  *s1 = z1 * (r * a);
}

__device__ __forceinline__ float softplusf(float z) {
  // np.logaddexp(z, 0)
  return fmaxf(z, 0.f) + log1pf(expf(-fabsf(z)));
}

struct MlpSmall {
  float xs[kMlpBatch][2];             // network inputs (before BatchNorm)
  float out[kMlpBatch][4];            // softplus heads
  float w0[2][256];
  float b0[256];
  float b1[256];
  float w2[256][4];
  float b2[4];
  float bn_a[2], bn_b[2];             // x_hat = x * a + b
};

struct MlpShared : MlpSmall {         // FP32 FMA path (double-buffered)
  __align__(16) float a[2][kChunk][kMlpBatch];  // layer-1 activations, k-major
  __align__(16) float b[2][kChunk][256];        // W1 chunk
};

// Loads the small layers once per CTA.
__device__ __forceinline__ void mlp_stage_small(const MlpView& w,
                                                MlpSmall& sh) {
  for (int i = threadIdx.x; i < w.h1; i += blockDim.x) {
    sh.w0[0][i] = w.w0[i];
    sh.w0[1][i] = w.w0[w.h1 + i];
    sh.b0[i] = w.b0[i];
  }
  for (int i = threadIdx.x; i < w.h2; i += blockDim.x) {
    sh.b1[i] = w.b1[i];
    sh.w2[i][0] = w.w2[4 * i + 0];
    sh.w2[i][1] = w.w2[4 * i + 1];
    sh.w2[i][2] = w.w2[4 * i + 2];
    sh.w2[i][3] = w.w2[4 * i + 3];
  }
  if (threadIdx.x < 4) sh.b2[threadIdx.x] = w.b2[threadIdx.x];
  if (threadIdx.x < 2) {
    const int i = threadIdx.x;
    if (w.batchnorm) {
      // hk.BatchNorm eval: (x - mean) * rsqrt(var + 1e-5) * scale + offset
      const float inv = 1.0f / sqrtf(w.bn_var[i] + 1e-5f);
      const float g = inv * w.bn_scale[i];
      sh.bn_a[i] = g;
      sh.bn_b[i] = w.bn_offset[i] - w.bn_mean[i] * g;
    } else {
      sh.bn_a[i] = 1.0f;
      sh.bn_b[i] = 0.0f;
    }
  }
  __syncthreads();
}

// One GEMM wave: out[b][0..3] = softplus(MLP(xs[b])) for b < kMlpBatch.
// NPT = H2 / 16 outputs per thread; thread (ty, tx) owns envs ty*4..ty*4+3 and
// the NPT / VEC column vectors (j * 16 + tx) * VEC .. + VEC - 1, so that the 16
// tx lanes read consecutive 16-byte words of the W1 chunk.  Chunks of 16 rows
// of W1 arrive with cp.async into the buffer the previous chunk has left
// while the current one is being contracted.
template <int NPT>
__device__ __forceinline__ void mlp_stage_chunk(const MlpView& w, MlpShared& sh,
                                                int k0, int buf) {
  const int tid = threadIdx.x;
  // layer-1 chunk: a[k][b] = swish(x0*W0[0][k] + x1*W0[1][k] + b0[k])
  for (int i = tid; i < kChunk * kMlpBatch; i += kMlpThreads) {
    const int k = i / kMlpBatch, b = i - k * kMlpBatch;
    const float x0 = sh.xs[b][0] * sh.bn_a[0] + sh.bn_b[0];
    const float x1 = sh.xs[b][1] * sh.bn_a[1] + sh.bn_b[1];
    const int kk = k0 + k;
    float z = sh.b0[kk];
    z = fmaf(x0, sh.w0[0][kk], z);
    z = fmaf(x1, sh.w0[1][kk], z);
    sh.a[buf][k][b] = swishf(z);
  }
  // W1 chunk: rows k0..k0+15, all H2 columns, 16 bytes per cp.async
  for (int i = tid; i < kChunk * (NPT * 16) / 4; i += kMlpThreads) {
    const int k = i / (NPT * 4), c4 = i - k * (NPT * 4);
    const float* src = w.w1 + static_cast<size_t>(k0 + k) * w.h2 + 4 * c4;
    const uint32_t dst = static_cast<uint32_t>(
        __cvta_generic_to_shared(&sh.b[buf][k][4 * c4]));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst),
                 "l"(src)
                 : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int NPT>
__device__ __forceinline__ void mlp_wave(const MlpView& w, MlpShared& sh) {
  constexpr int VEC = NPT < 4 ? NPT : 4;
  constexpr int NV = NPT / VEC;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[kEnvPerThread][NPT];
#pragma unroll
  for (int i = 0; i < kEnvPerThread; ++i)
#pragma unroll
    for (int j = 0; j < NPT; ++j) acc[i][j] = 0.f;

  __syncthreads();  // xs of this wave are in place; the last wave is done
  mlp_stage_chunk<NPT>(w, sh, 0, 0);
  const int n_chunks = w.h1 / kChunk;
  for (int c = 0; c < n_chunks; ++c) {
    const int buf = c & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // chunk c visible; everyone has left chunk c - 1
    if (c + 1 < n_chunks) mlp_stage_chunk<NPT>(w, sh, (c + 1) * kChunk, buf ^ 1);
#pragma unroll
    for (int k = 0; k < kChunk; ++k) {
      float bv[NPT];
      const float4 a0 =
          *reinterpret_cast<const float4*>(&sh.a[buf][k][ty * kEnvPerThread]);
      const float av[kEnvPerThread] = {a0.x, a0.y, a0.z, a0.w};
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float* src = &sh.b[buf][k][(j * 16 + tx) * VEC];
        if constexpr (VEC == 4) {
          const float4 v = *reinterpret_cast<const float4*>(src);
          bv[4 * j] = v.x; bv[4 * j + 1] = v.y;
          bv[4 * j + 2] = v.z; bv[4 * j + 3] = v.w;
        } else {
          const float2 v = *reinterpret_cast<const float2*>(src);
          bv[2 * j] = v.x; bv[2 * j + 1] = v.y;
        }
      }
#pragma unroll
      for (int i = 0; i < kEnvPerThread; ++i)
#pragma unroll
        for (int j = 0; j < NPT; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  // layer 2 activation, layer 3 partial products, reduce over the 16 tx lanes
#pragma unroll
  for (int i = 0; i < kEnvPerThread; ++i) {
    float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      const int col = ((j / VEC) * 16 + tx) * VEC + (j % VEC);
      const float h = swishf(acc[i][j] + sh.b1[col]);
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q] = fmaf(h, sh.w2[col][q], o[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int s = 8; s > 0; s >>= 1)
        o[q] += __shfl_xor_sync(0xffffffffu, o[q], s);
    }
    if (tx == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        sh.out[ty * kEnvPerThread + i][q] = softplusf(o[q] + sh.b2[q]);
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// Tensor-core wave (pd_mlp.tensor_core): the hidden contraction
//   acc[128 envs][H2] = h1[128][H1] (bf16) x W1[H1][H2] (bf16), FP32 in TMEM
// as H1/16 tcgen05.mma (cta_group::1, kind::f16, M = 128, N = H2, K = 16)
// issued by one thread.  Operands sit in shared memory in the canonical
// no-swizzle K-major layout (8-row x 16-byte core matrices, LBO = 128 B
// between the core matrices of one K step, SBO = H1 * 16 B between 8-row
// groups); W1 is copied there once per CTA, h1 is regenerated from the two
// network inputs every wave.  The epilogue reads the accumulator with
// tcgen05.ld (thread = TMEM lane = env), applies bias + swish, contracts with
// W2 in registers and finishes with softplus.
//
// pd_mlp.tensor_core == 2, the split form: both operands as fp16 hi + fp16 lo
// (x = hi + lo to ~2^-22 relative) and three MMAs per K step into the same
// accumulator, hi hi + hi lo + lo hi (the lo lo term is below 2^-22); the
// activations keep the SFU forms (ex2 / rcp, ~2^-21 relative).  The rates
// then agree with the FP32 path to ~1e-6 of the largest rate (bf16, one MMA:
// 2e-3), i.e. inside the FP32 path's own tolerance against the oracle (2e-5).  Twice
// the operand tiles: H1 (128 + H2) 4 bytes of shared memory, which H = 128
// fits and H = 256 does not: there a wave takes K in two halves, copying the
// W1 tiles of a half from global memory (L2: 256 KB per wave and CTA, ~1.5 TB/s
// over the device) and regenerating h1 for it, with the MMAs of the second
// half accumulating onto the first.
// ---------------------------------------------------------------------------
struct TcShared {
  unsigned long long mbar;
  uint32_t tmem_base;
  uint32_t pad_;
  float partial[4][kMlpBatch][4];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr,
                                              uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);         // [0,14)
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;   // [16,30)
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;   // [32,46)
  d |= static_cast<uint64_t>(1) << 46;                            // version
  return d;  // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

__device__ __forceinline__ uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4)                                  // D format F32
         | (1u << 7) | (1u << 10)                   // A, B = BF16
         | (static_cast<uint32_t>(n >> 3) << 17)    // N
         | (static_cast<uint32_t>(m >> 4) << 24);   // M; K-major A and B
}

__device__ __forceinline__ uint32_t umma_idesc_f16(int m, int n) {
  return (1u << 4)                                  // D format F32; A, B = F16
         | (static_cast<uint32_t>(n >> 3) << 17)    // N
         | (static_cast<uint32_t>(m >> 4) << 24);   // M; K-major A and B
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// x = hi + lo in fp16: (a, b) -> the packed hi pair and the packed lo pair.
// Packed conversions (F2FP on the ALU pipe): the scalar cvt.rn.f16.f32 is an
// F2F on the same quarter-rate unit as the wave's ex2 / rcp.
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t* hi,
                                            uint32_t* lo) {
  uint32_t h;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));  // a: low
  const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&h));
  uint32_t l;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(b - back.y),
      "f"(a - back.x));
  *hi = h;
  *lo = l;
}

struct TcCtx {
  unsigned char* a_tile;   // shared, [128][H1] bf16 / fp16 canonical
                           // (split: the lo tile follows, a_bytes further on)
  uint32_t a_addr, b_addr; // shared-window addresses (split: lo tiles follow)
  uint32_t a_bytes, b_bytes;  // size of one A / B tile
  TcShared* ts;
  uint32_t phase;
  int b_part;  // streamed W1: the part of K whose tiles are in shared memory
};

__device__ __forceinline__ void tc_setup(const MlpView& w, TcCtx& tc,
                                         unsigned char* a_tile,
                                         unsigned char* b_tile, TcShared* ts) {
  tc.a_tile = a_tile;
  tc.a_addr = smem_u32(a_tile);
  tc.b_addr = smem_u32(b_tile);
  tc.ts = ts;
  tc.phase = 0;
  tc.b_part = -1;
  tc.a_bytes = static_cast<uint32_t>(kMlpBatch) * w.h1 * 2 / w.tc_k_phases;
  tc.b_bytes = static_cast<uint32_t>(w.h2) * w.h1 * 2 / w.tc_k_phases;
  // W1^T in UMMA layout (split: hi tile, then lo tile): verbatim 16-byte
  // copies (when it is streamed, mlp_wave_tc copies it part by part)
  const int n16 = w.tc_k_phases > 1
                      ? 0
                      : w.h1 * w.h2 * 2 / 16 * (w.tensor_core == 2 ? 2 : 1);
  const uint4* src = reinterpret_cast<const uint4*>(w.w1_umma);
  uint4* dst = reinterpret_cast<uint4*>(b_tile);
  for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(
                     smem_u32(&ts->mbar)),
                 "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    int cols = 32;
    while (cols < w.h2) cols <<= 1;
    asm volatile(
        "tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::
            "r"(smem_u32(&ts->tmem_base)),
        "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
}

__device__ __forceinline__ void tc_teardown(const MlpView& w, TcCtx& tc) {
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) {
    int cols = 32;
    while (cols < w.h2) cols <<= 1;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(
                     tc.ts->tmem_base),
                 "r"(cols));
  }
}

template <int THREADS>
__device__ __forceinline__ void mlp_wave_tc(const MlpView& w, MlpSmall& sh,
                                            TcCtx& tc) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // 512 threads (one CTA per SM) or 256 (two: see learned_step)
  constexpr int n_threads = THREADS, n_warps = THREADS >> 5;
  const int kblocks_all = w.h1 >> 3;      // 16-byte blocks along K
  const int kblocks = kblocks_all / w.tc_k_phases;  // ... of one part of K
  const uint32_t sbo = static_cast<uint32_t>(kblocks) * 128u;
  const bool split = w.tensor_core == 2;
  const uint32_t tmem = tc.ts->tmem_base;
#ifdef PD_MLP_PHASE_CLOCKS
  long long sub_t0 = clock64();
#endif
  // (streamed W1: a wave starts with the part of K the previous wave ended
  // with, whose tiles are still there -- one copy per wave instead of two)
  const int first_part = tc.b_part >= 0 ? tc.b_part : 0;
  for (int pi = 0; pi < w.tc_k_phases; ++pi) {
    const int ph = (first_part + pi) % w.tc_k_phases;
    if (w.tc_k_phases > 1 && tc.b_part != ph) {
      // ---- this part's W1 tiles (hi, then lo), from global memory: per
      // 8-row group the part's k blocks are contiguous ----
      const int per_group = kblocks * 8;  // 16-byte units
      const int n_units = (w.h2 >> 3) * per_group;
      for (int t = 0; t < (split ? 2 : 1); ++t) {
        const uint4* src = reinterpret_cast<const uint4*>(w.w1_umma) +
                           static_cast<size_t>(t) * w.h1 * w.h2 * 2 / 16;
        uint4* dst = reinterpret_cast<uint4*>(
            tc.a_tile + (split ? 2 : 1) * tc.a_bytes + t * tc.b_bytes);
        const int pg_shift = 31 - __clz(per_group);
        const bool pg_pow2 = (per_group & (per_group - 1)) == 0;
        for (int i = tid; i < n_units; i += n_threads) {
          const int g = pg_pow2 ? i >> pg_shift : i / per_group;
          const int j = i - g * per_group;
          // (cp.async: the copy lands while this thread generates h1 below)
          asm volatile(
              "cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                  smem_u32(dst + i)),
              "l"(src + (g * kblocks_all + ph * kblocks) * 8 + j)
              : "memory");
        }
      }
      tc.b_part = ph;
    }
    // ---- h1 = swish(x_hat W0 + b0) as bf16 / fp16, canonical K-major ----
    // A warp writes 8 rows x 4 k-blocks (512 contiguous bytes) per trip; with
    // kblocks / 4 dividing the warp count a thread keeps its 8 columns of W0 /
    // b0 in registers for the whole wave.
    {
      const int m8 = lane & 7, kb_lo = lane >> 3;
      const int wcols = kblocks >> 2;  // warp columns
      const int n_cells = 16 * wcols;
      int cur_wc = -1;
      float w0a[8], w0b[8], b0v[8];
      const int wc_shift = 31 - __clz(wcols);
      const bool wc_pow2 = (wcols & (wcols - 1)) == 0;
      for (int cell = warp; cell < n_cells; cell += n_warps) {
        const int g = wc_pow2 ? cell >> wc_shift : cell / wcols;
        const int wc = cell - g * wcols;
        const int kb = wc * 4 + kb_lo;           // within this part of K
        const int kcol = (ph * kblocks + kb) * 8;  // column of W0 / b0
        if (wc != cur_wc) {
          cur_wc = wc;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            w0a[j] = sh.w0[0][kcol + j];
            w0b[j] = sh.w0[1][kcol + j];
            b0v[j] = sh.b0[kcol + j];
          }
        }
        const int m = g * 8 + m8;
        const float2 xr = *reinterpret_cast<const float2*>(sh.xs[m]);
        const float x0 = xr.x * sh.bn_a[0] + sh.bn_b[0];
        const float x1 = xr.y * sh.bn_a[1] + sh.bn_b[1];
        float h[8];
        unsigned char* cell_p =
            tc.a_tile + static_cast<size_t>(g) * sbo + kb * 128 + m8 * 16;
        if (split) {
#pragma unroll
          for (int j = 0; j < 8; j += 2)
            swish_fast2(fmaf(x1, w0b[j], fmaf(x0, w0a[j], b0v[j])),
                        fmaf(x1, w0b[j + 1], fmaf(x0, w0a[j + 1], b0v[j + 1])),
                        &h[j], &h[j + 1]);
          uint4 vh, vl;
          split_f16x2(h[0], h[1], &vh.x, &vl.x);
          split_f16x2(h[2], h[3], &vh.y, &vl.y);
          split_f16x2(h[4], h[5], &vh.z, &vl.z);
          split_f16x2(h[6], h[7], &vh.w, &vl.w);
          *reinterpret_cast<uint4*>(cell_p) = vh;
          *reinterpret_cast<uint4*>(cell_p + tc.a_bytes) = vl;
        } else {
#pragma unroll
          for (int j = 0; j < 8; j += 2)
            swish_fast2(fmaf(x1, w0b[j], fmaf(x0, w0a[j], b0v[j])),
                        fmaf(x1, w0b[j + 1], fmaf(x0, w0a[j + 1], b0v[j + 1])),
                        &h[j], &h[j + 1]);
          uint4 v;
          v.x = pack_bf16x2(h[0], h[1]);
          v.y = pack_bf16x2(h[2], h[3]);
          v.z = pack_bf16x2(h[4], h[5]);
          v.w = pack_bf16x2(h[6], h[7]);
          *reinterpret_cast<uint4*>(cell_p) = v;
        }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    // generic-proxy writes -> visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    PD_MLP_SUB(8);
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t idesc = split ? umma_idesc_f16(kMlpBatch, w.h2)
                                   : umma_idesc_bf16(kMlpBatch, w.h2);
      auto mma = [&](uint32_t a_addr, uint32_t b_addr, uint32_t accumulate) {
        const uint64_t da = umma_desc(a_addr, 128u, sbo);
        const uint64_t db = umma_desc(b_addr, 128u, sbo);
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}\n" ::"r"(tmem),
            "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
      };
      for (int kk = 0; kk < (kblocks >> 1); ++kk) {
        const uint32_t a_hi = tc.a_addr + kk * 256;
        const uint32_t b_hi = tc.b_addr + kk * 256;
        mma(a_hi, b_hi, (pi > 0 || kk > 0) ? 1u : 0u);
        if (split) {  // + hi lo + lo hi
          mma(a_hi, b_hi + tc.b_bytes, 1u);
          mma(a_hi + tc.a_bytes, b_hi, 1u);
        }
      }
      // arrives on the mbarrier when every MMA above has completed
      asm volatile(
          "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster."
          "b64 [%0];" ::"r"(smem_u32(&tc.ts->mbar)));
    }
    // ---- wait for the MMAs (the accumulator; the tiles are free again) ----
    {
      const uint32_t bar = smem_u32(&tc.ts->mbar);
      uint32_t done = 0;
      // try_wait suspends for a hardware-defined interval; the spin bound
      // turns a lost completion into a trap instead of a hung GPU.
      for (int spin = 0; !done && spin < (1 << 24); ++spin) {
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(tc.phase)
            : "memory");
      }
      if (!done) __trap();
      tc.phase ^= 1u;
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    PD_MLP_SUB(9);
  }
  // ---- epilogue: thread = TMEM lane = env row; up to four column slices,
  // dealt to the groups of four warps (the slices, and with them the order of
  // the float32 sums, are the same for 256 and 512 threads) ----
  const int n_slices = (w.h2 >> 4) < 4 ? (w.h2 >> 4) : 4;
  {
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    auto run_slice = [&](const int slice) {
      const int per = w.h2 / n_slices;
      const int c_lo = slice * per, c_hi = c_lo + per;
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
        uint32_t r[16];
        const uint32_t taddr =
            tmem + (static_cast<uint32_t>(quarter * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
            "%14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]),
              "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
              "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const int col = c0 + j;
          float hh[2];
          swish_fast2(__uint_as_float(r[j]) + sh.b1[col],
                      __uint_as_float(r[j + 1]) + sh.b1[col + 1], &hh[0],
                      &hh[1]);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const float4 w2v =
                *reinterpret_cast<const float4*>(sh.w2[col + t]);
            o[0] = fmaf(hh[t], w2v.x, o[0]);
            o[1] = fmaf(hh[t], w2v.y, o[1]);
            o[2] = fmaf(hh[t], w2v.z, o[2]);
            o[3] = fmaf(hh[t], w2v.w, o[3]);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) tc.ts->partial[slice][m][q] = o[q];
    };
    if constexpr (n_warps >= 16) {
      if ((warp >> 2) < n_slices) run_slice(warp >> 2);
    } else {
      for (int slice = warp >> 2; slice < n_slices; slice += n_warps >> 2)
        run_slice(slice);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  PD_MLP_SUB(10);
  // (128 envs x 4 heads over the CTA's threads: one or two per thread)
  for (int i = tid; i < kMlpBatch * 4; i += n_threads) {
    const int m = i >> 2, q = i & 3;
    float acc = sh.b2[q];
    for (int sl = 0; sl < n_slices; ++sl) acc += tc.ts->partial[sl][m][q];
    sh.out[m][q] = softplusf(acc);
  }
  __syncthreads();
  PD_MLP_SUB(11);
}

// Shared-memory carve-up of the tensor-core kernels (dynamic, 1 KB aligned).
// (tiles = 2 in the split form: hi and lo)
__host__ __device__ inline size_t tc_a_bytes(int h1, int tiles = 1) {
  return static_cast<size_t>(kMlpBatch) * h1 * 2 * tiles;
}
__host__ __device__ inline size_t tc_b_bytes(int h1, int h2, int tiles = 1) {
  return static_cast<size_t>(h2) * h1 * 2 * tiles;
}

// predict()'s frame canonicalisation for one env (float64 like the
// reference): returns the network input and, for each neighbour slot, which
// network head holds its rate.
struct Canonical {
  float x0, x1;
  int head[3];
};

// The reference's own formulation (atan2 / sincos / fmod), kept for geometries
// where the shortcut below cannot order the neighbours safely.
__device__ __noinline__ Canonical canonicalise_by_angles(
    const double bx, const double by, const double nx0, const double ny0,
    const double nx1, const double ny1, const double nx2, const double ny2,
    const int k) {
  const double kTwoPi = 6.283185307179586;
  const double nx[3] = {nx0, nx1, nx2}, ny[3] = {ny0, ny1, ny2};
  double ang[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) ang[i] = atan2(ny[i], nx[i]);
  const double rot = -(k == 0 ? ang[0] : (k == 1 ? ang[1] : ang[2]));
  double s, c;
  sincos(rot, &s, &c);
  Canonical out;
  // rotate_coordinates(beam, rot): (x c - y s, x s + y c)
  out.x0 = static_cast<float>(bx * c - by * s);
  out.x1 = static_cast<float>(bx * s + by * c);
  // order = argsort((ang + rot) % 2pi); rates = heads[argsort(order)]
  double pos[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double m = fmod(ang[i] + rot, kTwoPi);
    if (m < 0.0) m += kTwoPi;  // NumPy remainder takes the divisor's sign
    pos[i] = m;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j)
      rank += (pos[j] < pos[i]) || (pos[j] == pos[i] && j < i);
    out.head[i] = rank;
  }
  return out;
}

// The rotation by -atan2(n_k) is the unit vector of n_k itself, and the
// argsort of the rotated angles only asks on which side of n_k the other two
// neighbours lie; on a honeycomb lattice they sit near +-120 degrees, so two
// cross products decide it.  cos/sin taken this way differ from
// sincos(atan2(.)) by an ulp or two of float64, i.e. the float32 network
// input is the same number except on a rounding tie (~1e-8 of the inputs);
// when the two neighbours are not clearly on opposite sides the angle form
// above runs instead.
__device__ __forceinline__ Canonical canonicalise(const double2 beam,
                                                  const double2 psi,
                                                  const double2 pn[3]) {
  double nx[3], ny[3];
  // beam in bond lengths, neighbours in angstroms (learn_rates.py:952-955:
  // the reference's unit mix).
  const double bx = (beam.x - psi.x) / kBond, by = (beam.y - psi.y) / kBond;
  // nearest neighbour of the beam: first index of the smallest distance.
  // Squared distances order like their roots unless two of them are within
  // rounding of each other (the roots could then round to the same number,
  // which the strict comparison resolves by index): the roots decide there.
  int k = 0;
  double best = 0.0, d2[3];
  bool close_call = false;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    nx[i] = pn[i].x - psi.x;
    ny[i] = pn[i].y - psi.y;
    const double dx = nx[i] - bx, dy = ny[i] - by;
    d2[i] = dx * dx + dy * dy;
    if (i > 0) close_call |= fabs(d2[i] - best) <= 1e-15 * best;
    if (i == 0 || d2[i] < best) {
      best = d2[i];
      k = i;
    }
  }
  if (close_call) {
    k = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double dist = sqrt(d2[i]);
      if (i == 0 || dist < best) {
        best = dist;
        k = i;
      }
    }
  }
  const double kx = k == 0 ? nx[0] : (k == 1 ? nx[1] : nx[2]);
  const double ky = k == 0 ? ny[0] : (k == 1 ? ny[1] : ny[2]);
  // the other two, in index order
  const double px = k == 0 ? nx[1] : nx[0], py = k == 0 ? ny[1] : ny[0];
  const double qx = k == 2 ? nx[1] : nx[2], qy = k == 2 ? ny[1] : ny[2];
  const double len2 = kx * kx + ky * ky;
  const double cp = kx * py - ky * px, cq = kx * qy - ky * qx;
  // sin(120 deg) |n|^2 = 0.87 len2 on the undistorted lattice
  const double margin = 0.05 * len2;
  const bool p_first = cp > margin && cq < -margin;
  const bool q_first = cq > margin && cp < -margin;
  if (!(p_first || q_first) || !(len2 > 0.0))
    return canonicalise_by_angles(bx, by, nx[0], ny[0], nx[1], ny[1], nx[2],
                                  ny[2], k);
  const double inv = 1.0 / sqrt(len2);
  const double c = kx * inv, s = -ky * inv;
  Canonical out;
  out.x0 = static_cast<float>(bx * c - by * s);
  out.x1 = static_cast<float>(bx * s + by * c);
  const int hp = p_first ? 1 : 2, hq = p_first ? 2 : 1;
  out.head[0] = k == 0 ? 0 : hp;                   // p is index 0 unless k == 0
  out.head[1] = k == 1 ? 0 : (k == 0 ? hp : hq);   // index 1: p if k==0, q if k==2
  out.head[2] = k == 2 ? 0 : hq;                   // q is index 2 unless k == 2
  return out;
}

template <bool TC>
struct MlpStepSharedT {
  typename std::conditional<TC, MlpSmall, MlpShared>::type m;
  int q_env[kMlpBatch], q_ctl[kMlpBatch], q_si[kMlpBatch];
  uint32_t q_it[kMlpBatch];
  long long q_elapsed[kMlpBatch], q_total[kMlpBatch];
  int q_ev[kMlpBatch], q_tr[kMlpBatch], q_logn[kMlpBatch];
  uint32_t q_cc[kMlpBatch];
  // Philox key of each item's next event (written by the owner) and the two
  // variates drawn from it by a helper thread while the owner canonicalises
  double k_draw[kMlpBatch], k_choice[kMlpBatch];
  uint32_t k_env[kMlpBatch], k_seq[kMlpBatch], k_it[kMlpBatch];
  int warp_cnt[kMlpThreads / 32];
  int q_count;
};

// Carves the tensor-core operands out of the dynamic shared memory that
// follows `used` bytes of fixed structures.
__device__ __forceinline__ void tc_carve(const MlpView& w,
                                         unsigned char* smem_raw, size_t used,
                                         TcCtx& tc) {
  TcShared* ts = reinterpret_cast<TcShared*>(
      (reinterpret_cast<uintptr_t>(smem_raw + used) + 15) & ~uintptr_t(15));
  unsigned char* a_tile = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(ts + 1) + 127) & ~uintptr_t(127));
  tc_setup(w, tc, a_tile,
           a_tile + tc_a_bytes(w.h1 / w.tc_k_phases,
                               w.tensor_core == 2 ? 2 : 1),
           ts);
}

static size_t tc_smem_bytes(size_t used, int h1, int h2, int tiles,
                            int k_phases) {
  return used + sizeof(TcShared) + 16 + 128 +
         tc_a_bytes(h1 / k_phases, tiles) +
         tc_b_bytes(h1 / k_phases, h2, tiles);
}

// SLIM (tensor path, operand tiles <= ~100 KB): 256 threads and two CTAs per
// SM, so that one CTA's float64 item build / event phases (four warps,
// latency-bound) run under the other's wave.
constexpr int kMlpSlimThreads = 256;
template <int NPT, bool TC, bool SLIM = false>
__global__ void __launch_bounds__(SLIM ? kMlpSlimThreads : kMlpThreads,
                                  SLIM ? 2 : 1)
    k_step_learned(const StepArgs a, const MlpView w) {
  static_assert(!SLIM || TC, "the slim form is a tensor-path form");
  static_assert(kMlpSlimThreads >= 2 * kMlpBatch, "owners + variate helpers");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Shared = MlpStepSharedT<TC>;
  Shared& sh = *reinterpret_cast<Shared*>(smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  mlp_stage_small(w, sh.m);
  TcCtx tc{};
  if constexpr (TC) tc_carve(w, smem_raw, sizeof(Shared), tc);
  const double2* base = reinterpret_cast<const double2*>(a.lat.base_xy);
  const int4* nbr_tab = reinterpret_cast<const int4*>(a.lat.nbr);
  const LogSink log{a.out.log_count ? a.out.log_capacity : 0,
                    a.out.log_elapsed_us, a.out.log_site, a.out.log_ctrl};
  // contiguous env range of this CTA
  const int64_t n = a.st.n_envs;
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  int64_t cursor = per * blockIdx.x;
  const int64_t hi = cursor + per < n ? cursor + per : n;
  if (tid == 0) sh.q_count = 0;
  __syncthreads();
#ifdef PD_MLP_PHASE_CLOCKS
  long long ph_t0 = clock64();
#endif

  while (true) {
    const int n_pending = sh.q_count;
    int64_t avail = hi - cursor;
    if (avail < 0) avail = 0;
    const int n_fresh = static_cast<int>(
        avail < kMlpBatch - n_pending ? avail : kMlpBatch - n_pending);
    const int n_act = n_pending + n_fresh;
    if (n_act == 0) break;
    // ---- build items (threads 0..n_act-1 own one item each) ----
#ifdef PD_MLP_PHASE_CLOCKS
    long long bld_t0 = clock64();
#endif
    int env = 0, ctl = 0, si = 0, ev = 0, tr = 0, logn = 0;
    uint32_t it = 0, cc = 0;  // cc: the env's control counter at launch
    long long elapsed = 0, total = 0, dwell = 0;
    int nb[3] = {0, 0, 0};
    double2 pn[3];
    Canonical can;
    can.x0 = can.x1 = 0.f;
    can.head[0] = can.head[1] = can.head[2] = 0;
    Fov4 fov{};
    Lattice4 lt{};
    const bool own = tid < n_act;
    bool skipped = false;  // env sits this call out (pd_env_step resets it)
    if (own) {
      if (tid < n_pending) {
        env = sh.q_env[tid]; ctl = sh.q_ctl[tid]; si = sh.q_si[tid];
        it = sh.q_it[tid]; elapsed = sh.q_elapsed[tid];
        total = sh.q_total[tid]; ev = sh.q_ev[tid]; tr = sh.q_tr[tid];
        logn = sh.q_logn[tid]; cc = sh.q_cc[tid];
      } else {
        env = static_cast<int>(cursor + (tid - n_pending));
        si = a.st.si_idx[env];
        cc = a.st.ctrl_count[env];
      }
      skipped = a.skip && a.skip[env];
      lt = load_lattice4(a.st.lattice, env);
      fov = load_fov4(a.st.fov, env);
    }
    // controls with zero dwell do no rate evaluation (graphene.py:658)
    bool need_eval = false;
    double2 psi = make_double2(0.0, 0.0);
    if (own && !skipped) {
      while (ctl < a.n_controls) {
        dwell = a.dwell_us ? a.dwell_us[static_cast<int64_t>(env) *
                                            a.n_controls + ctl]
                           : a.dwell_us_scalar;
        if (elapsed < dwell) {
          need_eval = true;
          break;
        }
        total += dwell;
        ++ctl;
        elapsed = 0;
        it = 0;
      }
    }
    // The event's random variates depend on (env, control, iteration) only:
    // threads 128..255 draw them (Philox + the float64 log) while the owners
    // run the geometry below.
    PD_MLP_BLD(4);
    // (named barrier 1: the four owner warps arrive and go on, the four
    // helper warps wait for them; the other warps take no part)
    if (tid < kMlpBatch) {
      sh.k_it[tid] = need_eval ? it : 0xFFFFFFFFu;
      sh.k_env[tid] = static_cast<uint32_t>(env);
      sh.k_seq[tid] = cc + static_cast<uint32_t>(ctl);
      __threadfence_block();
      asm volatile("bar.arrive 1, %0;" ::"n"(2 * kMlpBatch) : "memory");
    } else if (tid < 2 * kMlpBatch) {
      asm volatile("bar.sync 1, %0;" ::"n"(2 * kMlpBatch) : "memory");
      const int m = tid - kMlpBatch;
      const uint32_t k_it = sh.k_it[m];
      if (k_it != 0xFFFFFFFFu) {
        const uint4 pw = philox4x32_10(a.st.env_offset + sh.k_env[m],
                                       sh.k_seq[m], k_it, PD_STREAM_KMC,
                                       a.st.seed);
        sh.k_draw[m] = -log1p(-u53(pw.x, pw.y));
        sh.k_choice[m] = u53(pw.z, pw.w);
      }
    }
    PD_MLP_BLD(5);
    if (own && !skipped) {
      psi = site_position(__ldg(base + si), lt);
      if (need_eval) {
        const double2 c2 = reinterpret_cast<const double2*>(
            a.controls_xy)[static_cast<int64_t>(env) * a.n_controls + ctl];
        const double2 beam = a.material_frame
                                 ? c2
                                 : microscope_to_material(fov, c2.x, c2.y);
        const int4 v = __ldg(nbr_tab + si);
        nb[0] = v.x; nb[1] = v.y; nb[2] = v.z;
#pragma unroll
        for (int i = 0; i < 3; ++i)
          pn[i] = site_position(__ldg(base + nb[i]), lt);
        PD_MLP_BLD(6);
        can = canonicalise(beam, psi, pn);
        PD_MLP_BLD(7);
      }
    }
    if (tid < kMlpBatch) {
      sh.m.xs[tid][0] = can.x0;
      sh.m.xs[tid][1] = can.x1;
    }
    __syncthreads();
    PD_MLP_PHASE(0);
    // ---- network for the whole batch ----
    if constexpr (TC) {
      mlp_wave_tc<SLIM ? kMlpSlimThreads : kMlpThreads>(w, sh.m, tc);
    } else {
      mlp_wave<NPT>(w, sh.m);
    }
    PD_MLP_PHASE(1);
    // ---- events ----
    bool survive = false;
#ifdef PD_MLP_PHASE_CLOCKS
    long long ev_t0 = clock64();
#endif
    if (own) {
      // the env's counters, asked for now so that they are here when an env
      // finalises below (loads issued after the first store there would
      // queue up behind one another: the pointers may alias)
      long long st_time = 0;
      int st_events = 0, st_transitions = 0;
      double st_scale = 0.0;
      if (!skipped) {
        st_time = a.st.sim_time_us[env];
        st_events = a.st.n_events[env];
        st_transitions = a.st.n_transitions[env];
        st_scale = a.st.fov_scale[env];
      }
      if (need_eval) {
        float r[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) r[i] = sh.m.out[tid][can.head[i]];
        PD_MLP_EV(12);
        int slot = 0;
        bool bad = false;
        const bool hit = kmc_event_drawn(r, sh.k_draw[tid], sh.k_choice[tid],
                                         dwell, &elapsed, &slot, &bad);
        PD_MLP_EV(13);
        if (bad) a.st.status[env] |= PD_ENV_BAD_RATE;
        ++ev;
        ++it;
        if (hit) {
          si = nb[slot];
          psi = slot == 0 ? pn[0] : (slot == 1 ? pn[1] : pn[2]);
          ++tr;
          if (log.capacity > 0) {
            if (logn < log.capacity) {
              const int64_t o = static_cast<int64_t>(env) * log.capacity + logn;
              log.elapsed_us[o] = elapsed;
              log.site[o] = si;
              if (log.ctrl) log.ctrl[o] = ctl;
            } else {
              a.st.status[env] |= PD_ENV_LOG_OVERFLOW;
            }
            ++logn;
          }
        }
        if (elapsed >= dwell) {  // control finished
          total += dwell;
          ++ctl;
          elapsed = 0;
          it = 0;
        }
        survive = ctl < a.n_controls;
      }
      PD_MLP_EV(14);
      if (!survive && !skipped) {
        // ---- finalise the env: simulator.py:152-169 ----
        uint8_t recentred = 0;
        if (!a.material_frame) {
          total += a.image_duration_us;
          if (silicon_outside_safe_area(fov, psi)) {
            store_fov4(a.st.fov, env, centred_fov(psi, st_scale));
            total += a.image_duration_us;
            recentred = 1;
          }
          a.st.sim_time_us[env] = st_time + total;
        }
        a.st.si_idx[env] = si;
        a.st.ctrl_count[env] = cc + static_cast<uint32_t>(a.n_controls);
        a.st.n_events[env] = st_events + ev;
        a.st.n_transitions[env] = st_transitions + tr;
        if (a.out.elapsed_us) a.out.elapsed_us[env] = total;
        if (a.out.transitions) a.out.transitions[env] = tr;
        if (a.out.events) a.out.events[env] = ev;
        if (a.out.recentred) a.out.recentred[env] = recentred;
        if (a.out.si_xy) reinterpret_cast<double2*>(a.out.si_xy)[env] = psi;
        if (a.out.log_count) a.out.log_count[env] = logn;
      }
      PD_MLP_EV(15);
    }
    // ---- ordered compaction of the survivors into the queue ----
    const unsigned bal = __ballot_sync(0xffffffffu, survive);
    if (lane == 0) sh.warp_cnt[warp] = __popc(bal);
    __syncthreads();
    PD_MLP_PHASE(2);
    // only the owners' warps (the first kMlpBatch / 32) hold survivors
    int offset = 0, total_surv = 0;
#pragma unroll
    for (int w2 = 0; w2 < kMlpBatch / 32; ++w2) {
      const int c = sh.warp_cnt[w2];
      offset += w2 < warp ? c : 0;
      total_surv += c;
    }
    // every thread read q_count at the top of the loop, barriers ago
    if (tid == 0) sh.q_count = total_surv;
    if (survive) {
      const int pos = offset + __popc(bal & ((1u << lane) - 1u));
      sh.q_env[pos] = env; sh.q_ctl[pos] = ctl; sh.q_si[pos] = si;
      sh.q_it[pos] = it; sh.q_elapsed[pos] = elapsed; sh.q_total[pos] = total;
      sh.q_ev[pos] = ev; sh.q_tr[pos] = tr; sh.q_logn[pos] = logn;
      sh.q_cc[pos] = cc;
    }
    cursor += n_fresh;
    __syncthreads();
    PD_MLP_PHASE(3);
  }
  if constexpr (TC) tc_teardown(w, tc);
}

// RateFunction seam with the learned model: rates + successor sites.
template <int NPT, bool TC>
__global__ void __launch_bounds__(kMlpThreads, 1)
    k_rates_learned(const pd_lattice lat, const pd_state st, const MlpView w,
                    const double* __restrict__ beam_xy,
                    float* __restrict__ rates_out,
                    int32_t* __restrict__ nbr_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Shared = typename std::conditional<TC, MlpSmall, MlpShared>::type;
  Shared& sh = *reinterpret_cast<Shared*>(smem_raw);
  mlp_stage_small(w, sh);
  TcCtx tc{};
  if constexpr (TC) tc_carve(w, smem_raw, sizeof(Shared), tc);
  const double2* base = reinterpret_cast<const double2*>(lat.base_xy);
  const int4* nbr_tab = reinterpret_cast<const int4*>(lat.nbr);
  for (int64_t t0 = static_cast<int64_t>(blockIdx.x) * kMlpBatch;
       t0 < st.n_envs; t0 += static_cast<int64_t>(gridDim.x) * kMlpBatch) {
    const int64_t e = t0 + threadIdx.x;
    const bool own = threadIdx.x < kMlpBatch && e < st.n_envs;
    Canonical can;
    can.x0 = can.x1 = 0.f;
    can.head[0] = can.head[1] = can.head[2] = 0;
    int nb[3] = {0, 0, 0};
    if (own) {
      const Lattice4 lt = load_lattice4(st.lattice, e);
      const int si = st.si_idx[e];
      const double2 psi = site_position(__ldg(base + si), lt);
      const int4 v = __ldg(nbr_tab + si);
      nb[0] = v.x; nb[1] = v.y; nb[2] = v.z;
      double2 pn[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) pn[i] = site_position(__ldg(base + nb[i]), lt);
      can = canonicalise(reinterpret_cast<const double2*>(beam_xy)[e], psi, pn);
    }
    __syncthreads();
    if (threadIdx.x < kMlpBatch) {
      sh.xs[threadIdx.x][0] = can.x0;
      sh.xs[threadIdx.x][1] = can.x1;
    }
    __syncthreads();
    if constexpr (TC) {
      mlp_wave_tc<kMlpThreads>(w, sh, tc);
    } else {
      mlp_wave<NPT>(w, sh);
    }
    if (own) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (rates_out) rates_out[3 * e + i] = sh.out[threadIdx.x][can.head[i]];
        if (nbr_out) nbr_out[3 * e + i] = nb[i];
      }
    }
  }
  if constexpr (TC) tc_teardown(w, tc);
}

// apply_model (learn_rates.py:704-732) for one model: softmax(o[:3]) * o[3].
template <int NPT>
__global__ void __launch_bounds__(kMlpThreads, 1)
    k_apply_model(const MlpView w, const float* __restrict__ x, int64_t n,
                  float* __restrict__ out, float scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MlpShared& sh = *reinterpret_cast<MlpShared*>(smem_raw);
  mlp_stage_small(w, sh);
  for (int64_t t0 = static_cast<int64_t>(blockIdx.x) * kMlpBatch; t0 < n;
       t0 += static_cast<int64_t>(gridDim.x) * kMlpBatch) {
    const int64_t e = t0 + threadIdx.x;
    const bool own = threadIdx.x < kMlpBatch && e < n;
    __syncthreads();
    if (threadIdx.x < kMlpBatch) {
      sh.xs[threadIdx.x][0] = own ? x[2 * e] : 0.f;
      sh.xs[threadIdx.x][1] = own ? x[2 * e + 1] : 0.f;
    }
    __syncthreads();
    mlp_wave<NPT>(w, sh);
    if (own) {
      const float* o = sh.out[threadIdx.x];
      const float mx = fmaxf(o[0], fmaxf(o[1], o[2]));
      const float e0 = expf(o[0] - mx), e1 = expf(o[1] - mx),
                  e2 = expf(o[2] - mx);
      const float s = o[3] / (e0 + e1 + e2) * scale;
      out[3 * e + 0] += e0 * s;
      out[3 * e + 1] += e1 * s;
      out[3 * e + 2] += e2 * s;
    }
  }
}

// PD_MLP_SLIM=0 keeps the one-CTA-per-SM form of the tensor step for every
// shape.  Default: two CTAs per SM where two sets of resident tiles fit.
// (Streaming W1 to make two CTAs fit at H = 128 was measured: 0.225 ms
// against 0.215 ms with one CTA and resident tiles.)
int& option_mlp_slim() {
  static int v = [] {
    const char* e = getenv("PD_MLP_SLIM");
    return e && e[0] == '0' ? 0 : 1;
  }();
  return v;
}

static int mlp_view(const pd_mlp* mlp, MlpView* v) {
  PD_REQUIRE(mlp != nullptr, "null pd_mlp");
  PD_REQUIRE(mlp->context_dim == 2,
             "only context_dim == 2 is a valid drop-in for predict() "
             "(learn_rates.py:961-964 fails with voltage/current context)");
  PD_REQUIRE(mlp->hidden1 % kChunk == 0 && mlp->hidden1 >= kChunk &&
                 mlp->hidden1 <= 256,
             "hidden1 must be a multiple of 16 in [16, 256]");
  PD_REQUIRE(mlp->hidden2 == 32 || mlp->hidden2 == 64 || mlp->hidden2 == 128 ||
                 mlp->hidden2 == 256,
             "hidden2 must be 32, 64, 128 or 256");
  PD_REQUIRE(mlp->w0 && mlp->b0 && mlp->w1 && mlp->b1 && mlp->w2 && mlp->b2,
             "null MLP weights");
  PD_REQUIRE(!mlp->batchnorm || (mlp->bn_scale && mlp->bn_offset &&
                                 mlp->bn_mean && mlp->bn_var),
             "null BatchNorm statistics");
  if (mlp->tensor_core) {
    PD_REQUIRE(mlp->w1_umma != nullptr, "tensor_core needs w1_umma");
    PD_REQUIRE(mlp->hidden2 % 32 == 0, "tensor_core needs hidden2 % 32 == 0");
    PD_REQUIRE(mlp->hidden1 % 32 == 0, "tensor_core needs hidden1 % 32 == 0");
    PD_REQUIRE(mlp->tensor_core == 1 || mlp->tensor_core == 2,
               "tensor_core: 0 (FP32), 1 (bf16) or 2 (fp16 hi + lo)");
  }
  // W1's tiles stay in shared memory when they fit beside the h1 tiles;
  // otherwise a wave takes K in two halves and streams them (split form at
  // H = 256)
  int k_phases = 1;
  if (mlp->tensor_core) {
    const int tiles = mlp->tensor_core == 2 ? 2 : 1;
    const size_t cap = 227 * 1024;
    if (tc_smem_bytes(sizeof(MlpStepSharedT<true>), mlp->hidden1, mlp->hidden2,
                      tiles, 1) > cap) {
      k_phases = 2;
      PD_REQUIRE(mlp->hidden1 % 64 == 0 &&
                     tc_smem_bytes(sizeof(MlpStepSharedT<true>), mlp->hidden1,
                                   mlp->hidden2, tiles, 2) <= cap,
                 "tensor_core: the operand tiles do not fit in shared memory");
    }
  }
  *v = MlpView{mlp->context_dim, mlp->hidden1, mlp->hidden2, mlp->batchnorm,
               mlp->bn_scale, mlp->bn_offset, mlp->bn_mean, mlp->bn_var,
               mlp->w0, mlp->b0, mlp->w1, mlp->b1, mlp->w2, mlp->b2,
               mlp->tensor_core, mlp->w1_umma, k_phases};
  return PD_OK;
}

template <typename F>
static int dispatch_npt(int h2, F&& f) {
  switch (h2 / 16) {
    case 2: return f(std::integral_constant<int, 2>());
    case 4: return f(std::integral_constant<int, 4>());
    case 8: return f(std::integral_constant<int, 8>());
    default: return f(std::integral_constant<int, 16>());
  }
}

int learned_step(const pd_lattice* lat, const pd_state* st, const pd_mlp* mlp,
                 const StepArgs& a, bool rollout, cudaStream_t stream) {
  (void)lat;
  if (rollout) {
    set_error("pd_rollout with PD_RATE_LEARNED is not implemented; call "
              "pd_step_and_image per step");
    return PD_ERR_UNSUPPORTED;
  }
  MlpView v;
  int rc = mlp_view(mlp, &v);
  if (rc != PD_OK) return rc;
  const int64_t tiles = (st->n_envs + kMlpBatch - 1) / kMlpBatch;
  const int grid = static_cast<int>(tiles < sm_count() ? tiles : sm_count());
  if (v.tensor_core) {
    const int smem = static_cast<int>(
        tc_smem_bytes(sizeof(MlpStepSharedT<true>), v.h1, v.h2,
                      v.tensor_core == 2 ? 2 : 1, v.tc_k_phases));
    // two CTAs per SM when two sets of tiles fit (1 KB per CTA is reserved)
    if (option_mlp_slim() && 2 * (smem + 1024) <= 228 * 1024) {
      auto kern = k_step_learned<2, true, true>;
      PD_CUDA_OK(cudaFuncSetAttribute(
          kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      PD_CUDA_OK(cudaFuncSetAttribute(
          kern, cudaFuncAttributePreferredSharedMemoryCarveout,
          cudaSharedmemCarveoutMaxShared));
      const int grid2 = static_cast<int>(
          tiles < 2 * sm_count() ? tiles : 2 * sm_count());
      kern<<<grid2, kMlpSlimThreads, smem, stream>>>(a, v);
      PD_CUDA_OK(cudaGetLastError());
      return PD_OK;
    }
    auto kern = k_step_learned<2, true>;
    PD_CUDA_OK(cudaFuncSetAttribute(
        kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, kMlpThreads, smem, stream>>>(a, v);
    PD_CUDA_OK(cudaGetLastError());
    return PD_OK;
  }
  const int smem = static_cast<int>(sizeof(MlpStepSharedT<false>));
  return dispatch_npt(v.h2, [&](auto npt) -> int {
    auto kern = k_step_learned<decltype(npt)::value, false>;
    PD_CUDA_OK(cudaFuncSetAttribute(
        kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, kMlpThreads, smem, stream>>>(a, v);
    PD_CUDA_OK(cudaGetLastError());
    return PD_OK;
  });
}

int learned_rates(const pd_lattice* lat, const pd_state* st, const pd_mlp* mlp,
                  const double* beam_xy, float* rates_out, int32_t* nbr_out,
                  cudaStream_t stream) {
  MlpView v;
  int rc = mlp_view(mlp, &v);
  if (rc != PD_OK) return rc;
  const int64_t tiles = (st->n_envs + kMlpBatch - 1) / kMlpBatch;
  const int grid = static_cast<int>(tiles < sm_count() ? tiles : sm_count());
  if (v.tensor_core) {
    const int smem =
        static_cast<int>(tc_smem_bytes(sizeof(MlpSmall), v.h1, v.h2,
                                       v.tensor_core == 2 ? 2 : 1,
                                       v.tc_k_phases));
    auto kern = k_rates_learned<2, true>;
    PD_CUDA_OK(cudaFuncSetAttribute(
        kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, kMlpThreads, smem, stream>>>(*lat, *st, v, beam_xy, rates_out,
                                              nbr_out);
    PD_CUDA_OK(cudaGetLastError());
    return PD_OK;
  }
  const int smem = static_cast<int>(sizeof(MlpShared));
  return dispatch_npt(v.h2, [&](auto npt) -> int {
    auto kern = k_rates_learned<decltype(npt)::value, false>;
    PD_CUDA_OK(cudaFuncSetAttribute(
        kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, kMlpThreads, smem, stream>>>(*lat, *st, v, beam_xy, rates_out,
                                              nbr_out);
    PD_CUDA_OK(cudaGetLastError());
    return PD_OK;
  });
}

}  // namespace pd

// learn_rates.py:704-732 apply_model: mean over an ensemble of
// softmax(out[:3]) * out[3].  x: device float [n][2]; out: device float
// [n][3] (overwritten).
extern "C" int pd_mlp_apply_model(const pd_mlp* models, int32_t n_models,
                                  const float* x, int64_t n, float* out,
                                  void* stream) {
  PD_REQUIRE(models != nullptr && n_models > 0, "no models");
  PD_REQUIRE(n >= 0 && (n == 0 || (x && out)), "bad arguments");
  if (n == 0) return PD_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PD_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * 3 * n, s));
  const int64_t tiles = (n + pd::kMlpBatch - 1) / pd::kMlpBatch;
  const int grid =
      static_cast<int>(tiles < pd::sm_count() ? tiles : pd::sm_count());
  const int smem = static_cast<int>(sizeof(pd::MlpShared));
  for (int i = 0; i < n_models; ++i) {
    pd::MlpView v;
    int rc = pd::mlp_view(&models[i], &v);
    if (rc != PD_OK) return rc;
    rc = pd::dispatch_npt(v.h2, [&](auto npt) -> int {
      auto kern = pd::k_apply_model<decltype(npt)::value>;
      PD_CUDA_OK(cudaFuncSetAttribute(
          kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      kern<<<grid, pd::kMlpThreads, smem, s>>>(v, x, n, out,
                                               1.0f / n_models);
      PD_CUDA_OK(cudaGetLastError());
      return PD_OK;
    });
    if (rc != PD_OK) return rc;
  }
  return PD_OK;
}

#ifdef PD_MLP_PHASE_CLOCKS
extern "C" int pd_debug_mlp_phases(unsigned long long* out16) {
  unsigned long long zero[16] = {0};
  PD_CUDA_OK(cudaDeviceSynchronize());
  PD_CUDA_OK(cudaMemcpyFromSymbol(out16, pd::g_mlp_phase, sizeof(zero)));
  PD_CUDA_OK(cudaMemcpyToSymbol(pd::g_mlp_phase, zero, sizeof(zero)));
  return PD_OK;
}
#endif
