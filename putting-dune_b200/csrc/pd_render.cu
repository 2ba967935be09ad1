// K3 (generic form): STEM frame renderer, imaging.py:117-265
// generate_stem_image, and the pd_render entry point.
//
// pd_render runs the cluster kernel (pd_render_cluster.cu) over all frames;
// that kernel keeps a frame in the shared memory of eight CTAs and therefore
// bounds the atoms in view (1024), the clean-image kernel radius (128 px) and
// the blur radius (4 px = blur_amount < 1.125; imaging.py:42-72 samples
// blur < 1).  Frames outside those bounds are flagged and rendered by the
// kernel in this file: one CTA (1024 threads) per frame, ten passes separated
// by the image-wide reductions, intermediates in a per-CTA global scratch slot
// (2 x S*S floats).  Same arithmetic, same noise fields.
//
//   P0 setup     atoms in view (graphene.py:600-644) -> pixel bins and Z^e
//                weights; Gaussian kernel tables; per-row jitter shifts
//   P1 clean     imaging.py:117-173: histogram + separable Gaussian (zero
//                padding) evaluated as a direct splat of the truncated,
//                renormalised kernel; max
//   P2 blur      :212-214 separable Gaussian, reflect padding; max
//   P3 poisson   :199-203 inverse-CDF Poisson of image*mult; max
//   P4 jitter..  :188-196 row roll, :206-209 s&p, :217-218 gamma, :231-236
//                uniform noise; max
//   P5 exp       :221-228 exponential noise; max
//   P6 gauss     :176-185 Gaussian noise, clip; min & max
//   P7 clahe-1   :264 quantise to 14 bit, bin (//65), tile histograms, clip
//                redistribution, CDF maps
//   P8 clahe-2   bilinear blend of the four neighbouring tile maps; min & max
//   P9 output    rescale to [0, 1], write the frame
//
// Noise fields follow the injected convention documented in
// include/pdune_b200.h: one Philox call per four consecutive pixels and stage.
#include "pd_render.cuh"

namespace pd {

constexpr int kRenderThreads = 1024;
constexpr int kRenderWarps = kRenderThreads / 32;
constexpr int kMaxAtoms = 2048;
constexpr int kMaxRadius = 255;   // clean-image kernel radius (4 sigma)
constexpr int kMaxBlurRadius = 16;
constexpr int kBands = 16;        // 512 / 32 row bands
constexpr int kBandCap = 1024;
constexpr int kUnroll = 4;        // independent loads per thread per trip

struct RenderShared {
  short2 atom_rc[kMaxAtoms];   // (row, col) pixel of each atom
  float atom_w[kMaxAtoms];     // Z^exponent
  float ky[kMaxRadius + 1];    // rows (uses fov width: imaging.py:159 quirk)
  float kx[kMaxRadius + 1];
  float kb[kMaxBlurRadius + 1];
  int shift[512];
  int warp_count[2][kRenderWarps];
  float red_a[kRenderWarps];
  float red_b[kRenderWarps];
  double red_d[kRenderWarps];
  float bcast[4];
  int n_atoms;
  int lwy, lwx, lwb;
  unsigned short maps[kTiles * kTiles][kBins];
  // P1: atoms that can touch each 32-row band, in atom order
  unsigned short band_list[kBands][kBandCap];
  int band_n[kBands];
  double inv_k[kInvTable];     // 1/k for the Poisson recurrence
  // union: row accumulators (P1/P2) or histograms (P7)
  union {
    float rows[kRenderWarps][512];
    int hist[kTiles * kTiles][kBins];
  } u;
};

__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[threadIdx.x & 31];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
  return r;
}

__device__ __forceinline__ float block_min(float v, float* red) {
  return -block_max(-v, red);
}

__device__ __forceinline__ int reflect_index(int i, int n) {
  // scipy 'reflect': d c b a | a b c d | d c b a
  while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - 1 - i;
  return i;
}

__global__ void __launch_bounds__(kRenderThreads, 1)
    k_render_generic(const RenderArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RenderShared& sh = *reinterpret_cast<RenderShared*>(smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int S = a.size;
  const int npix = S * S;
  const int mask = S - 1;
  float* v0 = a.scratch + static_cast<size_t>(blockIdx.x) * 2 * npix;
  float* v1 = v0 + npix;
  const double2* base = reinterpret_cast<const double2*>(a.lat.base_xy);

  for (int i = tid; i < kInvTable; i += kRenderThreads)
    sh.inv_k[i] = i > 0 ? 1.0 / static_cast<double>(i) : 0.0;

  if (a.generic != nullptr && *a.n_generic == 0) return;
  for (int f = blockIdx.x; f < a.m; f += gridDim.x) {
    if (a.generic != nullptr && a.generic[f] == 0) continue;
    const int e = a.env_ids ? a.env_ids[f] : f;
    const uint32_t env = a.st.env_offset + static_cast<uint32_t>(e);
    const uint32_t frame = a.st.frame_count[e];
    const uint64_t seed = a.st.seed;
    const Lattice4 lt = load_lattice4(a.st.lattice, e);
    const Fov4 fv = load_fov4(a.st.fov, e);
    const int si = a.st.si_idx[e];
    const double* ip = a.st.image_params + 9 * e;
    const float exponent = static_cast<float>(ip[0]);
    const float gauss_sd = sqrtf(static_cast<float>(ip[1]));
    const double jitter_rate = ip[2];
    const float poisson_mult = static_cast<float>(ip[3]);
    const float sp_amount = static_cast<float>(ip[4]);
    const double blur_amount = ip[5];
    const float gamma = static_cast<float>(ip[6]);
    const float exp_lambda = static_cast<float>(ip[7]);
    const float uniform_scale = static_cast<float>(ip[8]);
    float* out = a.out + static_cast<size_t>(f) * npix;
    __syncthreads();

    // ---------------------------------------------------------------- P0
    {
      const double fw = fv.urx - fv.llx, fh = fv.ury - fv.lly;
      const float w_c = powf(6.0f, exponent), w_si = powf(14.0f, exponent);
      bool keep[2];
      short2 rc[2];
      float wt[2];
#pragma unroll
      for (int round = 0; round < 2; ++round) {
        const int k = round * kRenderThreads + tid;
        keep[round] = false;
        if (k < a.lat.n_sites) {
          const double2 p = site_position(__ldg(base + k), lt);
          const AtomBin ab = atom_bin(p, fv, S, a.buffer, a.bw);
          if (ab.keep) {
            keep[round] = true;
            rc[round] = make_short2(static_cast<short>(ab.row),
                                    static_cast<short>(ab.col));
            wt[round] = k == si ? w_si : w_c;
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep[round]);
        if (lane == 0) sh.warp_count[round][warp] = __popc(m);
      }
      if (tid == 0) {
        const double sy = S / (2.15 * fw), sx = S / (2.15 * fh);
        sh.lwy = static_cast<int>(4.0 * sy + 0.5);
        sh.lwx = static_cast<int>(4.0 * sx + 0.5);
        sh.lwb = blur_amount > 1e-15 ? static_cast<int>(4.0 * blur_amount + 0.5)
                                     : -1;
        if (sh.lwy > kMaxRadius) sh.lwy = kMaxRadius;
        if (sh.lwx > kMaxRadius) sh.lwx = kMaxRadius;
        if (sh.lwb > kMaxBlurRadius) sh.lwb = kMaxBlurRadius;
      }
      __syncthreads();
      // exclusive prefix over (round, warp) in site order
      int offset = 0;
      {
        int total = 0;
        for (int r2 = 0; r2 < 2; ++r2)
          for (int w2 = 0; w2 < kRenderWarps; ++w2) {
            const int c = sh.warp_count[r2][w2];
            if (r2 == 0 && w2 == warp) offset = total;
            total += c;
          }
        if (tid == 0) sh.n_atoms = total < kMaxAtoms ? total : kMaxAtoms;
      }
      int off1 = 0;
      for (int w2 = 0; w2 < kRenderWarps; ++w2) off1 += sh.warp_count[0][w2];
      for (int w2 = 0; w2 < warp; ++w2) off1 += sh.warp_count[1][w2];
#pragma unroll
      for (int round = 0; round < 2; ++round) {
        const unsigned m = __ballot_sync(0xffffffffu, keep[round]);
        const int pos = (round == 0 ? offset : off1) +
                        __popc(m & ((1u << lane) - 1u));
        if (keep[round] && pos < kMaxAtoms) {
          sh.atom_rc[pos] = rc[round];
          sh.atom_w[pos] = wt[round];
        }
      }
      // Gaussian tables: w[x] = exp(-0.5 x^2 / sigma^2) / sum (scipy
      // _gaussian_kernel1d with radius int(4 sigma + 0.5)).
      if (warp < 3) {
        const double sigma = warp == 0 ? S / (2.15 * fw)
                             : warp == 1 ? S / (2.15 * fh) : blur_amount;
        const int lw = warp == 0 ? sh.lwy : warp == 1 ? sh.lwx : sh.lwb;
        float* tab = warp == 0 ? sh.ky : warp == 1 ? sh.kx : sh.kb;
        if (lw >= 0) {
          double sum = 0.0;
          for (int x = lane; x <= lw; x += 32) {
            const double v = exp(-0.5 / (sigma * sigma) * x * x);
            sum += x == 0 ? v : 2.0 * v;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
          for (int x = lane; x <= lw; x += 32)
            tab[x] = static_cast<float>(
                exp(-0.5 / (sigma * sigma) * x * x) / sum);
        }
      }
      // per-row jitter shifts (imaging.py:192)
      if (tid < S) {
        const uint4 w = philox4x32_10(env, frame, tid, PD_STREAM_JITTER, seed);
        sh.shift[tid] = poisson_icdf(jitter_rate, u24(w.x)) & mask;
      }
      __syncthreads();
    }
    const int n_atoms = sh.n_atoms;
    const int lwy = sh.lwy, lwx = sh.lwx, lwb = sh.lwb;
    // band lists: warp b collects, in atom order, the atoms whose kernel
    // footprint reaches rows [32 b, 32 b + 31]
    bool bands_ok = true;
    if (warp < kBands) {
      const int r_lo = warp * 32 - lwy, r_hi = warp * 32 + 31 + lwy;
      int cnt = 0;
      for (int i0 = 0; i0 < n_atoms; i0 += 32) {
        const int i = i0 + lane;
        const bool hit = i < n_atoms && sh.atom_rc[i].x >= r_lo &&
                         sh.atom_rc[i].x <= r_hi;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        const int pos = cnt + __popc(m & ((1u << lane) - 1u));
        if (hit && pos < kBandCap)
          sh.band_list[warp][pos] = static_cast<unsigned short>(i);
        cnt += __popc(m);
      }
      if (lane == 0) sh.band_n[warp] = cnt;
    }
    __syncthreads();
    for (int b = 0; b < kBands; ++b) bands_ok &= sh.band_n[b] <= kBandCap;

    // ---------------------------------------------------------------- P1
    float vmax = 0.f;
    {
      float* acc = sh.u.rows[warp];
      for (int r = warp; r < S; r += kRenderWarps) {
        for (int c = lane; c < S; c += 32) acc[c] = 0.f;
        __syncwarp();
        const int band = r >> 5;
        const int n_list = bands_ok ? sh.band_n[band] : n_atoms;
        for (int ii = 0; ii < n_list; ++ii) {
          const int i = bands_ok ? sh.band_list[band][ii] : ii;
          const short2 rc = sh.atom_rc[i];
          int dr = r - rc.x;
          dr = dr < 0 ? -dr : dr;
          if (dr > lwy) continue;
          const float wr = sh.atom_w[i] * sh.ky[dr];
          const int c0 = rc.y - lwx;
          for (int c = c0 + lane; c <= rc.y + lwx; c += 32) {
            if (c >= 0 && c < S) {
              const int dc = c - rc.y;
              acc[c] += wr * sh.kx[dc < 0 ? -dc : dc];
            }
          }
          __syncwarp();
        }
        for (int c = lane; c < S; c += 32) {
          const float v = acc[c];
          v0[r * S + c] = v;
          vmax = fmaxf(vmax, v);
        }
        __syncwarp();
      }
    }
    float m_prev = block_max(vmax, sh.red_a);  // max of the clean image
    float* cur = v0;
    float* other = v1;
    if (a.stop_stage == PD_RENDER_CLEAN) {
      const float inv = 1.0f / m_prev;
      __syncthreads();
      for (int p = tid; p < npix; p += kRenderThreads) out[p] = cur[p] * inv;
      continue;
    }

    // ---------------------------------------------------------------- P2
    // gaussian_filter(image / max, blur, mode='reflect'): axis 0 then axis 1.
    if (lwb >= 0) {
      __syncthreads();
      for (int p = tid; p < npix; p += kRenderThreads) {
        const int r = p >> a.log2_size, c = p & mask;
        float s = sh.kb[0] * cur[p];
        for (int k = 1; k <= lwb; ++k)
          s += sh.kb[k] * (cur[reflect_index(r - k, S) * S + c] +
                           cur[reflect_index(r + k, S) * S + c]);
        other[p] = s;
      }
      __syncthreads();
      vmax = 0.f;
      for (int p = tid; p < npix; p += kRenderThreads) {
        const int r = p >> a.log2_size, c = p & mask;
        const float* row = other + r * S;
        float s = sh.kb[0] * row[c];
        for (int k = 1; k <= lwb; ++k)
          s += sh.kb[k] * (row[reflect_index(c - k, S)] +
                           row[reflect_index(c + k, S)]);
        cur[p] = s;
        vmax = fmaxf(vmax, s);
      }
      m_prev = block_max(vmax, sh.red_a);
    }
    if (a.stop_stage == PD_RENDER_BLUR) {
      const float inv = 1.0f / m_prev;
      __syncthreads();
      for (int p = tid; p < npix; p += kRenderThreads) out[p] = cur[p] * inv;
      continue;
    }

    // ---------------------------------------------------------------- P3
    {
      const float scale = poisson_mult / m_prev;
      vmax = 0.f;
      __syncthreads();
      for (int g = tid; g < npix / 4; g += kRenderThreads) {
        const float4 v = reinterpret_cast<const float4*>(cur)[g];
        const uint4 w =
            philox4x32_10(env, frame, g, PD_STREAM_RENDER_POISSON, seed);
        float4 o;
        o.x = static_cast<float>(poisson_icdf_tab(
            static_cast<double>(v.x * scale), u24(w.x), sh.inv_k));
        o.y = static_cast<float>(poisson_icdf_tab(
            static_cast<double>(v.y * scale), u24(w.y), sh.inv_k));
        o.z = static_cast<float>(poisson_icdf_tab(
            static_cast<double>(v.z * scale), u24(w.z), sh.inv_k));
        o.w = static_cast<float>(poisson_icdf_tab(
            static_cast<double>(v.w * scale), u24(w.w), sh.inv_k));
        reinterpret_cast<float4*>(cur)[g] = o;
        vmax = fmaxf(vmax, fmaxf(fmaxf(o.x, o.y), fmaxf(o.z, o.w)));
      }
      m_prev = block_max(vmax, sh.red_a);
    }
    if (a.stop_stage == PD_RENDER_POISSON) {
      const float inv = 1.0f / m_prev;
      __syncthreads();
      for (int p = tid; p < npix; p += kRenderThreads) out[p] = cur[p] * inv;
      continue;
    }

    // ---------------------------------------------------------------- P4
    {
      const float inv = 1.0f / m_prev;
      const bool jitter_only = a.stop_stage == PD_RENDER_JITTER;
      vmax = 0.f;
      __syncthreads();
      for (int g = tid; g < npix / 4; g += kRenderThreads) {
        const int p = 4 * g;
        const int r = p >> a.log2_size, c0 = p & mask;
        float vv[4];
        // np.roll(row, k): out[(j + k) % S] = in[j]
#pragma unroll
        for (int j = 0; j < 4; ++j)
          vv[j] = cur[r * S + ((c0 + j - sh.shift[r]) & mask)] * inv;
        if (jitter_only) {
          reinterpret_cast<float4*>(out)[g] =
              make_float4(vv[0], vv[1], vv[2], vv[3]);
          continue;
        }
        const uint4 ws = philox4x32_10(env, frame, g, PD_STREAM_RENDER_SP, seed);
        const uint4 wu =
            philox4x32_10(env, frame, g, PD_STREAM_RENDER_UNIFORM, seed);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t sw = word_of(ws, j);
          float v = powf(fminf(fmaxf(vv[j], 0.f), 1.f), gamma);
          if (u24(sw) <= sp_amount) v = (sw & 255u) < 128u ? 1.0f : 0.0f;
          v += uniform_scale * u24(word_of(wu, j));
          vv[j] = v;
          vmax = fmaxf(vmax, v);
        }
        reinterpret_cast<float4*>(other)[g] =
            make_float4(vv[0], vv[1], vv[2], vv[3]);
      }
      if (jitter_only) continue;
      m_prev = block_max(vmax, sh.red_a);
      float* t = cur;
      cur = other;
      other = t;
    }
    if (a.stop_stage == PD_RENDER_UNIFORM) {
      const float inv = 1.0f / m_prev;
      __syncthreads();
      for (int p = tid; p < npix; p += kRenderThreads) out[p] = cur[p] * inv;
      continue;
    }

    // ---------------------------------------------------------------- P5
    {
      const float inv = 1.0f / m_prev;
      vmax = 0.f;
      __syncthreads();
      for (int g = tid; g < npix / 4; g += kRenderThreads) {
        float4 v = reinterpret_cast<const float4*>(cur)[g];
        const uint4 w = philox4x32_10(env, frame, g, PD_STREAM_RENDER_EXP, seed);
        v.x = v.x * inv - __logf(1.0f - u24(w.x)) * exp_lambda;
        v.y = v.y * inv - __logf(1.0f - u24(w.y)) * exp_lambda;
        v.z = v.z * inv - __logf(1.0f - u24(w.z)) * exp_lambda;
        v.w = v.w * inv - __logf(1.0f - u24(w.w)) * exp_lambda;
        reinterpret_cast<float4*>(cur)[g] = v;
        vmax = fmaxf(vmax, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
      }
      m_prev = block_max(vmax, sh.red_a);
    }
    if (a.stop_stage == PD_RENDER_EXPONENTIAL) {
      const float inv = 1.0f / m_prev;
      __syncthreads();
      for (int p = tid; p < npix; p += kRenderThreads) out[p] = cur[p] * inv;
      continue;
    }

    // ---------------------------------------------------------------- P6
    float g_min, g_max;
    {
      const float inv = 1.0f / m_prev;
      float lo = 1e30f, hi = -1e30f;
      __syncthreads();
      for (int g = tid; g < npix / 4; g += kRenderThreads) {
        float4 v = reinterpret_cast<const float4*>(cur)[g];
        const uint4 w =
            philox4x32_10(env, frame, g, PD_STREAM_RENDER_GAUSS, seed);
        const float ra = gauss_sd * sqrtf(-2.0f * __logf(u24_open(w.x)));
        const float rb = gauss_sd * sqrtf(-2.0f * __logf(u24_open(w.z)));
        float sa, ca, sb, cb;
        __sincosf(fmaf(u24(w.y), 6.28318530717958648f, -3.14159265358979324f),
                  &sa, &ca);
        __sincosf(fmaf(u24(w.w), 6.28318530717958648f, -3.14159265358979324f),
                  &sb, &cb);
        v.x = fminf(fmaxf(v.x * inv - ra * ca, 0.f), 1.f);
        v.y = fminf(fmaxf(v.y * inv - ra * sa, 0.f), 1.f);
        v.z = fminf(fmaxf(v.z * inv - rb * cb, 0.f), 1.f);
        v.w = fminf(fmaxf(v.w * inv - rb * sb, 0.f), 1.f);
        reinterpret_cast<float4*>(cur)[g] = v;
        lo = fminf(lo, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
        hi = fmaxf(hi, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
      }
      g_max = block_max(hi, sh.red_a);
      g_min = block_min(lo, sh.red_b);
    }
    if (a.stop_stage == PD_RENDER_GAUSSIAN) {
      __syncthreads();
      for (int p = tid; p < npix; p += kRenderThreads) out[p] = cur[p];
      continue;
    }

    // ---------------------------------------------------------------- P7
    const int ts = S / kTiles;
    const int log2_ts = a.log2_size - 3;
    {
      for (int i = tid; i < kTiles * kTiles * kBins; i += kRenderThreads)
        (&sh.u.hist[0][0])[i] = 0;
      __syncthreads();
      const float range = g_max - g_min;
      const float q_scale = range > 0.f ? (kGray - 1) / range : 0.f;
      for (int p0 = tid; p0 < npix; p0 += kUnroll * kRenderThreads) {
        float vv[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) vv[u] = cur[p0 + u * kRenderThreads];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int p = p0 + u * kRenderThreads;
          const int r = p >> a.log2_size, c = p & mask;
          // np.round(rescale_intensity(img, out_range=(0, 16383)))
          const float q = rintf((vv[u] - g_min) * q_scale);
          const int bin = static_cast<int>(q) / kBinSize;
          other[p] = __int_as_float(bin);
          atomicAdd(
              &sh.u.hist[(r >> log2_ts) * kTiles + (c >> log2_ts)][bin], 1);
        }
      }
      __syncthreads();
      if (tid < kTiles * kTiles) {
        int clim = static_cast<int>(0.01 * ts * ts);
        if (clim < 1) clim = 1;
        clahe_tile_map(sh.u.hist[tid], sh.maps[tid], clim, ts * ts);
      }
      __syncthreads();
    }

    // ---------------------------------------------------------------- P8
    int m_lo = 1 << 30, m_hi = -1;
    {
      const float inv_ts = 1.0f / ts;
      const int half = ts >> 1;
      for (int p0 = tid; p0 < npix; p0 += kUnroll * kRenderThreads) {
       int bins[kUnroll];
#pragma unroll
       for (int u = 0; u < kUnroll; ++u)
         bins[u] = __float_as_int(other[p0 + u * kRenderThreads]);
#pragma unroll
       for (int u = 0; u < kUnroll; ++u) {
        const int p = p0 + u * kRenderThreads;
        const int r = p >> a.log2_size, c = p & mask;
        const int bin = bins[u];
        const int pr = r + half, pc = c + half;
        const int bi = pr >> log2_ts, bj = pc >> log2_ts;
        const float cy = (pr & (ts - 1)) * inv_ts;
        const float cx = (pc & (ts - 1)) * inv_ts;
        const int t0r = bi - 1 < 0 ? 0 : bi - 1;
        const int t1r = bi > kTiles - 1 ? kTiles - 1 : bi;
        const int t0c = bj - 1 < 0 ? 0 : bj - 1;
        const int t1c = bj > kTiles - 1 ? kTiles - 1 : bj;
        // result += (mapped * coeff).astype(float32), edges in ndindex order
        float acc = __fmul_rn(sh.maps[t0r * kTiles + t0c][bin],
                              __fmul_rn(1.0f - cy, 1.0f - cx));
        acc = __fadd_rn(acc, __fmul_rn(sh.maps[t0r * kTiles + t1c][bin],
                                       __fmul_rn(1.0f - cy, cx)));
        acc = __fadd_rn(acc, __fmul_rn(sh.maps[t1r * kTiles + t0c][bin],
                                       __fmul_rn(cy, 1.0f - cx)));
        acc = __fadd_rn(acc, __fmul_rn(sh.maps[t1r * kTiles + t1c][bin],
                                       __fmul_rn(cy, cx)));
        const int mv = static_cast<int>(acc);  // astype(uint16) truncates
        cur[p] = __int_as_float(mv);
        m_lo = min(m_lo, mv);
        m_hi = max(m_hi, mv);
       }
      }
      m_hi = static_cast<int>(block_max(static_cast<float>(m_hi), sh.red_a));
      m_lo = static_cast<int>(block_min(static_cast<float>(m_lo), sh.red_b));
    }

    // ---------------------------------------------------------------- P9
    {
      const float denom = static_cast<float>(m_hi - m_lo);
      __syncthreads();
      const float inv_d = denom > 0.f ? 1.0f / denom : 0.f;
      for (int p0 = tid; p0 < npix; p0 += kUnroll * kRenderThreads) {
        int mvs[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
          mvs[u] = __float_as_int(cur[p0 + u * kRenderThreads]);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int mv = mvs[u];
          // rescale_intensity: (v - min) / (max - min)
          out[p0 + u * kRenderThreads] =
              denom > 0.f ? static_cast<float>(mv - m_lo) * inv_d
                          : fminf(fmaxf(static_cast<float>(mv), 0.f), 1.f);
        }
      }
    }
  }
}

__global__ void k_advance_frames(const pd_state st, const int32_t* env_ids,
                                 int32_t m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) st.frame_count[env_ids ? env_ids[i] : i] += 1;
}

int validate_common(const pd_lattice* lat, const pd_state* st,
                    const pd_rate_config* rc);

}  // namespace pd

// Workspace layout: [0, 256) counter of generic frames; [256, 256 + 2^18)
// per-frame flags; then kGenericGrid scratch slots of 2 S^2 floats.
static const int64_t kFlagOffset = 256;
static const int64_t kScratchOffset = 256 + pd::kMaxFramesPerLaunch;

extern "C" int pd_render_workspace_bytes(int32_t image_size,
                                         int64_t* out_bytes) {
  PD_REQUIRE(out_bytes != nullptr, "null output");
  PD_REQUIRE(image_size >= 64 && image_size <= 512 &&
                 (image_size & (image_size - 1)) == 0,
             "image_size must be a power of two in [64, 512]");
  *out_bytes = kScratchOffset + static_cast<int64_t>(pd::kGenericGrid) * 2 *
                                    image_size * image_size * sizeof(float);
  return PD_OK;
}

extern "C" int pd_render_clusters(int32_t image_size, int32_t* out_clusters) {
  PD_REQUIRE(out_clusters != nullptr, "null output");
  PD_REQUIRE(image_size >= 64 && image_size <= 512 &&
                 (image_size & (image_size - 1)) == 0,
             "image_size must be a power of two in [64, 512]");
  int n = 0;
  const int rc = pd::render_cluster_count(image_size, &n);
  *out_clusters = n;
  return rc;
}

extern "C" int pd_render(const pd_lattice* lat, const pd_state* st,
                         const int32_t* env_ids, int32_t m, int32_t image_size,
                         int32_t stop_stage, int32_t advance_frame_count,
                         double buffer_size, float* frames_out, void* workspace,
                         int64_t workspace_bytes, void* stream) {
  int rcode = pd::validate_common(lat, st, nullptr);
  if (rcode != PD_OK) return rcode;
  PD_REQUIRE(m >= 0, "negative frame count");
  PD_REQUIRE(env_ids != nullptr || m <= st->n_envs, "m exceeds n_envs");
  PD_REQUIRE(stop_stage >= PD_RENDER_CLEAN && stop_stage <= PD_RENDER_FINAL,
             "unknown stop_stage");
  PD_REQUIRE(buffer_size >= 0.0 && buffer_size <= 0.25,
             "buffer_size must lie in [0, 0.25]");
  PD_REQUIRE(lat->n_sites <= 2 * pd::kRenderThreads,
             "renderer supports lattices of at most 2048 sites");
  int64_t need = 0;
  rcode = pd_render_workspace_bytes(image_size, &need);
  if (rcode != PD_OK) return rcode;
  if (m == 0) return PD_OK;
  PD_REQUIRE(frames_out != nullptr && st->image_params && st->frame_count,
             "null frames / state arrays");
  PD_REQUIRE(workspace != nullptr && workspace_bytes >= need,
             "workspace too small (pd_render_workspace_bytes)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  int log2_size = 0;
  while ((1 << log2_size) < image_size) ++log2_size;
  const size_t frame_floats = static_cast<size_t>(image_size) * image_size;
  for (int32_t first = 0; first < m; first += pd::kMaxFramesPerLaunch) {
    const int32_t count = m - first < pd::kMaxFramesPerLaunch
                              ? m - first : pd::kMaxFramesPerLaunch;
    pd::RenderArgs a{};
    a.lat = *lat;
    a.st = *st;
    a.env_ids = env_ids ? env_ids + first : nullptr;
    if (env_ids == nullptr && first > 0) {
      // envs [first, first + count): shift the per-env arrays
      a.st.env_offset += first;
      a.st.si_idx += first;
      a.st.lattice += 4 * static_cast<size_t>(first);
      a.st.fov += 4 * static_cast<size_t>(first);
      a.st.image_params += 9 * static_cast<size_t>(first);
      a.st.frame_count += first;
    }
    a.m = count;
    a.size = image_size;
    a.log2_size = log2_size;
    a.stop_stage = stop_stage;
    a.advance = advance_frame_count;
    a.buffer = buffer_size;
    a.bw = static_cast<int32_t>(buffer_size * image_size);
    a.out = frames_out + static_cast<size_t>(first) * frame_floats;
    a.n_generic = reinterpret_cast<int32_t*>(ws);
    a.generic = reinterpret_cast<uint8_t*>(ws + kFlagOffset);
    a.scratch = reinterpret_cast<float*>(ws + kScratchOffset);
    PD_CUDA_OK(cudaMemsetAsync(ws, 0, kFlagOffset + count, s));
    rcode = pd::launch_render_cluster(a, s);
    if (rcode != PD_OK) return rcode;
    const int smem = static_cast<int>(sizeof(pd::RenderShared));
    PD_CUDA_OK(cudaFuncSetAttribute(
        pd::k_render_generic, cudaFuncAttributeMaxDynamicSharedMemorySize,
        smem));
    const int grid = count < pd::kGenericGrid ? count : pd::kGenericGrid;
    pd::k_render_generic<<<grid, pd::kRenderThreads, smem, s>>>(a);
    PD_CUDA_OK(cudaGetLastError());
  }
  if (advance_frame_count) {
    pd::k_advance_frames<<<(m + 255) / 256, 256, 0, s>>>(*st, env_ids, m);
    PD_CUDA_OK(cudaGetLastError());
  }
  return PD_OK;
}
