// K3 (cluster form): STEM frame renderer, imaging.py:117-265
// generate_stem_image, one thread-block cluster of 8 CTAs per frame.
//
// CTA k of the cluster owns image rows [k S/8, (k+1) S/8) -- exactly one row
// of CLAHE tiles -- and keeps them in shared memory (float32, 128 KB for
// S = 512) from the Gaussian splat to the finished frame; the band goes to HBM
// once, with one bulk (TMA) shared->global store per CTA.  The image-wide
// reductions the reference performs between stages ("/ max" five times, the
// CLAHE input and output ranges) are exchanged through distributed shared
// memory: every CTA pushes its partial into the other seven and one
// cluster barrier publishes them.  Neighbouring tile rows' CLAHE maps (for
// the bilinear blend) are pulled the same way.
//
//   P0  atoms in view (graphene.py:600-644) -> pixel bins, Z^e weights;
//       Gaussian tables; 32-column strip lists; per-row jitter shifts
//   P1  imaging.py:117-173 clean image: direct splat of the truncated,
//       renormalised separable Gaussian, 12 rows x 32 columns per warp in
//       registers (+ the blur halo rows above and below the band)
//   P2  :212-214 blur (reflect), vertical with a register window, horizontal
//   P3  :199-203 Poisson by inverse CDF (float32 search, float64 when close)
//   P4  :188-196 jitter roll, :206-209 s&p, :217-218 gamma (table over the
//       Poisson counts), :231-236 uniform noise
//   P5  :221-228 exponential noise     P6  :176-185 Gaussian noise, clip
//   P7  :264 CLAHE: 14-bit quantise, tile histograms (shared atomics), clip +
//       CDF maps by one warp per tile
//   P8  bilinear blend of the tile maps     P9  rescale, bulk store
//
// Noise fields: include/pdune_b200.h (one Philox call per four consecutive
// pixels and stage).
#include <cooperative_groups.h>

#include "pd_render.cuh"

namespace cg = cooperative_groups;

namespace pd {

// -DPD_RENDER_PHASE_CLOCKS: thread 0 of every CTA adds the cycles spent up to
// each phase boundary into RenderArgs::scratch (as long long [grid][16]);
// profiles/prof_render_phases.py reads them.  Off in the product build.
#ifdef PD_RENDER_PHASE_CLOCKS
#define PD_PHASE(i)                                                   \
  do {                                                                \
    if (tid == 0) {                                                   \
      const long long now_ = clock64();                               \
      reinterpret_cast<long long*>(a.scratch)[blockIdx.x * 16 + (i)] += \
          now_ - phase_t0;                                            \
      phase_t0 = now_;                                                \
    }                                                                 \
  } while (0)
#else
#define PD_PHASE(i) do { } while (0)
#endif

constexpr int kCluster = 8;          // CTAs per frame = CLAHE tile rows
constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;
constexpr int kFastAtoms = 1024;     // atoms in view the shared tables hold
constexpr int kFastRadius = 128;     // clean-image kernel radius (4 sigma)
constexpr int kHalo = 4;             // blur radius: blur_amount < 1.125
constexpr int kStripCap = 256;
constexpr int kMaxStrips = 16;       // S / 32
constexpr int kAcc = 12;             // rows per register-accumulator group
constexpr int kKyHalf = kFastRadius + kAcc;      // zero-padded half width
constexpr int kKyLen = 2 * kKyHalf + 8;
constexpr int kPowTab = 1024;
constexpr int kStages = 8;
constexpr int kBias = 1024;           // atom rows / columns are stored + kBias

struct ClusterShared {
  uint2 atoms[kFastAtoms];           // x = row << 16 | col, y = Z^e (float)
  unsigned short strip[kMaxStrips][kStripCap];
  int strip_n[kMaxStrips];
  float kx[kFastRadius + 1];
  // kys[s][i] = ky_padded[i + s]: a 12-row window starts 16-byte aligned in
  // one of the four copies
  __align__(16) float kys[4][kKyLen];
  float kb[kHalo + 1];
  __align__(16) float inv_k1[kInvTable];  // 1 / (i + 1)
  double inv_kd[kInvTable];
  float pow_tab[kPowTab];
  int shift[64];
  int warp_count[2][kWarps];
  float red[2][kWarps];
  float cred[2][kStages][2][kCluster];  // [frame parity][stage][value][src]
  int n_atoms, lwy, lwx, lwb;
  int work[2];                       // dynamic tile / chunk counters (P1, P3)
  double tab_sum[12];                // partial sums of the Gaussian tables
  int hist[kTiles][kBins];
  unsigned short maps[3][kTiles][kBins];  // tile rows rank-1, rank, rank+1
};

struct Ctx {
  ClusterShared* sh;
  cg::cluster_group cl;
  int rank, tid, lane, warp, parity;
};

// max over the whole frame of two values; stage selects the mailbox.
__device__ __forceinline__ float2 frame_max2(Ctx& c, float a, float b,
                                             int stage) {
  ClusterShared& sh = *c.sh;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
    b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
  }
  if (c.lane == 0) {
    sh.red[0][c.warp] = a;
    sh.red[1][c.warp] = b;
  }
  __syncthreads();
  if (c.tid < kCluster) {
    float ra = sh.red[0][0], rb = sh.red[1][0];
    for (int w = 1; w < kWarps; ++w) {
      ra = fmaxf(ra, sh.red[0][w]);
      rb = fmaxf(rb, sh.red[1][w]);
    }
    float* dst = c.cl.map_shared_rank(&sh.cred[c.parity][stage][0][0], c.tid);
    dst[c.rank] = ra;
    dst[kCluster + c.rank] = rb;
  }
  c.cl.sync();
  float ra = sh.cred[c.parity][stage][0][0];
  float rb = sh.cred[c.parity][stage][1][0];
#pragma unroll
  for (int j = 1; j < kCluster; ++j) {
    ra = fmaxf(ra, sh.cred[c.parity][stage][0][j]);
    rb = fmaxf(rb, sh.cred[c.parity][stage][1][j]);
  }
  return make_float2(ra, rb);
}

__device__ __forceinline__ void bulk_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// Scales the band in place and sends it to out with bulk stores.
__device__ __forceinline__ void emit_band(Ctx& c, float4* band4, int n_groups,
                                          float scale, float* dst) {
  for (int g = c.tid; scale != 1.0f && g < n_groups; g += kThreads) {
    float4 v = band4[g];
    v.x *= scale;
    v.y *= scale;
    v.z *= scale;
    v.w *= scale;
    band4[g] = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (c.tid == 0) {
    const uint32_t bytes = static_cast<uint32_t>(n_groups) * 16u;
    const uint32_t chunk = 32768u;
    for (uint32_t off = 0; off < bytes; off += chunk) {
      const uint32_t nb = bytes - off < chunk ? bytes - off : chunk;
      const uint32_t src = static_cast<uint32_t>(__cvta_generic_to_shared(
          reinterpret_cast<char*>(band4) + off));
      asm volatile(
          "cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
              reinterpret_cast<char*>(dst) + off),
          "r"(src), "r"(nb)
          : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}

// Vertical pass of the blur, in place.  A thread owns `rows` consecutive rows
// of one column: the originals it needs from above (R rows) stay in its
// register window, the R originals below its range -- which the thread that
// owns the next range overwrites -- are fetched before the barrier.
template <int R>
__device__ __forceinline__ void blur_vertical(float* img, int S, int first,
                                              int rows, const float* kb,
                                              int col, bool on) {
  // img row j <-> image row r0 - kHalo + j; outputs rows [first, first+rows)
  float w[2 * R + 1], tail[R];
  const float* kk = kb;
  if (on) {
#pragma unroll
    for (int i = 0; i < 2 * R; ++i) w[i] = img[(first - R + i) * S + col];
#pragma unroll
    for (int i = 0; i < R; ++i) tail[i] = img[(first + rows + i) * S + col];
  }
  __syncthreads();
  if (!on) return;
  for (int j = 0; j < rows - R; ++j) {
    w[2 * R] = img[(first + j + R) * S + col];
    float s = kk[0] * w[R];
#pragma unroll
    for (int k = 1; k <= R; ++k) s += kk[k] * (w[R - k] + w[R + k]);
    img[(first + j) * S + col] = s;
#pragma unroll
    for (int i = 0; i < 2 * R; ++i) w[i] = w[i + 1];
  }
#pragma unroll
  for (int t = 0; t < R; ++t) {
    w[2 * R] = tail[t];
    float s = kk[0] * w[R];
#pragma unroll
    for (int k = 1; k <= R; ++k) s += kk[k] * (w[R - k] + w[R + k]);
    img[(first + rows - R + t) * S + col] = s;
#pragma unroll
    for (int i = 0; i < 2 * R; ++i) w[i] = w[i + 1];
  }
}

__global__ void __launch_bounds__(kThreads, 1)
    k_render_cluster(const RenderArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ClusterShared& sh = *reinterpret_cast<ClusterShared*>(smem_raw);
  float* img = reinterpret_cast<float*>(
      smem_raw + ((sizeof(ClusterShared) + 127) / 128) * 128);
  Ctx c{&sh, cg::this_cluster(), 0, static_cast<int>(threadIdx.x),
        static_cast<int>(threadIdx.x) & 31, static_cast<int>(threadIdx.x) >> 5,
        0};
  c.rank = static_cast<int>(c.cl.block_rank());
  const int tid = c.tid, lane = c.lane, warp = c.warp, rank = c.rank;
  const int S = a.size;
  const int mask = S - 1;
  const int B = S / kCluster;           // band rows = CLAHE tile size
  const int log2_ts = a.log2_size - 3;
  const int r0 = rank * B;
  const int n_groups = B * S / 4;       // float4 groups of the band
  const int groups_per_row = S / 4;
  const int log2_gpr = a.log2_size - 2;
  const int gpr_mask = groups_per_row - 1;
  const uint32_t g_base = static_cast<uint32_t>(r0) * groups_per_row;
  float4* band4 = reinterpret_cast<float4*>(img + kHalo * S);
  float* band = img + kHalo * S;
  const double2* base = reinterpret_cast<const double2*>(a.lat.base_xy);
  const int n_clusters = gridDim.x / kCluster;
  const int cluster_id = blockIdx.x / kCluster;

  for (int i = tid; i < kInvTable; i += kThreads) {
    sh.inv_kd[i] = i > 0 ? 1.0 / static_cast<double>(i) : 0.0;
    sh.inv_k1[i] = 1.0f / static_cast<float>(i + 1);
  }

#ifdef PD_RENDER_PHASE_CLOCKS
  long long phase_t0 = clock64();
#endif
  for (int f = cluster_id; f < a.m; f += n_clusters) {
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
    float* out = a.out + static_cast<size_t>(f) * S * S +
                 static_cast<size_t>(r0) * S;
    __syncthreads();
    PD_PHASE(10);

    // ---------------------------------------------------------------- P0
    const double fw = fv.urx - fv.llx, fh = fv.ury - fv.lly;
    int n_total;
    {
      // (a) which lattice sites are in view (graphene.py:600-644): the test
      //     only; bins are computed after compaction
      bool keep[2];
#pragma unroll
      for (int round = 0; round < 2; ++round) {
        const int k = round * kThreads + tid;
        keep[round] = false;
        if (k < a.lat.n_sites) {
          const double2 p = site_position(__ldg(base + k), lt);
          // with a buffer the exact test is on the normalised position, in
          // (d); this box is a little generous
          const double mx = (a.buffer > 0.0 ? a.buffer + 1e-9 : 0.0) * fw;
          const double my = (a.buffer > 0.0 ? a.buffer + 1e-9 : 0.0) * fh;
          keep[round] = fv.llx - mx <= p.x && p.x <= fv.urx + mx &&
                        fv.lly - my <= p.y && p.y <= fv.ury + my;
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep[round]);
        if (lane == 0) sh.warp_count[round][warp] = __popc(m);
      }
      // (b) meanwhile twelve warps start the Gaussian tables
      //     w[x] = exp(-0.5 x^2 / sigma^2) / sum (scipy _gaussian_kernel1d
      //     with radius int(4 sigma + 0.5)), four warps per table, and two
      //     more draw the per-row jitter shifts (imaging.py:192)
      double tab_v = 0.0, tab_v128 = 0.0;
      const int tw = warp - (kWarps - 12);
      const int which = tw >> 2;  // 0: rows, 1: columns, 2: blur
      if (tw >= 0) {
        // rows use the FOV width, columns its height (imaging.py:159-160)
        const double sigma = which == 0 ? S / (2.15 * fw)
                             : which == 1 ? S / (2.15 * fh) : blur_amount;
        int lw = static_cast<int>(4.0 * sigma + 0.5);
        if (which == 2 && !(blur_amount > 1e-15)) lw = -1;
        const int x = lane + 32 * (tw & 3);
        if (x == 0) (which == 0 ? sh.lwy : which == 1 ? sh.lwx : sh.lwb) = lw;
        if (which == 0)
          for (int i = x; i < 4 * kKyLen; i += 128) (&sh.kys[0][0])[i] = 0.f;
        if (which == 2 && x <= kHalo) sh.kb[x] = 0.f;
        const double a2 = -0.5 / (sigma * sigma);
        if (x <= lw) tab_v = exp(a2 * x * x);
        if (x == 0 && lw >= 128) tab_v128 = exp(a2 * 128.0 * 128.0);
        double part = (x == 0 ? tab_v : 2.0 * tab_v) + 2.0 * tab_v128;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
          part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) sh.tab_sum[tw] = part;
      } else if (tid < B) {
        const uint4 w =
            philox4x32_10(env, frame, r0 + tid, PD_STREAM_JITTER, seed);
        sh.shift[tid] = poisson_icdf_tab(jitter_rate, u24(w.x), sh.inv_kd) & mask;
      }
      __syncthreads();
      PD_PHASE(11);
      if (tw >= 0) {
        const int lw = which == 0 ? sh.lwy : which == 1 ? sh.lwx : sh.lwb;
        const int cap = which == 2 ? kHalo : kFastRadius;
        if (lw >= 0 && lw <= cap) {
          const double inv = 1.0 / (sh.tab_sum[4 * which] +
                                    sh.tab_sum[4 * which + 1] +
                                    sh.tab_sum[4 * which + 2] +
                                    sh.tab_sum[4 * which + 3]);
          const int x0 = lane + 32 * (tw & 3);
#pragma unroll
          for (int rep = 0; rep < 2; ++rep) {
            const int x = rep == 0 ? x0 : 128;
            if (rep == 1 && (x0 != 0 || lw < 128)) break;
            if (x > lw) break;
            const float t =
                static_cast<float>((rep == 0 ? tab_v : tab_v128) * inv);
            if (which == 0) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                // padded index of offset d is d + kKyHalf
                const int ip_ = kKyHalf + x - q, im_ = kKyHalf - x - q;
                if (ip_ >= 0) sh.kys[q][ip_] = t;
                if (im_ >= 0) sh.kys[q][im_] = t;
              }
            } else if (which == 1) {
              sh.kx[x] = t;
            } else {
              sh.kb[x] = t;
            }
          }
        }
      }
      // (c) offsets of every warp's survivors, in site order
      if (warp == 0) {
        const int c0 = sh.warp_count[0][lane], c1 = sh.warp_count[1][lane];
        int i0 = c0, i1 = c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t0 = __shfl_up_sync(0xffffffffu, i0, o);
          const int t1 = __shfl_up_sync(0xffffffffu, i1, o);
          if (lane >= o) {
            i0 += t0;
            i1 += t1;
          }
        }
        const int tot0 = __shfl_sync(0xffffffffu, i0, 31);
        const int tot1 = __shfl_sync(0xffffffffu, i1, 31);
        sh.warp_count[0][lane] = i0 - c0;
        sh.warp_count[1][lane] = tot0 + i1 - c1;
        if (lane == 0) sh.n_atoms = tot0 + tot1;
      }
      __syncthreads();
      n_total = sh.n_atoms;
      const bool fast = n_total <= kFastAtoms && sh.lwy <= kFastRadius &&
                        sh.lwx <= kFastRadius && sh.lwb <= kHalo;
      if (!fast) {  // same decision in all eight CTAs
        if (rank == 0 && tid == 0) {
          a.generic[f] = 1;
          atomicAdd(a.n_generic, 1);
        }
        continue;
      }
      unsigned short* site_list = &sh.strip[0][0];  // free until the strips
#pragma unroll
      for (int round = 0; round < 2; ++round) {
        const unsigned m = __ballot_sync(0xffffffffu, keep[round]);
        const int pos = sh.warp_count[round][warp] +
                        __popc(m & ((1u << lane) - 1u));
        if (keep[round])
          site_list[pos] = static_cast<unsigned short>(round * kThreads + tid);
      }
      __syncthreads();
      // (d) pixel bins and Z^e weights of the atoms in view
      if (tid < n_total) {
        const int k = site_list[tid];
        const double2 p = site_position(__ldg(base + k), lt);
        const AtomBin ab = atom_bin(p, fv, S, a.buffer, a.bw);
        const float wt = powf(k == si ? 14.0f : 6.0f, exponent);
        // an atom the exact test drops keeps its slot with zero weight, far
        // from every strip
        const uint32_t r16 = static_cast<uint32_t>(ab.keep ? ab.row + kBias : 0);
        const uint32_t c16 = static_cast<uint32_t>(ab.keep ? ab.col + kBias : 0);
        sh.atoms[tid] = make_uint2((r16 << 16) | c16,
                                   __float_as_uint(ab.keep ? wt : 0.f));
      }
      if (tid == 0) {
        sh.work[0] = 0;
        sh.work[1] = 0;
        bulk_store_wait_read();  // the previous frame has left the band
      }
      __syncthreads();
      PD_PHASE(12);
    }
    const int n_atoms = n_total;
    const int lwy = sh.lwy, lwx = sh.lwx, lwb = sh.lwb;
    const int hb = lwb > 0 ? lwb : 0;  // halo rows the blur reads
    // strip lists: warp s collects, in atom order, the atoms whose footprint
    // reaches columns [32 s, 32 s + 31] and the rows this CTA computes
    const int n_strips = S / 32;
    if (warp < n_strips) {
      const int c_lo = warp * 32 - lwx, c_hi = warp * 32 + 31 + lwx;
      const int r_lo = r0 - hb - lwy, r_hi = r0 + B - 1 + hb + lwy;
      int cnt = 0;
      for (int i0 = 0; i0 < n_atoms; i0 += 32) {
        const int i = i0 + lane;
        bool hit = false;
        if (i < n_atoms) {
          const uint32_t rcv = sh.atoms[i].x;
          const int ar = static_cast<int>(rcv >> 16) - kBias;
          const int ac = static_cast<int>(rcv & 0xffffu) - kBias;
          hit = ar >= r_lo && ar <= r_hi && ac >= c_lo && ac <= c_hi;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        const int pos = cnt + __popc(m & ((1u << lane) - 1u));
        if (hit && pos < kStripCap)
          sh.strip[warp][pos] = static_cast<unsigned short>(i);
        cnt += __popc(m);
      }
      if (lane == 0) sh.strip_n[warp] = cnt;
    }
    __syncthreads();

    PD_PHASE(0);
    // ---------------------------------------------------------------- P1
    float vmax = 0.f;
    {
      // tiles of 32 columns x kAcc rows, handed to the warps dynamically
      // (strips differ in how many atoms reach them)
      const int rows_total = B + 2 * hb;
      const int jb_first = kHalo - hb;
      const int jb1 = jb_first + rows_total;
      const int n_tiles = n_strips * ((rows_total + kAcc - 1) / kAcc);
      for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&sh.work[0], 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_tiles) break;
        const int strip = t % n_strips;
        const int j0 = jb_first + (t / n_strips) * kAcc;
        const int col = strip * 32 + lane;
        const int cnt = sh.strip_n[strip];
        const bool listed = cnt <= kStripCap;
        const int n_list = listed ? cnt : n_atoms;
        const int img_r0 = r0 - kHalo + j0;  // image row of acc[0]
        float acc[kAcc];
#pragma unroll
        for (int j = 0; j < kAcc; ++j) acc[j] = 0.f;
        for (int ii = 0; ii < n_list; ++ii) {
          const int i = listed ? sh.strip[strip][ii] : ii;
          const uint2 at = sh.atoms[i];
          const int ar = static_cast<int>(at.x >> 16) - kBias;
          const int d0 = img_r0 - ar;
          if (d0 > lwy || d0 + kAcc - 1 < -lwy) continue;
          int dc = col - (static_cast<int>(at.x & 0xffffu) - kBias);
          dc = dc < 0 ? -dc : dc;
          const float wx = dc <= lwx ? __uint_as_float(at.y) * sh.kx[dc] : 0.f;
          const int idx = d0 + kKyHalf;  // >= 0: d0 >= -lwy - kAcc + 1
          const float4* ky4 =
              reinterpret_cast<const float4*>(&sh.kys[idx & 3][idx & ~3]);
#pragma unroll
          for (int q = 0; q < kAcc / 4; ++q) {
            const float4 k4 = ky4[q];
            acc[4 * q + 0] += wx * k4.x;
            acc[4 * q + 1] += wx * k4.y;
            acc[4 * q + 2] += wx * k4.z;
            acc[4 * q + 3] += wx * k4.w;
          }
        }
#pragma unroll
        for (int j = 0; j < kAcc; ++j) {
          const int jb = j0 + j;
          const int r = img_r0 + j;
          if (jb < jb1 && r >= 0 && r < S) {
            img[jb * S + col] = acc[j];
            if (jb >= kHalo && jb < kHalo + B) vmax = fmaxf(vmax, acc[j]);
          }
        }
      }
    }
    float m_prev = 1.f;
    if (a.stop_stage == PD_RENDER_CLEAN || lwb <= 0)
      m_prev = frame_max2(c, vmax, 0.f, 0).x;  // max of the clean image
    else
      __syncthreads();
    if (a.stop_stage == PD_RENDER_CLEAN) {
      emit_band(c, band4, n_groups, 1.0f / m_prev, out);
      c.parity ^= 1;
      continue;
    }

    PD_PHASE(1);
    // ---------------------------------------------------------------- P2
    // gaussian_filter(image / max, blur, mode='reflect'): axis 0 then axis 1
    // (a radius-0 kernel is [1.0]: the identity).
    if (lwb > 0) {
      // reflected rows above the first and below the last image row
      if (rank == 0 || rank == kCluster - 1) {
        for (int i = tid; i < hb * S; i += kThreads) {
          const int k = i / S, col = i - k * S;
          if (rank == 0)
            img[(kHalo - 1 - k) * S + col] = img[(kHalo + k) * S + col];
          if (rank == kCluster - 1)
            img[(kHalo + B + k) * S + col] = img[(kHalo + B - 1 - k) * S + col];
        }
        __syncthreads();
      }
      {
        // parts x S threads, each at least kHalo rows
        int parts = kThreads / S;
        if (parts > B / kHalo) parts = B / kHalo;
        const int rows = B / parts;
        const int part = tid / S, col = tid & mask;
        const bool on = part < parts;
        const int first = kHalo + part * rows;
        switch (lwb) {
          case 1: blur_vertical<1>(img, S, first, rows, sh.kb, col, on); break;
          case 2: blur_vertical<2>(img, S, first, rows, sh.kb, col, on); break;
          case 3: blur_vertical<3>(img, S, first, rows, sh.kb, col, on); break;
          default: blur_vertical<4>(img, S, first, rows, sh.kb, col, on); break;
        }
      }
      __syncthreads();
      vmax = 0.f;
      for (int g0 = 0; g0 < n_groups; g0 += kThreads) {
        const int g = g0 + tid;
        const bool on = g < n_groups;
        float x[4 + 2 * kHalo];  // columns cg0 - 4 .. cg0 + 7 (kHalo == 4)
        if (on) {
          const int gc = g & gpr_mask;
          const float4 own = band4[g];
          // scipy 'reflect' at the row ends: d c b a | a b c d | d c b a
          const float4 rev = make_float4(own.w, own.z, own.y, own.x);
          const float4 lft = gc == 0 ? rev : band4[g - 1];
          const float4 rgt = gc == gpr_mask ? rev : band4[g + 1];
          x[0] = lft.x; x[1] = lft.y; x[2] = lft.z; x[3] = lft.w;
          x[4] = own.x; x[5] = own.y; x[6] = own.z; x[7] = own.w;
          x[8] = rgt.x; x[9] = rgt.y; x[10] = rgt.z; x[11] = rgt.w;
        }
        __syncthreads();
        if (on) {
          float o4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float s = sh.kb[0] * x[kHalo + j];
#pragma unroll
            for (int k = 1; k <= kHalo; ++k)  // kb is zero beyond lwb
              s += sh.kb[k] * (x[kHalo + j - k] + x[kHalo + j + k]);
            o4[j] = s;
            vmax = fmaxf(vmax, s);
          }
          band4[g] = make_float4(o4[0], o4[1], o4[2], o4[3]);
        }
      }
      m_prev = frame_max2(c, vmax, 0.f, 1).x;
    }
    if (a.stop_stage == PD_RENDER_BLUR) {
      emit_band(c, band4, n_groups, 1.0f / m_prev, out);
      c.parity ^= 1;
      continue;
    }

    PD_PHASE(2);
    // ---------------------------------------------------------------- P3
    {
      const float scale = poisson_mult / m_prev;
      vmax = 0.f;
      // 128-pixel chunks, handed to the warps dynamically (the search length
      // follows the brightness).  Taking chunks of the other bands through
      // distributed shared memory was measured and dropped: the remote
      // counter traffic cost more than the imbalance it removed (-4 %).
      const int n_chunks = n_groups / 32;
      for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&sh.work[1], 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_chunks) break;
        // chunk t = a 16-column x 8-row patch (smaller than the atom
        // spacing), so the 128 searches of a warp have similar lengths
        const int g = (((t >> (a.log2_size - 4)) * 8 + (lane >> 2))
                       << log2_gpr) +
                      (t & ((S >> 4) - 1)) * 4 + (lane & 3);
        const float4 v = band4[g];
        const uint4 w = philox4x32_10(env, frame, g_base + g,
                                      PD_STREAM_RENDER_POISSON, seed);
        const float lam[4] = {v.x * scale, v.y * scale, v.z * scale,
                              v.w * scale};
        const float u[4] = {u24(w.x), u24(w.y), u24(w.z), u24(w.w)};
        int k[4];
        poisson4(lam, u, sh.inv_k1, sh.inv_kd, k);
        const float4 o = make_float4(
            static_cast<float>(k[0]), static_cast<float>(k[1]),
            static_cast<float>(k[2]), static_cast<float>(k[3]));
        band4[g] = o;
        vmax = fmaxf(vmax, fmaxf(fmaxf(o.x, o.y), fmaxf(o.z, o.w)));
      }
      m_prev = frame_max2(c, vmax, 0.f, 2).x;
    }
    if (a.stop_stage == PD_RENDER_POISSON) {
      emit_band(c, band4, n_groups, 1.0f / m_prev, out);
      c.parity ^= 1;
      continue;
    }

    PD_PHASE(3);
    // ---------------------------------------------------------------- P4
    {
      const float inv = 1.0f / m_prev;
      const bool jitter_only = a.stop_stage == PD_RENDER_JITTER;
      // adjust_gamma over the (few) distinct counts
      const bool tabled = m_prev < static_cast<float>(kPowTab);
      if (tabled && tid <= static_cast<int>(m_prev))
        sh.pow_tab[tid] = powf(
            fminf(fmaxf(static_cast<float>(tid) * inv, 0.f), 1.f), gamma);
      __syncthreads();
      vmax = 0.f;
      for (int g0 = 0; g0 < n_groups; g0 += kThreads) {
        const int g = g0 + tid;
        const bool on = g < n_groups;
        float kf[4];
        if (on) {
          const int row = g >> log2_gpr;
          const int cg0 = (g & gpr_mask) * 4;
          const int sft = sh.shift[row];
          // np.roll(row, k): out[(j + k) % S] = in[j]
#pragma unroll
          for (int j = 0; j < 4; ++j)
            kf[j] = band[row * S + ((cg0 + j - sft) & mask)];
        }
        __syncthreads();
        if (on) {
          float o4[4];
          if (jitter_only) {
#pragma unroll
            for (int j = 0; j < 4; ++j) o4[j] = kf[j];
          } else {
            const uint4 ws = philox4x32_10(env, frame, g_base + g,
                                           PD_STREAM_RENDER_SP, seed);
            const uint4 wu = philox4x32_10(env, frame, g_base + g,
                                           PD_STREAM_RENDER_UNIFORM, seed);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t s = word_of(ws, j);
              float v;
              if (tabled)
                v = sh.pow_tab[static_cast<int>(kf[j])];
              else
                v = powf(fminf(fmaxf(kf[j] * inv, 0.f), 1.f), gamma);
              if (u24(s) <= sp_amount) v = (s & 255u) < 128u ? 1.0f : 0.0f;
              v += uniform_scale * u24(word_of(wu, j));
              o4[j] = v;
              vmax = fmaxf(vmax, v);
            }
          }
          band4[g] = make_float4(o4[0], o4[1], o4[2], o4[3]);
        }
      }
      if (jitter_only) {
        emit_band(c, band4, n_groups, inv, out);
        c.parity ^= 1;
        continue;
      }
      m_prev = frame_max2(c, vmax, 0.f, 3).x;
    }
    if (a.stop_stage == PD_RENDER_UNIFORM) {
      emit_band(c, band4, n_groups, 1.0f / m_prev, out);
      c.parity ^= 1;
      continue;
    }

    PD_PHASE(4);
    // ---------------------------------------------------------------- P5
    {
      const float inv = 1.0f / m_prev;
      vmax = 0.f;
      for (int g = tid; g < n_groups; g += kThreads) {
        float4 v = band4[g];
        const uint4 w = philox4x32_10(env, frame, g_base + g,
                                      PD_STREAM_RENDER_EXP, seed);
        // -log1p(-u): 1 - u24 is exact in float32
        v.x = v.x * inv - __logf(1.0f - u24(w.x)) * exp_lambda;
        v.y = v.y * inv - __logf(1.0f - u24(w.y)) * exp_lambda;
        v.z = v.z * inv - __logf(1.0f - u24(w.z)) * exp_lambda;
        v.w = v.w * inv - __logf(1.0f - u24(w.w)) * exp_lambda;
        band4[g] = v;
        vmax = fmaxf(vmax, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
      }
      m_prev = frame_max2(c, vmax, 0.f, 4).x;
    }
    if (a.stop_stage == PD_RENDER_EXPONENTIAL) {
      emit_band(c, band4, n_groups, 1.0f / m_prev, out);
      c.parity ^= 1;
      continue;
    }

    PD_PHASE(5);
    // ---------------------------------------------------------------- P6
    float g_min, g_max;
    {
      const float inv = 1.0f / m_prev;
      float lo = 1e30f, hi = -1e30f;
      for (int g = tid; g < n_groups; g += kThreads) {
        float4 v = band4[g];
        const uint4 w = philox4x32_10(env, frame, g_base + g,
                                      PD_STREAM_RENDER_GAUSS, seed);
        // Box-Muller; cos(2 pi u) = -cos(2 pi u - pi) keeps the fast
        // sin/cos argument in [-pi, pi)
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
        band4[g] = v;
        lo = fminf(lo, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
        hi = fmaxf(hi, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
      }
      const float2 mm = frame_max2(c, hi, -lo, 5);
      g_max = mm.x;
      g_min = -mm.y;
    }
    if (a.stop_stage == PD_RENDER_GAUSSIAN) {
      emit_band(c, band4, n_groups, 1.0f, out);
      c.parity ^= 1;
      continue;
    }

    PD_PHASE(6);
    // ---------------------------------------------------------------- P7
    const int ts = B;
    {
      for (int i = tid; i < kTiles * kBins; i += kThreads)
        (&sh.hist[0][0])[i] = 0;
      __syncthreads();
      const float range = g_max - g_min;
      const float q_scale = range > 0.f ? (kGray - 1) / range : 0.f;
      for (int g = tid; g < n_groups; g += kThreads) {
        const float4 v = band4[g];
        const int tc = ((g & gpr_mask) * 4) >> log2_ts;
        // np.round(rescale_intensity(img, out_range=(0, 16383)))
        const int b0 = static_cast<int>(rintf((v.x - g_min) * q_scale)) /
                       kBinSize;
        const int b1 = static_cast<int>(rintf((v.y - g_min) * q_scale)) /
                       kBinSize;
        const int b2 = static_cast<int>(rintf((v.z - g_min) * q_scale)) /
                       kBinSize;
        const int b3 = static_cast<int>(rintf((v.w - g_min) * q_scale)) /
                       kBinSize;
        atomicAdd(&sh.hist[tc][b0], 1);
        atomicAdd(&sh.hist[tc][b1], 1);
        atomicAdd(&sh.hist[tc][b2], 1);
        atomicAdd(&sh.hist[tc][b3], 1);
        band4[g] = make_float4(__int_as_float(b0), __int_as_float(b1),
                               __int_as_float(b2), __int_as_float(b3));
      }
      __syncthreads();
      PD_PHASE(13);
      if (warp < kTiles) {
        int clim = static_cast<int>(0.01 * ts * ts);
        if (clim < 1) clim = 1;
        clahe_tile_map_warp(sh.hist[warp], sh.maps[1][warp], clim, ts * ts,
                            lane);
      }
      c.cl.sync();
      PD_PHASE(14);
      // tile rows above and below, from the neighbouring CTAs
      {
        constexpr int kWords = kTiles * kBins / 2;  // uint32 words per tile row
        const uint32_t* own = reinterpret_cast<const uint32_t*>(&sh.maps[1][0][0]);
        if (rank > 0) {
          const uint32_t* src = c.cl.map_shared_rank(own, rank - 1);
          uint32_t* dst = reinterpret_cast<uint32_t*>(&sh.maps[0][0][0]);
          for (int i = tid; i < kWords; i += kThreads) dst[i] = src[i];
        }
        if (rank < kCluster - 1) {
          const uint32_t* src = c.cl.map_shared_rank(own, rank + 1);
          uint32_t* dst = reinterpret_cast<uint32_t*>(&sh.maps[2][0][0]);
          for (int i = tid; i < kWords; i += kThreads) dst[i] = src[i];
        }
      }
      __syncthreads();
    }

    PD_PHASE(7);
    // ---------------------------------------------------------------- P8
    int m_lo = 1 << 30, m_hi = -1;
    {
      const float inv_ts = 1.0f / ts;
      const int half = ts >> 1;
      for (int g = tid; g < n_groups; g += kThreads) {
        const float4 v = band4[g];
        const int bins[4] = {__float_as_int(v.x), __float_as_int(v.y),
                             __float_as_int(v.z), __float_as_int(v.w)};
        const int r = r0 + (g >> log2_gpr);
        const int cg0 = (g & gpr_mask) * 4;
        const int pr = r + half;
        const int bi = pr >> log2_ts;
        const float cy = (pr & (ts - 1)) * inv_ts;
        const int t0r = (bi - 1 < 0 ? 0 : bi - 1) - rank + 1;
        const int t1r = (bi > kTiles - 1 ? kTiles - 1 : bi) - rank + 1;
        int mv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int pc = cg0 + j + half;
          const int bj = pc >> log2_ts;
          const float cx = (pc & (ts - 1)) * inv_ts;
          const int t0c = bj - 1 < 0 ? 0 : bj - 1;
          const int t1c = bj > kTiles - 1 ? kTiles - 1 : bj;
          const int bin = bins[j];
          // result += (mapped * coeff).astype(float32), edges in ndindex order
          float acc = __fmul_rn(sh.maps[t0r][t0c][bin],
                                __fmul_rn(1.0f - cy, 1.0f - cx));
          acc = __fadd_rn(acc, __fmul_rn(sh.maps[t0r][t1c][bin],
                                         __fmul_rn(1.0f - cy, cx)));
          acc = __fadd_rn(acc, __fmul_rn(sh.maps[t1r][t0c][bin],
                                         __fmul_rn(cy, 1.0f - cx)));
          acc = __fadd_rn(acc, __fmul_rn(sh.maps[t1r][t1c][bin],
                                         __fmul_rn(cy, cx)));
          mv[j] = static_cast<int>(acc);  // astype(uint16) truncates
          m_lo = min(m_lo, mv[j]);
          m_hi = max(m_hi, mv[j]);
        }
        band4[g] = make_float4(static_cast<float>(mv[0]),
                               static_cast<float>(mv[1]),
                               static_cast<float>(mv[2]),
                               static_cast<float>(mv[3]));
      }
      const float2 mm = frame_max2(c, static_cast<float>(m_hi),
                                   -static_cast<float>(m_lo), 6);
      m_hi = static_cast<int>(mm.x);
      m_lo = -static_cast<int>(mm.y);
    }

    PD_PHASE(8);
    // ---------------------------------------------------------------- P9
    {
      // rescale_intensity: (v - min) / (max - min)
      const float denom = static_cast<float>(m_hi - m_lo);
      const float inv_d = denom > 0.f ? 1.0f / denom : 0.f;
      const float lo = static_cast<float>(m_lo);
      for (int g = tid; g < n_groups; g += kThreads) {
        float4 v = band4[g];
        if (denom > 0.f) {
          v.x = (v.x - lo) * inv_d;
          v.y = (v.y - lo) * inv_d;
          v.z = (v.z - lo) * inv_d;
          v.w = (v.w - lo) * inv_d;
        } else {
          v.x = fminf(fmaxf(v.x, 0.f), 1.f);
          v.y = fminf(fmaxf(v.y, 0.f), 1.f);
          v.z = fminf(fmaxf(v.z, 0.f), 1.f);
          v.w = fminf(fmaxf(v.w, 0.f), 1.f);
        }
        band4[g] = v;
      }
      emit_band(c, band4, n_groups, 1.0f, out);
      c.parity ^= 1;
    }
    PD_PHASE(9);
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  // no CTA may exit while a neighbour can still read its shared memory
  c.cl.sync();
}

static size_t cluster_smem_bytes(int image_size) {
  const size_t head = ((sizeof(ClusterShared) + 127) / 128) * 128;
  return head + static_cast<size_t>(image_size / kCluster + 2 * kHalo) *
                    image_size * sizeof(float);
}

static int cluster_config(int image_size, int n_clusters,
                          cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr,
                          cudaStream_t stream) {
  const size_t smem = cluster_smem_bytes(image_size);
  PD_CUDA_OK(cudaFuncSetAttribute(
      k_render_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize,
      static_cast<int>(smem)));
  *cfg = cudaLaunchConfig_t{};
  cfg->gridDim = dim3(static_cast<unsigned>(n_clusters * kCluster));
  cfg->blockDim = dim3(kThreads);
  cfg->dynamicSmemBytes = smem;
  cfg->stream = stream;
  attr->id = cudaLaunchAttributeClusterDimension;
  attr->val.clusterDim.x = kCluster;
  attr->val.clusterDim.y = 1;
  attr->val.clusterDim.z = 1;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
  return PD_OK;
}

// Clusters of 8 CTAs the device can hold at once (each CTA fills an SM).
int render_cluster_count(int image_size, int* out_clusters) {
  static int cached[16] = {0};
  int slot = 0;
  while ((64 << slot) < image_size) ++slot;
  if (cached[slot] == 0) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr;
    const int rc = cluster_config(image_size, sm_count() / kCluster, &cfg,
                                  &attr, nullptr);
    if (rc != PD_OK) return rc;
    int n = 0;
    PD_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, k_render_cluster, &cfg));
    cached[slot] = n > 0 ? n : 1;
  }
  *out_clusters = cached[slot];
  return PD_OK;
}

int launch_render_cluster(const RenderArgs& a, cudaStream_t stream) {
  int n_clusters = 0;
  int rc = render_cluster_count(a.size, &n_clusters);
  if (rc != PD_OK) return rc;
  if (n_clusters > a.m) n_clusters = a.m;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr;
  rc = cluster_config(a.size, n_clusters, &cfg, &attr, stream);
  if (rc != PD_OK) return rc;
  PD_CUDA_OK(cudaLaunchKernelEx(&cfg, k_render_cluster, a));
  return PD_OK;
}

}  // namespace pd
