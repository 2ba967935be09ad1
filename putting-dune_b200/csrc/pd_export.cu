// Trajectory export: every env's observation as the protobuf wire bytes of
// putting_dune.proto's MicroscopeObservation, produced on the device, and the
// host-side framing of per-env Trajectory records into a TFRecord stream.
//
//   putting_dune.proto:7-45          Point2D, Atom, AtomicGrid, BeamControl,
//                                    FieldOfView, MicroscopeObservation
//   putting_dune.proto:47-49         Trajectory
//   microscope_utils.py:72-131       AtomicGrid.to_proto
//   microscope_utils.py:180-230      BeamControl.to_proto
//   microscope_utils.py:496-501      MicroscopeFieldOfView.to_proto
//   microscope_utils.py:589-604      MicroscopeObservation.to_proto
//   microscope_utils.py:737-757      Trajectory.to_proto
//   io.py:65-82                      write_records (tf.io.TFRecordWriter)
//
// The messages are proto2 with `optional` scalar fields the reference always
// sets, so every field is present on the wire and every message except the
// repeated ones has a fixed size:
//   Point2D      0D x:f32 15 y:f32                                   10 B
//   Atom         08 Z 12 0A <Point2D>                                14 B
//   AtomicGrid   (0A 0E <Atom>) x M                                  16 M B
//   FieldOfView  0A 0A <Point2D> 12 0A <Point2D>                     24 B
//   BeamControl  0A 0A <Point2D> 15 dwell 1D kV 25 nA                27 B
//   Observation  0A len(16 M) <grid> 12 18 <fov> (1A 1B <ctl>) x C
//                25 elapsed:f32
// (image / label_image, fields 5 and 6, follow field 4 and are appended by
// the host when frames were rendered).  float64 -> float32 conversions round
// to nearest even, as the protobuf runtime's do.
//
// Compiled with -fmad=false (positions as get_atoms_in_bounds forms them).
#include <stdlib.h>
#include <string.h>

#include "pd_kmc.cuh"

namespace pd {

constexpr int kExportThreads = 128;  // 4 warps = 4 envs per CTA

__host__ __device__ inline int varint_len(uint32_t v) {
  int n = 1;
  while (v >= 128u) {
    v >>= 7;
    ++n;
  }
  return n;
}

__device__ __forceinline__ int put_varint(uint8_t* p, uint32_t v) {
  int n = 0;
  while (v >= 128u) {
    p[n++] = static_cast<uint8_t>(v | 0x80u);
    v >>= 7;
  }
  p[n++] = static_cast<uint8_t>(v);
  return n;
}

__device__ __forceinline__ void put_f32(uint8_t* p, float v) {
  const uint32_t u = __float_as_uint(v);
  p[0] = static_cast<uint8_t>(u);
  p[1] = static_cast<uint8_t>(u >> 8);
  p[2] = static_cast<uint8_t>(u >> 16);
  p[3] = static_cast<uint8_t>(u >> 24);
}

__device__ __forceinline__ void put_point(uint8_t* p, double x, double y) {
  p[0] = 0x0D;
  put_f32(p + 1, __double2float_rn(x));
  p[5] = 0x15;
  put_f32(p + 6, __double2float_rn(y));
}

// row_run_in_view divides by cos / sin of the lattice angle.
__device__ __forceinline__ bool rows_usable(const Lattice4& t) {
  return fabs(t.c) >= 1e-6 && fabs(t.s) >= 1e-6;
}

__host__ __device__ inline int64_t observation_bytes(int32_t atoms,
                                                     int32_t n_controls) {
  const uint32_t grid = 16u * static_cast<uint32_t>(atoms);
  return 1 + varint_len(grid) + grid + 26 + 29LL * n_controls + 5;
}

// Pass 1: atoms in view (graphene.py:600-644, inclusive bounds) -> record
// length.  One warp per env.
__global__ void __launch_bounds__(kExportThreads)
    k_obs_sizes(const pd_lattice lat, const pd_state st, int32_t n_controls,
                int32_t* __restrict__ out_atoms, int32_t* __restrict__ out_len) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = static_cast<int64_t>(gridDim.x) * (kExportThreads / 32);
  const double2* base = reinterpret_cast<const double2*>(lat.base_xy);
  for (int64_t e = blockIdx.x * (kExportThreads / 32) + (threadIdx.x >> 5);
       e < st.n_envs; e += warps) {
    const Lattice4 t = load_lattice4(st.lattice, e);
    const Fov4 f = load_fov4(st.fov, e);
    int count = 0;
    if (rows_usable(t)) {
      // one lane per lattice row: the atoms in view are one run per row
      const int ce = lat.n_cols - (lat.n_cols + 2) / 3;
      const int co = lat.n_cols - (lat.n_cols + 1) / 3;
      for (int j0 = 0;; j0 += 32) {
        const int j = j0 + lane;
        const int k0 = (j >> 1) * (ce + co) + (j & 1) * ce;
        const int cnt_row = (j & 1) ? co : ce;
        const bool row_ok = k0 + cnt_row <= lat.n_sites;
        int m_lo = 0, m_hi = -1;
        if (row_ok) row_run_in_view(base, k0, cnt_row, j, lat.n_cols, t, f, &m_lo,
                                    &m_hi);
        if (m_hi >= m_lo) count += m_hi - m_lo + 1;
        if (__ballot_sync(0xffffffffu, row_ok) != 0xffffffffu) break;
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1)
        count += __shfl_xor_sync(0xffffffffu, count, d);
    } else {
      int c_lo, c_hi;
      fov_chunk_range(lat, t, f, &c_lo, &c_hi);
      for (int k0 = 32 * c_lo; k0 < 32 * c_hi; k0 += 32) {
        const int k = k0 + lane;
        bool keep = false;
        if (k < lat.n_sites) {
          const double2 p = site_position(__ldg(base + k), t);
          keep = (f.llx <= p.x) && (p.x <= f.urx) && (f.lly <= p.y) &&
                 (p.y <= f.ury);
        }
        count += __popc(__ballot_sync(0xffffffffu, keep));
      }
    }
    if (lane == 0) {
      out_atoms[e] = count;
      out_len[e] = static_cast<int32_t>(observation_bytes(count, n_controls));
    }
  }
}

// Exclusive scan of the 16-byte-aligned record sizes (one CTA, eight
// consecutive records per thread and trip; a small fraction of the encode).
__global__ void __launch_bounds__(1024)
    k_obs_offsets(const int32_t* __restrict__ len, int64_t n,
                  int64_t* __restrict__ offsets) {
  constexpr int kPer = 8;
  __shared__ int64_t warp_sum[32];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t i0 = 0; i0 < n; i0 += 1024 * kPer) {
    const int64_t first = i0 + static_cast<int64_t>(threadIdx.x) * kPer;
    int64_t v[kPer];
    int64_t mine = 0;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      v[k] = first + k < n
                 ? ((static_cast<int64_t>(len[first + k]) + 15) & ~15LL)
                 : 0;
      mine += v[k];
    }
    int64_t x = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) warp_sum[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int64_t s = warp_sum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int64_t y = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= d) s += y;
      }
      warp_sum[lane] = s;
    }
    __syncthreads();
    int64_t run = carry + (wid ? warp_sum[wid - 1] : 0) + x - mine;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      if (first + k < n) offsets[first + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry = run;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n] = carry;
}

// Pass 2: one warp per env builds the record in shared memory and streams it
// out in 16-byte words (records start 16-byte aligned in `out`).
//   A  in-view test over the candidate rows; the ids of the atoms in view are
//      compacted, in lattice order, into a list behind the record;
//   B  dense pass over the list: position, normalisation (two float64
//      divisions, as graphene.py:637-641 forms them), one 16-byte atom record
//      per lane as four aligned word stores -- the record is placed so that
//      the atom array (which follows a 2..4-byte header) is word aligned;
//   C  header, FOV, controls, elapsed time (byte stores, < 100 B);
//   D  copy-out: each 16-byte output word is funnel-shifted out of the
//      staged words to undo the placement offset.
__global__ void __launch_bounds__(kExportThreads)
    k_obs_encode(const pd_lattice lat, const pd_state st,
                 const double* __restrict__ controls_xy,
                 const int64_t* __restrict__ dwell_us,
                 int64_t dwell_us_scalar, int32_t n_controls,
                 const int64_t* __restrict__ elapsed_us, float voltage_kv,
                 float current_na, const int32_t* __restrict__ atoms,
                 const int32_t* __restrict__ len,
                 const int64_t* __restrict__ offsets, int32_t record_cap,
                 int32_t smem_per_warp, uint8_t* __restrict__ out,
                 int64_t capacity, uint8_t* __restrict__ overflow) {
  extern __shared__ __align__(16) uint8_t stage_all[];
  const int lane = threadIdx.x & 31;
  uint8_t* stage = stage_all + (threadIdx.x >> 5) * smem_per_warp;
  // ids of the atoms in view; behind the record (+16: placement shift, tail
  // word of the funnel shift)
  uint16_t* list = reinterpret_cast<uint16_t*>(stage + record_cap + 16);
  const int64_t warps = static_cast<int64_t>(gridDim.x) * (kExportThreads / 32);
  const double2* base = reinterpret_cast<const double2*>(lat.base_xy);
  for (int64_t e = blockIdx.x * (kExportThreads / 32) + (threadIdx.x >> 5);
       e < st.n_envs; e += warps) {
    const int32_t m_atoms = atoms[e];
    const int32_t bytes = len[e];
    const int64_t off = offsets[e];
    const int32_t padded = (bytes + 15) & ~15;
    if (padded > record_cap || off + padded > capacity) {
      if (lane == 0 && overflow) overflow[e] = 1;
      continue;
    }
    if (lane == 0 && overflow) overflow[e] = 0;
    const Lattice4 t = load_lattice4(st.lattice, e);
    const Fov4 f = load_fov4(st.fov, e);
    const int si = st.si_idx[e];
    const double w = __dsub_rn(f.urx, f.llx);
    const double h = __dsub_rn(f.ury, f.lly);
    const uint32_t grid_bytes = 16u * static_cast<uint32_t>(m_atoms);
    const int hdr = 1 + varint_len(grid_bytes);
    const int shift = (4 - (hdr & 3)) & 3;  // record starts at stage + shift
    uint8_t* rec = stage + shift;
    // ---- A: compact the ids of the atoms in view ----
    // (a lane-per-row variant of this pass -- runs from row_run_in_view, a
    // per-atom search of the row prefix in pass B -- was measured slower:
    // 324 vs 274 us per 65 536 records; the size pass does use the rows)
    int c_lo, c_hi;
    fov_chunk_range(lat, t, f, &c_lo, &c_hi);
    int count = 0;
#pragma unroll 2
    for (int k0 = 32 * c_lo; k0 < 32 * c_hi; k0 += 32) {
      const int k = k0 + lane;
      bool keep = false;
      if (k < lat.n_sites) {
        const double2 p = site_position(__ldg(base + k), t);
        keep = (f.llx <= p.x) && (p.x <= f.urx) && (f.lly <= p.y) &&
               (p.y <= f.ury);
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep)
        list[count + __popc(m & ((1u << lane) - 1u))] =
            static_cast<uint16_t>(k);
      count += __popc(m);
    }
    __syncwarp();
    // ---- B: atom records, AtomicGrid.atoms (field 1 of the grid) ----
    uint32_t* arec = reinterpret_cast<uint32_t*>(rec + hdr);
    for (int i = lane; i < m_atoms; i += 32) {
      const int k = list[i];
      const double2 p = site_position(__ldg(base + k), t);
      const uint32_t x = __float_as_uint(
          __double2float_rn(__ddiv_rn(__dsub_rn(p.x, f.llx), w)));
      const uint32_t y = __float_as_uint(
          __double2float_rn(__ddiv_rn(__dsub_rn(p.y, f.lly), h)));
      const uint32_t z = k == si ? kSilicon : kCarbon;
      // 0A 0E 08 Z | 12 0A 0D x0 | x1 x2 x3 15 | y0 y1 y2 y3
      arec[4 * i + 0] = 0x0A | (0x0E << 8) | (0x08 << 16) | (z << 24);
      arec[4 * i + 1] = 0x12 | (0x0A << 8) | (0x0D << 16) | (x << 24);
      arec[4 * i + 2] = (x >> 8) | (0x15u << 24);
      arec[4 * i + 3] = y;
    }
    // ---- C: grid header, fov (2), controls (3), elapsed time (4) ----
    uint8_t* tail = rec + hdr + grid_bytes;
    if (lane == 0) {
      for (int i = 0; i < shift; ++i) stage[i] = 0;
      rec[0] = 0x0A;
      put_varint(rec + 1, grid_bytes);
      tail[0] = 0x12;
      tail[1] = 0x18;
      tail[2] = 0x0A;
      tail[3] = 0x0A;
      put_point(tail + 4, f.llx, f.lly);
      tail[14] = 0x12;
      tail[15] = 0x0A;
      put_point(tail + 16, f.urx, f.ury);
      const int64_t us = elapsed_us ? elapsed_us[e] : st.sim_time_us[e];
      uint8_t* z = tail + 26 + 29 * n_controls;
      z[0] = 0x25;
      // timedelta.total_seconds(): microseconds / 10**6, correctly rounded
      put_f32(z + 1, __double2float_rn(
                         __ddiv_rn(static_cast<double>(us), 1e6)));
    }
    for (int c = lane; c < n_controls; c += 32) {
      uint8_t* b = tail + 26 + 29 * c;
      const double2 xy =
          reinterpret_cast<const double2*>(controls_xy)[e * n_controls + c];
      const int64_t us = dwell_us ? dwell_us[e * n_controls + c] : dwell_us_scalar;
      b[0] = 0x1A;
      b[1] = 0x1B;
      b[2] = 0x0A;
      b[3] = 0x0A;
      put_point(b + 4, xy.x, xy.y);
      b[14] = 0x15;
      put_f32(b + 15, __double2float_rn(
                          __ddiv_rn(static_cast<double>(us), 1e6)));
      b[19] = 0x1D;
      put_f32(b + 20, voltage_kv);
      b[24] = 0x25;
      put_f32(b + 25, current_na);
    }
    // zero the padding and the word the last funnel shift reads
    for (int i = shift + bytes + lane; i < padded + 8; i += 32) stage[i] = 0;
    __syncwarp();
    // ---- D: copy out ----
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(stage);
    uint4* dst = reinterpret_cast<uint4*>(out + off);
    const int sh = 8 * shift;
    for (int i = lane; i < padded / 16; i += 32) {
      const uint32_t w0 = sw[4 * i], w1 = sw[4 * i + 1], w2 = sw[4 * i + 2],
                     w3 = sw[4 * i + 3], w4 = sw[4 * i + 4];
      dst[i] = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh),
                          __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
    }
    __syncwarp();
  }
}

int validate_common(const pd_lattice* lat, const pd_state* st,
                    const pd_rate_config* rc);

// ---- CRC-32C (Castagnoli), the checksum of the TFRecord framing ----
static uint32_t g_crc_table[8][256];
static bool g_crc_ready = false;

static void crc_init() {
  if (g_crc_ready) return;
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
    g_crc_table[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int s = 1; s < 8; ++s)
      g_crc_table[s][i] =
          (g_crc_table[s - 1][i] >> 8) ^ g_crc_table[0][g_crc_table[s - 1][i] & 255u];
  g_crc_ready = true;
}

// slicing-by-8
static uint32_t crc32c_update(uint32_t crc, const uint8_t* p, size_t n) {
  crc = ~crc;
  while (n && (reinterpret_cast<uintptr_t>(p) & 7u)) {
    crc = (crc >> 8) ^ g_crc_table[0][(crc ^ *p++) & 255u];
    --n;
  }
  while (n >= 8) {
    uint64_t v;
    memcpy(&v, p, 8);
    v ^= crc;
    crc = g_crc_table[7][v & 255u] ^ g_crc_table[6][(v >> 8) & 255u] ^
          g_crc_table[5][(v >> 16) & 255u] ^ g_crc_table[4][(v >> 24) & 255u] ^
          g_crc_table[3][(v >> 32) & 255u] ^ g_crc_table[2][(v >> 40) & 255u] ^
          g_crc_table[1][(v >> 48) & 255u] ^ g_crc_table[0][v >> 56];
    p += 8;
    n -= 8;
  }
  while (n--) crc = (crc >> 8) ^ g_crc_table[0][(crc ^ *p++) & 255u];
  return ~crc;
}

static uint32_t masked_crc(uint32_t crc) {
  return ((crc >> 15) | (crc << 17)) + 0xA282EAD8u;
}

static void put_le32(uint8_t* p, uint32_t v) {
  for (int i = 0; i < 4; ++i) p[i] = static_cast<uint8_t>(v >> (8 * i));
}

static int host_put_varint(uint8_t* p, uint64_t v) {
  int n = 0;
  while (v >= 128u) {
    if (p) p[n] = static_cast<uint8_t>(v | 0x80u);
    ++n;
    v >>= 7;
  }
  if (p) p[n] = static_cast<uint8_t>(v);
  return n + 1;
}

}  // namespace pd

extern "C" int64_t pd_observation_bytes(int32_t atoms, int32_t n_controls) {
  return pd::observation_bytes(atoms, n_controls);
}

extern "C" int pd_encode_observations(
    const pd_lattice* lat, const pd_state* st, const double* controls_xy,
    const int64_t* dwell_us, int64_t dwell_us_scalar, int32_t n_controls,
    const int64_t* elapsed_us, float voltage_kv, float current_na,
    int32_t max_atoms, uint8_t* out_bytes, int64_t capacity,
    int64_t* out_offsets, int32_t* out_len, int32_t* out_atoms,
    uint8_t* out_overflow, void* stream) {
  int rcode = pd::validate_common(lat, st, nullptr);
  if (rcode != PD_OK) return rcode;
  PD_REQUIRE(n_controls >= 0 && (n_controls == 0 || controls_xy),
             "bad controls");
  PD_REQUIRE(out_bytes && out_offsets && out_len && out_atoms, "null outputs");
  PD_REQUIRE(max_atoms >= 0 && capacity >= 0, "bad capacity");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t n = st->n_envs;
  const int warps_per_cta = pd::kExportThreads / 32;
  const int64_t want = (n + warps_per_cta - 1) / warps_per_cta;
  const int64_t cap = static_cast<int64_t>(pd::sm_count()) * 16;
  const int grid = static_cast<int>(want < 1 ? 1 : (want < cap ? want : cap));
  if (n > 0) {
    pd::k_obs_sizes<<<grid, pd::kExportThreads, 0, s>>>(*lat, *st, n_controls,
                                                       out_atoms, out_len);
    PD_CUDA_OK(cudaGetLastError());
  }
  pd::k_obs_offsets<<<1, 1024, 0, s>>>(out_len, n, out_offsets);
  PD_CUDA_OK(cudaGetLastError());
  if (n == 0) return PD_OK;
  PD_REQUIRE(lat->n_sites <= 65535, "lattice too large for the export ids");
  const int64_t record_cap =
      (pd::observation_bytes(max_atoms, n_controls) + 15) & ~15LL;
  // record + placement shift / funnel tail + the list of atom ids
  const int64_t per_warp = (record_cap + 16 + 2LL * max_atoms + 15) & ~15LL;
  const int64_t smem = per_warp * warps_per_cta;
  PD_REQUIRE(smem <= 200 * 1024, "max_atoms too large for one staging buffer");
  PD_CUDA_OK(cudaFuncSetAttribute(pd::k_obs_encode,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
  pd::k_obs_encode<<<grid, pd::kExportThreads, smem, s>>>(
      *lat, *st, controls_xy, dwell_us, dwell_us_scalar, n_controls, elapsed_us,
      voltage_kv, current_na, out_atoms, out_len, out_offsets,
      static_cast<int32_t>(record_cap), static_cast<int32_t>(per_warp),
      out_bytes, capacity, out_overflow);
  PD_CUDA_OK(cudaGetLastError());
  return PD_OK;
}

extern "C" uint32_t pd_crc32c(const void* data, int64_t size) {
  pd::crc_init();
  return pd::crc32c_update(0u, static_cast<const uint8_t*>(data),
                           static_cast<size_t>(size));
}

// One TFRecord per env: the env's Trajectory = its observations in step
// order, each as field 1 (0A len bytes).
extern "C" int pd_tfrecord_trajectories(
    int32_t n_steps, int64_t n_envs, const uint8_t* const* step_bytes,
    const int64_t* const* step_offsets, const int32_t* const* step_len,
    uint8_t* out, int64_t out_capacity, int64_t* out_size) {
  PD_REQUIRE(n_steps >= 0 && n_envs >= 0 && out_size, "bad arguments");
  PD_REQUIRE(n_steps == 0 || (step_bytes && step_offsets && step_len),
             "null inputs");
  pd::crc_init();
  int64_t pos = 0;
  for (int64_t e = 0; e < n_envs; ++e) {
    uint64_t payload = 0;
    for (int t = 0; t < n_steps; ++t) {
      const uint64_t l = static_cast<uint64_t>(step_len[t][e]);
      payload += 1 + pd::host_put_varint(nullptr, l) + l;
    }
    const int64_t need = 8 + 4 + static_cast<int64_t>(payload) + 4;
    if (out && pos + need <= out_capacity) {
      uint8_t* p = out + pos;
      for (int i = 0; i < 8; ++i) p[i] = static_cast<uint8_t>(payload >> (8 * i));
      pd::put_le32(p + 8, pd::masked_crc(pd::crc32c_update(0u, p, 8)));
      uint8_t* d = p + 12;
      for (int t = 0; t < n_steps; ++t) {
        const uint64_t l = static_cast<uint64_t>(step_len[t][e]);
        *d++ = 0x0A;  // Trajectory.observations
        d += pd::host_put_varint(d, l);
        memcpy(d, step_bytes[t] + step_offsets[t][e], l);
        d += l;
      }
      pd::put_le32(d, pd::masked_crc(pd::crc32c_update(0u, p + 12, payload)));
    }
    pos += need;
  }
  *out_size = pos;
  if (out && pos > out_capacity) {
    pd::set_error("pd_tfrecord_trajectories: %lld bytes needed, capacity %lld",
                  static_cast<long long>(pos),
                  static_cast<long long>(out_capacity));
    return PD_ERR_INVALID_ARGUMENT;
  }
  return PD_OK;
}
