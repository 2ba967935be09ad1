// Microbenchmark behind the streamed host rollout (DESIGN section 4,
// k_rollout_pre<STREAM>): how fast can SMs move data over PCIe themselves,
// one direction and both at once, with plain 16-byte loads / stores and with
// bulk (TMA) copies through shared memory?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pcie_sm_copy pcie_sm_copy.cu
//   ./pcie_sm_copy
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                              \
  do {                                                                     \
    cudaError_t e_ = (x);                                                  \
    if (e_ != cudaSuccess) {                                               \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(1);                                                             \
    }                                                                      \
  } while (0)

constexpr int kThreads = 128;

// plain: 8 x 16-byte loads in flight per thread, then the stores
__global__ void __launch_bounds__(kThreads)
    k_copy_ldst(const uint4* __restrict__ src, uint4* __restrict__ dst,
                int64_t units) {
  const int64_t block = kThreads * 8;
  for (int64_t base = block * blockIdx.x; base < units;
       base += block * gridDim.x) {
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int64_t i = base + k * kThreads + threadIdx.x;
      if (i < units) v[k] = __ldcs(src + i);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int64_t i = base + k * kThreads + threadIdx.x;
      if (i < units) dst[i] = v[k];
    }
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// bulk: global -> shared (mbarrier complete_tx), shared -> global (bulk
// group), `piece` bytes per copy, two buffers of `piece` bytes per CTA
__global__ void __launch_bounds__(kThreads)
    k_copy_bulk(const unsigned char* __restrict__ src,
                unsigned char* __restrict__ dst, int64_t bytes, int piece) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar[2];
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(
          smem_u32(&bar[b])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int64_t pieces = bytes / piece;
  int it = 0;
  // software pipeline of depth 2: load piece i + 1 while piece i is stored
  int64_t p = blockIdx.x;
  auto issue_load = [&](int64_t q, int b) {
    asm volatile(
        "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
            smem_u32(&bar[b])),
        "r"(piece));
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + b * piece)),
        "l"(src + q * piece), "r"(piece), "r"(smem_u32(&bar[b]))
        : "memory");
  };
  if (p < pieces) issue_load(p, 0);
  for (; p < pieces; p += gridDim.x, ++it) {
    const int b = it & 1;
    const int64_t nxt = p + gridDim.x;
    if (nxt < pieces) {
      // buffer b^1 was stored two iterations ago: make sure that read is done
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      issue_load(nxt, b ^ 1);
    }
    const uint32_t parity = (it >> 1) & 1;
    uint32_t ok = 0;
    while (!ok)
      asm volatile(
          "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], "
          "%2; selp.u32 %0, 1, 0, p; }"
          : "=r"(ok)
          : "r"(smem_u32(&bar[b])), "r"(parity)
          : "memory");
    asm volatile(
        "cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
            dst + p * piece),
        "r"(smem_u32(sm + b * piece)), "r"(piece)
        : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static float time_ms(cudaStream_t s0, cudaStream_t s1, void (*f0)(cudaStream_t),
                     void (*f1)(cudaStream_t)) {
  cudaEvent_t a, b, j;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  CK(cudaEventCreateWithFlags(&j, cudaEventDisableTiming));
  float best = 1e9f;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a, s0));
    if (f1) {
      CK(cudaStreamWaitEvent(s1, a, 0));
      f1(s1);
      CK(cudaEventRecord(j, s1));
    }
    if (f0) f0(s0);
    if (f1) CK(cudaStreamWaitEvent(s0, j, 0));
    CK(cudaEventRecord(b, s0));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  return best;
}

static unsigned char *h_in, *h_out, *d_in, *d_out;
static int64_t g_bytes = 16 << 20;
static int g_ctas = 32, g_piece = 8192;

static void rd_ldst(cudaStream_t s) {
  k_copy_ldst<<<g_ctas, kThreads, 0, s>>>((const uint4*)h_in, (uint4*)d_in,
                                          g_bytes / 16);
}
static void wr_ldst(cudaStream_t s) {
  k_copy_ldst<<<g_ctas, kThreads, 0, s>>>((const uint4*)d_out, (uint4*)h_out,
                                          g_bytes / 16);
}
static void rd_bulk(cudaStream_t s) {
  k_copy_bulk<<<g_ctas, kThreads, 2 * g_piece, s>>>(h_in, d_in, g_bytes,
                                                    g_piece);
}
static void wr_bulk(cudaStream_t s) {
  k_copy_bulk<<<g_ctas, kThreads, 2 * g_piece, s>>>(d_out, h_out, g_bytes,
                                                    g_piece);
}
static void rd_ce(cudaStream_t s) {
  CK(cudaMemcpyAsync(d_in, h_in, g_bytes, cudaMemcpyHostToDevice, s));
}
static void wr_ce(cudaStream_t s) {
  CK(cudaMemcpyAsync(h_out, d_out, g_bytes, cudaMemcpyDeviceToHost, s));
}
// H2D in g_chunks pieces, a 4-byte "rows arrived" copy behind each (what a
// copy-engine producer for a running kernel would do)
static int g_chunks = 16;
static uint32_t *h_flag_src, *d_flag;
static void rd_ce_flags(cudaStream_t s) {
  const int64_t piece = g_bytes / g_chunks;
  for (int c = 0; c < g_chunks; ++c) {
    CK(cudaMemcpyAsync(d_in + c * piece, h_in + c * piece, piece,
                       cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_flag, h_flag_src + c, 4, cudaMemcpyHostToDevice, s));
  }
}

int main() {
  CK(cudaHostAlloc(&h_in, g_bytes, cudaHostAllocDefault));
  CK(cudaHostAlloc(&h_out, g_bytes, cudaHostAllocDefault));
  CK(cudaMalloc(&d_in, g_bytes));
  CK(cudaMalloc(&d_out, g_bytes));
  for (int64_t i = 0; i < g_bytes; ++i) h_in[i] = (unsigned char)(i * 7 + 3);
  CK(cudaMemset(d_out, 5, g_bytes));
  CK(cudaFuncSetAttribute(k_copy_bulk,
                          cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
  cudaStream_t s0, s1;
  CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
  const double mb = g_bytes / 1e6;
  auto report = [&](const char* name, float ms, int dirs) {
    printf("%-34s %7.3f ms  %6.1f GB/s per direction%s\n", name, ms,
           mb / ms, dirs == 2 ? " (both busy)" : "");
  };
  report("copy engine H2D", time_ms(s0, s1, rd_ce, nullptr), 1);
  report("copy engine D2H", time_ms(s0, s1, wr_ce, nullptr), 1);
  report("copy engine duplex", time_ms(s0, s1, rd_ce, wr_ce), 2);
  CK(cudaHostAlloc(&h_flag_src, 4096, cudaHostAllocDefault));
  CK(cudaMalloc(&d_flag, 4));
  for (int i = 0; i < 1024; ++i) h_flag_src[i] = i + 1;
  for (int chunks : {8, 16, 32, 64}) {
    g_chunks = chunks;
    char nm[96];
    snprintf(nm, sizeof nm, "CE H2D, %d chunks + 4 B flags", chunks);
    report(nm, time_ms(s0, s1, rd_ce_flags, nullptr), 1);
  }
  g_ctas = 32;
  report("CE H2D + ld/st write 32 CTAs", time_ms(s0, s1, rd_ce, wr_ldst), 2);
  report("ld/st read 32 CTAs + CE D2H", time_ms(s0, s1, rd_ldst, wr_ce), 2);
  g_chunks = 16;
  report("CE H2D 16 chunks+flags + ld/st write",
         time_ms(s0, s1, rd_ce_flags, wr_ldst), 2);
  for (int ctas : {16, 32, 64}) {
    g_ctas = ctas;
    char nm[96];
    snprintf(nm, sizeof nm, "ld/st read   %3d CTAs", ctas);
    report(nm, time_ms(s0, s1, rd_ldst, nullptr), 1);
    snprintf(nm, sizeof nm, "ld/st write  %3d CTAs", ctas);
    report(nm, time_ms(s0, s1, wr_ldst, nullptr), 1);
    snprintf(nm, sizeof nm, "ld/st duplex %3d + %3d CTAs", ctas, ctas);
    report(nm, time_ms(s0, s1, rd_ldst, wr_ldst), 2);
  }
  for (int piece : {2048, 8192, 32768}) {
    for (int ctas : {16, 64}) {
      g_ctas = ctas;
      g_piece = piece;
      char nm[96];
      snprintf(nm, sizeof nm, "bulk read   %3d CTAs x %5d B", ctas, piece);
      report(nm, time_ms(s0, s1, rd_bulk, nullptr), 1);
      snprintf(nm, sizeof nm, "bulk write  %3d CTAs x %5d B", ctas, piece);
      report(nm, time_ms(s0, s1, wr_bulk, nullptr), 1);
      snprintf(nm, sizeof nm, "bulk duplex %3d CTAs x %5d B", ctas, piece);
      report(nm, time_ms(s0, s1, rd_bulk, wr_bulk), 2);
      snprintf(nm, sizeof nm, "bulk read + ld/st write %3d CTAs", ctas);
      report(nm, time_ms(s0, s1, rd_bulk, wr_ldst), 2);
    }
  }
  // check the bulk copies moved the right bytes
  CK(cudaDeviceSynchronize());
  unsigned char* chk = (unsigned char*)malloc(g_bytes);
  CK(cudaMemcpy(chk, d_in, g_bytes, cudaMemcpyDeviceToHost));
  int64_t bad = 0;
  for (int64_t i = 0; i < g_bytes; ++i) bad += chk[i] != h_in[i];
  for (int64_t i = 0; i < g_bytes; ++i) bad += h_out[i] != 5;
  printf("mismatches: %lld\n", (long long)bad);
  return bad != 0;
}
