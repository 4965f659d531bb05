// Micro-benchmark (GPU box): sustained cycles per tcgen05.mma (SS mode, K-major SWIZZLE_128B operands, kind::f16,
// K = 16) as a function of cta_group, M, N and the A-operand address pattern, with every SM busy.  It answers the
// question the conv kernels' profiles raise: is an M = 256 / N = 128 (or N = 64) instruction paced by the tensor
// pipe (M*N*K / 4096 MAC per clock and SM) or by the shared-memory operand fetch?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_rate_test tools/umma_rate_test.cu && tools/umma_rate_test
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

struct RateParams {
  int n;             // MMA N
  int m;             // MMA M (128 for cta_group::1, 256 for cta_group::2)
  int nacc;          // accumulators cycled through per group (as the conv kernels do)
  int iters;         // groups; one group = nacc x 4 MMAs (the four K = 16 steps of one 128-byte slab)
  int sbo;           // A 8-row-group stride in bytes (1024 = contiguous rows, 2560 = 20-pixel patch pitch)
  int shift;         // 1: the A start address walks over 5 x 5 tap shifts like the conv kernels
  int b_rows;        // B rows held by this CTA (N for cta_group::1, N/2 for cta_group::2)
  int commit_each;   // 1: tcgen05.commit after every group (stage release), as in the conv kernels
  int tf32;          // 1: kind::tf32 (K = 8 per instruction)
  int data;          // operand data: 0 dense random mantissas, 1 all zeros, 2 ReLU-like (half zeros, small bf16 values)
  int interleave;    // 1: K-step outer, accumulator inner (consecutive MMAs hit different accumulators)
  int ovh;           // issue-loop overhead elements added per block (what the conv kernels' issuer does between MMA groups):
                     // 1 mbarrier.test_wait, 2 mbarrier.try_wait, 4 tcgen05.fence::after_thread_sync, 8 a second commit,
                     // 16 a second try_wait, 32 a polling-style branch around the try_wait (call-free slow path)
};

template <int CG, int NACC, int ORDER>
__global__ void __launch_bounds__(128, 1) rate_kernel(RateParams p, unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - smem_u32(raw));
  const uint32_t s_a = base, s_b = base + 64 * 1024;       // A: 64 KB patch area, B: 128 KB ring
  __shared__ uint64_t bar_done, bar_dummy, bar_ready;
  __shared__ uint32_t tslot;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));

  // operands: finite 16-bit patterns around 1.0 (bf16 0x3F80 .. 0x3FFF); the values do not matter, the toggling does
  for (uint32_t i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (i + blockIdx.x * 7919u) * 2654435761u;
    uint32_t v = p.tf32 ? (0x3F800000u | (h & 0x007FE000u)) : (0x3F803F80u | (h & 0x007F007Fu));
    if (p.data == 1) v = 0;
    if (p.data == 2) {      // bf16 pairs: each half zero with probability 1/2, else 2^-3 .. 2^-1 with random mantissa
      const uint32_t lo = (h & 0x100u) ? 0u : (0x3E00u + ((h >> 9) & 0xFFu)), hi = (h & 0x1000000u) ? 0u : (0x3E00u + ((h >> 13) & 0xFFu));
      v = lo | (hi << 16);
    }
    reinterpret_cast<uint32_t*>(sm)[i] = v;
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_done)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_dummy)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_ready)) : "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_ready)) : "memory");   // phase 0 complete for good
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;

  long long t0 = 0, t1 = 0;
  if (threadIdx.x < 32 && rank == 0) {
    const uint64_t hi_a = ((uint64_t)1 << 16) | ((uint64_t)(p.sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint64_t hi_b = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint32_t fmt = p.tf32 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.n >> 3) << 17) | ((uint32_t)(p.m >> 4) << 24);
    const uint32_t b_stage = (uint32_t)p.b_rows * 128u;
    const int nstage = (int)(128u * 1024u / b_stage) > 8 ? 8 : (int)(128u * 1024u / b_stage);
    const uint32_t pitch = (uint32_t)p.sbo;
    const uint32_t sub_off = p.sbo == 1024 ? 16384u : 8u * 128u;     // second sub-tile: next 16 KB / 8 pixels to the right
    int tap = 0, st = 0;
    uint32_t fused_done = 0;
    t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
      const int dx = p.shift ? tap % 5 : 0, dy = p.shift ? tap / 5 : 0;
      const uint64_t ad0 = hi_a | (uint64_t)(((s_a + (uint32_t)dy * pitch + (uint32_t)dx * 128u) & 0x3FFFFu) >> 4);
      const uint64_t bd = hi_b | (uint64_t)(((s_b + (uint32_t)st * b_stage) & 0x3FFFFu) >> 4);
      if (p.ovh) {
        const uint32_t rb = smem_u32(&bar_ready);
        uint32_t ok = 1;
        if (p.ovh & 1) asm volatile("{\n\t.reg .pred q;\n\tmbarrier.test_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok) : "r"(rb), "r"(0) : "memory");
        if (p.ovh & 2) { if (!(p.ovh & 32) || !ok) mbar_wait(rb, 0); }
        if (p.ovh & 16) mbar_wait(rb, 0);
        if (p.ovh & 4) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      if (ORDER == 4) {
        // ONE asm block per issue block (cta_group::2, two accumulators x four K steps + the stage commit).  ovh bit 64:
        // the phase test of the next stage is the block's FIRST instruction and its result is consumed by the block's
        // LAST one, so that the poll's latency overlaps the MMA issue; the next iteration waits only if it failed.
        if (elect_one()) {
          const uint32_t rb = smem_u32(&bar_ready), cb = smem_u32(&bar_dummy);
          const uint64_t a0 = ad0, a1 = ad0 + (uint64_t)(sub_off >> 4);
          const uint32_t d0 = tmem, d1 = tmem + (uint32_t)p.n, acc = it ? 1u : 0u;
          if (p.ovh & 64) {
            if (!fused_done) mbar_wait(rb, 0);
#define MMA4(D, A) \
            "tcgen05.mma.cta_group::2.kind::f16 [" D "], " A ", %5, %6, p;\n\t" \
            "add.u64 a, " A ", 2;\n\tadd.u64 b, %5, 2;\n\ttcgen05.mma.cta_group::2.kind::f16 [" D "], a, b, %6, t;\n\t" \
            "add.u64 a, " A ", 4;\n\tadd.u64 b, %5, 4;\n\ttcgen05.mma.cta_group::2.kind::f16 [" D "], a, b, %6, t;\n\t" \
            "add.u64 a, " A ", 6;\n\tadd.u64 b, %5, 6;\n\ttcgen05.mma.cta_group::2.kind::f16 [" D "], a, b, %6, t;\n\t"
            asm volatile("{\n\t.reg .pred q, p, t;\n\t.reg .b64 a, b;\n\t"
                         "mbarrier.test_wait.parity.shared::cta.b64 q, [%8], %9;\n\t"
                         "setp.ne.b32 p, %7, 0;\n\tsetp.eq.u32 t, %6, %6;\n\t"
                         MMA4("%1", "%3") MMA4("%2", "%4")
                         "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%10], %11;\n\t"
                         "selp.u32 %0, 1, 0, q;\n\t}"
                         : "=r"(fused_done)
                         : "r"(d0), "r"(d1), "l"(a0), "l"(a1), "l"(bd), "r"(idesc), "r"(acc), "r"(rb), "r"(0), "r"(cb), "h"((uint16_t)3)
                         : "memory");
          } else {
            asm volatile("{\n\t.reg .pred p, t;\n\t.reg .b64 a, b;\n\t"
                         "setp.ne.b32 p, %7, 0;\n\tsetp.eq.u32 t, %6, %6;\n\t"
                         MMA4("%1", "%3") MMA4("%2", "%4")
                         "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%10], %11;\n\t"
                         "mov.u32 %0, 1;\n\t}"
                         : "=r"(fused_done)
                         : "r"(d0), "r"(d1), "l"(a0), "l"(a1), "l"(bd), "r"(idesc), "r"(acc), "r"(rb), "r"(0), "r"(cb), "h"((uint16_t)3)
                         : "memory");
#undef MMA4
          }
        }
        __syncwarp();
      } else
      if (elect_one()) {
#pragma unroll
        for (int q = 0; q < 4 * NACC; ++q) {
          const int j = ORDER ? q % NACC : q / 4, k = ORDER ? q / NACC : q % 4;
          const uint64_t ad = ad0 + (uint64_t)((j * sub_off) >> 4);
          const uint32_t d = tmem + (uint32_t)(j * p.n);
          const uint32_t acc = (it | k) ? 1u : 0u;
          if (ORDER == 3) {
          } else if (CG == 1)
            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}"
                         ::"r"(d), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(acc) : "memory");
          else
            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, q;\n\t}"
                         ::"r"(d), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(acc) : "memory");
          if (ORDER >= 2) {
            if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_dummy)) : "memory");
            else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                              ::"r"(smem_u32(&bar_dummy)), "h"((uint16_t)3) : "memory");
          }
        }
        if (p.ovh & 8) {
          if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_dummy)) : "memory");
          else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                            ::"r"(smem_u32(&bar_dummy)), "h"((uint16_t)3) : "memory");
        }
        if (p.commit_each) {
          if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_dummy)) : "memory");
          else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                            ::"r"(smem_u32(&bar_dummy)), "h"((uint16_t)3) : "memory");
        }
      }
      __syncwarp();
      if (++tap == 25) tap = 0;
      if (++st == nstage) st = 0;
    }
    if (elect_one()) {
      if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_done)) : "memory");
      else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                        ::"r"(smem_u32(&bar_done)), "h"((uint16_t)3) : "memory");
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar_done), 0);
  t1 = clock64();
  if (threadIdx.x == 0 && rank == 0) out[blockIdx.x / CG] = (unsigned long long)(t1 - t0);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < 32) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int CG, int NACC, int ORDER>
static int run(const char* label, RateParams p, int grid, unsigned long long* d_out) {
  const size_t smem = 193 * 1024 + 1024;
  cudaFuncSetAttribute(rate_kernel<CG, NACC, ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best_ms = 1e30f;
  std::vector<unsigned long long> h(grid);
  double cyc = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(d_out, 0, grid * sizeof(unsigned long long));
    cudaEventRecord(e0);
    cudaError_t le = cudaLaunchKernelEx(&cfg, rate_kernel<CG, NACC, ORDER>, p, d_out);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (le != cudaSuccess || e != cudaSuccess) {
      printf("%s: CUDA error %s / %s\n", label, cudaGetErrorString(le), cudaGetErrorString(e));
      return 1;
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best_ms) {
      best_ms = ms;
      cudaMemcpy(h.data(), d_out, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      std::vector<unsigned long long> v(h.begin(), h.begin() + grid / CG);
      std::sort(v.begin(), v.end());
      cyc = (double)v[v.size() / 2];
    }
  }
  const double mmas = (double)p.iters * p.nacc * 4;
  const int kk = p.tf32 ? 8 : 16;
  const double mac_per_sm = (double)p.m / CG * p.n * kk;                      // per instruction and SM
  const double nominal = mac_per_sm / (p.tf32 ? 2048.0 : 4096.0);
  const double flops = 2.0 * p.m * p.n * kk * mmas * (grid / CG);
  const double a_bytes = (double)p.m / CG * 32, b_bytes = (double)p.b_rows * 32;
  printf("%-44s cyc/MMA %7.1f  nominal %5.0f  eff %5.1f%%  smem B/clk %6.1f  kernel %.3f ms  %.0f TFLOP/s  (clk %.0f MHz)\n",
         label, cyc / mmas, nominal, 100.0 * nominal / (cyc / mmas), (a_bytes + b_bytes) / (cyc / mmas), best_ms,
         flops / (best_ms * 1e-3) / 1e12, cyc / (best_ms * 1e-3) / 1e6);
  return 0;
}

template <int CG>
static int sweep(int n, int iters, int sms, unsigned long long* d_out) {
  char label[160];
  int rc = 0;
  RateParams p = {};
  p.n = n; p.m = 128 * CG; p.b_rows = n / CG; p.sbo = 2560; p.shift = 1; p.commit_each = 1; p.data = 2;
  const int grid = CG == 2 ? (sms & ~1) : sms;
#define ONE(NACC, ORDER)                                                                                  \
  if (NACC * n <= 512) {                                                                                   \
    p.nacc = NACC; p.iters = iters * 2 / NACC; p.interleave = ORDER;                                      \
    snprintf(label, sizeof label, "bf16 cg%d M%d N%-3d nacc%d %s", CG, 128 * CG, n, NACC, ORDER == 0 ? "acc-outer" : ORDER == 1 ? "k-outer  " : ORDER == 2 ? "commit/MMA" : "commits only"); \
    rc |= run<CG, NACC, ORDER>(label, p, grid, d_out);                                                    \
  }
  ONE(1, 0) ONE(2, 0) ONE(2, 1) ONE(4, 0) ONE(4, 1) ONE(1, 2) ONE(1, 3) ONE(2, 2) ONE(2, 3)
#undef ONE
  return rc;
}

int main(int argc, char** argv) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = argc > 1 ? atoi(argv[1]) : 10000;
  unsigned long long* d_out;
  cudaMalloc(&d_out, 1024 * sizeof(unsigned long long));
  int rc = 0;
  if (argc > 2 && atoi(argv[2]) == 1) { rc |= sweep<2>(128, iters, sms, d_out); rc |= sweep<1>(128, iters, sms, d_out); return rc; }
  if (argc > 2 && atoi(argv[2]) == 3) {
    // A-patch row pitch (the descriptor's 8-row-group stride): 1024 = contiguous, 1536 = 12-pixel patch (8-pixel-wide
    // tile + 5x5 halo: the NACC = 1 geometry), 1280 (8 + 3x3 halo), 2560 = 20-pixel patch (NACC = 2), 2304 (16 + 3x3)
    for (int sbo : {1024, 1280, 1536, 2048, 2304, 2560, 3072}) {
      char label[160];
      RateParams p = {};
      p.n = 128; p.m = 256; p.b_rows = 64; p.sbo = sbo; p.shift = 1; p.commit_each = 1; p.data = 2;
      p.nacc = 1; p.iters = iters * 2;
      snprintf(label, sizeof label, "bf16 cg2 M256 N128 nacc1 sbo %d shift", sbo);
      rc |= run<2, 1, 0>(label, p, sms & ~1, d_out);
      p.shift = 0;
      snprintf(label, sizeof label, "bf16 cg2 M256 N128 nacc1 sbo %d fixed", sbo);
      rc |= run<2, 1, 0>(label, p, sms & ~1, d_out);
      p.n = 64; p.b_rows = 32; p.shift = 1;
      snprintf(label, sizeof label, "bf16 cg2 M256 N64  nacc1 sbo %d shift", sbo);
      rc |= run<2, 1, 0>(label, p, sms & ~1, d_out);
    }
    return rc;
  }
  if (argc > 2 && atoi(argv[2]) == 4) {
    // issue-loop overhead elements: cycles per BLOCK (4 MMAs of N = 128 = 256 cycles of tensor work, or no MMA at all)
    for (int with_mma = 1; with_mma >= 0; --with_mma)
      for (int ovh : {0, 1, 2, 4, 8, 16, 1 | 2 | 32, 2 | 4, 2 | 16, 2 | 4 | 8 | 16, 1 | 2 | 32 | 4 | 8 | 16}) {
        char label[160];
        RateParams p = {};
        p.n = 128; p.m = 256; p.b_rows = 64; p.sbo = 2560; p.shift = 1; p.commit_each = 1; p.data = 2; p.nacc = 1; p.iters = iters * 2;
        p.ovh = ovh;
        snprintf(label, sizeof label, "cg2 N128 4 MMAs/block %s ovh %2d (x4 = cyc/block)", with_mma ? "MMA " : "none", ovh);
        rc |= with_mma ? run<2, 1, 0>(label, p, sms & ~1, d_out) : run<2, 1, 3>(label, p, sms & ~1, d_out);
      }
    // one asm block per issue block, without / with the next stage's phase test fused into it
    for (int n : {64, 128})
      for (int ovh : {0, 64}) {
        char label[160];
        RateParams p = {};
        p.n = n; p.m = 256; p.b_rows = n / 2; p.sbo = 2560; p.shift = 1; p.commit_each = 0; p.data = 2; p.nacc = 2; p.iters = iters;
        p.ovh = ovh;
        snprintf(label, sizeof label, "cg2 N%d 8 MMAs in ONE asm block, test %s (x8 = cyc/block)", n, ovh ? "fused" : "none ");
        rc |= run<2, 2, 4>(label, p, sms & ~1, d_out);
      }
    // the same with N = 64 blocks of 8 MMAs (pair kernel border taps: 344 cycles of tensor work per block)
    for (int ovh : {0, 2, 2 | 4, 2 | 4 | 8 | 16, 1 | 2 | 32 | 4 | 8 | 16}) {
      char label[160];
      RateParams p = {};
      p.n = 64; p.m = 256; p.b_rows = 32; p.sbo = 2560; p.shift = 1; p.commit_each = 1; p.data = 2; p.nacc = 2; p.iters = iters;
      p.ovh = ovh;
      snprintf(label, sizeof label, "cg2 N64 8 MMAs/block MMA ovh %2d (x8 = cyc/block)", ovh);
      rc |= run<2, 2, 0>(label, p, sms & ~1, d_out);
    }
    return rc;
  }
  if (argc > 2 && atoi(argv[2]) == 2) { for (int n : {16, 32, 48, 64, 96}) { rc |= sweep<2>(n, iters, sms, d_out); rc |= sweep<1>(n, iters, sms, d_out); } return rc; }
  for (int n : {64, 128, 256}) rc |= sweep<2>(n, iters, sms, d_out);
  for (int n : {64, 128, 256}) rc |= sweep<1>(n, iters, sms, d_out);
  return rc;
}
