// Micro-test (GPU box): does a K-major SWIZZLE_128B UMMA descriptor whose start address is offset by a
// non-multiple-of-8 number of 128-byte rows read the rows TMA-style (XOR by absolute address bits 7-9)?
// Tries base_offset = 0 and base_offset = (addr >> 7) & 7, and SBO = 1024 / 2560.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) test_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D,
                                                      int row0, int sbo_bytes, int use_base_off, int nrowsA) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - smem_u32(raw));
  uint8_t* sA = sm;                      // nrowsA rows x 128 B, swizzled by absolute address
  uint8_t* sB = sm + 64 * 1024;          // 64 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < nrowsA * 8; i += 128) {     // 16-byte chunks
    const int r = i >> 3, c = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)r * 64 + c * 8);
    *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  for (int i = threadIdx.x; i < 64 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(B + (size_t)r * 64 + c * 8);
    *reinterpret_cast<uint4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    auto desc = [&](uint32_t addr, uint32_t sbo, uint32_t boff) {
      uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
      d |= (uint64_t)1 << 16;
      d |= (uint64_t)(sbo >> 4) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)(boff & 7) << 49;
      d |= (uint64_t)2 << 61;
      return d;
    };
    const uint32_t a_addr = smem_u32(sA) + row0 * 128;
    const uint32_t boff = use_base_off ? ((a_addr >> 7) & 7) : 0;
    const uint64_t ad = desc(a_addr, sbo_bytes, boff), bd = desc(smem_u32(sB), 1024, 0);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      const uint32_t acc = k ? 1u : 0u;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  const int warp = threadIdx.x >> 5;
  for (int c0 = 0; c0 < 64; c0 += 32) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[(size_t)threadIdx.x * 64 + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
  const int NR = 400;
  std::vector<__nv_bfloat16> hA((size_t)NR * 64), hB(64 * 64);
  std::vector<float> fA((size_t)NR * 64), fB(64 * 64);
  uint32_t s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 9) & 0xFF) / 64.0f - 2.0f; };
  for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16(rnd()); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16(rnd()); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  const int row0s[] = {0, 8, 1, 3, 5, 21};
  const int sbos[] = {1024, 2560};
  for (int sbo : sbos)
    for (int boff = 0; boff < 2; ++boff)
      for (int row0 : row0s) {
        cudaMemset(dD, 0, 128 * 64 * 4);
        test_kernel<<<1, 128, 80 * 1024>>>(dA, dB, dD, row0, sbo, boff, NR);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("row0=%d sbo=%d boff=%d: CUDA error %s\n", row0, sbo, boff, cudaGetErrorString(e)); return 1; }
        std::vector<float> hD(128 * 64);
        cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        // expected: M row m = group g (m / 8), r = m % 8 -> smem row row0 + g * (sbo / 128) + r
        double maxerr = 0;
        for (int m = 0; m < 128; ++m) {
          const int ar = row0 + (m / 8) * (sbo / 128) + (m % 8);
          for (int n = 0; n < 64; ++n) {
            float acc = 0;
            for (int k = 0; k < 64; ++k) acc += fA[(size_t)ar * 64 + k] * fB[n * 64 + k];
            maxerr = fmax(maxerr, fabs(acc - hD[m * 64 + n]));
          }
        }
        printf("row0=%2d sbo=%4d base_offset_field=%s : max err %.4f %s\n", row0, sbo, boff ? "(addr>>7)&7" : "0", maxerr,
               maxerr < 0.05 ? "OK" : "MISMATCH");
      }
  return 0;
}
