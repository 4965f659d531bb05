// fp32 CUDA-core direct convolution, NHWC, stride 1, zero padding KS/2, no bias.
//
// This is the parity-mode trunk (CODON_MODE_FP32): every multiply-add of the reference's
// nn.Conv2d layers (CODON_X4/CODON_x4.py:24-47) is an fp32 FFMA with fp32 accumulation, so the
// result differs from the reference's CPU fp32 forward only by summation order.
//
// Tiling: one CTA = 8 x 16 output pixels x 64 output channels, 128 threads; a thread owns one
// pixel column (8 pixels) x 8 channels = 64 accumulators.  The input patch (tile + halo) and
// the weights of all taps are staged in shared memory KC = 8 input channels at a time;
// A operands are broadcast 128-bit loads, B operands conflict-free 128-bit loads
// (16 LDS.128 per 256 FFMA).
#include <atomic>
#include "common.cuh"
#include "kernels.h"

namespace codon {

namespace {

constexpr int kTH = 8, kTW = 16, kKC = 8, kNC = 64;

struct DirectParams {
  ConvJob job[2];
  int B, H, W, Cin, Cout, tiles_x, nchunks, relu;
};

template <int KS>
__global__ void __launch_bounds__(128, 3) conv_direct_f32_kernel(const DirectParams p) {
  constexpr int PAD = KS / 2, PH = kTH + KS - 1, PW = kTW + KS - 1, T = KS * KS;
  extern __shared__ __align__(16) float smem[];
  float* sp = smem;                       // [PH][PW][KC]
  float* sw = smem + PH * PW * kKC;       // [T][KC][64]

  const int tid = threadIdx.x, tn = tid & 7, tm = tid >> 3;
  const int jidx = blockIdx.y / p.nchunks, nch = blockIdx.y % p.nchunks;
  const ConvJob& job = p.job[jidx];
  const int b = blockIdx.z;
  const int ty0 = (blockIdx.x / p.tiles_x) * kTH, tx0 = (blockIdx.x % p.tiles_x) * kTW;
  const float* __restrict__ in = static_cast<const float*>(job.in);
  const float* __restrict__ wg = static_cast<const float*>(job.w);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int c0 = 0; c0 < p.Cin; c0 += kKC) {
    for (int idx = tid; idx < PH * PW * (kKC / 4); idx += 128) {
      const int v = idx % (kKC / 4), px = idx / (kKC / 4);
      const int gy = ty0 + px / PW - PAD, gx = tx0 + px % PW - PAD;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)
        val = __ldg(reinterpret_cast<const float4*>(
            in + (size_t)((size_t)(b * p.H + gy) * p.W + gx) * job.in_stride + job.in_off + c0 + v * 4));
      *reinterpret_cast<float4*>(sp + px * kKC + v * 4) = val;
    }
    for (int idx = tid; idx < T * kKC * (kNC / 4); idx += 128) {
      const int v = idx % (kNC / 4), k = (idx / (kNC / 4)) % kKC, t = idx / ((kNC / 4) * kKC);
      const float4 val = __ldg(reinterpret_cast<const float4*>(
          wg + (size_t)(t * p.Cin + c0 + k) * p.Cout + nch * kNC + v * 4));
      *reinterpret_cast<float4*>(sw + (t * kKC + k) * kNC + v * 4) = val;
    }
    __syncthreads();
#pragma unroll 1
    for (int dy = 0; dy < KS; ++dy) {
#pragma unroll
      for (int dx = 0; dx < KS; ++dx) {
        const int t = dy * KS + dx;
#pragma unroll
        for (int k4 = 0; k4 < kKC / 4; ++k4) {
          float4 a[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            a[i] = *reinterpret_cast<const float4*>(sp + ((i + dy) * PW + tm + dx) * kKC + k4 * 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const float* wrow = sw + (t * kKC + k4 * 4 + kk) * kNC;
            const float4 b0 = *reinterpret_cast<const float4*>(wrow + tn * 4);
            const float4 b1 = *reinterpret_cast<const float4*>(wrow + 32 + tn * 4);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
              acc[i][0] = fmaf(av, b0.x, acc[i][0]);
              acc[i][1] = fmaf(av, b0.y, acc[i][1]);
              acc[i][2] = fmaf(av, b0.z, acc[i][2]);
              acc[i][3] = fmaf(av, b0.w, acc[i][3]);
              acc[i][4] = fmaf(av, b1.x, acc[i][4]);
              acc[i][5] = fmaf(av, b1.y, acc[i][5]);
              acc[i][6] = fmaf(av, b1.z, acc[i][6]);
              acc[i][7] = fmaf(av, b1.w, acc[i][7]);
            }
          }
        }
      }
    }
    __syncthreads();
  }

  float* __restrict__ out = static_cast<float*>(job.out);
  const float* __restrict__ res = static_cast<const float*>(job.res);
  const int gx = tx0 + tm;
  if (gx >= p.W) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gy = ty0 + i;
    if (gy >= p.H) break;
    const size_t pix = (size_t)(b * p.H + gy) * p.W + gx;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = nch * kNC + h * 32 + tn * 4;
      float4 v = make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
      if (p.relu) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      }
      if (res) {
        const float4 r = *reinterpret_cast<const float4*>(res + pix * job.res_stride + job.res_off + c);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      *reinterpret_cast<float4*>(out + pix * job.out_stride + job.out_off + c) = v;
    }
  }
}

template <int KS>
cudaError_t launch_ks(const DirectParams& p, int njobs, cudaStream_t st) {
  constexpr int PH = kTH + KS - 1, PW = kTW + KS - 1, T = KS * KS;
  const size_t smem = sizeof(float) * (PH * PW * kKC + T * kKC * kNC);
  static std::atomic<int> configured[64];      // function attributes are per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!configured[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(conv_direct_f32_kernel<KS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[dev].store(1, std::memory_order_release);
  }
  dim3 grid(p.tiles_x * cdiv(p.H, kTH), p.nchunks * njobs, p.B);
  conv_direct_f32_kernel<KS><<<grid, 128, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_conv_direct_f32(const ConvJob* jobs, int njobs, int B, int H, int W, int Cin,
                                   int Cout, int KS, bool relu, cudaStream_t st) {
  if (njobs < 1 || njobs > 2 || Cin % kKC || Cout % kNC) return cudaErrorInvalidValue;
  DirectParams p{};
  for (int i = 0; i < njobs; ++i) p.job[i] = jobs[i];
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.tiles_x = cdiv(W, kTW); p.nchunks = Cout / kNC; p.relu = relu ? 1 : 0;
  switch (KS) {
    case 1: return launch_ks<1>(p, njobs, st);
    case 3: return launch_ks<3>(p, njobs, st);
    case 5: return launch_ks<5>(p, njobs, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace codon
