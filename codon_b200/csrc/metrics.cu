// Driver post-processing and evaluation metrics on the GPU.
//
//   quantise : np.clip(out, 0, 1); (out * 255).astype(np.uint8)   CODON_X4/test.py:130,132.  The
//              reference evaluates this on a float16 array (test.py:52,127-128), so `via_half`
//              reproduces the fp16 rounding of the input and of the product before truncation.
//   rmse     : EvaluationResults, CODON_X4/test.py:148-164 -- RMSE in grey levels over the pixels
//              whose label is non-zero.  Integer-exact sums (uint64), one double per frame.
//   ssim     : ssim_exact, CODON_X4/ssim_2.py:36-52 -- Gaussian-window SSIM (sigma 1.5, scipy
//              gaussian_filter defaults: radius int(4*sigma+0.5) = 6, 'reflect' boundary, axis 0 then
//              axis 1), float64 throughout, mean of the SSIM map.  Deterministic fixed-order sums.
#include <cmath>
#include "common.cuh"
#include "kernels.h"

namespace codon {
namespace {

constexpr int kMaxRadius = 32;

__global__ void __launch_bounds__(256) quantise_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst,
                                                       size_t n, int via_half) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = src[i];
    if (via_half) {
      v = __half2float(__float2half_rn(v));
      v = fminf(fmaxf(v, 0.f), 1.f);
      v = __half2float(__float2half_rn(v * 255.f));
    } else {
      v = fminf(fmaxf(v, 0.f), 1.f) * 255.f;
    }
    dst[i] = (uint8_t)(int)v;   // truncation; NaN -> 0
  }
}

// one CTA per frame
__global__ void __launch_bounds__(1024) masked_rmse_kernel(const uint8_t* __restrict__ label,
                                                           const uint8_t* __restrict__ out, int HW,
                                                           double* __restrict__ rmse) {
  __shared__ unsigned long long s_sq[32], s_n[32];
  const uint8_t* l = label + (size_t)blockIdx.x * HW;
  const uint8_t* o = out + (size_t)blockIdx.x * HW;
  unsigned long long sq = 0, n = 0;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const int lv = l[i];
    if (lv != 0) { const int d = lv - (int)o[i]; sq += (unsigned long long)(d * d); ++n; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    sq += __shfl_xor_sync(0xffffffffu, sq, off);
    n += __shfl_xor_sync(0xffffffffu, n, off);
  }
  if ((threadIdx.x & 31) == 0) { s_sq[threadIdx.x >> 5] = sq; s_n[threadIdx.x >> 5] = n; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long tsq = 0, tn = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { tsq += s_sq[w]; tn += s_n[w]; }
    rmse[blockIdx.x] = sqrt((double)tsq / (double)tn);   // tn == 0 -> NaN, as the reference's 0/0 would raise
  }
}

struct GaussTaps { double w[2 * kMaxRadius + 1]; int radius; };

// scipy 'reflect': (d c b a | a b c d | d c b a)
__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (n == 1) return 0;
  const int period = 2 * n;
  i %= period;
  if (i < 0) i += period;
  return i < n ? i : period - 1 - i;
}

// pass 1: filter along axis 0 (rows) the five planes a, b, a*a, b*b, a*b  -> tmp [5][H][W]
__device__ __forceinline__ double img_val(const uint8_t* p, size_t i) { return (double)p[i] / 255.0; }
__device__ __forceinline__ double img_val(const double* p, size_t i) { return p[i]; }

template <typename T>
__global__ void __launch_bounds__(256) ssim_vertical_kernel(const T* __restrict__ A, const T* __restrict__ Bm,
                                                            int H, int W, const GaussTaps taps,
                                                            double* __restrict__ tmp) {
  const int b = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const T* a = A + (size_t)b * H * W;
  const T* bb = Bm + (size_t)b * H * W;
  double s[5] = {0, 0, 0, 0, 0};
  for (int k = -taps.radius; k <= taps.radius; ++k) {
    const int yy = reflect_idx(y + k, H);
    const double w = taps.w[k + taps.radius];
    const double va = img_val(a, (size_t)yy * W + x), vb = img_val(bb, (size_t)yy * W + x);
    s[0] += w * va; s[1] += w * vb; s[2] += w * (va * va); s[3] += w * (vb * vb); s[4] += w * (va * vb);
  }
  double* t = tmp + (size_t)b * 5 * H * W;
#pragma unroll
  for (int i = 0; i < 5; ++i) t[((size_t)i * H + y) * W + x] = s[i];
}

// pass 2: filter along axis 1, SSIM map, per-row partial sums (one CTA per row)
__global__ void __launch_bounds__(256) ssim_horizontal_kernel(const double* __restrict__ tmp, int H, int W,
                                                              const GaussTaps taps, double c1, double c2,
                                                              double* __restrict__ rowsum) {
  const int b = blockIdx.y, y = blockIdx.x;
  const double* t = tmp + (size_t)b * 5 * H * W;
  __shared__ double sh[256];
  double acc = 0.0;
  for (int x0 = 0; x0 < W; x0 += 256) {      // fixed order: x ascending per thread
    const int x = x0 + threadIdx.x;
    if (x < W) {
      double s[5] = {0, 0, 0, 0, 0};
      for (int k = -taps.radius; k <= taps.radius; ++k) {
        const int xx = reflect_idx(x + k, W);
        const double w = taps.w[k + taps.radius];
#pragma unroll
        for (int i = 0; i < 5; ++i) s[i] += w * t[((size_t)i * H + y) * W + xx];
      }
      const double mu1 = s[0], mu2 = s[1];
      const double s11 = s[2] - mu1 * mu1, s22 = s[3] - mu2 * mu2, s12 = s[4] - mu1 * mu2;
      const double num = (2 * mu1 * mu2 + c1) * (2 * s12 + c2);
      const double den = (mu1 * mu1 + mu2 * mu2 + c1) * (s11 + s22 + c2);
      acc += num / den;
    }
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) rowsum[(size_t)b * H + y] = sh[0];
}

__global__ void __launch_bounds__(256) ssim_final_kernel(const double* __restrict__ rowsum, int H, int W,
                                                         double* __restrict__ ssim) {
  __shared__ double sh[256];
  const int b = blockIdx.x;
  double acc = 0.0;
  for (int y = threadIdx.x; y < H; y += 256) acc += rowsum[(size_t)b * H + y];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) ssim[b] = sh[0] / ((double)H * (double)W);
}

}  // namespace

cudaError_t launch_quantise_u8(const float* src, uint8_t* dst, size_t n, int via_half, cudaStream_t st) {
  size_t g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g == 0) g = 1;
  quantise_kernel<<<(int)g, 256, 0, st>>>(src, dst, n, via_half);
  return cudaGetLastError();
}

cudaError_t launch_masked_rmse(const uint8_t* label, const uint8_t* out, int B, int HW, double* rmse, cudaStream_t st) {
  masked_rmse_kernel<<<B, 1024, 0, st>>>(label, out, HW, rmse);
  return cudaGetLastError();
}

size_t ssim_workspace_bytes(int B, int H, int W) { return ((size_t)5 * H * W + H) * B * sizeof(double); }

cudaError_t launch_ssim_gauss(const void* a, const void* b, int img_dtype, int B, int H, int W, double sd, double c1, double c2,
                              double* ssim, double* ws, cudaStream_t st) {
  GaussTaps taps;
  // scipy.ndimage._gaussian_kernel1d: radius = int(truncate * sd + 0.5), truncate = 4.0
  taps.radius = (int)(4.0 * sd + 0.5);
  if (taps.radius > kMaxRadius || taps.radius < 0) return cudaErrorInvalidValue;
  double sum = 0.0;
  for (int k = -taps.radius; k <= taps.radius; ++k) {
    taps.w[k + taps.radius] = std::exp(-0.5 / (sd * sd) * (double)(k * k));
    sum += taps.w[k + taps.radius];
  }
  for (int k = 0; k <= 2 * taps.radius; ++k) taps.w[k] /= sum;
  double* tmp = ws;
  double* rowsum = ws + (size_t)5 * H * W * B;
  if (img_dtype == 0)
    ssim_vertical_kernel<uint8_t><<<dim3(cdiv(W, 256), H, B), 256, 0, st>>>((const uint8_t*)a, (const uint8_t*)b, H, W, taps, tmp);
  else
    ssim_vertical_kernel<double><<<dim3(cdiv(W, 256), H, B), 256, 0, st>>>((const double*)a, (const double*)b, H, W, taps, tmp);
  ssim_horizontal_kernel<<<dim3(H, B), 256, 0, st>>>(tmp, H, W, taps, c1, c2, rowsum);
  ssim_final_kernel<<<B, 256, 0, st>>>(rowsum, H, W, ssim);
  return cudaGetLastError();
}

}  // namespace codon
