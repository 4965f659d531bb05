// Edge layers of the trunk that are not tensor-core shaped, plus layout/dtype utilities.
//   input / input_c : 1 -> 64, 3x3, ReLU   (CODON_X4/CODON_x4.py:24,31,68,71)   K = 9
//   output          : 64 -> 1, 3x3, + x    (CODON_X4/CODON_x4.py:47,130-131)    N = 1
// Both are HBM-bound: 128-bit coalesced NHWC stores / loads, weights in shared memory.
#include "common.cuh"
#include "kernels.h"

namespace codon {
namespace {

// ---------------------------------------------------------------------------------------------
// first conv: a thread owns one 16-byte channel vector (kVec channels of one branch) for ALL the pixels
// it visits, so its 9 x kVec weights live in registers (the r01 version re-read them from shared memory
// with 4-way bank conflicts and ran at ~0.8 TB/s).  out is NHWC with 128 channels: [depth | colour].
template <typename T>
__global__ void __launch_bounds__(256) conv_first_kernel(const float* __restrict__ x,
                                                         const float* __restrict__ y,
                                                         const float* __restrict__ w_d,
                                                         const float* __restrict__ w_c,
                                                         T* __restrict__ out, int B, int H, int W, int rnd_tf32) {
  constexpr int V = Act<T>::kVec, LPP = 128 / V, PPB = 256 / LPP;   // lanes per pixel, pixels per CTA pass
  const int g = threadIdx.x % LPP, sub = threadIdx.x / LPP;
  const int c0 = g * V, br = c0 >> 6, c = c0 & 63;
  float w[9][V];
  {
    const float* wsrc = br ? w_c : w_d;
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int j = 0; j < V; ++j) w[t][j] = __ldg(wsrc + t * 64 + c + j);
  }
  const float* img = br ? y : x;
  // 32-bit pixel arithmetic: codon_forward bounds B*H*W below 2^30 (64-bit div/mod dominated the r01 loop)
  const uint32_t npix = (uint32_t)B * H * W;
  for (uint32_t pix = blockIdx.x * PPB + sub; pix < npix; pix += gridDim.x * PPB) {
    const uint32_t row = pix / (uint32_t)W;
    const int gx = (int)(pix - row * W), gy = (int)(row % (uint32_t)H);
    const float* src = img + (size_t)(pix - (uint32_t)gy * W - gx);   // frame base
    float in[9];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int yy = gy + dy - 1, xx = gx + dx - 1;
        in[dy * 3 + dx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(src + (size_t)yy * W + xx) : 0.f;
      }
    float v[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float a = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) a = fmaf(in[t], w[t][j], a);
      v[j] = fmaxf(a, 0.f);
      if (rnd_tf32) v[j] = round_tf32(v[j]);
    }
    Act<T>::store(out + (size_t)pix * 128 + c0, v);
  }
}

// ---------------------------------------------------------------------------------------------
// last conv: LPP lanes cooperate on one pixel (each lane owns one 16-byte channel vector for all
// 9 taps, weights in registers), then a shuffle reduction; lane 0 of the group adds the global
// residual and writes.
template <typename T>
__global__ void __launch_bounds__(256) conv_last_kernel(const T* __restrict__ in, int in_stride,
                                                        const float* __restrict__ w,
                                                        const float* __restrict__ x,
                                                        float* __restrict__ out, int B, int H, int W) {
  constexpr int V = Act<T>::kVec, LPP = 64 / V, PPB = 256 / LPP;
  const int g = threadIdx.x % LPP, sub = threadIdx.x / LPP;
  float wr[9][V];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < V; ++j) wr[t][j] = __ldg(w + t * 64 + g * V + j);
  const uint32_t npix = (uint32_t)B * H * W;
  const uint32_t npass = (npix + PPB - 1) / PPB;            // whole warps stay in the loop (shuffles)
  for (uint32_t pass = blockIdx.x; pass < npass; pass += gridDim.x) {
    const uint32_t pix = pass * PPB + sub;
    float a = 0.f;
    if (pix < npix) {
      const uint32_t row = pix / (uint32_t)W;
      const int gx = (int)(pix - row * W), gy = (int)(row % (uint32_t)H);
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int yy = gy + dy - 1, xx = gx + dx - 1;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            float v[V];
            Act<T>::load(in + ((size_t)pix + (size_t)((dy - 1) * W + (dx - 1))) * in_stride + g * V, v);
#pragma unroll
            for (int j = 0; j < V; ++j) a = fmaf(v[j], wr[dy * 3 + dx][j], a);
          }
        }
    }
#pragma unroll
    for (int o = LPP / 2; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (g == 0 && pix < npix) out[pix] = a + x[pix];
  }
}

template <typename S>
__global__ void convert_to_f32_kernel(const S* __restrict__ src, float* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = Act<S>::to_float(src[i]);
}
template <typename D>
__global__ void convert_from_f32_kernel(const float* __restrict__ src, D* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = Act<D>::from_float(src[i]);
}

// NHWC slice -> NCHW fp32 through a 32x32 shared-memory transpose (debug taps only).
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, int stride, int off, int C, int HW,
                                    float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (p < HW && c < C) ? Act<T>::to_float(src[((size_t)b * HW + p) * stride + off + c]) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    if (p < HW && c < C) dst[((size_t)b * C + c) * HW + p] = tile[threadIdx.x][r];
  }
}

inline int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = 148 * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace

cudaError_t launch_conv_first(const float* x, const float* y, const float* w_d, const float* w_c,
                              void* out, int act, int B, int H, int W, cudaStream_t st, int rnd_tf32) {
  const size_t pix = (size_t)B * H * W;
  if (act == ACT_F32)
    conv_first_kernel<float><<<grid_for(pix * 32, 256), 256, 0, st>>>(x, y, w_d, w_c, (float*)out, B, H, W, rnd_tf32);
  else if (act == ACT_BF16)
    conv_first_kernel<__nv_bfloat16><<<grid_for(pix * 16, 256), 256, 0, st>>>(x, y, w_d, w_c, (__nv_bfloat16*)out, B, H, W, 0);
  else
    conv_first_kernel<__half><<<grid_for(pix * 16, 256), 256, 0, st>>>(x, y, w_d, w_c, (__half*)out, B, H, W, 0);
  return cudaGetLastError();
}

cudaError_t launch_conv_last(const void* in, int in_stride, int act, const float* w, const float* x,
                             float* out, int B, int H, int W, cudaStream_t st) {
  const size_t pix = (size_t)B * H * W;
  if (act == ACT_F32)
    conv_last_kernel<float><<<grid_for(pix * 16, 256), 256, 0, st>>>((const float*)in, in_stride, w, x, out, B, H, W);
  else if (act == ACT_BF16)
    conv_last_kernel<__nv_bfloat16><<<grid_for(pix * 8, 256), 256, 0, st>>>((const __nv_bfloat16*)in, in_stride, w, x, out, B, H, W);
  else
    conv_last_kernel<__half><<<grid_for(pix * 8, 256), 256, 0, st>>>((const __half*)in, in_stride, w, x, out, B, H, W);
  return cudaGetLastError();
}

cudaError_t launch_convert_to_f32(const void* src, int dtype, float* dst, size_t n, cudaStream_t st) {
  const int g = grid_for(n, 256);
  if (dtype == 0) return cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
  if (dtype == 1) convert_to_f32_kernel<__half><<<g, 256, 0, st>>>((const __half*)src, dst, n);
  else if (dtype == 2) convert_to_f32_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, dst, n);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_convert_from_f32(const float* src, void* dst, int dtype, size_t n, cudaStream_t st) {
  const int g = grid_for(n, 256);
  if (dtype == 0) return cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
  if (dtype == 1) convert_from_f32_kernel<__half><<<g, 256, 0, st>>>(src, (__half*)dst, n);
  else if (dtype == 2) convert_from_f32_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(src, (__nv_bfloat16*)dst, n);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_nhwc_to_nchw_f32(const void* src, int act, int stride, int off, int C, int B, int HW,
                                    float* dst, cudaStream_t st) {
  dim3 grid(cdiv(HW, 32), cdiv(C, 32), B), block(32, 8);
  if (act == ACT_F32) nhwc_to_nchw_kernel<float><<<grid, block, 0, st>>>((const float*)src, stride, off, C, HW, dst);
  else if (act == ACT_BF16) nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)src, stride, off, C, HW, dst);
  else nhwc_to_nchw_kernel<__half><<<grid, block, 0, st>>>((const __half*)src, stride, off, C, HW, dst);
  return cudaGetLastError();
}

}  // namespace codon
