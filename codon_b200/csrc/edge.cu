// Edge layers of the trunk that are not tensor-core shaped, plus layout/dtype utilities.
//   input / input_c : 1 -> 64, 3x3, ReLU   (CODON_X4/CODON_x4.py:24,31,68,71)   K = 9
//   output          : 64 -> 1, 3x3, + x    (CODON_X4/CODON_x4.py:47,130-131)    N = 1
// Both are HBM-bound: 128-bit coalesced NHWC stores / loads, weights in shared memory.
#include "common.cuh"
#include "kernels.h"

namespace codon {
namespace {

// ---------------------------------------------------------------------------------------------
// Both edge layers work on strips of kStrip consecutive pixels of one image row.  A thread owns one 16-byte
// channel vector (kVec channels) for the whole strip: its 9 x kVec weights live in registers, the 3 x (kStrip + 2)
// input window is loaded once and reused by the three horizontal taps, and the pixel / row arithmetic is done once
// per strip.  (The r01 kernels handled one pixel per thread group and were issue-bound at ~80 us per 640x480 frame,
// 7-10x their HBM time.)
constexpr int kStrip = 8;

// first conv: out is NHWC with 128 channels [depth | colour]; the LPP lanes of a strip write 128 * sizeof(T)
// contiguous bytes per pixel.
template <typename T>
__global__ void __launch_bounds__(256) conv_first_kernel(const float* __restrict__ x,
                                                         const float* __restrict__ y,
                                                         const float* __restrict__ w_d,
                                                         const float* __restrict__ w_c,
                                                         T* __restrict__ out, int B, int H, int W, int rnd_tf32) {
  constexpr int V = Act<T>::kVec, LPP = 128 / V, SPB = 256 / LPP;   // lanes per pixel, strips per CTA pass
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
  // 32-bit arithmetic: codon_forward bounds B*H*W below 2^30
  const uint32_t spr = (uint32_t)(W + kStrip - 1) / kStrip;          // strips per row
  const uint32_t nstrips = (uint32_t)B * H * spr;
  for (uint32_t sidx = blockIdx.x * SPB + sub; sidx < nstrips; sidx += gridDim.x * SPB) {
    const uint32_t row = sidx / spr;                                   // global row index n * H + gy
    const int x0 = (int)(sidx - row * spr) * kStrip, gy = (int)(row % (uint32_t)H);
    const float* src = img + (size_t)row * W;                          // this row of this frame
    float in[3][kStrip + 2];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = gy + dy - 1;
      const bool rok = yy >= 0 && yy < H;
#pragma unroll
      for (int i = 0; i < kStrip + 2; ++i) {
        const int xx = x0 + i - 1;
        in[dy][i] = (rok && xx >= 0 && xx < W) ? __ldg(src + (ptrdiff_t)(dy - 1) * W + xx) : 0.f;
      }
    }
    T* dst = out + ((size_t)row * W + x0) * 128 + c0;
#pragma unroll
    for (int i = 0; i < kStrip; ++i) {
      if (x0 + i >= W) break;
      float v[V];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float a = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) a = fmaf(in[t / 3][i + t % 3], w[t][j], a);
        v[j] = fmaxf(a, 0.f);
        if (rnd_tf32) v[j] = round_tf32(v[j]);
      }
      Act<T>::store(dst + (size_t)i * 128, v);
    }
  }
}

// last conv: the LPP lanes of a strip each reduce their kVec channels over the 9 taps for all kStrip pixels, then a
// fixed-order butterfly leaves lane l with the sum of pixel l (kStrip == LPP for 16-bit activations; fp32 has 16
// lanes, two per pixel); that lane adds the global residual and writes, so a strip is one coalesced 32-byte store.
template <typename T>
__global__ void __launch_bounds__(256) conv_last_kernel(const T* __restrict__ in, int in_stride,
                                                        const float* __restrict__ w,
                                                        const float* __restrict__ x,
                                                        float* __restrict__ out, int B, int H, int W) {
  constexpr int V = Act<T>::kVec, LPP = 64 / V, SPB = 256 / LPP;
  const int g = threadIdx.x % LPP, sub = threadIdx.x / LPP;
  float wr[9][V];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < V; ++j) wr[t][j] = __ldg(w + t * 64 + g * V + j);
  const uint32_t spr = (uint32_t)(W + kStrip - 1) / kStrip;
  const uint32_t nstrips = (uint32_t)B * H * spr;
  const uint32_t npass = (nstrips + SPB - 1) / SPB;            // whole warps stay in the loop (shuffles)
  for (uint32_t pass = blockIdx.x; pass < npass; pass += gridDim.x) {
    const uint32_t sidx = pass * SPB + sub;
    const bool live = sidx < nstrips;
    const uint32_t row = live ? sidx / spr : 0;
    const int x0 = live ? (int)(sidx - row * spr) * kStrip : 0, gy = (int)(row % (uint32_t)H);
    float acc[kStrip];
#pragma unroll
    for (int i = 0; i < kStrip; ++i) acc[i] = 0.f;
    if (live) {
      const T* src = in + ((size_t)row * W) * in_stride + g * V;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int yy = gy + dy - 1;
        if (yy < 0 || yy >= H) continue;
        // all kStrip + 2 loads of the row are issued back to back (clamped addresses, no branches between them);
        // out-of-image columns are zeroed afterwards
        const T* rowp = src + (ptrdiff_t)(dy - 1) * W * in_stride;
        typename Act<T>::Raw raw[kStrip + 2];
#pragma unroll
        for (int i = 0; i < kStrip + 2; ++i) {
          int xx = x0 + i - 1;
          xx = xx < 0 ? 0 : (xx >= W ? W - 1 : xx);
          raw[i] = Act<T>::ldg(rowp + (size_t)xx * in_stride);
        }
#pragma unroll
        for (int i = 0; i < kStrip + 2; ++i) {
          const int xx = x0 + i - 1;
          float v[V];
          Act<T>::unpack(raw[i], v);
          const float m = (xx >= 0 && xx < W) ? 1.f : 0.f;
          // input column i feeds output pixel i - dx for the horizontal taps dx = 0, 1, 2
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const int o = i - dx;
            if (o < 0 || o >= kStrip) continue;
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < V; ++j) a = fmaf(v[j], wr[dy * 3 + dx][j], a);
            acc[o] = fmaf(a, m, acc[o]);
          }
        }
      }
    }
    // reduce over the LPP lanes: lanes that differ in bit k exchange the half of the pixels they do not keep
    float r = 0.f;
    if (LPP == 8) {
      float a4[4], a2[2];
      const bool hi4 = (g & 4) != 0, hi2 = (g & 2) != 0, hi1 = (g & 1) != 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float keep = hi4 ? acc[4 + i] : acc[i], give = hi4 ? acc[i] : acc[4 + i];
        a4[i] = keep + __shfl_xor_sync(0xffffffffu, give, 4);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float keep = hi2 ? a4[2 + i] : a4[i], give = hi2 ? a4[i] : a4[2 + i];
        a2[i] = keep + __shfl_xor_sync(0xffffffffu, give, 2);
      }
      const float keep = hi1 ? a2[1] : a2[0], give = hi1 ? a2[0] : a2[1];
      r = keep + __shfl_xor_sync(0xffffffffu, give, 1);
    } else {
      // 16 lanes (fp32 activations): pair-sum first, then the same butterfly over lanes / 2
#pragma unroll
      for (int i = 0; i < kStrip; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 1);
      const int h = g >> 1;
      float a4[4], a2[2];
      const bool hi4 = (h & 4) != 0, hi2 = (h & 2) != 0, hi1 = (h & 1) != 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float keep = hi4 ? acc[4 + i] : acc[i], give = hi4 ? acc[i] : acc[4 + i];
        a4[i] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float keep = hi2 ? a4[2 + i] : a4[i], give = hi2 ? a4[i] : a4[2 + i];
        a2[i] = keep + __shfl_xor_sync(0xffffffffu, give, 4);
      }
      const float keep = hi1 ? a2[1] : a2[0], give = hi1 ? a2[0] : a2[1];
      r = keep + __shfl_xor_sync(0xffffffffu, give, 2);
    }
    const int px = LPP == 8 ? g : (g >> 1);
    const bool writer = LPP == 8 ? true : ((g & 1) == 0);
    if (live && writer && x0 + px < W) {
      const size_t pix = (size_t)row * W + x0 + px;
      out[pix] = r + x[pix];
    }
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
    tile[r][threadIdx.x] = (p < HW && c < C) ? Act<T>::get(src + ((size_t)b * HW + p) * stride + off + c) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    if (p < HW && c < C) dst[((size_t)b * C + c) * HW + p] = tile[threadIdx.x][r];
  }
}

inline int grid_for(size_t total, int block, int ctas_per_sm = 16) {
  size_t g = (total + block - 1) / block;
  const size_t cap = (size_t)148 * ctas_per_sm;
  return (int)(g < cap ? (g ? g : 1) : cap);
}
// The strip kernels preload 9 x kVec weights per thread: persistent CTAs (two fit per SM at 128 registers) loop over
// many strips so that the preload is paid once, not once per strip.
constexpr int kEdgeCtasPerSm = 2;

}  // namespace

cudaError_t launch_conv_first(const float* x, const float* y, const float* w_d, const float* w_c,
                              void* out, int act, int B, int H, int W, cudaStream_t st, int rnd_tf32) {
  const size_t strips = (size_t)B * H * cdiv(W, kStrip);
  if (act == ACT_F32)
    conv_first_kernel<float><<<grid_for(strips * 32, 256, kEdgeCtasPerSm), 256, 0, st>>>(x, y, w_d, w_c, (float*)out, B, H, W, rnd_tf32);
  else if (act == ACT_BF16)
    conv_first_kernel<__nv_bfloat16><<<grid_for(strips * 16, 256, kEdgeCtasPerSm), 256, 0, st>>>(x, y, w_d, w_c, (__nv_bfloat16*)out, B, H, W, 0);
  else if (act == ACT_SPLIT16)
    conv_first_kernel<split16><<<grid_for(strips * 16, 256, kEdgeCtasPerSm), 256, 0, st>>>(x, y, w_d, w_c, (split16*)out, B, H, W, 0);
  else
    conv_first_kernel<__half><<<grid_for(strips * 16, 256, kEdgeCtasPerSm), 256, 0, st>>>(x, y, w_d, w_c, (__half*)out, B, H, W, 0);
  return cudaGetLastError();
}

cudaError_t launch_conv_last(const void* in, int in_stride, int act, const float* w, const float* x,
                             float* out, int B, int H, int W, cudaStream_t st) {
  const size_t strips = (size_t)B * H * cdiv(W, kStrip);
  if (act == ACT_F32)
    conv_last_kernel<float><<<grid_for(strips * 16, 256, kEdgeCtasPerSm), 256, 0, st>>>((const float*)in, in_stride, w, x, out, B, H, W);
  else if (act == ACT_BF16)
    conv_last_kernel<__nv_bfloat16><<<grid_for(strips * 8, 256, kEdgeCtasPerSm), 256, 0, st>>>((const __nv_bfloat16*)in, in_stride, w, x, out, B, H, W);
  else if (act == ACT_SPLIT16)
    conv_last_kernel<split16><<<grid_for(strips * 8, 256, kEdgeCtasPerSm), 256, 0, st>>>((const split16*)in, in_stride, w, x, out, B, H, W);
  else
    conv_last_kernel<__half><<<grid_for(strips * 8, 256, kEdgeCtasPerSm), 256, 0, st>>>((const __half*)in, in_stride, w, x, out, B, H, W);
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
  else if (act == ACT_SPLIT16) nhwc_to_nchw_kernel<split16><<<grid, block, 0, st>>>((const split16*)src, stride, off, C, HW, dst);
  else nhwc_to_nchw_kernel<__half><<<grid, block, 0, st>>>((const __half*)src, stride, off, C, HW, dst);
  return cudaGetLastError();
}

}  // namespace codon
