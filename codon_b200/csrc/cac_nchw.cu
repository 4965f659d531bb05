// Stand-alone NCHW fp32 versions of the attention pieces, behind codon_cac_channel /
// codon_cac_spatial / codon_cac_apply.  They serve the module-level API of the reference
// (CAC_module.CAC_channel / CAC_spatial, attention/ResCBAM.py ChannelGate / SpatialGate) for
// arbitrary channel counts; the fused forward uses the NHWC kernels in cac.cu instead.
#include "common.cuh"
#include "kernels.h"

namespace codon {
namespace {

// one CTA per (b, c) plane: mean and max over HW (CAC_module.py:43,47; ResCBAM.py:42,45)
__global__ void __launch_bounds__(256) nchw_channel_stats_kernel(const float* __restrict__ x, int HW,
                                                                 float* __restrict__ avg,
                                                                 float* __restrict__ mx) {
  const float* src = x + (size_t)blockIdx.x * HW;
  float s = 0.f, m = -INFINITY;
  for (int i = threadIdx.x; i < HW; i += 256) { const float v = src[i]; s += v; m = fmaxf(m, v); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  }
  __shared__ float rs[8], rm[8];
  if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rm[threadIdx.x >> 5] = m; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f, tm = -INFINITY;
    for (int w = 0; w < 8; ++w) { ts += rs[w]; tm = fmaxf(tm, rm[w]); }
    avg[blockIdx.x] = ts / (float)HW;
    mx[blockIdx.x] = tm;
  }
}

// one CTA per sample: scale = sigmoid(mlp(avg) + mlp(max)), mlp = Linear(C,hidden)+ReLU+Linear(hidden,c_out)
__global__ void __launch_bounds__(128) gate_mlp_kernel(const float* __restrict__ avg, const float* __restrict__ mx,
                                                       int C, const float* __restrict__ w1,
                                                       const float* __restrict__ b1,
                                                       const float* __restrict__ w2,
                                                       const float* __restrict__ b2, int hidden, int c_out,
                                                       float* __restrict__ scale) {
  extern __shared__ float hid[];   // [2][hidden]
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < 2 * hidden; i += 128) {
    const int h = i % hidden;
    const float* v = (i < hidden ? avg : mx) + (size_t)b * C;
    float a = b1[h];
    for (int j = 0; j < C; ++j) a = fmaf(w1[h * C + j], v[j], a);
    hid[i] = fmaxf(a, 0.f);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < c_out; o += 128) {
    float za = b2[o], zm = b2[o];
    for (int h = 0; h < hidden; ++h) {
      za = fmaf(w2[o * hidden + h], hid[h], za);
      zm = fmaf(w2[o * hidden + h], hid[hidden + h], zm);
    }
    scale[(size_t)b * c_out + o] = sigmoidf_exact(za + zm);
  }
}

// ChannelPool (CAC_module.py:78-81): pooled[b,0] = max_c, pooled[b,1] = mean_c
__global__ void __launch_bounds__(256) nchw_channel_pool_kernel(const float* __restrict__ x, int C, int HW,
                                                                float* __restrict__ pooled) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= HW) return;
  const float* src = x + (size_t)b * C * HW + p;
  float s = 0.f, m = -INFINITY;
  for (int c = 0; c < C; ++c) { const float v = src[(size_t)c * HW]; s += v; m = fmaxf(m, v); }
  pooled[((size_t)b * 2) * HW + p] = m;
  pooled[((size_t)b * 2 + 1) * HW + p] = s / (float)C;
}

// sigmoid(conv5x5 2->1, zero pad 2, no bias) (CAC_module.py:88-93)
__global__ void __launch_bounds__(256) nchw_spatial_scale_kernel(const float* __restrict__ pooled,
                                                                 const float* __restrict__ w, int H, int W,
                                                                 float* __restrict__ scale) {
  __shared__ float sw[50];
  if (threadIdx.x < 50) sw[threadIdx.x] = w[threadIdx.x];
  __syncthreads();
  const int b = blockIdx.y, HW = H * W;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= HW) return;
  const int gy = p / W, gx = p % W;
  const float* pm = pooled + (size_t)b * 2 * HW;
  float q = 0.f;
#pragma unroll
  for (int dy = 0; dy < 5; ++dy)
#pragma unroll
    for (int dx = 0; dx < 5; ++dx) {
      const int yy = gy + dy - 2, xx = gx + dx - 2;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        q = fmaf(sw[dy * 5 + dx], pm[yy * W + xx], q);
        q = fmaf(sw[25 + dy * 5 + dx], pm[HW + yy * W + xx], q);
      }
    }
  scale[(size_t)b * HW + p] = sigmoidf_exact(q);
}

__global__ void __launch_bounds__(256) nchw_apply_kernel(const float* __restrict__ x, const float* __restrict__ sc,
                                                         const float* __restrict__ ss,
                                                         const float* __restrict__ res, int C, int HW,
                                                         int c_gate, float* __restrict__ y) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= HW) return;
  const size_t o = ((size_t)b * C + c) * HW + p;
  float g = 1.f;
  if (sc) g *= sc[(size_t)b * c_gate + (c % c_gate)];
  if (ss) g *= ss[(size_t)b * HW + p];
  float v = x[o] * g;
  if (res) v += res[o];
  y[o] = v;
}

}  // namespace

cudaError_t launch_nchw_channel_stats(const float* x, int B, int C, int HW, float* avg, float* mx,
                                      cudaStream_t st) {
  nchw_channel_stats_kernel<<<B * C, 256, 0, st>>>(x, HW, avg, mx);
  return cudaGetLastError();
}
cudaError_t launch_gate_mlp(const float* avg, const float* mx, int B, int C, const float* w1, const float* b1,
                            const float* w2, const float* b2, int hidden, int c_out, float* scale,
                            cudaStream_t st) {
  gate_mlp_kernel<<<B, 128, 2 * hidden * sizeof(float), st>>>(avg, mx, C, w1, b1, w2, b2, hidden, c_out, scale);
  return cudaGetLastError();
}
cudaError_t launch_nchw_channel_pool(const float* x, int B, int C, int HW, float* pooled, cudaStream_t st) {
  nchw_channel_pool_kernel<<<dim3(cdiv(HW, 256), B), 256, 0, st>>>(x, C, HW, pooled);
  return cudaGetLastError();
}
cudaError_t launch_nchw_spatial_scale(const float* pooled, const float* w, int B, int H, int W, float* scale,
                                      cudaStream_t st) {
  nchw_spatial_scale_kernel<<<dim3(cdiv(H * W, 256), B), 256, 0, st>>>(pooled, w, H, W, scale);
  return cudaGetLastError();
}
cudaError_t launch_nchw_apply(const float* x, const float* sc, const float* ss, const float* res, int B, int C,
                              int HW, int c_gate, float* y, cudaStream_t st) {
  nchw_apply_kernel<<<dim3(cdiv(HW, 256), C, B), 256, 0, st>>>(x, sc, ss, res, C, HW, c_gate, y);
  return cudaGetLastError();
}

}  // namespace codon
