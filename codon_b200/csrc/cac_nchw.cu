// Stand-alone NCHW fp32 versions of the attention pieces, behind codon_cac_channel /
// codon_cac_spatial / codon_cac_apply.  They serve the module-level API of the reference
// (CAC_module.CAC_channel / CAC_spatial, attention/ResCBAM.py ChannelGate / SpatialGate) for
// arbitrary channel counts; the fused forward uses the NHWC kernels in cac.cu instead.
#include "common.cuh"
#include "kernels.h"

namespace codon {
namespace {

// one CTA per (b, c) plane: the pooled statistics CAC_channel / ChannelGate can select
// (CAC_module.py:43,47,50-55; ResCBAM.py:42-52): stats[0] mean, [1] max, [2] lp (p=2:
// sqrt(sum x^2)), [3] lse (max + log sum exp(x - max)), each [B*C].
__device__ __forceinline__ float block_sum(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < 8; ++w) t += sh[w];
  return t;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = -INFINITY;
  for (int w = 0; w < 8; ++w) t = fmaxf(t, sh[w]);
  return t;
}
__global__ void __launch_bounds__(256) nchw_channel_stats_kernel(const float* __restrict__ x, int HW, int BC,
                                                                 int pool_mask, float* __restrict__ stats) {
  __shared__ float sh[8];
  const float* src = x + (size_t)blockIdx.x * HW;
  float s = 0.f, m = -INFINITY, q = 0.f;
  for (int i = threadIdx.x; i < HW; i += 256) { const float v = src[i]; s += v; m = fmaxf(m, v); q = fmaf(v, v, q); }
  s = block_sum(s, sh); m = block_max(m, sh); q = block_sum(q, sh);
  float l = 0.f;
  if (pool_mask & 8) {
    for (int i = threadIdx.x; i < HW; i += 256) l += expf(src[i] - m);
    l = block_sum(l, sh);
  }
  if (threadIdx.x == 0) {
    stats[blockIdx.x] = s / (float)HW;
    stats[BC + blockIdx.x] = m;
    stats[2 * BC + blockIdx.x] = sqrtf(q);
    stats[3 * BC + blockIdx.x] = m + logf(l);
  }
}

// one CTA per sample: scale = sigmoid(sum over the selected pools of mlp(pool)),
// mlp = Linear(C,hidden)+ReLU+Linear(hidden,c_out) (CAC_module.py:29-35,57-62)
__global__ void __launch_bounds__(128) gate_mlp_kernel(const float* __restrict__ stats, int BC, int C, int pool_mask,
                                                       const float* __restrict__ w1,
                                                       const float* __restrict__ b1,
                                                       const float* __restrict__ w2,
                                                       const float* __restrict__ b2, int hidden, int c_out,
                                                       float* __restrict__ scale) {
  extern __shared__ float hid[];   // [4][hidden]
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < 4 * hidden; i += 128) {
    const int h = i % hidden, k = i / hidden;
    const float* v = stats + (size_t)k * BC + (size_t)b * C;
    float a = b1[h];
    for (int j = 0; j < C; ++j) a = fmaf(w1[h * C + j], v[j], a);
    hid[i] = fmaxf(a, 0.f);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < c_out; o += 128) {
    float z = 0.f;
    for (int k = 0; k < 4; ++k) {
      if (!(pool_mask >> k & 1)) continue;
      float zk = b2[o];
      for (int h = 0; h < hidden; ++h) zk = fmaf(w2[o * hidden + h], hid[k * hidden + h], zk);
      z += zk;
    }
    scale[(size_t)b * c_out + o] = sigmoidf_exact(z);
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

// Generic NCHW fp32 convolution for the small helper convs of the module-level API
// (BasicConv, CAC_module.py:6-20): one thread per output element, cross-correlation.
struct Conv2dParams {
  int B, Cin, H, W, Cout, OH, OW, kh, kw, sh, sw, ph, pw, dh, dw, groups, relu;
};
__global__ void __launch_bounds__(256) conv2d_nchw_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias, const Conv2dParams p,
                                                          float* __restrict__ y) {
  const size_t total = (size_t)p.B * p.Cout * p.OH * p.OW;
  const int cin_g = p.Cin / p.groups, cout_g = p.Cout / p.groups;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % p.OW), oy = (int)((idx / p.OW) % p.OH);
    const int co = (int)((idx / ((size_t)p.OW * p.OH)) % p.Cout), b = (int)(idx / ((size_t)p.OW * p.OH * p.Cout));
    const int g = co / cout_g;
    float acc = bias ? bias[co] : 0.f;
    for (int ci = 0; ci < cin_g; ++ci) {
      const float* xp = x + ((size_t)b * p.Cin + g * cin_g + ci) * p.H * p.W;
      const float* wp = w + ((size_t)co * cin_g + ci) * p.kh * p.kw;
      for (int ky = 0; ky < p.kh; ++ky) {
        const int iy = oy * p.sh - p.ph + ky * p.dh;
        if (iy < 0 || iy >= p.H) continue;
        for (int kx = 0; kx < p.kw; ++kx) {
          const int ix = ox * p.sw - p.pw + kx * p.dw;
          if (ix < 0 || ix >= p.W) continue;
          acc = fmaf(xp[(size_t)iy * p.W + ix], wp[ky * p.kw + kx], acc);
        }
      }
    }
    y[idx] = p.relu ? fmaxf(acc, 0.f) : acc;
  }
}

}  // namespace

cudaError_t launch_nchw_channel_stats(const float* x, int B, int C, int HW, int pool_mask, float* stats,
                                      cudaStream_t st) {
  nchw_channel_stats_kernel<<<B * C, 256, 0, st>>>(x, HW, B * C, pool_mask, stats);
  return cudaGetLastError();
}
cudaError_t launch_gate_mlp(const float* stats, int B, int C, int pool_mask, const float* w1, const float* b1,
                            const float* w2, const float* b2, int hidden, int c_out, float* scale,
                            cudaStream_t st) {
  gate_mlp_kernel<<<B, 128, 4 * hidden * sizeof(float), st>>>(stats, B * C, C, pool_mask, w1, b1, w2, b2, hidden,
                                                              c_out, scale);
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

cudaError_t launch_conv2d_nchw(const float* x, const float* w, const float* bias, int B, int Cin, int H, int W,
                               int Cout, int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int groups,
                               int relu, float* y, cudaStream_t st) {
  Conv2dParams p;
  p.B = B; p.Cin = Cin; p.H = H; p.W = W; p.Cout = Cout; p.kh = kh; p.kw = kw; p.sh = sh; p.sw = sw;
  p.ph = ph; p.pw = pw; p.dh = dh; p.dw = dw; p.groups = groups; p.relu = relu;
  p.OH = (H + 2 * ph - dh * (kh - 1) - 1) / sh + 1;
  p.OW = (W + 2 * pw - dw * (kw - 1) - 1) / sw + 1;
  if (p.OH < 1 || p.OW < 1) return cudaErrorInvalidValue;
  size_t total = (size_t)B * Cout * p.OH * p.OW, g = (total + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  conv2d_nchw_kernel<<<(int)g, 256, 0, st>>>(x, w, bias, p, y);
  return cudaGetLastError();
}

}  // namespace codon
