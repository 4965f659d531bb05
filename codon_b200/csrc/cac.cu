// CAC cross-domain attention (CODON_X4/CAC_module.py; call sites CODON_X4/CODON_x4.py:85-118),
// three HBM-bound kernels per stage on the NHWC 128-channel feature buffer F = [depth | colour]:
//
//   stats : one read of F.  Per pixel max / mean over the 128 channels (ChannelPool,
//           CAC_module.py:78-81) -> pooled [B,H,W,2]; per 1024-pixel chunk the per-channel sum
//           and max (avg_pool2d / max_pool2d over the frame, CAC_module.py:43,47) -> partials.
//           16-byte loads, fp32 accumulation, warp-shuffle reductions.
//   mlp   : fixed-order reduction of the partials (deterministic, independent of batch and GPU
//           count), the shared 128->8->64 MLP on avg and max, sum, sigmoid (CAC_module.py:29-35,
//           59-62) -> s_c [B,64].
//   apply : s_s = sigmoid(conv5x5(pooled)) (CAC_module.py:88-93) from a shared-memory halo tile,
//           then F = F * s_c[c % 64] * s_s + E in one pass (CODON_x4.py:88-91,117-118).
//
// Algorithmic HBM bytes per pixel per stage (e = bytes per element): stats 128e, apply
// 128e (F) + 128e (E) + 128e (store) = 512e in total.
#include "common.cuh"
#include "kernels.h"

namespace codon {
namespace {

constexpr int kChunkPx = 256;    // pixels per stats CTA; fixed so that results are batch-invariant

// 256 threads; every thread keeps kU 16-byte loads in flight (memory-latency bound otherwise: the
// r01 version had 4 loads per thread and 0.5 waves at B=1 and reached 1.5-2.4 TB/s).
template <typename T>
__global__ void __launch_bounds__(256, 4) cac_stats_kernel(const T* __restrict__ F, int HW, int chunks,
                                                        float* __restrict__ pooled,
                                                        float* __restrict__ part) {
  constexpr int V = Act<T>::kVec, LPP = 128 / V, PPW = 32 / LPP, kU = sizeof(typename Act<T>::Raw) > 16 ? 4 : 8;
  constexpr int ROWPX = 8 * PPW;                 // pixels covered by one load instruction of the CTA
  constexpr int NIT = kChunkPx / (ROWPX * kU);   // 2 (16-bit) or 4 (fp32) batches of kU loads
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane % LPP, sub = lane / LPP;
  const int p_begin = chunk * kChunkPx, p_end = min(p_begin + kChunkPx, HW);
  const T* base = F + (size_t)b * HW * 128;
  float* pl = pooled + (size_t)b * HW * 2;

  float csum[V], cmax[V];
#pragma unroll
  for (int j = 0; j < V; ++j) { csum[j] = 0.f; cmax[j] = -INFINITY; }

#pragma unroll 1
  for (int it = 0; it < NIT; ++it) {
    const int p0 = p_begin + it * ROWPX * kU + warp * PPW + sub;
    typename Act<T>::Raw raw[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int p = p0 + u * ROWPX;
      raw[u] = Act<T>::ldg(base + (size_t)(p < p_end ? p : p_begin) * 128 + g * V);   // out-of-range lanes are masked below
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int p = p0 + u * ROWPX;
      const bool ok = p < p_end;
      float v[V];
      Act<T>::unpack(raw[u], v);
      float s = 0.f, m = -INFINITY;
      if (ok) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
          csum[j] += v[j];
          cmax[j] = fmaxf(cmax[j], v[j]);
          s += v[j];
          m = fmaxf(m, v[j]);
        }
      }
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      }
      if (ok && g == 0) *reinterpret_cast<float2*>(pl + (size_t)p * 2) = make_float2(m, s * (1.0f / 128.0f));
    }
  }

  __shared__ float rs[8 * PPW][128], rm[8 * PPW][128];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    rs[warp * PPW + sub][g * V + j] = csum[j];
    rm[warp * PPW + sub][g * V + j] = cmax[j];
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    float s = 0.f, m = -INFINITY;
#pragma unroll
    for (int r = 0; r < 8 * PPW; ++r) { s += rs[r][threadIdx.x]; m = fmaxf(m, rm[r][threadIdx.x]); }
    float* dst = part + ((size_t)(b * chunks + chunk) * 2) * 128;
    dst[threadIdx.x] = s;
    dst[128 + threadIdx.x] = m;
  }
}

// Lean variant for the tensor-core modes: the per-pixel ChannelPool partials are emitted by the producing
// 1x1 convolution's epilogue (conv_tc.cu, TcJob::pool), so this kernel only accumulates the per-channel
// sum / max of its 256-pixel chunk: no cross-lane work in the loop, the max runs on packed 16-bit pairs.
template <typename T> struct PackedMax;
template <> struct PackedMax<float> {
  using Acc = uint4;
  __device__ static inline void fin(const uint4& a, float (&m)[4]) { Act<float>::unpack(a, m); }
  __device__ static inline uint4 init() { const uint32_t n = 0xff800000u; return make_uint4(n, n, n, n); }
  __device__ static inline void upd(uint4& a, const uint4& v) {
    a.x = __float_as_uint(fmaxf(__uint_as_float(a.x), __uint_as_float(v.x)));
    a.y = __float_as_uint(fmaxf(__uint_as_float(a.y), __uint_as_float(v.y)));
    a.z = __float_as_uint(fmaxf(__uint_as_float(a.z), __uint_as_float(v.z)));
    a.w = __float_as_uint(fmaxf(__uint_as_float(a.w), __uint_as_float(v.w)));
  }
};
template <> struct PackedMax<__nv_bfloat16> {
  using Acc = uint4;
  __device__ static inline void fin(const uint4& a, float (&m)[8]) { Act<__nv_bfloat16>::unpack(a, m); }
  __device__ static inline uint4 init() { const uint32_t n = 0xff80ff80u; return make_uint4(n, n, n, n); }
  __device__ static inline uint32_t mx(uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  __device__ static inline void upd(uint4& a, const uint4& v) { a.x = mx(a.x, v.x); a.y = mx(a.y, v.y); a.z = mx(a.z, v.z); a.w = mx(a.w, v.w); }
};
template <> struct PackedMax<split16> {   // hi + lo must be recombined: the maximum is taken on the fp32 values
  struct Acc { float m[8]; };
  __device__ static inline Acc init() { Acc a; for (int j = 0; j < 8; ++j) a.m[j] = -INFINITY; return a; }
  __device__ static inline void upd(Acc& a, const RawSplit& v) {
    float f[8];
    Act<split16>::unpack(v, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) a.m[j] = fmaxf(a.m[j], f[j]);
  }
  __device__ static inline void fin(const Acc& a, float (&m)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = a.m[j];
  }
};
template <> struct PackedMax<__half> {
  using Acc = uint4;
  __device__ static inline void fin(const uint4& a, float (&m)[8]) { Act<__half>::unpack(a, m); }
  __device__ static inline uint4 init() { const uint32_t n = 0xfc00fc00u; return make_uint4(n, n, n, n); }
  __device__ static inline uint32_t mx(uint32_t a, uint32_t b) {
    __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  __device__ static inline void upd(uint4& a, const uint4& v) { a.x = mx(a.x, v.x); a.y = mx(a.y, v.y); a.z = mx(a.z, v.z); a.w = mx(a.w, v.w); }
};

template <typename T>
__global__ void __launch_bounds__(256, 4) cac_chan_stats_kernel(const T* __restrict__ F, int HW, int chunks,
                                                                float* __restrict__ part) {
  constexpr int V = Act<T>::kVec, LPP = 128 / V, PPW = 32 / LPP, kU = sizeof(typename Act<T>::Raw) > 16 ? 4 : 8;
  constexpr int ROWPX = 8 * PPW;
  constexpr int NIT = kChunkPx / (ROWPX * kU);
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane % LPP, sub = lane / LPP;
  const int p_begin = chunk * kChunkPx, p_end = min(p_begin + kChunkPx, HW);
  const T* base = F + (size_t)b * HW * 128;

  float csum[V];
#pragma unroll
  for (int j = 0; j < V; ++j) csum[j] = 0.f;
  typename PackedMax<T>::Acc cmax = PackedMax<T>::init();

#pragma unroll 1
  for (int it = 0; it < NIT; ++it) {
    const int p0 = p_begin + it * ROWPX * kU + warp * PPW + sub;
    typename Act<T>::Raw raw[kU];
    bool ok[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int p = p0 + u * ROWPX;
      ok[u] = p < p_end;
      if (ok[u]) raw[u] = Act<T>::ldg(base + (size_t)p * 128 + g * V);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (ok[u]) {
        float v[V];
        Act<T>::unpack(raw[u], v);
#pragma unroll
        for (int j = 0; j < V; ++j) csum[j] += v[j];
        PackedMax<T>::upd(cmax, raw[u]);
      }
    }
  }
  float cmx[V];
  PackedMax<T>::fin(cmax, cmx);
  __shared__ float rs[8 * PPW][128], rm[8 * PPW][128];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    rs[warp * PPW + sub][g * V + j] = csum[j];
    rm[warp * PPW + sub][g * V + j] = cmx[j];
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    float s = 0.f, m = -INFINITY;
#pragma unroll
    for (int r = 0; r < 8 * PPW; ++r) { s += rs[r][threadIdx.x]; m = fmaxf(m, rm[r][threadIdx.x]); }
    float* dst = part + ((size_t)(b * chunks + chunk) * 2) * 128;
    dst[threadIdx.x] = s;
    dst[128 + threadIdx.x] = m;
  }
}

// Tensor-core modes with the fused 5x5 + 1x1 kernel: the conv epilogue already left per-channel (sum, max) partials
// per 8 x 16-pixel cell and TMEM lane quarter (TcJob::cstat: [frame][cell][4][64] float2 per branch).  This kernel
// folds kCellsPerChunk consecutive cells (row-major cell order of the frame) into one chunk of the `part` layout the
// MLP kernel reduces, in a fixed order; thread = F channel (depth 0..63 from cstat_d, colour 64..127 from cstat_c).
// It reads 32 B per pixel instead of the 128 * e B per pixel the stand-alone statistics pass re-reads from F.
constexpr int kCellsPerChunk = 8;
__global__ void __launch_bounds__(128) cac_cell_reduce_kernel(const float2* __restrict__ cstat_d,
                                                              const float2* __restrict__ cstat_c, int cells,
                                                              int chunks, float* __restrict__ part) {
  const int b = blockIdx.y, chunk = blockIdx.x, ch = threadIdx.x;
  const float2* src = (ch < 64 ? cstat_d : cstat_c) + (size_t)b * cells * 256 + (ch & 63);
  const int c0 = chunk * kCellsPerChunk, c1 = min(c0 + kCellsPerChunk, cells);
  float s = 0.f, m = -INFINITY;
  for (int c = c0; c < c1; ++c) {
    float2 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = __ldg(src + ((size_t)c * 4 + q) * 64);
#pragma unroll
    for (int q = 0; q < 4; ++q) { s += v[q].x; m = fmaxf(m, v[q].y); }
  }
  float* dst = part + ((size_t)(b * chunks + chunk) * 2) * 128;
  dst[ch] = s;
  dst[128 + ch] = m;
}

// One CTA per frame, 1024 threads.  F channel c (depth 0..63 | colour 64..127) is Fcat channel
// (c + 64) % 128 (Fcat = [colour | depth], CODON_x4.py:85); w1 is indexed by Fcat channel.
// The chunk partials ([chunks][sum 128 | max 128]) are reduced in a fixed order -- 16 interleaved groups of
// chunks, each thread one float4 of columns, then a fixed 16-way combine -- so the result does not depend
// on the batch size or the GPU count.
__global__ void __launch_bounds__(1024) cac_mlp_kernel(const float* __restrict__ part, int chunks, int HW,
                                                       const float* __restrict__ w1,
                                                       const float* __restrict__ b1,
                                                       const float* __restrict__ w2,
                                                       const float* __restrict__ b2,
                                                       float* __restrict__ sc) {
  __shared__ float4 red[16][64];
  __shared__ float avg[128], mx[128], hid[2][8];
  const int b = blockIdx.x, t = threadIdx.x, cg = t & 63, grp = t >> 6;
  const float4* src = reinterpret_cast<const float4*>(part + (size_t)b * chunks * 256) + cg;
  const bool is_max = cg >= 32;                  // float4 columns 32..63 hold the per-channel maxima
  float4 acc = is_max ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c0 = grp; c0 < chunks; c0 += 16 * 8) {     // 8 independent 16-byte loads in flight per thread
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = c0 + 16 * u;
      v[u] = c < chunks ? __ldg(src + (size_t)c * 64) : acc;
      if (c >= chunks) v[u] = is_max ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (is_max) { acc.x = fmaxf(acc.x, v[u].x); acc.y = fmaxf(acc.y, v[u].y); acc.z = fmaxf(acc.z, v[u].z); acc.w = fmaxf(acc.w, v[u].w); }
      else { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
  }
  red[grp][cg] = acc;
  __syncthreads();
  if (t < 256) {
    const int col = t, q4 = col >> 2, e = col & 3;
    const bool mxcol = col >= 128;
    float r = reinterpret_cast<const float*>(&red[0][q4])[e];
#pragma unroll
    for (int g = 1; g < 16; ++g) {
      const float v = reinterpret_cast<const float*>(&red[g][q4])[e];
      r = mxcol ? fmaxf(r, v) : r + v;
    }
    const int fc = ((col & 127) + 64) & 127;
    if (mxcol) mx[fc] = r; else avg[fc] = r / (float)HW;
  }
  __syncthreads();
  if (t < 512) {
    // 16 hidden units x 32 lanes: warp w computes hidden unit (w & 7) of the avg (w < 8) or max branch
    const int w = t >> 5, lane = t & 31, h = w & 7;
    const float* v = (w < 8) ? avg : mx;
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) a = fmaf(w1[h * 128 + lane * 4 + j], v[lane * 4 + j], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) hid[w >> 3][h] = fmaxf(a + b1[h], 0.f);
  }
  __syncthreads();
  if (t < 64) {
    float za = b2[t], zm = b2[t];
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      za = fmaf(w2[t * 8 + h], hid[0][h], za);
      zm = fmaf(w2[t * 8 + h], hid[1][h], zm);
    }
    sc[b * 64 + t] = sigmoidf_exact(za + zm);
  }
}

// apply tile: up to 8 x 32 pixels.  The tile HEIGHT is chosen per launch (4 .. 8 rows) so that the tile count comes close
// to a whole number of waves of resident CTAs: at 640x480 x 1 frame, 8-row tiles are 1200 CTAs = 2.03 waves of 592 (the
// kernel ran three waves' worth of time for two waves of work: 0.70 of the HBM peak), 6-row tiles are 1600 = 2.7 waves.
constexpr int kATH = 8, kATW = 32;

template <typename T>
__global__ void __launch_bounds__(256, 4) cac_apply_kernel(T* __restrict__ F, const T* __restrict__ E,
                                                           const float* __restrict__ pooled,
                                                           const float* __restrict__ sc,
                                                           const float* __restrict__ ws, int H, int W,
                                                        int tiles_x, int rnd_tf32, int pool_parts, size_t part_stride, int th) {
  constexpr int V = Act<T>::kVec, LPP = 128 / V, PH = kATH + 4, PW = kATW + 4;
  __shared__ float2 sp[PH][PW];
  __shared__ float sw[50], ssc[64], sss[kATH * kATW];
  const int b = blockIdx.y, t = threadIdx.x;
  const int ty0 = (blockIdx.x / tiles_x) * th, tx0 = (blockIdx.x % tiles_x) * kATW;   // th <= kATH rows per tile
  const size_t fb = (size_t)b * H * W;
  // The tile's F / E vectors do not depend on the gate: the first batch of kU + kU 16-byte loads is issued before
  // the pooled-halo staging and the 5x5 gate convolution, so HBM is busy while the CTA computes s_s (r01h: every
  // CTA of a wave sat in that prologue at the same time with no load in flight).  Thread t owns vector
  // i = t + k * 256 (pixel i / LPP, 16-byte group i % LPP) for k < LPP, fetched in batches of kU.
  constexpr int kU = sizeof(typename Act<T>::Raw) > 16 ? 2 : 4, NB = LPP / kU;
  typename Act<T>::Raw rf[kU], re[kU];
  int pix[kU];                                   // pixel index inside the frame, -1 = outside the image
  auto fetch = [&](int bt) {
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int i = t + (bt * kU + u) * 256, g = i % LPP, px = i / LPP;
      const int gy = ty0 + px / kATW, gx = tx0 + px % kATW;
      pix[u] = (px < th * kATW && gy < H && gx < W) ? gy * W + gx : -1;
      if (pix[u] >= 0) {
        const size_t o = (fb + (size_t)pix[u]) * 128 + g * V;
        rf[u] = Act<T>::ld(F + o);
        re[u] = Act<T>::ldg(E + o);
      }
    }
  };
  fetch(0);
  if (t < 50) sw[t] = ws[t];
  if (t >= 64 && t < 128) ssc[t - 64] = sc[b * 64 + t - 64];
  for (int i = t; i < (th + 4) * PW; i += 256) {
    const int gy = ty0 + i / PW - 2, gx = tx0 + i % PW - 2;
    float2 v = make_float2(0.f, 0.f);
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      v = __ldg(reinterpret_cast<const float2*>(pooled + (fb + (size_t)gy * W + gx) * 2));
      if (pool_parts > 1) {
        // (max, sum) partials from the 1x1 conv epilogues (2: one per branch; 4: per branch and 32-channel half)
        // -> (max, mean) over the 128 channels
        for (int k = 1; k < pool_parts; ++k) {
          const float2 w = __ldg(reinterpret_cast<const float2*>(pooled + ((size_t)k * part_stride + fb + (size_t)gy * W + gx) * 2));
          v = make_float2(fmaxf(v.x, w.x), v.y + w.y);
        }
        v.y *= (1.0f / 128.0f);
      }
    }
    sp[i / PW][i % PW] = v;
  }
  __syncthreads();
  if (t < th * kATW) {
    const int r = t / kATW, c = t % kATW;
    float q = 0.f;
#pragma unroll
    for (int dy = 0; dy < 5; ++dy)
#pragma unroll
      for (int dx = 0; dx < 5; ++dx) {
        const float2 pv = sp[r + dy][c + dx];
        q = fmaf(sw[dy * 5 + dx], pv.x, q);
        q = fmaf(sw[25 + dy * 5 + dx], pv.y, q);
      }
    sss[t] = sigmoidf_exact(q);
  }
  __syncthreads();
#pragma unroll 1
  for (int bt = 0; bt < NB; ++bt) {
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int i = t + (bt * kU + u) * 256, px = i / LPP;
      if (pix[u] >= 0) {
        float f[V], e[V];
        Act<T>::unpack(rf[u], f);
        Act<T>::unpack(re[u], e);
        const float s = sss[px];
        const int c0 = ((i % LPP) * V) & 63;
#pragma unroll
        for (int j = 0; j < V; ++j) f[j] = fmaf(f[j], ssc[c0 + j] * s, e[j]);
        if (rnd_tf32) {
#pragma unroll
          for (int j = 0; j < V; ++j) f[j] = round_tf32(f[j]);
        }
        Act<T>::store(F + (fb + (size_t)pix[u]) * 128 + (i % LPP) * V, f);
      }
    }
    if (bt + 1 < NB) fetch(bt + 1);
  }
}

}  // namespace

int cac_stats_chunks(int /*B*/, int H, int W) { return cdiv(H * W, kChunkPx); }

cudaError_t launch_cac_stats(const void* F, int act, int B, int H, int W, float* pooled, float* part,
                             int chunks, cudaStream_t st) {
  dim3 grid(chunks, B);
  const int HW = H * W;
  if (act == ACT_F32) cac_stats_kernel<float><<<grid, 256, 0, st>>>((const float*)F, HW, chunks, pooled, part);
  else if (act == ACT_BF16) cac_stats_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)F, HW, chunks, pooled, part);
  else if (act == ACT_SPLIT16) cac_stats_kernel<split16><<<grid, 256, 0, st>>>((const split16*)F, HW, chunks, pooled, part);
  else cac_stats_kernel<__half><<<grid, 256, 0, st>>>((const __half*)F, HW, chunks, pooled, part);
  return cudaGetLastError();
}

cudaError_t launch_cac_chan_stats(const void* F, int act, int B, int H, int W, float* part, int chunks,
                                  cudaStream_t st) {
  dim3 grid(chunks, B);
  const int HW = H * W;
  if (act == ACT_F32) cac_chan_stats_kernel<float><<<grid, 256, 0, st>>>((const float*)F, HW, chunks, part);
  else if (act == ACT_BF16) cac_chan_stats_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)F, HW, chunks, part);
  else if (act == ACT_SPLIT16) cac_chan_stats_kernel<split16><<<grid, 256, 0, st>>>((const split16*)F, HW, chunks, part);
  else cac_chan_stats_kernel<__half><<<grid, 256, 0, st>>>((const __half*)F, HW, chunks, part);
  return cudaGetLastError();
}

int cac_cell_chunks(int cells) { return cdiv(cells, kCellsPerChunk); }

cudaError_t launch_cac_cell_reduce(const void* cstat_d, const void* cstat_c, int B, int cells, float* part, int chunks,
                                   cudaStream_t st) {
  cac_cell_reduce_kernel<<<dim3(chunks, B), 128, 0, st>>>(static_cast<const float2*>(cstat_d), static_cast<const float2*>(cstat_c),
                                                            cells, chunks, part);
  return cudaGetLastError();
}

cudaError_t launch_cac_mlp(const float* part, int chunks, int B, int HW, const float* w1, const float* b1,
                           const float* w2, const float* b2, float* sc, cudaStream_t st) {
  cac_mlp_kernel<<<B, 1024, 0, st>>>(part, chunks, HW, w1, b1, w2, b2, sc);
  return cudaGetLastError();
}

cudaError_t launch_cac_apply(void* F, const void* E, int act, const float* pooled, const float* sc,
                             const float* ws, int B, int H, int W, cudaStream_t st, int rnd_tf32, int pool_parts,
                             size_t part_stride) {
  const int tiles_x = cdiv(W, kATW);
  if (part_stride == 0) part_stride = (size_t)B * H * W;
  // tile height: the one that needs the least (waves of 148 x 4 resident CTAs) x (rows per tile)
  int th = kATH;
  {
    long best = -1;
    for (int h = kATH; h >= 4; --h) {
      const long tiles = (long)tiles_x * cdiv(H, h) * B;
      // a tile costs its rows plus a fixed part (gate prologue, partially idle lanes of a short tile) worth ~2 rows:
      // with many waves the full 8-row tile wins (measured: B = 8 bf16 0.90 of the HBM peak with 8 rows, 0.73 with 6)
      const long cost = ((tiles + 591) / 592) * (h + 2);
      if (best < 0 || cost < best) { best = cost; th = h; }
    }
  }
  dim3 grid(tiles_x * cdiv(H, th), B);
  if (act == ACT_F32) cac_apply_kernel<float><<<grid, 256, 0, st>>>((float*)F, (const float*)E, pooled, sc, ws, H, W, tiles_x, rnd_tf32, pool_parts, part_stride, th);
  else if (act == ACT_BF16) cac_apply_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16*)F, (const __nv_bfloat16*)E, pooled, sc, ws, H, W, tiles_x, rnd_tf32, pool_parts, part_stride, th);
  else if (act == ACT_SPLIT16) cac_apply_kernel<split16><<<grid, 256, 0, st>>>((split16*)F, (const split16*)E, pooled, sc, ws, H, W, tiles_x, 0, pool_parts, part_stride, th);
  else cac_apply_kernel<__half><<<grid, 256, 0, st>>>((__half*)F, (const __half*)E, pooled, sc, ws, H, W, tiles_x, rnd_tf32, pool_parts, part_stride, th);
  return cudaGetLastError();
}

}  // namespace codon
