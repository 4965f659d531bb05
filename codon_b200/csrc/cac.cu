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
#include <atomic>
#include <type_traits>
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

// Spatial gate map s_s = sigmoid(conv5x5(ChannelPool)) (CAC_module.py:90-94), one 32 x 8-pixel tile per 128-thread
// block: pooled halo (zero padding at the image border) in shared memory, two pixels per thread.  The gate used to be
// the prologue of every cac_apply tile; computed by extra blocks of the statistics-fold launch (or by its own small
// launch on the paths without that fold) it leaves cac_apply a pure streaming kernel with no per-tile fixed cost and no
// wave quantisation.  Arithmetic and summation order are those of the r01 apply prologue (bit-identical gates).
constexpr int kGTW = 32, kGTH = 8;
__device__ __forceinline__ void gate_tile(const CacGate& g, int b, int tile) {
  __shared__ float2 sp[kGTH + 4][kGTW + 4];
  __shared__ float sw[50];
  const int t = threadIdx.x;
  const int tiles_x = (g.W + kGTW - 1) / kGTW;
  const int ty0 = (tile / tiles_x) * kGTH, tx0 = (tile % tiles_x) * kGTW;
  const int H = g.H, W = g.W;
  const size_t fb = (size_t)b * H * W;
  if (t < 50) sw[t] = g.ws[t];
  // Every load of the block is issued before the first one is used (a loop over halo pixels and a run-time part count
  // issues one dependent load after the other: a dozen memory latencies per block instead of one).
  auto stage = [&](auto PARTS) {
    constexpr int parts = decltype(PARTS)::value;
    constexpr int kHalo = (kGTH + 4) * (kGTW + 4), kIt = (kHalo + 127) / 128;
    float2 w[kIt][parts];
    bool in[kIt];
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int i = t + it * 128;
      const int gy = ty0 + i / (kGTW + 4) - 2, gx = tx0 + i % (kGTW + 4) - 2;
      in[it] = i < kHalo && gy >= 0 && gy < H && gx >= 0 && gx < W;
#pragma unroll
      for (int k = 0; k < parts; ++k)
        w[it][k] = in[it] ? __ldg(reinterpret_cast<const float2*>(g.pooled + ((size_t)k * g.part_stride + fb + (size_t)gy * W + gx) * 2))
                          : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int i = t + it * 128;
      float2 v = w[it][0];
      if (parts > 1) {
        // (max, sum) partials from the 1x1 conv epilogues (2: one per branch; 4: per branch and 32-channel half)
        // -> (max, mean) over the 128 channels
#pragma unroll
        for (int k = 1; k < parts; ++k) v = make_float2(fmaxf(v.x, w[it][k].x), v.y + w[it][k].y);
        v.y *= (1.0f / 128.0f);
      }
      if (!in[it]) v = make_float2(0.f, 0.f);
      if (i < kHalo) sp[i / (kGTW + 4)][i % (kGTW + 4)] = v;
    }
  };
  if (g.pool_parts == 4) stage(std::integral_constant<int, 4>{});
  else if (g.pool_parts == 2) stage(std::integral_constant<int, 2>{});
  else stage(std::integral_constant<int, 1>{});
  __syncthreads();
#pragma unroll
  for (int h = 0; h < kGTH / 4; ++h) {            // rows r, r + 4
    const int r = t / kGTW + h * 4, c = t % kGTW;
    if (ty0 + r < H && tx0 + c < W) {
      float acc = 0.f;
#pragma unroll
      for (int dy = 0; dy < 5; ++dy)
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
          const float2 pv = sp[r + dy][c + dx];
          acc = fmaf(sw[dy * 5 + dx], pv.x, acc);
          acc = fmaf(sw[25 + dy * 5 + dx], pv.y, acc);
        }
      g.gate[fb + (size_t)(ty0 + r) * W + tx0 + c] = sigmoidf_exact(acc);
    }
  }
}
__global__ void __launch_bounds__(128) cac_gate_kernel(const CacGate g) { gate_tile(g, blockIdx.y, blockIdx.x); }

// Tensor-core modes with the fused 5x5 + 1x1 kernel: the conv epilogue already left per-channel (sum, max) partials
// per 8 x 16-pixel cell and TMEM lane quarter (TcJob::cstat: [frame][cell][4][64] float2 per branch).  This kernel
// folds kCellsPerChunk consecutive cells (row-major cell order of the frame) into one chunk of the `part` layout the
// MLP kernel reduces, in a fixed order; thread = F channel (depth 0..63 from cstat_d, colour 64..127 from cstat_c).
// It reads 32 B per pixel instead of the 128 * e B per pixel the stand-alone statistics pass re-reads from F.
constexpr int kCellsPerChunk = 8;
__global__ void __launch_bounds__(128) cac_cell_reduce_kernel(const float2* __restrict__ cstat_d,
                                                              const float2* __restrict__ cstat_c, int cells,
                                                              int chunks, float* __restrict__ part, const CacGate g) {
  if ((int)blockIdx.x >= chunks) {               // blocks [chunks, chunks + gate tiles): the spatial gate map
    gate_tile(g, blockIdx.y, (int)blockIdx.x - chunks);
    return;
  }
  const int b = blockIdx.y, chunk = blockIdx.x, ch = threadIdx.x;
  const float2* src = (ch < 64 ? cstat_d : cstat_c) + (size_t)b * cells * 256 + (ch & 63);
  const int c0 = chunk * kCellsPerChunk, c1 = min(c0 + kCellsPerChunk, cells);
  float s = 0.f, m = -INFINITY;
  // all loads of the chunk first (8 cells x 4 lane quarters), then the fixed-order fold
  float2 v[kCellsPerChunk][4];
#pragma unroll
  for (int k = 0; k < kCellsPerChunk; ++k)
#pragma unroll
    for (int q = 0; q < 4; ++q)
      v[k][q] = c0 + k < c1 ? __ldg(src + ((size_t)(c0 + k) * 4 + q) * 64) : make_float2(0.f, -INFINITY);
#pragma unroll
  for (int k = 0; k < kCellsPerChunk; ++k)
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (c0 + k < c1) { s += v[k][q].x; m = fmaxf(m, v[k][q].y); }
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

// F = F * s_c * s_s + E in place (CODON_x4.py:86-118): a pure streaming kernel over the 16-byte vectors of F / E.  The
// gate map s_s comes from gate_tile (above), s_c from cac_mlp_kernel.  Persistent CTAs (4 per SM) walk the vectors in
// units of 256 (one per thread) with a grid stride -- no tiles, no prologue, no partial last wave (r01 .. r02a: 8 x
// 32-pixel tiles with the gate convolution as their prologue; one 640x480 frame was 2.03 or 2.7 waves of 592 resident
// CTAs and ran at 0.68-0.70 of the HBM peak in bf16; this kernel: 0.80 bf16, 0.95 with fp32 storage, same-session A/B
// profiles/r02_ab_cac_stream.txt).  The loads of the next batch are issued before the current batch is computed and
// stored.  The walk starts at the END of F: the convolution that produced F wrote its tiles top to bottom, so the
// bottom rows are the part of F most likely still in L2 (2-4 % over the forward walk on one frame; no effect at 8).
#ifndef CODON_APPLY_KU
#define CODON_APPLY_KU 2          // 16-byte vectors per thread and batch (two batches in flight)
#endif
#ifndef CODON_APPLY_ORDER
#define CODON_APPLY_ORDER 2       // 0: one contiguous range per CTA; 1 / 2: grid-stride units, forward / from the end
#endif
#ifndef CODON_APPLY_CTAS
#define CODON_APPLY_CTAS 4        // resident CTAs per SM
#endif
template <typename T>
__global__ void __launch_bounds__(256, CODON_APPLY_CTAS) cac_apply_kernel(T* __restrict__ F, const T* __restrict__ E,
                                                           const float* __restrict__ gate,
                                                           const float* __restrict__ sc, size_t nvec, int HW,
                                                           int rnd_tf32) {
  constexpr int V = Act<T>::kVec, LPP = 128 / V;          // vectors per pixel
  constexpr int kU = sizeof(typename Act<T>::Raw) > 16 ? CODON_APPLY_KU / 2 : CODON_APPLY_KU;   // vectors per thread and batch; two batches in flight
  // this CTA's units of 256 vectors (one per thread): unit_of(i), i < n
  const size_t units = (nvec + 255) / 256;
#if CODON_APPLY_ORDER == 0
  const size_t u0 = units * blockIdx.x / gridDim.x;
  const size_t n = units * (blockIdx.x + 1) / gridDim.x - u0;
  auto unit_of = [&](size_t i) { return u0 + i; };
#else
  const size_t n = (units + gridDim.x - 1 - blockIdx.x) / gridDim.x;
  auto unit_of = [&](size_t i) {
    const size_t u = blockIdx.x + i * gridDim.x;
    return CODON_APPLY_ORDER == 2 ? units - 1 - u : u;
  };
#endif
  const int t = threadIdx.x;
  typename Act<T>::Raw af[kU], ae[kU], bf[kU], be[kU];
  auto fetch = [&](size_t i, typename Act<T>::Raw (&f)[kU], typename Act<T>::Raw (&e)[kU]) {
#pragma unroll
    for (int k = 0; k < kU; ++k) {
      const size_t v = unit_of(i + k) * 256 + t;
      if (i + k < n && v < nvec) {
        f[k] = Act<T>::ld(F + v * V);
        e[k] = Act<T>::ldg(E + v * V);
      }
    }
  };
  auto finish = [&](size_t i, const typename Act<T>::Raw (&rf)[kU], const typename Act<T>::Raw (&re)[kU]) {
#pragma unroll
    for (int k = 0; k < kU; ++k) {
      const size_t v = unit_of(i + k) * 256 + t;
      if (i + k < n && v < nvec) {
        const size_t pix = v / LPP;
        const int c0 = ((int)(v % LPP) * V) & 63;
        const float s = __ldg(gate + pix);
        const float* ssc = sc + (pix / (size_t)HW) * 64 + c0;
        float f[V], e[V];
        Act<T>::unpack(rf[k], f);
        Act<T>::unpack(re[k], e);
#pragma unroll
        for (int j = 0; j < V; ++j) f[j] = fmaf(f[j], __ldg(ssc + j) * s, e[j]);
        if (rnd_tf32) {
#pragma unroll
          for (int j = 0; j < V; ++j) f[j] = round_tf32(f[j]);
        }
        Act<T>::store(F + v * V, f);
      }
    }
  };
  if (n > 0) fetch(0, af, ae);
#pragma unroll 1
  for (size_t i = 0; i < n; i += 2 * kU) {
    if (i + kU < n) fetch(i + kU, bf, be);
    finish(i, af, ae);
    if (i + 2 * kU < n) fetch(i + 2 * kU, af, ae);
    if (i + kU < n) finish(i + kU, bf, be);
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

static int gate_tiles(const CacGate& g) { return cdiv(g.W, kGTW) * cdiv(g.H, kGTH); }

cudaError_t launch_cac_cell_reduce(const void* cstat_d, const void* cstat_c, int B, int cells, float* part, int chunks,
                                   cudaStream_t st, const CacGate* gate) {
  CacGate g = {};
  if (gate) { g = *gate; if (g.part_stride == 0) g.part_stride = (size_t)B * g.H * g.W; }
  cac_cell_reduce_kernel<<<dim3(chunks + (gate ? gate_tiles(g) : 0), B), 128, 0, st>>>(
      static_cast<const float2*>(cstat_d), static_cast<const float2*>(cstat_c), cells, chunks, part, g);
  return cudaGetLastError();
}

cudaError_t launch_cac_gate(const CacGate& gate, int B, cudaStream_t st) {
  CacGate g = gate;
  if (g.part_stride == 0) g.part_stride = (size_t)B * g.H * g.W;
  cac_gate_kernel<<<dim3(gate_tiles(g), B), 128, 0, st>>>(g);
  return cudaGetLastError();
}

cudaError_t launch_cac_mlp(const float* part, int chunks, int B, int HW, const float* w1, const float* b1,
                           const float* w2, const float* b2, float* sc, cudaStream_t st) {
  cac_mlp_kernel<<<B, 1024, 0, st>>>(part, chunks, HW, w1, b1, w2, b2, sc);
  return cudaGetLastError();
}

cudaError_t launch_cac_apply(void* F, const void* E, int act, const float* gate, const float* sc, int B, int H, int W,
                             cudaStream_t st, int rnd_tf32) {
  static std::atomic<int> sms_of_dev[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  int num_sms = sms_of_dev[dev].load(std::memory_order_acquire);
  if (num_sms == 0) {
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    sms_of_dev[dev].store(num_sms, std::memory_order_release);
  }
  const int HW = H * W;
  const size_t P = (size_t)B * HW;
  auto go = [&](auto* f, const auto* e) {
    using T = std::remove_pointer_t<decltype(f)>;
    const size_t nvec = P * (128 / Act<T>::kVec);
    const size_t units = (nvec + 255) / 256;
    size_t grid = (size_t)num_sms * CODON_APPLY_CTAS;
    if (grid > units) grid = units;
    cac_apply_kernel<T><<<(unsigned)grid, 256, 0, st>>>(f, e, gate, sc, nvec, HW, rnd_tf32);
  };
  if (act == ACT_F32) go((float*)F, (const float*)E);
  else if (act == ACT_BF16) go((__nv_bfloat16*)F, (const __nv_bfloat16*)E);
  else if (act == ACT_SPLIT16) { rnd_tf32 = 0; go((split16*)F, (const split16*)E); }
  else go((__half*)F, (const __half*)E);
  return cudaGetLastError();
}

}  // namespace codon
