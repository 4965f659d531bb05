// Shared device/host helpers for libcodon_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace codon {

// Activation element types stored in HBM (NHWC): float (FP32/TF32 modes), bf16, fp16, and the split fp16 pair of
// the F16X3 mode (see split16 below).
enum ActType : int { ACT_F32 = 0, ACT_BF16 = 1, ACT_F16 = 2, ACT_SPLIT16 = 3 };

__host__ __device__ inline int act_bytes(int t) { return (t == ACT_F32 || t == ACT_SPLIT16) ? 4 : 2; }

// Split-fp16 activations (F16X3 mode): a value v is carried as hi = fp16(v) and lo = fp16(v - hi), i.e. ~22
// mantissa bits.  A 4-byte container per element so that all element offsets / strides are the fp32 ones, but
// the two halves are stored PLANAR per 64-channel slab -- [64 x hi][64 x lo] = 2 x 128 B -- so that each plane of
// a slab is one K-major 128-byte row for TMA / UMMA (conv_tc.cu).  Pixel bases are 256-byte aligned (strides and
// channel offsets are multiples of 64 elements, buffers 1024-byte aligned), so the address of element (slab s,
// channel c) follows from the element pointer alone: slab base = p & ~255, hi at + 2c, lo at + 128 + 2c.
struct split16 { uint32_t container; };

// fp32 -> (hi, lo) fp16 pairs, round-to-nearest, saturating (|v| > 65504 clamps instead of producing inf).
__device__ __forceinline__ uint32_t cvt_f16x2_sat(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = cvt_f16x2_sat(a, b);
  const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = cvt_f16x2_sat(a - h.x, b - h.y);
}

struct RawSplit { uint4 hi, lo; };

template <typename T> struct Act;
template <> struct Act<float> {
  static constexpr int kVec = 4;  // elements per 16-byte vector
  using Raw = uint4;              // one vector as loaded from memory
  __device__ static inline Raw ldg(const float* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ static inline Raw ld(const float* p) { return *reinterpret_cast<const uint4*>(p); }
  __device__ static inline void load(const float* p, float (&v)[4]) {
    float4 r = *reinterpret_cast<const float4*>(p);
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
  __device__ static inline void unpack(const uint4& r, float (&v)[4]) {
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
  }
  __device__ static inline void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  __device__ static inline float get(const float* p) { return *p; }
  __device__ static inline float to_float(float v) { return v; }
  __device__ static inline float from_float(float v) { return v; }
};
template <> struct Act<__nv_bfloat16> {
  static constexpr int kVec = 8;
  using Raw = uint4;
  __device__ static inline Raw ldg(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ static inline Raw ld(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
  __device__ static inline void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static inline void unpack(const uint4& r, float (&v)[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static inline void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  __device__ static inline float get(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  __device__ static inline float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static inline __nv_bfloat16 from_float(float v) { return __float2bfloat16_rn(v); }
};
template <> struct Act<__half> {
  static constexpr int kVec = 8;
  using Raw = uint4;
  __device__ static inline Raw ldg(const __half* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ static inline Raw ld(const __half* p) { return *reinterpret_cast<const uint4*>(p); }
  __device__ static inline void load(const __half* p, float (&v)[8]) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ static inline void unpack(const uint4& r, float (&v)[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ static inline void store(__half* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  __device__ static inline float get(const __half* p) { return __half2float(*p); }
  __device__ static inline float to_float(__half v) { return __half2float(v); }
  __device__ static inline __half from_float(float v) { return __float2half_rn(v); }
};

template <> struct Act<split16> {
  static constexpr int kVec = 8;   // 8 channels = one 16-byte vector of each plane
  using Raw = RawSplit;
  // p addresses element (pixel, channel c) in container units; c % 8 == 0 for the vector accessors
  __device__ static inline const uint8_t* hi_addr(const split16* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    return reinterpret_cast<const uint8_t*>((a & ~(uintptr_t)255) + ((a & 255) >> 1));
  }
  __device__ static inline Raw ldg(const split16* p) {
    const uint8_t* h = hi_addr(p);
    Raw r;
    r.hi = __ldg(reinterpret_cast<const uint4*>(h));
    r.lo = __ldg(reinterpret_cast<const uint4*>(h + 128));
    return r;
  }
  __device__ static inline Raw ld(const split16* p) {
    const uint8_t* h = hi_addr(p);
    Raw r;
    r.hi = *reinterpret_cast<const uint4*>(h);
    r.lo = *reinterpret_cast<const uint4*>(h + 128);
    return r;
  }
  __device__ static inline void unpack(const Raw& r, float (&v)[8]) {
    const uint32_t h[4] = {r.hi.x, r.hi.y, r.hi.z, r.hi.w}, l[4] = {r.lo.x, r.lo.y, r.lo.z, r.lo.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
      const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&l[i]));
      v[2 * i] = fh.x + fl.x; v[2 * i + 1] = fh.y + fl.y;
    }
  }
  __device__ static inline void load(const split16* p, float (&v)[8]) { unpack(ld(p), v); }
  __device__ static inline void store(split16* p, const float (&v)[8]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_pair(v[2 * i], v[2 * i + 1], h[i], l[i]);
    uint8_t* dst = const_cast<uint8_t*>(hi_addr(p));
    *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(dst + 128) = make_uint4(l[0], l[1], l[2], l[3]);
  }
  // single element (debug taps)
  __device__ static inline float get(const split16* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const __half* h = reinterpret_cast<const __half*>((a & ~(uintptr_t)255) + ((a & 255) >> 1));
    return __half2float(h[0]) + __half2float(h[64]);
  }
};

// Round-to-nearest fp32 -> tf32 (10-bit mantissa), kept in an fp32 container.  The TF32 MMA
// ignores the low 13 mantissa bits (truncation); rounding at the producer halves that error.
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

__device__ __forceinline__ float sigmoidf_exact(float z) { return 1.0f / (1.0f + expf(-z)); }

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace codon
