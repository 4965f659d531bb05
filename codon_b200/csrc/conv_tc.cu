// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (BF16, FP16 and TF32 modes).
//
// Implements the trunk convolutions of CODONNet (CODON_X4/CODON_x4.py:25-46, call sites :69-84,
// :120-129): NHWC activations, stride 1, zero padding, no bias; ReLU and the residual add are
// fused into the epilogue, and the 3x3 || 5x5 multi-scale pair that the reference concatenates
// (:75-80, :123-125) is ONE launch that writes the 128-channel concat directly.
//
// GEMM view: M = pixels, N = output channels (64 or 128), K = taps x input channels.
//   * A (activations): an accumulator covers a sub-tile of 8 x 16 pixels (M = 128, row m = y*8 + x);
//     a CTA tile is NAX x NAY sub-tiles (NACC = 1, 2 or 4).  For every 128-byte channel slab ONE 4-D
//     TMA box {slab, tile_w + ks - 1 px, tile_h + ks - 1 rows} lands a halo patch in shared memory
//     (SWIZZLE_128B, one pixel = one 128-B row; out-of-image pixels are zero-filled by TMA, which is
//     exactly the layer's zero padding).  Every tap (dx, dy) of every sub-tile then reads that one
//     patch through a UMMA descriptor whose start address is shifted by (dy * pitch + dx) pixels and
//     whose 8-row-group stride (SBO) is the patch row pitch: the hardware applies the 128-B swizzle
//     on absolute shared-memory address bits (verified by tools/umma_shift_test.cu), so a shifted
//     start address reads exactly what TMA wrote.  A traffic is one patch per slab and tile
//     ((T+k-1)^2 / T^2 of the input) instead of k*k im2col tiles.
//   * B (weights): host-packed per (slab, tap) K-major blocks already in the SWIZZLE_128B image,
//     streamed through a ring of stages; each block feeds NACC x 4 MMAs.
//   * D: NACC accumulators of 128 lanes x N fp32 columns in TMEM; double-buffered across tiles
//     when 2*NACC*N <= 512 so the epilogue of tile i overlaps the main loop of tile i+1.
//   * conv_tc_kernel (1 CTA): warp 0 producer, warp 1 TMEM owner + MMA issuer, warps 2..5 epilogue.
//     conv_tc2_kernel (cluster of 2 CTAs, tcgen05 cta_group::2, M = 256): see below.
//   * The producer / MMA loops are warp-uniform and one elect.sync lane issues (no per-thread
//     waterfall loops around UTCHMMA / UTMALDG); mbarrier pipelines throughout; persistent CTAs
//     stride over the tile list, and the tiles of the last partial wave are split into single
//     sub-tile items.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <atomic>
#include <utility>
#include <type_traits>
#include "common.cuh"
#include "kernels.h"
#include "conv_tc.h"

namespace codon {

namespace {

constexpr int kMaxDevices = 64;
constexpr int kMaxNPB = 8;                     // patch stages: as many as fit (runtime, TcKParams::npb)
constexpr int kBStages = 4;
constexpr uint32_t kBStageBytes = 128 * 128;   // up to 128 rows x 128 B
constexpr int kThreads = 320;                  // warp 0 producer, 1 MMA, 2-9 epilogue (two warps per TMEM lane quarter)
// Cluster kernel: warps 0-7 epilogue (two per TMEM lane quarter), 8 A producer, 9 MMA issuer (+ TMEM owner), 10 B producer.
// The single-thread roles are the HIGHEST warp ids of their scheduler partition (warp % 4): the arbiter favours the
// highest warp id, so the latency-critical issue loop never queues behind the epilogue's ALU work.
constexpr int kThreads2 = 352;
constexpr int kWarpA = 8, kWarpMma = 9, kWarpB = 10;
constexpr int kB2MaxStages = 10;               // barrier slots reserved for the weight ring of the cluster kernel
// Perf-experiment knobs of the convolution kernels (CODON_TC_DEBUG bits, cycle accounting, CODON_TC_PDL) exist only in
// builds with -DCODON_TC_EXPERIMENT: the MMA issue loop is latency-bound, every extra instruction in it costs
// throughput, and a product library does not read the environment.
#ifdef CODON_TC_EXPERIMENT
#define TC2_DBG(p, bit) (((p).debug & (bit)) != 0)
#else
#define TC2_DBG(p, bit) false
#endif
constexpr uint32_t kBarBytes = 1024;           // barrier block at the start of the dynamic smem

struct TcKParams {
  TcJob job[2];
  int njobs, B, H, W, tiles_x, tiles_y, tiles_per_job, total_tiles;
  int main_tiles, total_items;   // items [0, main_tiles) are full tiles; the tail tiles are split into single sub-tiles
  int nslab, slab_elems, ks, pad, ndx, ndy;
  int dx_ord[kTcMaxTaps], dy_ord[kTcMaxTaps];
  uint32_t b_bytes[kTcMaxTaps][kTcMaxTaps], b_off[kTcMaxTaps][kTcMaxTaps];
  uint32_t slab_bytes;
  int n_cols;
  uint32_t idesc_full, idesc_half;
  int relu, out_act, is_tf32, nbuf;
  int pw;                        // patch width in pixels (tile_w + ks - 1); the patch row pitch is pw * 128 B
  uint32_t patch_tx, patch_stage;   // bytes of one patch (TMA transaction) and of one patch stage (1024-aligned)
  int npb;                       // patch stages in the ring (2 .. kMaxNPB)
  uint32_t idesc_1x1;            // fused 1x1: kind::f16, M = 256, N = 64
  int fuse_njobs;
  unsigned long long pool_stride;   // fused mode: pixels between the two half-channel pool maps of a job
  int cells_x, cells_y;             // fused mode: 8 x 16-pixel cells per frame (TcJob::cstat)
  int core_y0, core_y1;             // fused mode: rows that count towards cstat (row-band mode; whole image otherwise)
  int debug;   // CODON_TC_DEBUG bits (perf experiments only; results are garbage): 1 no epilogue stores, 2 no B loads, 4 no A loads,
               // 8 no MMAs (1-CTA), 16 no waits (1-CTA), 32 no epilogue TMEM loads / Y staging (cluster kernel)
};

template <int NACC> struct Geo {
  static constexpr int NAX = NACC >= 2 ? 2 : 1, NAY = NACC / NAX;
  static constexpr int TW = NAX * kTcSubW, TH = NAY * kTcSubH;
};

// Tap schedule of the cluster kernel, known at compile time so that the MMA issuer and the weight producer are
// straight-line code per tap (descriptor offsets and byte counts are immediates).  Must match fill_orders() /
// tc_make_plan() / tc_make_pair_plan() below (checked on the host before every launch).
enum TcKind : int { TK_3X3 = 0, TK_5X5 = 1, TK_PAIR = 2 };
// split-fp16 mode: taps per accumulation chunk (see conv_tc2_kernel); must divide 9 resp. 25.  Measured on B200
// (profiles/r02_split_chunking.txt): one tap per chunk gives the smallest error (mean |err| 7.2e-8 vs the fp64
// reference, fp32 FFMA mode: 3.6e-8) at 20 MP/s; one column of taps per chunk 1.05e-7 at 25 MP/s, and both reproduce
// the reference's RMSE / SSIM to three decimals on all 30 bundled images -- the faster one is the default.
#ifndef CODON_SPLIT_TPC_3X3
#define CODON_SPLIT_TPC_3X3 3
#endif
#ifndef CODON_SPLIT_TPC_5X5
#define CODON_SPLIT_TPC_5X5 5
#endif
template <int KIND> struct Taps {
  static constexpr int KS = KIND == TK_3X3 ? 3 : 5;
  static constexpr int NT = KS * KS;
  // weight ring of the cluster kernel: stage count (divides NT) and bytes per stage = this CTA's half of the largest
  // block of the kind (3x3: 64 rows -> 32 x 128 B; 5x5 / pair: 128 rows -> 64 x 128 B)
  static constexpr int NST = KIND == TK_3X3 ? 9 : 5;
  static constexpr uint32_t kStageBytes = KIND == TK_3X3 ? 4096u : 8192u;
  __host__ __device__ static constexpr int ord(int i) {
    return KIND == TK_PAIR ? (i == 0 ? 2 : i == 1 ? 1 : i == 2 ? 3 : i == 3 ? 0 : 4) : i;
  }
  __host__ __device__ static constexpr int dx(int t) { return ord(t / KS); }     // issue order: dx outer, dy inner
  __host__ __device__ static constexpr int dy(int t) { return ord(t % KS); }
  // pair plans: the 16 border taps belong to the 5x5 convolution only (64 weight rows, N = 64)
  __host__ __device__ static constexpr bool outer(int t) {
    return KIND == TK_PAIR && (dx(t) == 0 || dx(t) == 4 || dy(t) == 0 || dy(t) == 4);
  }
  // pair plans, split mode: bit c = accumulation chunk c (taps [c * tpc, (c + 1) * tpc)) starts with a border tap
  __host__ __device__ static constexpr uint32_t outer_chunk_mask(int tpc) {
    uint32_t m = 0;
    for (int c = 0; c * tpc < NT; ++c) if (outer(c * tpc)) m |= 1u << c;
    return m;
  }
  // pair plans: 64-row (8 KB) units of one slab's weight stream that precede tap t
  __host__ __device__ static constexpr int units_before(int t) {
    int u = 0;
    for (int i = 0; i < t; ++i) u += outer(i) ? 1 : 2;
    return u;
  }
};
// Column structure of the 5-wide kinds (TK_5X5, TK_PAIR).  The issue order is dx outer / dy inner, so a "column" is the
// five taps 5c .. 5c + 4 that share dx.  Every column of a kind looks the same to the weight ring (its uses are a whole
// number of ring revolutions) and to the split mode's accumulation chunks, so the MMA issuer and the weight producer
// are LOOPS over columns whose bodies are straight-line per tap: only dx (one descriptor add), one ring parity bit and
// the first / last-column flags are run-time values.  Why: the fully unrolled 25-tap loops were 43 KB (issuer) +
// 17-32 KB (weight producer) of SASS per kernel next to ~25 KB of epilogue code, against a 32 KB instruction cache
// (L1.5) -- the latency-critical single-warp loops ran out of L2.  Pair kernels have two column types: type 0 (columns
// 0-2 of the centre-first order: three inner taps feeding both convolutions, then two 5x5-only border taps) and type 1
// (columns 3-4: border taps only).
#ifndef CODON_TC_COLROLL
#define CODON_TC_COLROLL 1
#endif
// Taps (ring stages) per elect / fence / branch block of the non-split issuer inside a column.  The issue loop's
// bookkeeping limits the tensor pipe: two taps per block measured +5.5 % (bf16) / +2.5 % (tf32) on the fused 5x5 kernels
// (profiles/r02_ab_tap_group.txt); the 3x3 || 5x5 pair does not move (tf32 -2 %), so only the 5x5 kind groups.
#ifndef CODON_TC_COL_GROUP
#define CODON_TC_COL_GROUP 0      // 0: per kind -- 5x5: 2 (blocks of 2 + 2 + 1 taps per column); pair: 1
#endif
template <int KIND> struct Cols {
  using TP = Taps<KIND>;
  __host__ __device__ static constexpr int ntypes() { return KIND == TK_PAIR ? 2 : 1; }
  __host__ __device__ static constexpr int type_of(int c) { return KIND == TK_PAIR && c >= 3 ? 1 : 0; }
  __host__ __device__ static constexpr bool outer(int ct, int i) { return KIND == TK_PAIR && (ct == 1 || i >= 3); }
  // pair plans: 64-row units of a column's weight stream that precede its tap i / that precede column c
  __host__ __device__ static constexpr int units_in_col(int ct, int i) { return ct == 1 ? i : (i < 3 ? 2 * i : 6 + (i - 3)); }
  __host__ __device__ static constexpr int units_before_col(int c) { return c < 3 ? 8 * c : 24 + 5 * (c - 3); }
  // the column view must be the tap schedule the weights were packed with
  __host__ __device__ static constexpr bool consistent() {
    if (KIND == TK_3X3) return true;
    for (int c = 0; c < 5; ++c)
      for (int i = 0; i < 5; ++i) {
        const int t = 5 * c + i;
        if (TP::dx(t) != TP::ord(c) || TP::dy(t) != TP::ord(i)) return false;
        if (TP::outer(t) != outer(type_of(c), i)) return false;
        if (KIND == TK_PAIR && TP::units_before(t) != units_before_col(c) + units_in_col(type_of(c), i)) return false;
      }
    return true;
  }
};
static_assert(Cols<TK_5X5>::consistent() && Cols<TK_PAIR>::consistent(), "column view of the tap schedule");

// Weight-ring depth of the cluster kernel.  The split-fp16 pair kernel streams two short blocks per tap (258-384 cycles
// of MMA work each): five stages are ~1500 cycles of look-ahead, less than the L2 -> shared-memory latency under load
// (its issuer spent 47 % of its time waiting for weights, profiles/r02_pair_ring.txt), and its 30 KB patches leave room
// for ten.
// The pair kernels of the other modes wait for weights too (95 of 490 cycles per tap at 5 stages), but ten stages there
// measured SLOWER (bf16 0.78 -> 0.87 ms, tf32 1.54 -> 1.59 ms per step, profiles/r02_pair_ring.txt): 25 taps do not
// divide by ten, so the tap loops exist twice (the ring pattern repeats every two slabs, kPeriod below) and the 51 KB
// patches drop from three stages to two.  The two-phase machinery stays (any stage count whose pattern repeats within
// two slabs keeps compile-time stage indices); only the split pair kernel uses a deeper ring.
template <int KIND, int OPERAND> struct Ring {
  static constexpr int NST = (KIND == TK_PAIR && OPERAND == TC_SPLIT16) ? 10 : Taps<KIND>::NST;
};
__host__ __device__ constexpr int cgcd(int a, int b) { return b == 0 ? a : cgcd(b, a % b); }
template <typename F, int... Is>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, Is...>) {
  (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, typename F>
__device__ __forceinline__ void static_for(F&& f) { static_for_impl(f, std::make_integer_sequence<int, N>{}); }

// ------------------------------------------------------------------------------------------------
// PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// slow path of a wait, out of line: spins with a watchdog (a protocol bug must fault, never hang the GPU)
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("conv_tc: mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try(bar, parity)) mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// One lane of a converged warp (the canonical way to issue tcgen05 / TMA instructions: the
// surrounding control flow stays warp-uniform, so the compiler emits plain uniform-datapath
// UTCHMMA / UTMALDG instead of a per-thread "waterfall" loop around each of them).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int OPERAND> struct OperandTraits;
template <> struct OperandTraits<TC_F16> { using Out = __half; };
template <> struct OperandTraits<TC_BF16> { using Out = __nv_bfloat16; };
template <> struct OperandTraits<TC_TF32> { using Out = float; };
template <> struct OperandTraits<TC_SPLIT16> { using Out = split16; };

struct Tile { int job, n, y0, x0, nacc, valid; };
// Work item -> tile.  Full tiles hold NACC sub-tiles (NAX x NAY).  The tiles of the last, partial
// wave are handed out as NACC separate single-sub-tile items, so the tail of the launch costs ~1/NACC
// of a wave.
template <int NACC>
__device__ __forceinline__ Tile decode_tile(const TcKParams& p, int item) {
  using G = Geo<NACC>;
  Tile r;
  int t = item, sub = 0;
  r.nacc = NACC;
  if (item >= p.main_tiles) {
    const int k = item - p.main_tiles;
    t = p.main_tiles + k / NACC;
    sub = k - (k / NACC) * NACC;
    r.nacc = 1;
  }
  r.job = t / p.tiles_per_job;
  int q = t - r.job * p.tiles_per_job;
  const int per_frame = p.tiles_x * p.tiles_y;
  r.n = q / per_frame;
  q -= r.n * per_frame;
  r.y0 = (q / p.tiles_x) * G::TH + (sub / G::NAX) * kTcSubH;
  r.x0 = (q % p.tiles_x) * G::TW + (sub % G::NAX) * kTcSubW;
  r.valid = (r.y0 < p.H) && (r.x0 < p.W);
  return r;
}

// K-major SWIZZLE_128B descriptor with a zero start address and the given 8-row-group stride.
__device__ __forceinline__ uint64_t umma_desc_hi(uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(sbo_bytes >> 4) << 32;          // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint64_t desc_addr(uint32_t saddr) { return (uint64_t)((saddr & 0x3FFFFu) >> 4); }

// Epilogue of one 32-channel chunk held by one thread (one pixel).
template <typename T>
__device__ __forceinline__ void store_chunk(const uint32_t (&r)[32], const TcJob& job, size_t pix, int c0,
                                            bool relu, bool rnd_tf32 = false, float scale = 1.f) {
  constexpr int V = Act<T>::kVec;
  T* out = static_cast<T*>(job.out) + pix * job.out_stride + job.out_off + c0;
  const T* res = job.res ? static_cast<const T*>(job.res) + pix * job.res_stride + job.res_off + c0 : nullptr;
#pragma unroll
  for (int v = 0; v < 32 / V; ++v) {
    float f[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      f[j] = __uint_as_float(r[v * V + j]);
      if (std::is_same<T, split16>::value) f[j] *= scale;
      if (relu) f[j] = fmaxf(f[j], 0.f);
    }
    if (res) {
      float e[V];
      Act<T>::load(res + v * V, e);
#pragma unroll
      for (int j = 0; j < V; ++j) f[j] += e[j];
    }
    if (rnd_tf32) {
#pragma unroll
      for (int j = 0; j < V; ++j) f[j] = round_tf32(f[j]);
    }
    Act<T>::store(out + v * V, f);
  }
}

template <int NACC, int OPERAND>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1,
               const __grid_constant__ TcKParams p) {
  using OutT = typename OperandTraits<OPERAND>::Out;
  using G = Geo<NACC>;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_bar = sbase;
  const uint32_t s_patch = sbase + kBarBytes;
  const uint32_t s_b = s_patch + p.npb * p.patch_stage;
  // barrier map (8 B each)
  const uint32_t bar_patch_full = s_bar, bar_patch_empty = s_bar + 8 * kMaxNPB;
  const uint32_t bar_b_full = s_bar + 16 * kMaxNPB, bar_b_empty = bar_b_full + 8 * kBStages;
  const uint32_t bar_acc_full = bar_b_empty + 8 * kBStages, bar_acc_empty = bar_acc_full + 16;
  const uint32_t s_tmem_slot = bar_acc_empty + 16;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (s_tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t pitch = (uint32_t)p.pw * 128u;    // patch row pitch in bytes

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap1) : "memory");
    for (int i = 0; i < p.npb; ++i) { mbar_init(bar_patch_full + 8 * i, 1); mbar_init(bar_patch_empty + 8 * i, 1); }
    for (int i = 0; i < kBStages; ++i) { mbar_init(bar_b_full + 8 * i, 1); mbar_init(bar_b_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_acc_full + 8 * i, 1); mbar_init(bar_acc_empty + 8 * i, 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tmem_slot), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================================ producer (warp-uniform loops, one elected lane issues) ====
    int ps = 0, bs = 0;
    uint32_t pph = 0, bph = 0;
    for (int t = blockIdx.x; t < p.total_items; t += gridDim.x) {
      const Tile tl = decode_tile<NACC>(p, t);
      if (!tl.valid) continue;                    // empty sub-tile of a split tail tile
      const TcJob& job = p.job[tl.job];
      for (int s = 0; s < p.nslab; ++s) {
        const uint8_t* wslab = job.w + (size_t)s * p.slab_bytes;
        mbar_wait(bar_patch_empty + 8 * ps, pph ^ 1);
        if (elect_one()) {
          if (TC2_DBG(p, 4)) mbar_arrive(bar_patch_full + 8 * ps);
          else {
            mbar_expect_tx(bar_patch_full + 8 * ps, p.patch_tx);
            tma_load_4d(s_patch + ps * p.patch_stage, tl.job ? &tmap1 : &tmap0, bar_patch_full + 8 * ps,
                        job.in_coff + s * p.slab_elems, tl.x0 - p.pad, tl.y0 - p.pad, tl.n);
          }
        }
        __syncwarp();
        if (++ps == p.npb) { ps = 0; pph ^= 1; }
        for (int dxi = 0; dxi < p.ndx; ++dxi) {
          for (int dyi = 0; dyi < p.ndy; ++dyi) {
            mbar_wait(bar_b_empty + 8 * bs, bph ^ 1);
            if (elect_one()) {
              const uint32_t bytes = p.b_bytes[dxi][dyi];
              if (TC2_DBG(p, 2)) mbar_arrive(bar_b_full + 8 * bs);
              else {
                mbar_expect_tx(bar_b_full + 8 * bs, bytes);
                bulk_load(s_b + bs * kBStageBytes, wslab + p.b_off[dxi][dyi], bytes, bar_b_full + 8 * bs);
              }
            }
            __syncwarp();
            if (++bs == kBStages) { bs = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (warp-uniform loops, one elected lane issues) ==
    int ps = 0, bs = 0;
    uint32_t pph = 0, bph = 0;
    int it = 0;
    const uint64_t desc_a = umma_desc_hi(pitch), desc_b = umma_desc_hi(1024);
    for (int t = blockIdx.x; t < p.total_items; t += gridDim.x) {
      const Tile tl = decode_tile<NACC>(p, t);
      if (!tl.valid) continue;
      const int outer_col = p.job[tl.job].outer_col;
      const int buf = it % p.nbuf;
      const uint32_t aph = (uint32_t)(it / p.nbuf) & 1u;
      mbar_wait(bar_acc_empty + 8 * buf, aph ^ 1);
      tc_fence_after();
      const uint32_t d_base = tmem_base + (uint32_t)(buf * NACC * p.n_cols);
      uint32_t acc0 = 0;                        // 0 only for the very first MMA group of the tile
      for (int s = 0; s < p.nslab; ++s) {
        if (!TC2_DBG(p, 16)) mbar_wait(bar_patch_full + 8 * ps, pph);
        const uint32_t patch = s_patch + ps * p.patch_stage;
        for (int dxi = 0; dxi < p.ndx; ++dxi) {
          for (int dyi = 0; dyi < p.ndy; ++dyi) {
            if (!TC2_DBG(p, 16)) mbar_wait(bar_b_full + 8 * bs, bph);
            tc_fence_after();
            const bool half = p.b_bytes[dxi][dyi] < (uint32_t)p.n_cols * 128u;
            const uint32_t idesc = half ? p.idesc_half : p.idesc_full;
            const uint32_t d0 = d_base + (half ? (uint32_t)outer_col : 0u);
            const uint64_t bdesc = desc_b | desc_addr(s_b + bs * kBStageBytes);
            // tap (dx, dy): the patch shifted by dy rows and dx pixels
            const uint64_t adesc0 = desc_a | desc_addr(patch + (uint32_t)p.dy_ord[dyi] * pitch + (uint32_t)p.dx_ord[dxi] * 128u);
            const bool last = (dxi == p.ndx - 1) && (dyi == p.ndy - 1);
            if (elect_one()) {
              if (!TC2_DBG(p, 8)) {
#pragma unroll
                for (int j = 0; j < NACC; ++j) {
                  if (j >= tl.nacc) break;
                  // sub-tile j = (jx, jy): + jy*16 patch rows + jx*8 pixels
                  const uint64_t adesc = adesc0 + (uint64_t)(((j / G::NAX) * kTcSubH * pitch + (j % G::NAX) * kTcSubW * 128u) >> 4);
                  const uint32_t d = d0 + (uint32_t)(j * p.n_cols);
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    // +32 B per K step inside the 128-B swizzled row == +2 in the 16-B address field
                    if (OPERAND == TC_TF32) umma_tf32(d, adesc + 2 * k, bdesc + 2 * k, idesc, k == 0 ? acc0 : 1u);
                    else                    umma_f16(d, adesc + 2 * k, bdesc + 2 * k, idesc, k == 0 ? acc0 : 1u);
                  }
                }
              }
              umma_commit(bar_b_empty + 8 * bs);
              if (last) umma_commit(bar_patch_empty + 8 * ps);
              if (last && s == p.nslab - 1) umma_commit(bar_acc_full + 8 * buf);
            }
            __syncwarp();
            acc0 = 1;
            if (++bs == kBStages) { bs = 0; bph ^= 1; }
          }
        }
        if (++ps == p.npb) { ps = 0; pph ^= 1; }
      }
      ++it;
    }
  } else {
    // ================================ epilogue =================================================
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int ehalf = (warp - 2) >> 2;           // 0 / 1: which half of the chunks this warp drains
    const int m = q * 32 + lane;                 // accumulator row == pixel inside the sub-tile
    const int my = m / kTcSubW, mx = m % kTcSubW;
    int it = 0;
    for (int t = blockIdx.x; t < p.total_items; t += gridDim.x) {
      const Tile tl = decode_tile<NACC>(p, t);
      if (!tl.valid) continue;
      const TcJob& job = p.job[tl.job];
      const int buf = it % p.nbuf;
      const uint32_t aph = (uint32_t)(it / p.nbuf) & 1u;
      mbar_wait(bar_acc_full + 8 * buf, aph);
      tc_fence_after();
      // software-pipelined drain: the TMEM load of chunk i+1 is in flight while chunk i is converted
      // and stored (chunk = 32 accumulator columns of one 128-pixel sub-tile)
      const int cpa_sh = p.n_cols == 128 ? 2 : 1, nchunk = tl.nacc << cpa_sh;    // chunks per accumulator: 4 or 2
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * NACC * p.n_cols);
      auto issue = [&](int i, uint32_t (&r)[32]) { tmem_ld32(lane_base + (uint32_t)(i << 5), r); };
      const bool pool_mode = job.pool != nullptr;           // 64-column launches only (host-checked)
      float pmax = -INFINITY, psum = 0.f;
      auto drain = [&](int i, const uint32_t (&r)[32]) {
        const int j = i >> cpa_sh, c0 = (i - (j << cpa_sh)) << 5;
        const int py = tl.y0 + (j / G::NAX) * kTcSubH + my, px = tl.x0 + (j % G::NAX) * kTcSubW + mx;
        if ((py < p.H) && (px < p.W) && !TC2_DBG(p, 1)) {
          const size_t pix = ((size_t)tl.n * p.H + py) * p.W + px;
          store_chunk<OutT>(r, job, pix, c0, p.relu != 0, OPERAND == TC_TF32);
          if (pool_mode) {
            // ChannelPool partial over this job's 64 channels (the thread sees both 32-column chunks of its pixel)
            float m = __uint_as_float(r[0]), sacc = 0.f;
#pragma unroll
            for (int e = 0; e < 32; ++e) { const float v = __uint_as_float(r[e]); m = fmaxf(m, v); sacc += v; }
            if (c0 == 0) { pmax = m; psum = sacc; }
            else job.pool[pix] = make_float2(fmaxf(pmax, m), psum + sacc);
          }
        }
      };
      // Chunk schedule of this warp.  Normally the two warps of a lane quarter take the even / odd chunks;
      // in pool mode they take alternate accumulators (both chunks), so a thread owns all 64 channels.
      auto nxt = [&](int i) { return pool_mode ? ((i & 1) ? i + 3 : i + 1) : i + 2; };
      uint32_t ra[32], rb[32];
      int i = pool_mode ? 2 * ehalf : ehalf;
      if (i < nchunk) {
        issue(i, ra);
#pragma unroll 1
        while (true) {
          tmem_ld_wait();
          const int i1 = nxt(i);
          if (i1 < nchunk) issue(i1, rb);
          drain(i, ra);
          if (i1 >= nchunk) break;
          tmem_ld_wait();
          const int i2 = nxt(i1);
          if (i2 < nchunk) issue(i2, ra);
          drain(i1, rb);
          if (i2 >= nchunk) break;
          i = i2;
        }
      }
      tc_fence_before();
      mbar_arrive(bar_acc_empty + 8 * buf);
      ++it;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ================================================================================================
// 2-CTA variant: a cluster of two CTAs (one SM pair) computes two tiles with ONE stream of
// tcgen05.mma.cta_group::2 instructions (M = 256: 128 pixels of each CTA's tile).  Each CTA stages its
// own activation patch and only HALF of every weight block (N/2 rows), so the per-SM shared-memory
// operand traffic per MMA drops from A+B to A+B/2, the weight fetches from L2 halve, and the MMA
// instruction count per pixel halves.  The leader CTA (cluster rank 0) issues the MMAs; both CTAs'
// TMA loads complete on the leader's mbarriers (cp.async.bulk.tensor ... .cta_group::2), and
// tcgen05.commit multicasts the "stage free" / "accumulator ready" arrivals to both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on a (possibly remote) mbarrier of the cluster.  Relaxed: the only thing the waiter (the MMA
// issuer) may touch afterwards is TMEM, whose reads were completed by tcgen05.wait::ld and ordered by
// tcgen05.fence::before_thread_sync; a release at cluster scope would drain every outstanding global
// store of the epilogue first (MEMBAR.ALL + ERRBAR, ~15 % of the epilogue in the r01b profile).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0,
                                                int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}


__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Non-blocking phase test.  CTA-scope acquire, also for barriers the peer CTA arrives on with a cluster-scope
// release: a cluster-scope acquire here costs a few thousand cycles per call (CCTL.IVALL + fence, measured with
// CODON_TC_DEBUG=64), and what the waiter goes on to touch is TMEM / the async proxy, not generic-proxy data.
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
template <int OP16> __device__ __forceinline__ uint32_t pack16(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack16<TC_BF16>(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack16<TC_F16>(float lo, float hi) {
  // saturating: a post-ReLU activation above 65504 (possible with fp32 storage in the tf32 mode) clamps to the
  // largest finite fp16 instead of turning into inf and poisoning the 1x1 product
  return cvt_f16x2_sat(lo, hi);
}

// Split mode, fused 1x1: state of the hand-offs the MMA issuer still owes the epilogue, and their out-of-line
// service routine (the issue loop is ~50 straight-line "uses" per slab; inlining the 12 MMAs of a hand-off into each
// of them made the loop ~10x larger than the instruction cache and doubled the cycles per MMA).
// All state is passed by value (it stays in the caller's registers); returns the number of hand-offs issued.
__device__ __noinline__ int fuse_service_split(int left, int j, int job, uint32_t d, uint32_t y_uses, bool wc_ready,
                                               uint32_t bar_y_full, uint32_t bar_y_done, uint32_t bar_wc_full,
                                               uint32_t s_y, uint32_t s_wc, uint32_t idesc_1x1, bool block);

struct Tile2 { int job, n, y0, x0, valid, nacc; };
// Work item -> this CTA's tile.  Items [0, main_tiles) are pair-tiles (tiles 2q and 2q+1 of a job);
// the pair-tiles of the last partial wave are handed out as NACC single-sub-tile items each.
template <int NACC>
__device__ __forceinline__ Tile2 decode_tile2(const TcKParams& p, int item, int rank) {
  using G = Geo<NACC>;
  Tile2 r;
  int pt = item, sub = 0;
  r.nacc = NACC;
  if (item >= p.main_tiles) {
    const int k = item - p.main_tiles;
    pt = p.main_tiles + k / NACC;
    sub = k - (k / NACC) * NACC;
    r.nacc = 1;
  }
  const int ppj = (p.tiles_per_job + 1) >> 1;        // pair-tiles per job
  r.job = pt / ppj;
  int t = 2 * (pt - r.job * ppj) + rank;
  r.valid = t < p.tiles_per_job;
  if (!r.valid) t = p.tiles_per_job - 1;             // odd tile count: the peer recomputes the last tile, stores nothing
  const int per_frame = p.tiles_x * p.tiles_y;
  r.n = t / per_frame;
  t -= r.n * per_frame;
  r.y0 = (t / p.tiles_x) * G::TH + (sub / G::NAX) * kTcSubH;
  r.x0 = (t % p.tiles_x) * G::TW + (sub % G::NAX) * kTcSubW;
  if (r.y0 >= p.H || r.x0 >= p.W) { r.valid = 0; r.y0 = 0; r.x0 = 0; }   // empty sub-tile: compute, store nothing
  return r;
}

// What the MMA issuer needs of a work item: its job and how many accumulators it covers (no tile coordinates: the
// issuer's loop is the critical path of the kernel, and the full decode is half a dozen integer divisions).
template <int NACC>
__device__ __forceinline__ void decode_job_nacc(const TcKParams& p, int item, int& job, int& nacc) {
  int pt = item;
  nacc = NACC;
  if (item >= p.main_tiles) { pt = p.main_tiles + (item - p.main_tiles) / NACC; nacc = 1; }
  const int ppj = (p.tiles_per_job + 1) >> 1;
  job = pt >= ppj ? 1 : 0;                           // at most two jobs per launch
}

// FUSE: the launch is a 5x5 128->128 (+ReLU) convolution followed by a 1x1 128->64 convolution (confuse /
// confuse_c / confuse_fuse, CODON_x4.py:83-84,127).  The epilogue drains each accumulator through ReLU into a
// 16-bit K-major SWIZZLE_128B tile Y in shared memory (fence.proxy.async), the MMA issuer multiplies it with
// the resident 1x1 weights (8 MMAs, N = 64) into the first 64 columns of the same accumulator, and the
// epilogue drains those (+ the fusion-stage residual) to global memory.  The 128-channel intermediate never
// touches HBM and the stand-alone 1x1 launch disappears.
template <int NACC, int OPERAND, bool FUSE, int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1,
                const __grid_constant__ CUtensorMap bmap0, const __grid_constant__ CUtensorMap bmap1,
                const __grid_constant__ CUtensorMap wmap0, const __grid_constant__ CUtensorMap wmap1,
                const __grid_constant__ TcKParams p) {
  using OutT = typename OperandTraits<OPERAND>::Out;
  using G = Geo<NACC>;
  using TP = Taps<KIND>;
  // Weight ring: NST stages whose count divides the taps of a slab, so that the stage of tap t (t % NST) and the
  // use count inside a slab (t / NST) are compile-time constants; the only run-time state is one parity bit per slab.
  constexpr int NST = Ring<KIND, OPERAND>::NST;
  constexpr uint32_t kStageBytes = TP::kStageBytes;
  constexpr int kUsesOfSlab = (OPERAND == TC_SPLIT16 ? 2 : 1) * TP::NT;          // ring uses per slab
  constexpr int kPeriod = NST / cgcd(kUsesOfSlab, NST);          // slabs after which the ring pattern repeats (1 or 2)
  constexpr int kUsesPerSlab = kPeriod * kUsesOfSlab / NST;      // uses of one stage per period; odd: its parity flips per period
  static_assert(NST <= kB2MaxStages && kPeriod <= 2, "weight ring geometry");
  constexpr int Y16 = OPERAND == TC_BF16 ? TC_BF16 : TC_F16;   // tf32 mode stages Y in fp16 (same 10-bit mantissa)
  // Split-fp16 operands (F16X3 mode): per 64-channel slab TWO activation patches (hi, lo planes) and per tap TWO
  // weight blocks (hi, lo) -- ring "uses" u = 2 * tap + plane -- and three MMA groups per tap: A_hi*B_hi and
  // A_lo*B_hi on the hi block, A_hi*B_lo on the lo block, all into the same accumulator.
  //   The tensor core truncates (round-toward-zero) the fp32 accumulator after every MMA; over the 600 MMAs of a
  // 5x5 x 128-channel output that bias is ~40x the error of an fp32 FFMA loop (measured, and reproduced by a CPU
  // model of RZ accumulation).  Split mode therefore keeps two kinds of TMEM accumulators:
  //   "big"   (hi*hi) in CHUNKS of kTPC taps of one slab: a chunk starts a fresh accumulator, and the epilogue warps
  //           promote every finished chunk to round-to-nearest fp32 running sums in registers.  Three rotating
  //           buffers of n_cols columns (kSplitBufs): two chunks can be in flight while one is being promoted.
  //   "small" (lo*hi + hi*lo, 2^-11 of the magnitude: its own truncation is irrelevant, and its additions no longer
  //           truncate at the big sum's ulp) once per tile, n_cols columns behind the big buffers; read together
  //           with the tile's last chunk and handed back through bar_small_empty.
  constexpr bool SPLIT = OPERAND == TC_SPLIT16;
  constexpr int NU = SPLIT ? 2 * TP::NT : TP::NT;                // weight-ring uses per slab
  constexpr int kTPC = KIND == TK_3X3 ? CODON_SPLIT_TPC_3X3 : CODON_SPLIT_TPC_5X5;   // taps per accumulation chunk
  constexpr int kSplitBufs = 3;
  // column-rolled tap loops (5x5 and pair kinds, see Cols): uses of the weight ring per column and the change of the
  // ring's revolution parity per column
  constexpr bool kColRoll = CODON_TC_COLROLL != 0 && KIND != TK_3X3;
  using CL = Cols<KIND>;
  constexpr int kNpl = SPLIT ? 2 : 1;
  constexpr int CU = 5 * kNpl;
  constexpr uint32_t kColFlip = (uint32_t)((CU / NST) & 1);
  constexpr int kColGroup = CODON_TC_COL_GROUP ? CODON_TC_COL_GROUP : (KIND == TK_5X5 ? 2 : 1);   // taps per issue block
  using G_ = Geo<NACC>;
  static_assert(!kColRoll || (CU % NST == 0 && (kTPC == 1 || kTPC == 5)), "a column is whole ring revolutions and whole chunks");
  static_assert(TP::NT % kTPC == 0, "chunks must tile the taps of a slab");
  static_assert(!SPLIT || NACC == 1, "split mode: one accumulator (big + small, double-buffered) per tile");
  constexpr uint32_t kPitch = (uint32_t)(G::TW + TP::KS - 1) * 128u;   // patch row pitch in bytes (== p.pw * 128)

  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_bar = sbase;
  const uint32_t s_patch = sbase + kBarBytes;
  const uint32_t s_b = s_patch + p.npb * p.patch_stage;
  const uint32_t bar_patch_full = s_bar, bar_patch_empty = s_bar + 8 * kMaxNPB;
  const uint32_t bar_b_full = s_bar + 16 * kMaxNPB, bar_b_empty = bar_b_full + 8 * kB2MaxStages;
  const uint32_t bar_acc_full = bar_b_empty + 8 * kB2MaxStages, bar_acc_empty = bar_acc_full + 32;   // up to 4 TMEM buffers
  const uint32_t bar_y_full = bar_acc_empty + 32, bar_y_done = bar_y_full + 8, bar_wc_full = bar_y_done + 8;
  const uint32_t bar_small_empty = bar_wc_full + 8;     // split mode: the per-tile "small" accumulator has been read
  const uint32_t s_tmem_slot = bar_small_empty + 8;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (s_tmem_slot - smem_u32(smem_raw)));
  // fused mode: [1x1 weights: 2 jobs x 2 slabs x 32 rows x 128 B = 16 KB][Y: 2 slabs x 128 rows x 128 B = 32 KB]
  // split:      [2 jobs x 2 slabs x (hi, lo) x 32 rows x 128 B = 32 KB][Y: (hi, lo) x 128 rows x 128 B = 32 KB, ONE
  //             64-channel slab at a time: an accumulator is handed over in two rounds]
  const uint32_t s_wc = s_b + NST * kStageBytes;
  const uint32_t s_y = s_wc + (SPLIT ? 32768u : 16384u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int total_items = p.total_items;

  if (warp == kWarpA && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&bmap0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&bmap1) : "memory");
    for (int i = 0; i < p.npb; ++i) { mbar_init(bar_patch_full + 8 * i, 1); mbar_init(bar_patch_empty + 8 * i, 1); }
    for (int i = 0; i < NST; ++i) { mbar_init(bar_b_full + 8 * i, 1); mbar_init(bar_b_empty + 8 * i, 1); }
    // accumulator hand-backs: every epilogue thread of both CTAs arrives (512); split mode hands a buffer back per
    // accumulation chunk, so there one lane per epilogue warp arrives (2 x 8) -- 512 serialised remote arrivals per
    // chunk were most of the promotion's turn-around time
    for (int i = 0; i < 4; ++i) { mbar_init(bar_acc_full + 8 * i, 1); mbar_init(bar_acc_empty + 8 * i, SPLIT ? 16 : 512); }
    mbar_init(bar_y_full, 512); mbar_init(bar_y_done, 1); mbar_init(bar_wc_full, 1); mbar_init(bar_small_empty, 16);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWarpMma) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tmem_slot), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // Programmatic dependent launch: the kernel may start while the previous kernel of the stream is still in its tail
  // wave.  The set-up above and the weight producer's first ring fill (weights are constants) overlap that tail; every
  // warp that reads activations or writes results waits here until the previous grid has completed and flushed.
  if (warp != kWarpB) asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == kWarpA) {
    // ================================ A producer (both CTAs): one activation patch per slab ======
    int ps = 0;
    uint32_t pph = 0;
    const uint32_t full_leader = mapa_u32(bar_patch_full, 0);
    for (int item = cluster_id; item < total_items; item += nclusters) {
      const Tile2 tl = decode_tile2<NACC>(p, item, (int)rank);
      const int coff = p.job[tl.job].in_coff;
      const int npatch = SPLIT ? 2 * p.nslab : p.nslab;   // split: patch 2s = hi plane, 2s + 1 = lo plane of slab s
      for (int s = 0; s < npatch; ++s) {
        // channel coordinate in the tensor map (split: fp16 units, a slab is [64 hi | 64 lo])
        const int ccoord = SPLIT ? 2 * coff + (s >> 1) * 128 + (s & 1) * 64 : coff + s * p.slab_elems;
        mbar_wait(bar_patch_empty + 8 * ps, pph ^ 1);
        if (elect_one()) {
          if (TC2_DBG(p, 4)) { if (leader) mbar_arrive(bar_patch_full + 8 * ps); }
          else {
            if (leader) mbar_expect_tx(bar_patch_full + 8 * ps, 2 * p.patch_tx);
            tma_load_4d_2sm(s_patch + ps * p.patch_stage, tl.job ? &tmap1 : &tmap0, full_leader + 8 * ps,
                            ccoord, tl.x0 - p.pad, tl.y0 - p.pad, tl.n);
          }
        }
        __syncwarp();
        if (++ps == p.npb) { ps = 0; pph ^= 1; }
      }
    }
  } else if (warp == kWarpB) {
    // ================================ B producer (both CTAs): this CTA's half of every weight block
    uint32_t slab_par = 0;                       // parity of the ring uses of the current period (flips per period)
    int phase = 0;                               // slab inside the period (kPeriod == 2: pair kernels with ten stages)
    const uint32_t full_leader = mapa_u32(bar_b_full, 0);
    if (FUSE && elect_one()) {
      // resident 1x1 weights: this CTA's 32 of the 64 output rows, per job and 128-byte K slab
      constexpr uint32_t kPl = SPLIT ? 2u : 1u;      // planes per slab of the 1x1 stream: [hi 64 rows][lo 64 rows]
      if (leader) mbar_expect_tx(bar_wc_full, (uint32_t)p.fuse_njobs * 2u * kPl * 8192u);
      const uint32_t wc_leader = mapa_u32(bar_wc_full, 0);
      for (int jb = 0; jb < p.fuse_njobs; ++jb)
        for (uint32_t sp = 0; sp < 2u * kPl; ++sp)     // sp = slab * planes + plane
          tma_load_2d_2sm(s_wc + ((uint32_t)jb * 2u * kPl + sp) * 4096u, jb ? &wmap1 : &wmap0, wc_leader, 0,
                          (int)((sp * 8192u + rank * 4096u) >> 7));
    }
    __syncwarp();
    // rows (128 B) of one tap's weight block: the pair plan mixes 128-row (inner) and 64-row (outer) blocks
    const int tap_rows = p.n_cols;
    uint32_t par = 0;                            // column-rolled loops: parity of the ring revolution at the column's start
    for (int item = cluster_id; item < total_items; item += nclusters) {
      const Tile2 tl = decode_tile2<NACC>(p, item, (int)rank);
      const CUtensorMap* bm = tl.job ? &bmap1 : &bmap0;
      for (int s = 0; s < p.nslab; ++s) {
        const int slab_row0 = (int)(((uint32_t)s * p.slab_bytes) >> 7);
        if constexpr (kColRoll) {
          // one column of taps: CU ring uses in issue order (split: use v = 2 * tap + plane); `first` = 128-byte rows of
          // the slab's weight stream that precede the column, per plane
          auto produce_column = [&](auto CT, const int first) {
            constexpr int ct = decltype(CT)::value;
            static_for<CU>([&](auto V) {
              constexpr int v = decltype(V)::value, i = v / kNpl, plane = v % kNpl, st = v % NST;
              constexpr uint32_t pf = (uint32_t)((v / NST) & 1);
              constexpr bool outer = CL::outer(ct, i);
              mbar_wait(bar_b_empty + 8 * st, par ^ pf ^ 1u);
              if (elect_one()) {
                // this CTA's half of the block: rows [rank * rows/2, (rank + 1) * rows/2) in 32-row (4 KB) boxes
                const int rows = KIND == TK_PAIR ? (outer ? 64 : 128) : tap_rows;
                const int row0 = slab_row0 + kNpl * (first + (KIND == TK_PAIR ? CL::units_in_col(ct, i) * 64 : i * tap_rows)) +
                                 plane * rows + (int)rank * (rows >> 1);
                if (TC2_DBG(p, 2)) { if (leader) mbar_arrive(bar_b_full + 8 * st); }
                else {
                  if (leader) mbar_expect_tx(bar_b_full + 8 * st, (uint32_t)rows << 7);
                  const uint32_t dst = s_b + (uint32_t)st * kStageBytes;
                  tma_load_2d_2sm(dst, bm, full_leader + 8 * st, 0, row0);
                  if (rows > 64) tma_load_2d_2sm(dst + 4096, bm, full_leader + 8 * st, 0, row0 + 32);
                }
              }
              __syncwarp();
            });
            par ^= kColFlip;
          };
          if constexpr (KIND == TK_PAIR) {
#pragma unroll 1
            for (int c = 0; c < 3; ++c) produce_column(std::integral_constant<int, 0>{}, CL::units_before_col(c) * 64);
#pragma unroll 1
            for (int c = 3; c < 5; ++c) produce_column(std::integral_constant<int, 1>{}, CL::units_before_col(c) * 64);
          } else {
#pragma unroll 1
            for (int c = 0; c < 5; ++c) produce_column(std::integral_constant<int, 0>{}, 5 * c * tap_rows);
          }
        } else {
        auto produce_slab = [&](auto PH) {
        constexpr int ph = decltype(PH)::value;
        static_for<NU>([&](auto U) {
          constexpr int u = decltype(U)::value;
          constexpr int t = SPLIT ? u / 2 : u, plane = SPLIT ? (u & 1) : 0, npl = SPLIT ? 2 : 1;
          constexpr bool outer = TP::outer(t);
          constexpr int ru = ph * NU + u;          // use index inside the period
          constexpr int st = ru % NST;
          mbar_wait(bar_b_empty + 8 * st, slab_par ^ (uint32_t)((ru / NST) & 1) ^ 1u);
          if (elect_one()) {
            // this CTA's half of the block: rows [rank * rows/2, (rank + 1) * rows/2) in 32-row (4 KB) boxes
            const int rows = KIND == TK_PAIR ? (outer ? 64 : 128) : tap_rows;
            const int row0 = slab_row0 + npl * (KIND == TK_PAIR ? TP::units_before(t) * 64 : t * tap_rows) + plane * rows +
                             (int)rank * (rows >> 1);
            if (TC2_DBG(p, 2)) { if (leader) mbar_arrive(bar_b_full + 8 * st); }
            else {
              if (leader) mbar_expect_tx(bar_b_full + 8 * st, (uint32_t)rows << 7);
              const uint32_t dst = s_b + (uint32_t)st * kStageBytes;
              tma_load_2d_2sm(dst, bm, full_leader + 8 * st, 0, row0);
              if (rows > 64) tma_load_2d_2sm(dst + 4096, bm, full_leader + 8 * st, 0, row0 + 32);
            }
          }
          __syncwarp();
        });
        };
        if (kPeriod == 1 || phase == 0) produce_slab(std::integral_constant<int, 0>{});
        else produce_slab(std::integral_constant<int, kPeriod - 1>{});
        if (++phase == kPeriod) { phase = 0; slab_par ^= (uint32_t)(kUsesPerSlab & 1); }
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ================================ MMA issuer (leader CTA only) ===============================
    if (leader) {
      int ps = 0;
      uint32_t pph = 0, slab_par = 0;
      uint32_t par = 0;                          // column-rolled loops: parity of the ring revolution at the column's start
      int phase = 0;
      int it = 0;
      const uint64_t desc_a = umma_desc_hi(kPitch), desc_b = umma_desc_hi(1024);
      const uint64_t b_base = desc_b | desc_addr(s_b);
      // fused 1x1: uses (accumulators) of the previous tile that are still to be multiplied
      int prev_left = 0, prev_j = 0, prev_job = 0;
      uint32_t prev_d = 0, y_uses = 0;
      bool wc_ready = false;
      auto service = [&](bool block) {
        if (SPLIT) {
          // split mode: hand-off prev_j covers the 64-channel slab (prev_j & 1) of the tile's accumulator
          if (prev_left > 0 && (block || mbar_test(bar_y_full, y_uses & 1u))) {
            const int n = fuse_service_split(prev_left, prev_j, prev_job, prev_d, y_uses, wc_ready, bar_y_full, bar_y_done,
                                             bar_wc_full, s_y, s_wc, p.idesc_1x1, block);
            prev_left -= n; prev_j += n; y_uses += (uint32_t)n;
            wc_ready = true;
          }
          return;
        }
        // issues the 1x1 MMAs of the next pending hand-off once its Y tile is complete in both CTAs.  A hand-off is
        // one accumulator (prev_j), or in split mode one 64-channel slab of it (accumulator prev_j >> 1, slab prev_j & 1:
        // Y holds the hi and lo planes of that slab, and hi*hi + lo*hi + hi*lo accumulate into the 64 result columns).
        while (prev_left > 0) {
          if (block) mbar_wait(bar_y_full, y_uses & 1u);
          else if (!mbar_test(bar_y_full, y_uses & 1u)) return;
          if (!wc_ready) { mbar_wait(bar_wc_full, 0); wc_ready = true; }
          tc_fence_after();
          if (elect_one()) {
            {
              const uint32_t d = prev_d + (uint32_t)(prev_j * 128);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const uint64_t ad = desc_b | desc_addr(s_y + (uint32_t)(k >> 2) * 16384u + (uint32_t)(k & 3) * 32u);
                const uint64_t bd = desc_b | desc_addr(s_wc + (uint32_t)(prev_job * 2 + (k >> 2)) * 4096u + (uint32_t)(k & 3) * 32u);
                umma_f16_2sm(d, ad, bd, p.idesc_1x1, k ? 1u : 0u);
              }
            }
            umma_commit_2sm(bar_y_done);
          }
          __syncwarp();
          ++y_uses; ++prev_j; --prev_left;
        }
      };
      const uint32_t idesc_full = p.idesc_full, idesc_half = p.idesc_half;
      const uint32_t n_cols = (uint32_t)p.n_cols;
#ifdef CODON_TC_EXPERIMENT
      // CODON_TC_DEBUG bit 64: cycle accounting of this warp (cluster 0 prints it when the kernel ends)
      const bool prof = (p.debug & 64) != 0;
      long long c_wait = 0, c_acc = 0, c_blk = 0, c_all = prof ? clock64() : 0, c_t = 0;
      unsigned long long g_all = 0;
      if (prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_all));
      int n_taps = 0;
#endif
      // waits until the epilogue has released TMEM buffer `b` (use count -> parity)
      auto acquire_acc = [&](int b, uint32_t par) {
#ifdef CODON_TC_EXPERIMENT
        const long long t_in = prof ? clock64() : 0;
#endif
        if (FUSE) {
          // the epilogue of the tile that last used this TMEM buffer needs our 1x1 MMAs to finish: keep serving
          while (!mbar_test(bar_acc_empty + 8 * b, par)) service(false);
        } else {
          mbar_wait(bar_acc_empty + 8 * b, par);
        }
#ifdef CODON_TC_EXPERIMENT
        if (prof) c_acc += clock64() - t_in;
#endif
        tc_fence_after();
      };
      int sbuf = 0;                 // split mode: next big buffer, its use parity, tiles issued so far
      uint32_t sbuf_par = 0, tile_it = 0;
      for (int item = cluster_id; item < total_items; item += nclusters) {
        struct { int job, nacc; } tl;
        decode_job_nacc<NACC>(p, item, tl.job, tl.nacc);
        const uint32_t outer_col = (uint32_t)p.job[tl.job].outer_col;
        int buf = it % p.nbuf;
        uint32_t d_base = 0;
#ifdef CODON_TC_EXPERIMENT
        if (prof) c_t = clock64();
#endif
        if (!SPLIT) {
          acquire_acc(buf, ((uint32_t)(it / p.nbuf) & 1u) ^ 1u);
          d_base = tmem_base + (uint32_t)buf * (uint32_t)NACC * n_cols;
        }
        const int nacc_rt = tl.nacc;
        uint32_t acc0 = 0;                         // 0 only for the first K step of the tile
        for (int s = 0; s < p.nslab; ++s) {
          mbar_wait(bar_patch_full + 8 * ps, pph);
          const uint64_t a_base = desc_a | desc_addr(s_patch + ps * p.patch_stage);
          const int ps_hi = ps;
          uint64_t a_base_lo = 0;
          if (SPLIT) {                             // the lo-plane patch of the slab sits in the next ring stage
            if (++ps == p.npb) { ps = 0; pph ^= 1; }
            mbar_wait(bar_patch_full + 8 * ps, pph);
            a_base_lo = desc_a | desc_addr(s_patch + ps * p.patch_stage);
          }
          const bool last_slab = s == p.nslab - 1;
          if constexpr (kColRoll) {
            // Column-rolled issue loop (see Cols): a run-time loop over the five tap columns (dx), straight-line code per
            // tap inside a column -- dy offsets, ring stages, barrier addresses and the split mode's chunk boundaries are
            // immediates; dx is one descriptor add per column and `par` the parity of the ring revolution the column
            // starts in.  The phase test of the NEXT block's first stage is issued before this block's MMAs.
            bool ready = mbar_test(bar_b_full, par);
            auto issue_column = [&](auto CT, const int dxv, const bool first_col, const bool last_col) {
              constexpr int ct = decltype(CT)::value;
              const uint64_t a_col = a_base + (uint64_t)((uint32_t)dxv * 8u);      // + dx pixels of 128 B, in 16-byte units
              const uint32_t acc_first = first_col ? acc0 : 1u;                    // 0 only for the tile's very first K step
              if constexpr (SPLIT) {
                // ONE issue block per tap covers both of its ring uses (hi block: A_hi -> big and A_lo -> small,
                // interleaved K step by K step; lo block: A_hi -> small): 12 MMAs per elect / fence / branch
                const uint64_t a_col_lo = a_base_lo + (uint64_t)((uint32_t)dxv * 8u);
                static_for<5>([&](auto I) {
                  constexpr int i = decltype(I)::value;
                  constexpr int v0 = 2 * i, v1 = 2 * i + 1, vn = v1 + 1;
                  constexpr int st0 = v0 % NST, st1 = v1 % NST;
                  constexpr uint32_t pf0 = (uint32_t)((v0 / NST) & 1), pf1 = (uint32_t)((v1 / NST) & 1);
                  constexpr bool outer = CL::outer(ct, i);
                  constexpr uint32_t dy_off = ((uint32_t)TP::ord(i) * kPitch) >> 4;
                  constexpr bool chunk_start = (i % kTPC) == 0, chunk_end = (i % kTPC) == kTPC - 1;
                  if (chunk_start) {
                    buf = sbuf;
                    acquire_acc(buf, sbuf_par ^ 1u);
                    d_base = tmem_base + (uint32_t)buf * n_cols;
                    if (++sbuf == kSplitBufs) { sbuf = 0; sbuf_par ^= 1u; }
                    if (i == 0 && first_col && s == 0) {
                      // the tile's first small MMA overwrites the small accumulator: the previous tile's must have been read
                      if (FUSE) { while (!mbar_test(bar_small_empty, (tile_it & 1u) ^ 1u)) service(false); }
                      else mbar_wait(bar_small_empty, (tile_it & 1u) ^ 1u);
                      tc_fence_after();
                      ++tile_it;
                    }
                  }
                  if (FUSE) service(false);
#ifdef CODON_TC_EXPERIMENT
                  if (prof) c_t = clock64();
#endif
                  if (!ready) mbar_wait(bar_b_full + 8 * st0, par ^ pf0);
                  mbar_wait(bar_b_full + 8 * st1, par ^ pf1);
#ifdef CODON_TC_EXPERIMENT
                  if (prof) { c_wait += clock64() - c_t; n_taps += 2; }
#endif
                  if constexpr (vn < CU) ready = mbar_test(bar_b_full + 8 * (vn % NST), par ^ (uint32_t)((vn / NST) & 1));
                  else ready = mbar_test(bar_b_full, par ^ kColFlip);
                  tc_fence_after();
                  const uint64_t bd_hi = b_base + (uint64_t)(((uint32_t)st0 * kStageBytes) >> 4);
                  const uint64_t bd_lo = b_base + (uint64_t)(((uint32_t)st1 * kStageBytes) >> 4);
                  const uint32_t idesc = outer ? idesc_half : idesc_full;
                  const uint32_t d_big = d_base + (outer ? outer_col : 0u);
                  const uint32_t d_small = tmem_base + (uint32_t)kSplitBufs * n_cols + (outer ? outer_col : 0u);
                  if (elect_one()) {
                    const uint64_t a_hi = a_col + (uint64_t)dy_off, a_lo = a_col_lo + (uint64_t)dy_off;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      umma_f16_2sm(d_big, a_hi + 2 * k, bd_hi + 2 * k, idesc, (chunk_start && k == 0) ? 0u : 1u);
                      umma_f16_2sm(d_small, a_lo + 2 * k, bd_hi + 2 * k, idesc, (i == 0 && k == 0) ? acc_first : 1u);
                    }
                    umma_commit_2sm(bar_b_empty + 8 * st0);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_2sm(d_small, a_hi + 2 * k, bd_lo + 2 * k, idesc, 1u);
                    umma_commit_2sm(bar_b_empty + 8 * st1);
                    if (i == 4 && last_col) {
                      umma_commit_2sm(bar_patch_empty + 8 * ps_hi);
                      umma_commit_2sm(bar_patch_empty + 8 * ps);
                    }
                    if (chunk_end) umma_commit_2sm(bar_acc_full + 8 * buf);
                  }
                  __syncwarp();
                });
              } else {
                // kColGroup taps (ring stages) per issue block
                constexpr int G = kColGroup;
                static_for<(5 + G - 1) / G>([&](auto BK) {
                  constexpr int i0 = decltype(BK)::value * G, cnt = (5 - i0) < G ? (5 - i0) : G, in = i0 + cnt;
                  if (FUSE) service(false);
#ifdef CODON_TC_EXPERIMENT
                  if (prof) c_t = clock64();
#endif
                  if (!ready) mbar_wait(bar_b_full + 8 * i0, par);
                  static_for<cnt - 1>([&](auto Q) { mbar_wait(bar_b_full + 8 * (i0 + 1 + decltype(Q)::value), par); });
#ifdef CODON_TC_EXPERIMENT
                  if (prof) { c_wait += clock64() - c_t; n_taps += cnt; }
#endif
                  if constexpr (in < 5) ready = mbar_test(bar_b_full + 8 * in, par);
                  else ready = mbar_test(bar_b_full, par ^ kColFlip);
                  tc_fence_after();
                  if (elect_one()) {
                    static_for<cnt>([&](auto Q) {
                      constexpr int i = i0 + decltype(Q)::value;          // tap of the column == ring stage
                      constexpr bool outer = CL::outer(ct, i);
                      constexpr uint32_t dy_off = ((uint32_t)TP::ord(i) * kPitch) >> 4;
                      const uint64_t bdesc = b_base + (uint64_t)(((uint32_t)i * kStageBytes) >> 4);
                      const uint32_t idesc = outer ? idesc_half : idesc_full;
                      const uint32_t d0 = d_base + (outer ? outer_col : 0u);
#pragma unroll
                      for (int j = 0; j < NACC; ++j) {
                        if (j < nacc_rt && !TC2_DBG(p, 8)) {
                          // sub-tile j = (jx, jy): + jy*16 patch rows + jx*8 pixels
                          constexpr uint32_t kSubY = ((uint32_t)kTcSubH * kPitch) >> 4, kSubX = ((uint32_t)kTcSubW * 128u) >> 4;
                          const uint64_t adesc = a_col + (uint64_t)(dy_off + (uint32_t)(j / G_::NAX) * kSubY + (uint32_t)(j % G_::NAX) * kSubX);
                          const uint32_t d = d0 + (uint32_t)j * n_cols;
#pragma unroll
                          for (int k = 0; k < 4; ++k) {
                            // +32 B per K step inside the 128-B swizzled row == +2 in the 16-B address field
                            const uint32_t acc = (i == 0 && k == 0) ? acc_first : 1u;
                            if (OPERAND == TC_TF32) umma_tf32_2sm(d, adesc + 2 * k, bdesc + 2 * k, idesc, acc);
                            else                    umma_f16_2sm(d, adesc + 2 * k, bdesc + 2 * k, idesc, acc);
                          }
                        }
                      }
                      umma_commit_2sm(bar_b_empty + 8 * i);
                      if (i == 4 && last_col) {
                        umma_commit_2sm(bar_patch_empty + 8 * ps_hi);
                        if (last_slab) umma_commit_2sm(bar_acc_full + 8 * buf);
                      }
                    });
                  }
                  __syncwarp();
                });
              }
              par ^= kColFlip;
            };
            if constexpr (KIND == TK_PAIR) {
#pragma unroll 1
              for (int c = 0; c < 3; ++c) issue_column(std::integral_constant<int, 0>{}, c == 0 ? 2 : (c == 1 ? 1 : 3), c == 0, false);
#pragma unroll 1
              for (int c = 3; c < 5; ++c) issue_column(std::integral_constant<int, 1>{}, c == 3 ? 0 : 4, false, c == 4);
            } else {
#pragma unroll 1
              for (int c = 0; c < 5; ++c) issue_column(std::integral_constant<int, 0>{}, c, c == 0, c == 4);
            }
          } else {
          const uint32_t par_even = slab_par, par_odd = slab_par ^ 1u;
          // Straight-line code per tap: the A start address is the patch shifted by (dy rows, dx pixels), the ring
          // stage and its barriers are immediates.  The phase test of the NEXT tap's weights is issued before this
          // tap's MMAs so that its latency hides behind their issue.
          static_assert(kPeriod == 1 || !SPLIT, "two-phase ring: generic tap loop only");
          bool ready = (kPeriod == 1 || phase == 0) ? mbar_test(bar_b_full, par_even)
                                                    : mbar_test(bar_b_full + 8 * (((kPeriod - 1) * NU) % NST),
                                                                ((((kPeriod - 1) * NU) / NST) & 1) ? par_odd : par_even);
          // Split mode: ONE issue block per tap covers both of its ring uses (hi block: A_hi -> big and A_lo -> small,
          // interleaved K step by K step; lo block: A_hi -> small) -- 12 MMAs per elect / fence / branch instead of 8 + 4.
          // The issue loop's bookkeeping is what limits the tensor pipe here, so fewer, longer blocks win.
          if constexpr (SPLIT) {
            static_for<TP::NT>([&](auto T) {
              constexpr int t = decltype(T)::value;
              constexpr int u0 = 2 * t, u1 = 2 * t + 1;
              constexpr bool outer = TP::outer(t);
              constexpr int st0 = u0 % NST, st1 = u1 % NST;
              constexpr uint32_t tap_off = ((uint32_t)TP::dy(t) * kPitch + (uint32_t)TP::dx(t) * 128u) >> 4;
              constexpr bool chunk_start = (t % kTPC) == 0, chunk_end = (t % kTPC) == kTPC - 1;
              if (chunk_start) {
                buf = sbuf;
                acquire_acc(buf, sbuf_par ^ 1u);
                d_base = tmem_base + (uint32_t)buf * n_cols;
                if (++sbuf == kSplitBufs) { sbuf = 0; sbuf_par ^= 1u; }
                if (t == 0 && s == 0) {
                  // the tile's first small MMA overwrites the small accumulator: the previous tile's must have been read
                  if (FUSE) { while (!mbar_test(bar_small_empty, (tile_it & 1u) ^ 1u)) service(false); }
                  else mbar_wait(bar_small_empty, (tile_it & 1u) ^ 1u);
                  tc_fence_after();
                  ++tile_it;
                }
              }
              if (FUSE) service(false);
#ifdef CODON_TC_EXPERIMENT
              if (prof) c_t = clock64();
#endif
              if (!ready) mbar_wait(bar_b_full + 8 * st0, ((u0 / NST) & 1) ? par_odd : par_even);
              mbar_wait(bar_b_full + 8 * st1, ((u1 / NST) & 1) ? par_odd : par_even);
#ifdef CODON_TC_EXPERIMENT
              if (prof) { c_wait += clock64() - c_t; n_taps += 2; }
#endif
              if (t + 1 < TP::NT) ready = mbar_test(bar_b_full + 8 * ((u1 + 1) % NST), (((u1 + 1) / NST) & 1) ? par_odd : par_even);
              tc_fence_after();
              const uint64_t bd_hi = b_base + (uint64_t)(((uint32_t)st0 * kStageBytes) >> 4);
              const uint64_t bd_lo = b_base + (uint64_t)(((uint32_t)st1 * kStageBytes) >> 4);
              const uint32_t idesc = outer ? idesc_half : idesc_full;
              const uint32_t d_big = d_base + (outer ? outer_col : 0u);
              const uint32_t d_small = tmem_base + (uint32_t)kSplitBufs * n_cols + (outer ? outer_col : 0u);
              if (elect_one()) {
                const uint64_t a_hi = a_base + (uint64_t)tap_off, a_lo = a_base_lo + (uint64_t)tap_off;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_f16_2sm(d_big, a_hi + 2 * k, bd_hi + 2 * k, idesc, (chunk_start && k == 0) ? 0u : 1u);
                  umma_f16_2sm(d_small, a_lo + 2 * k, bd_hi + 2 * k, idesc, (t == 0 && k == 0) ? acc0 : 1u);
                }
                umma_commit_2sm(bar_b_empty + 8 * st0);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_2sm(d_small, a_hi + 2 * k, bd_lo + 2 * k, idesc, 1u);
                umma_commit_2sm(bar_b_empty + 8 * st1);
                if (t == TP::NT - 1) {
                  umma_commit_2sm(bar_patch_empty + 8 * ps_hi);
                  umma_commit_2sm(bar_patch_empty + 8 * ps);
                }
                if (chunk_end) umma_commit_2sm(bar_acc_full + 8 * buf);
              }
              __syncwarp();
            });
          } else {
          auto issue_slab = [&](auto PH) {
          constexpr int ph = decltype(PH)::value;
          static_for<TP::NT>([&](auto T) {
            constexpr int t = decltype(T)::value;
            constexpr bool outer = TP::outer(t);
            constexpr int ru = ph * TP::NT + t;      // use index inside the ring period
            constexpr int st = ru % NST;
            constexpr uint32_t tap_off = ((uint32_t)TP::dy(t) * kPitch + (uint32_t)TP::dx(t) * 128u) >> 4;
            if (FUSE) service(false);
#ifdef CODON_TC_EXPERIMENT
            if (prof) c_t = clock64();
#endif
            if (!ready) mbar_wait(bar_b_full + 8 * st, ((ru / NST) & 1) ? par_odd : par_even);
#ifdef CODON_TC_EXPERIMENT
            if (prof) { c_wait += clock64() - c_t; ++n_taps; }
#endif
            if (t + 1 < TP::NT) ready = mbar_test(bar_b_full + 8 * ((ru + 1) % NST), (((ru + 1) / NST) & 1) ? par_odd : par_even);
            tc_fence_after();
            const uint64_t bdesc = b_base + (uint64_t)(((uint32_t)st * kStageBytes) >> 4);
            const uint32_t idesc = outer ? idesc_half : idesc_full;
            const uint32_t d0 = d_base + (outer ? outer_col : 0u);
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < NACC; ++j) {
                if (j < nacc_rt && !TC2_DBG(p, 8)) {
                  // sub-tile j = (jx, jy): + jy*16 patch rows + jx*8 pixels
                  const uint32_t sub_off = ((uint32_t)(j / G::NAX) * kTcSubH * kPitch + (uint32_t)(j % G::NAX) * kTcSubW * 128u) >> 4;
                  const uint64_t adesc = a_base + (uint64_t)(tap_off + sub_off);
                  const uint32_t d = d0 + (uint32_t)j * n_cols;
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    // +32 B per K step inside the 128-B swizzled row == +2 in the 16-B address field
                    const uint32_t acc = (t == 0 && k == 0) ? acc0 : 1u;
                    if (OPERAND == TC_TF32) umma_tf32_2sm(d, adesc + 2 * k, bdesc + 2 * k, idesc, acc);
                    else                    umma_f16_2sm(d, adesc + 2 * k, bdesc + 2 * k, idesc, acc);
                  }
                }
              }
              umma_commit_2sm(bar_b_empty + 8 * st);
              if (t == TP::NT - 1) {
                umma_commit_2sm(bar_patch_empty + 8 * ps_hi);
                if (last_slab) umma_commit_2sm(bar_acc_full + 8 * buf);
              }
            }
            __syncwarp();
          });
          };
          if (kPeriod == 1 || phase == 0) issue_slab(std::integral_constant<int, 0>{});
          else issue_slab(std::integral_constant<int, kPeriod - 1>{});
          }
          }
          acc0 = 1;
          if (++phase == kPeriod) { phase = 0; slab_par ^= (uint32_t)(kUsesPerSlab & 1); }
          if (++ps == p.npb) { ps = 0; pph ^= 1; }
        }
        if (FUSE) {
#ifdef CODON_TC_EXPERIMENT
          if (prof) c_t = clock64();
#endif
          service(true);                       // anything still pending belongs to the tile before this one
#ifdef CODON_TC_EXPERIMENT
          if (prof) c_blk += clock64() - c_t;
#endif
          prev_left = SPLIT ? 2 : tl.nacc; prev_j = 0; prev_job = tl.job; prev_d = d_base;
        }
        if (!SPLIT) ++it;
      }
      if (FUSE) service(true);
#ifdef CODON_TC_EXPERIMENT
      if (prof && (cluster_id % 24) == 0 && lane == 0) {
        unsigned long long g_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_end));
        const double cyc = (double)(clock64() - c_all);
        printf("conv_tc2<%d,%d,%d,%d> cluster %d issuer: %d taps, cycles/tap: total %.0f, of which b_full wait %.0f, acc_empty wait %.0f, end-of-tile service %.0f; %.3f ms at %.0f MHz\n",
               NACC, OPERAND, (int)FUSE, KIND, cluster_id, n_taps, cyc / n_taps, (double)c_wait / n_taps,
               (double)c_acc / n_taps, (double)c_blk / n_taps, (double)(g_end - g_all) * 1e-6, cyc / (double)(g_end - g_all) * 1e3);
      }
#endif
    }
  } else {
    // ================================ epilogue (both CTAs, own tile) ==============================
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int ehalf = warp >> 2;                 // 0 / 1: which half of the chunks this warp drains
    const int m = q * 32 + lane;
    const int my = m / kTcSubW, mx = m % kTcSubW;
    const uint32_t acc_empty_leader = mapa_u32(bar_acc_empty, 0);
    const uint32_t y_full_leader = mapa_u32(bar_y_full, 0);
    const uint32_t small_empty_leader = mapa_u32(bar_small_empty, 0);
    uint32_t y_uses = 0;                         // fused mode: Y hand-offs so far (phase of y_full / y_done)
    int ep_sbuf = 0; uint32_t ep_sbuf_par = 0;   // split mode: next big buffer to promote and its use parity
    int it = 0;
    for (int item = cluster_id; item < total_items; item += nclusters, ++it) {
      const Tile2 tl = decode_tile2<NACC>(p, item, (int)rank);
      const TcJob& job = p.job[tl.job];
      int buf = it % p.nbuf;
      // split mode: this warp's 32-column groups (columns g * 64 + ehalf * 32 ...), promoted chunk by chunk into
      // round-to-nearest fp32 running sums
      constexpr int NG = SPLIT ? (KIND == TK_3X3 ? 1 : 2) : 1;
      float tot[NG][32];
      if (SPLIT) {
        constexpr int NCH = TP::NT / kTPC;                       // chunks per slab
        const uint32_t n_cols = (uint32_t)p.n_cols;
        const int g_outer = job.outer_col >> 6;                  // pair plans: the group the 5x5-only taps feed
        int& sbuf = ep_sbuf; uint32_t& sbuf_par = ep_sbuf_par;
        for (int s = 0; s < p.nslab; ++s) {
#pragma unroll 1
          for (int c = 0; c < NCH; ++c) {
            // pair plans: a chunk made of border taps only carries the 5x5 half of the columns (a chunk starts with
            // a border tap only if all of its taps are border taps: the issue order is centre-first)
            const bool outer_only = KIND == TK_PAIR && ((TP::outer_chunk_mask(kTPC) >> c) & 1u);
            buf = sbuf;
            mbar_wait(bar_acc_full + 8 * buf, sbuf_par);
            if (++sbuf == kSplitBufs) { sbuf = 0; sbuf_par ^= 1u; }
            tc_fence_after();
            const uint32_t lq = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)ehalf * 32u;
            const uint32_t cb = lq + (uint32_t)buf * n_cols;
            const bool last_chunk = (s == p.nslab - 1) && (c == NCH - 1);
            uint32_t ra[NG][32];
#pragma unroll
            for (int g = 0; g < NG; ++g)
              if ((!outer_only || g == g_outer) && !TC2_DBG(p, 32)) tmem_ld32(cb + (uint32_t)g * 64u, ra[g]);      // warp-uniform predicate
            tmem_ld_wait();
            // the chunk is in registers: hand the buffer back before the additions (the fused kernel keeps the last
            // one of a tile: the 1x1 result is computed into its first 64 columns)
            if (!(FUSE && last_chunk)) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster(acc_empty_leader + 8 * buf);
            }
#pragma unroll
            for (int g = 0; g < NG; ++g) {
              if ((!outer_only || g == g_outer) && !TC2_DBG(p, 32)) {
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                  const float v = __uint_as_float(ra[g][e]);
                  tot[g][e] = (c == 0 && s == 0) ? v : tot[g][e] + v;
                }
              }
            }
            if (last_chunk) {
              // the tile's small accumulator (complete: the chunk's commit covers every earlier MMA)
#pragma unroll
              for (int g = 0; g < NG; ++g) tmem_ld32(lq + (uint32_t)kSplitBufs * n_cols + (uint32_t)g * 64u, ra[g]);
              tmem_ld_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster(small_empty_leader);
#pragma unroll
              for (int g = 0; g < NG; ++g)
#pragma unroll
                for (int e = 0; e < 32; ++e) tot[g][e] += __uint_as_float(ra[g][e]);
            }
          }
        }
      } else {
        mbar_wait(bar_acc_full + 8 * buf, (uint32_t)(it / p.nbuf) & 1u);
        tc_fence_after();
      }
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (SPLIT ? (uint32_t)buf * (uint32_t)p.n_cols : (uint32_t)(buf * NACC * p.n_cols));
      if (FUSE) {
        // accumulator j -> ReLU -> 16-bit -> Y (this warp: the 64 channels of slab `ehalf`), then the 1x1 result
        // (64 columns at the start of the same accumulator; this warp: 32 of them) -> (+ res2) -> out2
        auto stage_y = [&](int j) {
          uint32_t ra[32], rb[32];
          if (!TC2_DBG(p, 32)) {
          tmem_ld32(lane_base + (uint32_t)(j * 128 + ehalf * 64), ra);
          tmem_ld32(lane_base + (uint32_t)(j * 128 + ehalf * 64 + 32), rb);
          tmem_ld_wait();
          }
          const uint32_t row = s_y + (uint32_t)ehalf * 16384u + (uint32_t)m * 128u;
#pragma unroll
          for (int pc = 0; pc < 8; ++pc) {
            if (TC2_DBG(p, 32)) break;
            const uint32_t* r = pc < 4 ? ra : rb;
            const int e0 = (pc & 3) * 8;
            uint32_t w[4];
#pragma unroll
            for (int t2 = 0; t2 < 4; ++t2)
              w[t2] = pack16<Y16>(fmaxf(__uint_as_float(r[e0 + 2 * t2]), 0.f), fmaxf(__uint_as_float(r[e0 + 2 * t2 + 1]), 0.f));
            st_shared_v4(row + (uint32_t)((pc ^ (m & 7)) << 4), w[0], w[1], w[2], w[3]);
          }
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive_cluster_release(y_full_leader);
        };
        // split mode: one 64-channel slab of accumulator j per hand-off; this warp converts 32 of its channels
        // (columns slab * 64 + ehalf * 32 ...) of its 32 pixels: x = relu(acc) / scale -> hi = fp16(x), lo = fp16(x - hi)
        // -> rows m of the hi tile (s_y) and of the lo tile (s_y + 16 KB), K-major SWIZZLE_128B
        auto stage_y_split = [&](const float (&acc)[32]) {
          const float ds = job.descale;
          const uint32_t row = s_y + (uint32_t)m * 128u;
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int t2 = 0; t2 < 4; ++t2)
              split_pair(fmaxf(acc[pc * 8 + 2 * t2], 0.f) * ds, fmaxf(acc[pc * 8 + 2 * t2 + 1], 0.f) * ds, h[t2], l[t2]);
            const uint32_t off = (uint32_t)(((ehalf * 4 + pc) ^ (m & 7)) << 4);
            st_shared_v4(row + off, h[0], h[1], h[2], h[3]);
            st_shared_v4(row + 16384u + off, l[0], l[1], l[2], l[3]);
          }
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive_cluster_release(y_full_leader);
        };
        auto drain_d2 = [&](int j) {
          uint32_t r[32];
          if (TC2_DBG(p, 32)) return;
          tmem_ld32(lane_base + (uint32_t)(j * 128 + ehalf * 32), r);
          tmem_ld_wait();
          if (SPLIT) {
            const float ds2 = job.descale2;
#pragma unroll
            for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * ds2);
          }
          const int py = tl.y0 + (j / G::NAX) * kTcSubH + my, px = tl.x0 + (j % G::NAX) * kTcSubW + mx;
          if (tl.valid && (py < p.H) && (px < p.W) && !TC2_DBG(p, 1)) {
            const size_t pix = ((size_t)tl.n * p.H + py) * p.W + px;
            TcJob o2 = job;
            o2.out = job.out2; o2.out_stride = job.out2_stride; o2.out_off = job.out2_off;
            o2.res = job.res2; o2.res_stride = job.res2_stride; o2.res_off = job.res2_off;
            store_chunk<OutT>(r, o2, pix, ehalf * 32, false, OPERAND == TC_TF32);
            if (job.pool) {
              // ChannelPool partial (max, sum) over this thread's 32 of the job's 64 channels: pool[half][pixel]
              float mxv = __uint_as_float(r[0]), sacc = 0.f;
#pragma unroll
              for (int e = 0; e < 32; ++e) { const float v = __uint_as_float(r[e]); mxv = fmaxf(mxv, v); sacc += v; }
              job.pool[(size_t)ehalf * (size_t)p.pool_stride + pix] = make_float2(mxv, sacc);
            }
          }
          if (job.cstat && tl.valid && tl.y0 + (j / G::NAX) * kTcSubH < p.H && tl.x0 + (j % G::NAX) * kTcSubW < p.W) {
            // Per-channel (sum, max) over this warp's 32 pixels.  Sum: a fixed-order butterfly in which the lanes that
            // differ in bit k exchange the half of the channels they do not keep (31 shuffles; lane l ends with channel
            // l).  Max: one warp-collective redux.sync.max.f32 per channel (CREDUX, uniform datapath) -- shuffles go
            // through the shared-memory crossbar, which the MMA operand reads already keep ~80 % busy.
            const bool live = (py < p.core_y1) && (py >= p.core_y0) && (px < p.W);   // core_y1 <= H
            float sv[32];
            float mv0 = -INFINITY;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const float v = __uint_as_float(r[e]);
              sv[e] = live ? v : 0.f;
              float red;
              asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(red) : "f"(live ? v : -INFINITY));
              if (lane == e) mv0 = red;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
              const bool hi = (lane & off) != 0;
#pragma unroll
              for (int e = 0; e < off; ++e) {
                const float ks = hi ? sv[off + e] : sv[e], gs = hi ? sv[e] : sv[off + e];
                sv[e] = ks + __shfl_xor_sync(0xffffffffu, gs, off);
              }
            }
            float mv[1] = {mv0};
            const int cy = (tl.y0 + (j / G::NAX) * kTcSubH) / kTcSubH, cx = (tl.x0 + (j % G::NAX) * kTcSubW) / kTcSubW;
            const size_t cell = ((size_t)tl.n * p.cells_y + cy) * p.cells_x + cx;
            job.cstat[(cell * 4 + q) * 64 + ehalf * 32 + lane] = make_float2(sv[0], mv[0]);
          }
        };
        // Y is free here: the previous hand-off's y_done was awaited before its drain
        if (SPLIT) {
          // two hand-offs: the promoted sums of channel slab 0, then (once the 1x1 MMAs of slab 0 have read Y) slab 1
          stage_y_split(tot[0]);
          mbar_wait(bar_y_done, y_uses & 1u);
          ++y_uses;
          tc_fence_after();
          stage_y_split(tot[NG - 1]);
        } else {
        stage_y(0);
        for (int j = 1; j < tl.nacc; ++j) {
          mbar_wait(bar_y_done, y_uses & 1u);    // 1x1 of accumulator j-1 finished: Y is free, its result is in TMEM
          ++y_uses;
          tc_fence_after();
          stage_y(j);
          drain_d2(j - 1);
        }
        }
        mbar_wait(bar_y_done, y_uses & 1u);
        ++y_uses;
        tc_fence_after();
        drain_d2(tl.nacc - 1);
      } else if (SPLIT) {
        const int py = tl.y0 + my, px = tl.x0 + mx;
        if (tl.valid && (py < p.H) && (px < p.W)) {
          const size_t pix = ((size_t)tl.n * p.H + py) * p.W + px;
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            uint32_t r[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(tot[g][e]);
            store_chunk<OutT>(r, job, pix, g * 64 + ehalf * 32, p.relu != 0, false, job.descale);
          }
        }
      } else {
        const int cpa_sh = p.n_cols == 128 ? 2 : 1, nchunk = tl.nacc << cpa_sh;    // chunks per accumulator: 4 or 2
        auto issue = [&](int i, uint32_t (&r)[32]) { tmem_ld32(lane_base + (uint32_t)(i << 5), r); };
        auto drain = [&](int i, const uint32_t (&r)[32]) {
          const int j = i >> cpa_sh, c0 = (i - (j << cpa_sh)) << 5;
          const int py = tl.y0 + (j / G::NAX) * kTcSubH + my, px = tl.x0 + (j % G::NAX) * kTcSubW + mx;
          if (tl.valid && (py < p.H) && (px < p.W) && !TC2_DBG(p, 1)) {
            const size_t pix = ((size_t)tl.n * p.H + py) * p.W + px;
            store_chunk<OutT>(r, job, pix, c0, p.relu != 0, OPERAND == TC_TF32, job.descale);
          }
        };
        // the two warps of a lane quarter take the even / odd chunks
        uint32_t ra[32], rb[32];
        int i = ehalf;
        if (!TC2_DBG(p, 32)) {
        issue(i, ra);
#pragma unroll 1
        while (true) {
          tmem_ld_wait();
          if (i + 2 < nchunk) issue(i + 2, rb);
          drain(i, ra);
          if (i + 2 >= nchunk) break;
          tmem_ld_wait();
          if (i + 4 < nchunk) issue(i + 4, ra);
          drain(i + 2, rb);
          if (i + 4 >= nchunk) break;
          i += 4;
        }
        }
      }
      if (SPLIT) {
        if (FUSE) {                // (split, not fused: every chunk buffer was handed back right after its promotion)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc_empty_leader + 8 * buf);
        }
      } else {
        tc_fence_before();
        mbar_arrive_cluster(acc_empty_leader + 8 * buf);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // nobody frees TMEM or exits while the pair may still touch it
  if (warp == kWarpMma) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

__device__ __noinline__ int fuse_service_split(int left, int j, int job, uint32_t d, uint32_t y_uses, bool wc_ready,
                                               uint32_t bar_y_full, uint32_t bar_y_done, uint32_t bar_wc_full,
                                               uint32_t s_y, uint32_t s_wc, uint32_t idesc_1x1, bool block) {
  // Hand-off j covers the 64-channel slab (j & 1) of the tile's accumulator: Y holds that slab's hi plane (s_y) and
  // lo plane (s_y + 16 KB); hi*hi + lo*hi + hi*lo accumulate into the 64 result columns at the start of the buffer.
  const uint64_t desc_b = umma_desc_hi(1024);
  int n = 0;
  while (left > 0) {
    if (block) mbar_wait(bar_y_full, y_uses & 1u);
    else if (!mbar_test(bar_y_full, y_uses & 1u)) break;
    if (!wc_ready) { mbar_wait(bar_wc_full, 0); wc_ready = true; }
    tc_fence_after();
    if (elect_one()) {
      const uint32_t slab = (uint32_t)(j & 1);
      const uint32_t wc = s_wc + ((uint32_t)job * 4u + slab * 2u) * 4096u;      // [hi 4 KB][lo 4 KB] of this slab
#pragma unroll
      for (int g = 0; g < 3; ++g) {            // 0: Y_hi * W_hi, 1: Y_lo * W_hi, 2: Y_hi * W_lo
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = desc_b | desc_addr(s_y + (g == 1 ? 16384u : 0u) + (uint32_t)k * 32u);
          const uint64_t bd = desc_b | desc_addr(wc + (g == 2 ? 4096u : 0u) + (uint32_t)k * 32u);
          umma_f16_2sm(d, ad, bd, idesc_1x1, (slab == 0 && g == 0 && k == 0) ? 0u : 1u);
        }
      }
      umma_commit_2sm(bar_y_done);
    }
    __syncwarp();
    ++y_uses; ++j; --left; ++n;
  }
  return n;
}

// ------------------------------------------------------------------------------------------------
// host helpers

uint32_t make_idesc(int operand, int n, int m = 128) {
  // UMMA instruction descriptor: D = F32 (bits 4-5 = 1), A/B format (bits 7-9 / 10-12), both K-major,
  // N >> 3 at bit 17, M >> 4 at bit 24 (M = 128, or 256 for cta_group::2).
  const uint32_t fmt = (uint32_t)tc_umma_format(operand);
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

inline uint16_t f32_to_bf16(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline uint16_t f32_to_f16(float f) {
  __half h = __float2half_rn(f);
  uint16_t r; memcpy(&r, &h, 2);
  return r;
}
inline uint32_t f32_to_tf32(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return u;
  u += 0xfffu + ((u >> 13) & 1u);   // round to nearest even at bit 13
  return u & ~0x1fffu;
}

// Writes element (row, k) of a K-major block with 128-B rows into its SWIZZLE_128B position.
// TC_SPLIT16: `lo_block` is the lo-plane twin of the block; the value is scaled first (TcConvPlan::scale).
inline void put_elem(uint8_t* block, int row, int k, int esize, float v, int operand, uint8_t* lo_block = nullptr,
                     float scale = 1.f) {
  const int byte_in_row = k * esize;
  const int chunk = byte_in_row >> 4, within = byte_in_row & 15;
  const size_t off = (size_t)row * 128 + (size_t)((chunk ^ (row & 7)) << 4) + within;
  uint8_t* dst = block + off;
  if (operand == TC_TF32) { const uint32_t u = f32_to_tf32(v); memcpy(dst, &u, 4); }
  else if (operand == TC_SPLIT16) {
    const float vs = v * scale;                       // exact: scale is a power of two
    const __half h = __float2half_rn(vs);
    const __half l = __float2half_rn(vs - __half2float(h));
    memcpy(dst, &h, 2);
    memcpy(lo_block + off, &l, 2);
  }
  else { const uint16_t u = operand == TC_BF16 ? f32_to_bf16(v) : f32_to_f16(v); memcpy(dst, &u, 2); }
}

void fill_orders(TcConvPlan& p, bool centre_first) {
  p.ndx = p.ndy = p.ks;
  if (centre_first && p.ks == 5) {
    const int ord[5] = {2, 1, 3, 0, 4};
    for (int i = 0; i < 5; ++i) p.dx_ord[i] = p.dy_ord[i] = ord[i];
  } else {
    for (int i = 0; i < p.ks; ++i) p.dx_ord[i] = p.dy_ord[i] = i;
  }
}

}  // namespace

float tc_pick_scale(const float* w, size_t n, const float* w2, size_t n2) {
  float mx = 0.f;
  for (size_t i = 0; i < n; ++i) mx = std::fmax(mx, std::fabs(w[i]));
  for (size_t i = 0; i < n2; ++i) mx = std::fmax(mx, std::fabs(w2[i]));
  if (!(mx > 0.f) || !std::isfinite(mx)) return 1.f;
  int e = 0;
  std::frexp(mx, &e);                 // mx = f * 2^e, f in [0.5, 1)  ->  mx * 2^(11 - e) in [1024, 2048)
  int k = 11 - e;
  if (k > 60) k = 60;
  if (k < -60) k = -60;
  return std::ldexp(1.f, k);
}

TcConvPlan tc_make_plan(int ks, int cin, int cout, int operand, float scale) {
  TcConvPlan p;
  p.ks = ks; p.operand = operand;
  const int esize = operand == TC_TF32 ? 4 : 2;
  p.split = operand == TC_SPLIT16 ? 1 : 0;
  p.scale = p.split ? scale : 1.f;
  p.slab_elems = 128 / esize;
  p.nslab = cin / p.slab_elems;
  p.n_cols = cout;
  p.pair = 0;
  fill_orders(p, false);
  uint32_t off = 0;
  for (int dxi = 0; dxi < p.ndx; ++dxi)
    for (int dyi = 0; dyi < p.ndy; ++dyi) {
      p.b_bytes[dxi][dyi] = (uint32_t)cout * 128u;
      p.b_off[dxi][dyi] = off;
      off += p.b_bytes[dxi][dyi] * (p.split ? 2u : 1u);
    }
  p.slab_bytes = off;
  return p;
}

TcConvPlan tc_make_pair_plan(int cin, int operand, float scale) {
  TcConvPlan p;
  p.ks = 5; p.operand = operand;
  const int esize = operand == TC_TF32 ? 4 : 2;
  p.split = operand == TC_SPLIT16 ? 1 : 0;
  p.scale = p.split ? scale : 1.f;
  p.slab_elems = 128 / esize;
  p.nslab = cin / p.slab_elems;
  p.n_cols = 128;
  p.pair = 1;
  fill_orders(p, true);   // the first tap issued must cover all 128 columns (accumulate = 0)
  uint32_t off = 0;
  for (int dxi = 0; dxi < 5; ++dxi)
    for (int dyi = 0; dyi < 5; ++dyi) {
      const bool inner = std::abs(p.dx_ord[dxi] - 2) <= 1 && std::abs(p.dy_ord[dyi] - 2) <= 1;
      p.b_bytes[dxi][dyi] = (inner ? 128u : 64u) * 128u;
      p.b_off[dxi][dyi] = off;
      off += p.b_bytes[dxi][dyi] * (p.split ? 2u : 1u);
    }
  p.slab_bytes = off;
  return p;
}

void tc_pack_weights(const TcConvPlan& p, const float* w, std::vector<uint8_t>& dst) {
  const int esize = p.operand == TC_TF32 ? 4 : 2;
  const int cin = p.nslab * p.slab_elems, ks = p.ks, cout = p.n_cols;
  dst.assign(p.total_bytes(), 0);
  for (int s = 0; s < p.nslab; ++s)
    for (int dxi = 0; dxi < p.ndx; ++dxi)
      for (int dyi = 0; dyi < p.ndy; ++dyi) {
        uint8_t* block = dst.data() + (size_t)s * p.slab_bytes + p.b_off[dxi][dyi];
        uint8_t* lo = block + p.b_bytes[dxi][dyi];     // split plans only
        const int dx = p.dx_ord[dxi], dy = p.dy_ord[dyi];
        for (int n = 0; n < cout; ++n)
          for (int k = 0; k < p.slab_elems; ++k) {
            const int ci = s * p.slab_elems + k;
            put_elem(block, n, k, esize, w[(((size_t)n * cin + ci) * ks + dy) * ks + dx], p.operand, lo, p.scale);
          }
      }
}

void tc_pack_pair_weights(const TcConvPlan& p, const float* w3, const float* w5, bool three_first,
                          std::vector<uint8_t>& dst) {
  const int esize = p.operand == TC_TF32 ? 4 : 2;
  const int cin = p.nslab * p.slab_elems;
  dst.assign(p.total_bytes(), 0);
  const int row3 = three_first ? 0 : 64, row5 = three_first ? 64 : 0;
  for (int s = 0; s < p.nslab; ++s)
    for (int dxi = 0; dxi < 5; ++dxi)
      for (int dyi = 0; dyi < 5; ++dyi) {
        uint8_t* block = dst.data() + (size_t)s * p.slab_bytes + p.b_off[dxi][dyi];
        uint8_t* lo = block + p.b_bytes[dxi][dyi];     // split plans only
        const int dx = p.dx_ord[dxi], dy = p.dy_ord[dyi];
        const bool inner = p.b_bytes[dxi][dyi] == 128u * 128u;
        for (int n = 0; n < 64; ++n)
          for (int k = 0; k < p.slab_elems; ++k) {
            const int ci = s * p.slab_elems + k;
            const float v5 = w5[(((size_t)n * cin + ci) * 5 + dy) * 5 + dx];
            put_elem(block, inner ? row5 + n : n, k, esize, v5, p.operand, lo, p.scale);
            if (inner) {
              const float v3 = w3[(((size_t)n * cin + ci) * 3 + (dy - 1)) * 3 + (dx - 1)];
              put_elem(block, row3 + n, k, esize, v3, p.operand, lo, p.scale);
            }
          }
      }
}

cudaError_t tc_encode_tmap(CUtensorMap* map, const void* base, int act, int C, int W, int H, int B,
                           int slab_elems, int box_w, int box_h) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  // resolved once per process; one host thread per GPU may race here, hence the atomic (same value from every thread)
  static std::atomic<EncodeFn> encode_cache{nullptr};
  EncodeFn encode = encode_cache.load(std::memory_order_acquire);
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return e;
    if (!fn || qres != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
    encode = reinterpret_cast<EncodeFn>(fn);
    encode_cache.store(encode, std::memory_order_release);
  }
  // split activations: the map sees the fp16 planes, a pixel of C containers = 2C fp16 ([64 hi | 64 lo] per slab)
  const int es = act == ACT_SPLIT16 ? 2 : act_bytes(act);
  if (act == ACT_SPLIT16) C *= 2;
  const CUtensorMapDataType dt = act == ACT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                               : act == ACT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
  const cuuint32_t box[4] = {(cuuint32_t)slab_elems, (cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = encode(map, dt, 4, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// 2-D view of a packed weight stream for the 2-CTA kernel: rows of 128 bytes, box = 32 rows (4 KB),
// no swizzle (the stream already holds the SWIZZLE_128B shared-memory image).
cudaError_t tc_encode_bmap(CUtensorMap* map, const void* base, size_t bytes) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess) return e;
  if (!fn || qres != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
  const cuuint64_t dims[2] = {128, (cuuint64_t)(bytes / 128)};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {128, 32};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = reinterpret_cast<EncodeFn>(fn)(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims,
                                                    strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

namespace {
// shared geometry set-up; returns the dynamic shared memory the launch needs (0 if it does not fit)
template <int NACC>
size_t setup_geometry(TcKParams& kp, int b_stage_bytes_total, bool split = false) {
  using G = Geo<NACC>;
  kp.tiles_x = cdiv(kp.W, G::TW);
  kp.tiles_y = cdiv(kp.H, G::TH);
  kp.tiles_per_job = kp.B * kp.tiles_x * kp.tiles_y;
  kp.total_tiles = kp.tiles_per_job * kp.njobs;
  kp.pw = G::TW + kp.ks - 1;
  const int ph = G::TH + kp.ks - 1;
  kp.patch_tx = (uint32_t)kp.pw * ph * 128u;
  kp.patch_stage = (kp.patch_tx + 1023u) & ~1023u;
  kp.nbuf = (2 * NACC * kp.n_cols <= 512) ? 2 : 1;
  if (split) kp.nbuf = 2;        // two chunk buffers of [big n_cols][small n_cols] columns
  const size_t fixed = 1024 + kBarBytes + (size_t)b_stage_bytes_total;
  int npb = (int)((232448 - fixed) / kp.patch_stage);
  if (npb > kMaxNPB) npb = kMaxNPB;
  if (split) npb &= ~1;          // hi / lo patches of a slab occupy two consecutive stages
  if (npb < 2) return 0;
  kp.npb = npb;
  return fixed + (size_t)npb * kp.patch_stage;
}

// the compile-time tap schedule of the cluster kernel must be the one the weights were packed with
template <int KIND>
bool plan_matches_kind(const TcConvPlan& plan) {
  using TP = Taps<KIND>;
  if (plan.ks != TP::KS || plan.ndx != TP::KS || plan.ndy != TP::KS || (plan.pair != 0) != (KIND == TK_PAIR)) return false;
  if (KIND == TK_PAIR && plan.n_cols != 128) return false;
  uint32_t off = 0;
  for (int t = 0; t < TP::NT; ++t) {
    const int dxi = t / TP::KS, dyi = t % TP::KS;
    if (plan.dx_ord[dxi] != TP::dx(t) || plan.dy_ord[dyi] != TP::dy(t)) return false;
    const uint32_t bytes = (KIND == TK_PAIR ? (TP::outer(t) ? 64u : 128u) : (uint32_t)plan.n_cols) * 128u;
    const uint32_t planes = plan.split ? 2u : 1u;
    if (plan.b_bytes[dxi][dyi] != bytes || plan.b_off[dxi][dyi] != off) return false;
    if (KIND == TK_PAIR && off != (uint32_t)TP::units_before(t) * 8192u * planes) return false;
    off += bytes * planes;
  }
  return plan.slab_bytes == off;
}

template <int NACC, int OPERAND, bool FUSE, int KIND>
cudaError_t launch_nacc2(const CUtensorMap& tmap, const CUtensorMap& tmapj1, const CUtensorMap& b0, const CUtensorMap& b1,
                         const CUtensorMap& w0, const CUtensorMap& w1, TcKParams& kp, cudaStream_t st) {
  // function attributes are per device: configure once per (kernel instantiation, device); one host thread per GPU
  static std::atomic<int> sms_of_dev[kMaxDevices];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  int num_sms = sms_of_dev[dev].load(std::memory_order_acquire);
  if (num_sms == 0) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<NACC, OPERAND, FUSE, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    sms_of_dev[dev].store(num_sms, std::memory_order_release);
  }
  // fused mode adds the resident 1x1 weights (16 KB; split: 32 KB) and the Y staging tile (32 KB) behind the B ring
  const size_t smem = setup_geometry<NACC>(kp, Ring<KIND, OPERAND>::NST * Taps<KIND>::kStageBytes + (FUSE ? (OPERAND == TC_SPLIT16 ? 65536 : 49152) : 0),
                                           OPERAND == TC_SPLIT16);
  if (!smem) return cudaErrorInvalidConfiguration;
  if (kp.ks != Taps<KIND>::KS || (uint32_t)kp.n_cols * 64u > Taps<KIND>::kStageBytes) return cudaErrorInvalidValue;
  const int items = ((kp.tiles_per_job + 1) / 2) * kp.njobs;     // pair-tiles
  int clusters = num_sms / 2;
  if (items < clusters) clusters = items;
  {
    const int rem = items % clusters;
    const bool split = NACC > 1 && items > clusters && cdiv(rem * NACC, clusters) < NACC;
    kp.main_tiles = split ? items - rem : items;
    kp.total_items = kp.main_tiles + (items - kp.main_tiles) * NACC;
  }
#ifdef CODON_TC_EXPERIMENT
  const char* pdl_env = getenv("CODON_TC_PDL");          // 0: plain stream order (perf experiments)
  const int pdl = pdl_env ? atoi(pdl_env) : 1;
#else
  const int pdl = 1;
#endif
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(kThreads2); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, conv_tc2_kernel<NACC, OPERAND, FUSE, KIND>, tmap, tmapj1, b0, b1, w0, w1, kp);
}

template <int NACC, int OPERAND>
cudaError_t launch_nacc(const CUtensorMap& tmap, const CUtensorMap& tmapj1, TcKParams& kp, cudaStream_t st) {
  static std::atomic<int> sms_of_dev[kMaxDevices];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  int num_sms = sms_of_dev[dev].load(std::memory_order_acquire);
  if (num_sms == 0) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<NACC, OPERAND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    sms_of_dev[dev].store(num_sms, std::memory_order_release);
  }
  const size_t smem = setup_geometry<NACC>(kp, kBStages * kBStageBytes);
  if (!smem) return cudaErrorInvalidConfiguration;
  {
    const int g = kp.total_tiles < num_sms ? kp.total_tiles : num_sms;
    const int rem = kp.total_tiles % g;
    // split the tail only when that shortens it: rem*NACC single sub-tiles over g CTAs
    const bool split = NACC > 1 && kp.total_tiles > g && cdiv(rem * NACC, g) < NACC;
    kp.main_tiles = split ? kp.total_tiles - rem : kp.total_tiles;
    kp.total_items = kp.main_tiles + (kp.total_tiles - kp.main_tiles) * NACC;
  }
  const int grid = kp.total_tiles < num_sms ? kp.total_tiles : num_sms;
  conv_tc_kernel<NACC, OPERAND><<<grid, kThreads, smem, st>>>(tmap, tmapj1, kp);
  return cudaGetLastError();
}
}  // namespace

int tc_selftest() {
  int bad = 0;
  const int ops[4] = {TC_F16, TC_BF16, TC_TF32, TC_SPLIT16};
  for (int oi = 0; oi < 4; ++oi) {
    const int op = ops[oi];
    if (!plan_matches_kind<TK_3X3>(tc_make_plan(3, 64, 64, op))) bad |= 1;
    if (!plan_matches_kind<TK_3X3>(tc_make_plan(3, 128, 64, op))) bad |= 1;
    if (!plan_matches_kind<TK_5X5>(tc_make_plan(5, 128, 128, op))) bad |= 2;
    if (!plan_matches_kind<TK_5X5>(tc_make_plan(5, 64, 64, op))) bad |= 2;
    if (!plan_matches_kind<TK_PAIR>(tc_make_pair_plan(64, op))) bad |= 4;
    if (plan_matches_kind<TK_5X5>(tc_make_pair_plan(64, op)) || plan_matches_kind<TK_PAIR>(tc_make_plan(5, 64, 128, op))) bad |= 8;
    // packing: element (n, ci, dy, dx) of an OIHW tensor must be found at row n, K index ci % slab of block (slab, tap)
    const TcConvPlan p = tc_make_plan(3, 128, 64, op, 8.f);
    const int es = op == TC_TF32 ? 4 : 2, cin = 128;
    std::vector<float> w((size_t)64 * cin * 9);
    for (size_t i = 0; i < w.size(); ++i) w[i] = (float)((int)(i % 251) - 125);      // exactly representable in every operand type
    // split plans: 19 significant bits, only representable as hi + lo; the scale (8) must come back out
    if (op == TC_SPLIT16) for (size_t i = 0; i < w.size(); ++i) w[i] += 1.f / 4096.f;
    std::vector<uint8_t> packed;
    tc_pack_weights(p, w.data(), packed);
    if (packed.size() != p.total_bytes()) { bad |= 16; continue; }
    for (int n = 0; n < 64; n += 7)
      for (int ci = 0; ci < cin; ci += 5)
        for (int t = 0; t < 9; ++t) {
          const int dxi = t / 3, dyi = t % 3, s = ci / p.slab_elems, k = ci % p.slab_elems;
          const uint8_t* block = packed.data() + (size_t)s * p.slab_bytes + p.b_off[dxi][dyi];
          const int byte = k * es;
          const uint8_t* src = block + (size_t)n * 128 + (size_t)(((byte >> 4) ^ (n & 7)) << 4) + (byte & 15);
          const float want = w[(((size_t)n * cin + ci) * 3 + p.dy_ord[dyi]) * 3 + p.dx_ord[dxi]];
          float got;
          if (op == TC_TF32) memcpy(&got, src, 4);
          else {
            uint16_t u; memcpy(&u, src, 2);
            if (op == TC_BF16) { uint32_t v = (uint32_t)u << 16; memcpy(&got, &v, 4); }
            else { __half h; memcpy(&h, &u, 2); got = __half2float(h); }
            if (op == TC_SPLIT16) {
              __half l; memcpy(&l, src + p.b_bytes[dxi][dyi], 2);
              got = (got + __half2float(l)) / p.scale;
            }
          }
          if (got != want) bad |= 32;
        }
  }
  return bad;
}

cudaError_t launch_conv_tc(const CUtensorMap& tmap, const CUtensorMap& tmapj1, const TcConvPlan& plan, const TcLaunch& L,
                           cudaStream_t st) {
  TcKParams kp;
  memset(&kp, 0, sizeof(kp));
  for (int i = 0; i < L.njobs; ++i) kp.job[i] = L.job[i];
  kp.njobs = L.njobs; kp.B = L.B; kp.H = L.H; kp.W = L.W;
  kp.nslab = plan.nslab; kp.slab_elems = plan.slab_elems; kp.ks = plan.ks; kp.pad = plan.ks / 2;
  kp.ndx = plan.ndx; kp.ndy = plan.ndy;
  for (int i = 0; i < kTcMaxTaps; ++i) { kp.dx_ord[i] = plan.dx_ord[i]; kp.dy_ord[i] = plan.dy_ord[i]; }
  memcpy(kp.b_bytes, plan.b_bytes, sizeof(kp.b_bytes));
  memcpy(kp.b_off, plan.b_off, sizeof(kp.b_off));
  kp.slab_bytes = plan.slab_bytes;
  kp.n_cols = plan.n_cols;
  const int mma_m = L.two_cta ? 256 : 128;
  kp.idesc_full = make_idesc(plan.operand, plan.n_cols, mma_m);
  kp.idesc_half = make_idesc(plan.operand, 64, mma_m);
  kp.relu = L.relu; kp.out_act = L.out_act; kp.is_tf32 = plan.operand == TC_TF32;
#ifdef CODON_TC_EXPERIMENT
  {
    const char* e = getenv("CODON_TC_DEBUG");      // perf experiments only: results are garbage
    kp.debug = e ? atoi(e) : 0;
  }
#endif
  if (L.out_act != (plan.operand == TC_TF32 ? ACT_F32 : plan.operand == TC_BF16 ? ACT_BF16 : plan.operand == TC_SPLIT16 ? ACT_SPLIT16 : ACT_F16))
    return cudaErrorInvalidValue;   // activations are stored in the operand type
  if (plan.operand == TC_SPLIT16) {
    // split-fp16 operands: cluster kernel only; tile sizes whose two-plane patch ring fits shared memory
    if (!L.two_cta || !L.bmap[0] || (L.njobs > 1 && !L.bmap[1])) return cudaErrorInvalidValue;
    for (int i = 0; i < L.njobs; ++i) if (L.job[i].pool && !L.fuse) return cudaErrorInvalidValue;
    const CUtensorMap& b0 = *L.bmap[0];
    const CUtensorMap& b1 = *L.bmap[L.njobs > 1 ? 1 : 0];
    if (L.fuse) {
      if (plan.ks != 5 || plan.n_cols != 128 || plan.pair || L.nacc != 1 || !L.wmap[0] || (L.njobs > 1 && !L.wmap[1]) ||
          L.y16_operand != TC_F16 || !plan_matches_kind<TK_5X5>(plan))
        return cudaErrorInvalidValue;
      kp.idesc_1x1 = make_idesc(TC_F16, 64, 256);
      kp.fuse_njobs = L.njobs;
      kp.pool_stride = L.pool_stride ? L.pool_stride : (unsigned long long)L.B * L.H * L.W;
      kp.cells_x = cdiv(L.W, kTcSubW); kp.cells_y = cdiv(L.H, kTcSubH);
      kp.core_y0 = L.core_y1 > 0 ? L.core_y0 : 0; kp.core_y1 = L.core_y1 > 0 ? (L.core_y1 < L.H ? L.core_y1 : L.H) : L.H;
      return launch_nacc2<1, TC_SPLIT16, true, TK_5X5>(tmap, tmapj1, b0, b1, *L.wmap[0], *L.wmap[L.njobs > 1 ? 1 : 0], kp, st);
    }
    if (plan.pair) {
      if (L.nacc != 1 || !plan_matches_kind<TK_PAIR>(plan)) return cudaErrorInvalidValue;
      return launch_nacc2<1, TC_SPLIT16, false, TK_PAIR>(tmap, tmapj1, b0, b1, b0, b1, kp, st);
    }
    if (plan.ks == 3 && L.nacc == 1 && plan_matches_kind<TK_3X3>(plan))
      return launch_nacc2<1, TC_SPLIT16, false, TK_3X3>(tmap, tmapj1, b0, b1, b0, b1, kp, st);
    return cudaErrorInvalidValue;
  }
  for (int i = 0; i < L.njobs; ++i)
    if (L.job[i].pool && !L.fuse && (plan.n_cols != 64 || L.two_cta)) return cudaErrorInvalidValue;
  if (L.two_cta) {
    if (!L.bmap[0] || (L.njobs > 1 && !L.bmap[1])) return cudaErrorInvalidValue;
    const CUtensorMap& b0 = *L.bmap[0];
    const CUtensorMap& b1 = *L.bmap[L.njobs > 1 ? 1 : 0];
    if (L.fuse) {
      if (plan.ks != 5 || plan.n_cols != 128 || plan.pair || L.nacc > 2 || !L.wmap[0] || (L.njobs > 1 && !L.wmap[1]) ||
          L.y16_operand != (plan.operand == TC_BF16 ? TC_BF16 : TC_F16))
        return cudaErrorInvalidValue;
      kp.idesc_1x1 = make_idesc(L.y16_operand, 64, 256);
      kp.fuse_njobs = L.njobs;
      kp.pool_stride = L.pool_stride ? L.pool_stride : (unsigned long long)L.B * L.H * L.W;
      kp.cells_x = cdiv(L.W, kTcSubW); kp.cells_y = cdiv(L.H, kTcSubH);
      kp.core_y0 = L.core_y1 > 0 ? L.core_y0 : 0; kp.core_y1 = L.core_y1 > 0 ? (L.core_y1 < L.H ? L.core_y1 : L.H) : L.H;
      const CUtensorMap& w0 = *L.wmap[0];
      const CUtensorMap& w1 = *L.wmap[L.njobs > 1 ? 1 : 0];
      if (!plan_matches_kind<TK_5X5>(plan)) return cudaErrorInvalidValue;
#define CODON_TC2F_DISPATCH(N)                                                          \
  switch (plan.operand) {                                                               \
    case TC_F16: return launch_nacc2<N, TC_F16, true, TK_5X5>(tmap, tmapj1, b0, b1, w0, w1, kp, st);   \
    case TC_BF16: return launch_nacc2<N, TC_BF16, true, TK_5X5>(tmap, tmapj1, b0, b1, w0, w1, kp, st); \
    case TC_TF32: return launch_nacc2<N, TC_TF32, true, TK_5X5>(tmap, tmapj1, b0, b1, w0, w1, kp, st); \
    default: return cudaErrorInvalidValue;                                              \
  }
      switch (L.nacc) {
        case 1: CODON_TC2F_DISPATCH(1)
        case 2: CODON_TC2F_DISPATCH(2)
        default: return cudaErrorInvalidValue;
      }
#undef CODON_TC2F_DISPATCH
    }
#define CODON_TC2_DISPATCH_K(N, K)                                                          \
  if (!plan_matches_kind<K>(plan)) return cudaErrorInvalidValue;                               \
  switch (plan.operand) {                                                                      \
    case TC_F16: return launch_nacc2<N, TC_F16, false, K>(tmap, tmapj1, b0, b1, b0, b1, kp, st);  \
    case TC_BF16: return launch_nacc2<N, TC_BF16, false, K>(tmap, tmapj1, b0, b1, b0, b1, kp, st);\
    case TC_TF32: return launch_nacc2<N, TC_TF32, false, K>(tmap, tmapj1, b0, b1, b0, b1, kp, st);\
    default: return cudaErrorInvalidValue;                                                     \
  }
#define CODON_TC2_DISPATCH(N)                                                                  \
  if (plan.pair) { CODON_TC2_DISPATCH_K(N, TK_PAIR) }                                          \
  else if (plan.ks == 5) { CODON_TC2_DISPATCH_K(N, TK_5X5) }                                   \
  else if (plan.ks == 3) { CODON_TC2_DISPATCH_K(N, TK_3X3) }                                   \
  else return cudaErrorInvalidValue;
    switch (L.nacc) {
      case 1: CODON_TC2_DISPATCH(1)
      case 2: CODON_TC2_DISPATCH(2)
      case 4: CODON_TC2_DISPATCH(4)
      default: return cudaErrorInvalidValue;
    }
#undef CODON_TC2_DISPATCH
#undef CODON_TC2_DISPATCH_K
  }
#define CODON_TC_DISPATCH(N)                                                   \
  switch (plan.operand) {                                                      \
    case TC_F16: return launch_nacc<N, TC_F16>(tmap, tmapj1, kp, st);                  \
    case TC_BF16: return launch_nacc<N, TC_BF16>(tmap, tmapj1, kp, st);                \
    case TC_TF32: return launch_nacc<N, TC_TF32>(tmap, tmapj1, kp, st);                \
    default: return cudaErrorInvalidValue;                                     \
  }
  switch (L.nacc) {
    case 1: CODON_TC_DISPATCH(1)
    case 2: CODON_TC_DISPATCH(2)
    case 4: CODON_TC_DISPATCH(4)
    default: return cudaErrorInvalidValue;
  }
#undef CODON_TC_DISPATCH
}

}  // namespace codon
