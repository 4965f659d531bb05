// tcgen05 implicit-GEMM convolution: host-visible plan / launch structures (see conv_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

namespace codon {

constexpr int kTcSubW = 8;            // an accumulator (128 GEMM rows) covers a sub-tile of 8 x 16 pixels
constexpr int kTcSubH = 16;
constexpr int kTcMaxTaps = 5;

// TC_F16 / TC_BF16 / TC_TF32 == UMMA F16F32Format.  TC_SPLIT16 (F16X3 mode): every operand is carried as two fp16
// planes hi = fp16(v), lo = fp16(v - hi) (~22 mantissa bits) and every K step issues hi*hi + lo*hi + hi*lo as three
// kind::f16 MMAs into the same fp32 accumulator -- fp32-grade products at a third of the fp16 tensor rate.
enum TcOperand : int { TC_F16 = 0, TC_BF16 = 1, TC_TF32 = 2, TC_SPLIT16 = 3 };
inline int tc_umma_format(int operand) { return operand == TC_SPLIT16 ? TC_F16 : operand; }

// Geometry of the packed B (weight) stream of one convolution: for every 128-byte input-channel
// slab, for every tap in issue order, one K-major block of `rows` x 128 B, already in the
// SWIZZLE_128B shared-memory image, so the kernel fetches it with one cp.async.bulk.
struct TcConvPlan {
  int ks = 1;                  // 1, 3, 5
  int nslab = 1;               // Cin * elem_bytes / 128
  int slab_elems = 64;         // 64 (16-bit operands) or 32 (tf32)
  int n_cols = 64;             // accumulator columns per pixel (Cout of the fused layer(s))
  int pair = 0;                // 1: fused 3x3 + 5x5 pair (inner taps 128 rows, outer taps 64)
  int ndx = 1, ndy = 1;
  int dx_ord[kTcMaxTaps] = {0, 0, 0, 0, 0};
  int dy_ord[kTcMaxTaps] = {0, 0, 0, 0, 0};
  uint32_t b_bytes[kTcMaxTaps][kTcMaxTaps] = {};   // [dxi][dyi]
  uint32_t b_off[kTcMaxTaps][kTcMaxTaps] = {};     // byte offset inside one slab group
  uint32_t slab_bytes = 0;
  int operand = TC_BF16;
  // TC_SPLIT16: every tap's block is followed by its lo-plane twin (b_off = offset of the hi block, the lo block
  // sits b_bytes later; slab_bytes covers both), and the weights are multiplied by `scale` (a power of two that
  // moves them into the fp16 normal range so that the lo plane keeps its 11 bits) before the split; the kernel
  // epilogue multiplies the accumulator by 1 / scale (TcJob::descale).
  int split = 0;
  float scale = 1.f;
  size_t total_bytes() const { return (size_t)slab_bytes * nslab; }
};

// Power-of-two scale that puts max|w| into [1024, 2048) (1 if all weights are zero).
float tc_pick_scale(const float* w, size_t n, const float* w2 = nullptr, size_t n2 = 0);

// Plans.  `cout`/`cin` in elements; the pair plan is 64 -> (64 | 64) with kernel sizes 3 and 5.
TcConvPlan tc_make_plan(int ks, int cin, int cout, int operand, float scale = 1.f);
TcConvPlan tc_make_pair_plan(int cin, int operand, float scale = 1.f);

// Pack OIHW fp32 weights into the plan's byte stream.  For a pair plan, w3 (3x3) and w5 (5x5)
// are both [64][cin][k][k]; `three_first` puts the 3x3 result in columns 0..63 (depth branch,
// CODON_x4.py:79) and the 5x5 result in 64..127, otherwise the other way round (colour branch
// :80 and fusion stages :125).  `in_perm` (may be null) maps packed input channel -> OIHW input
// channel.
void tc_pack_weights(const TcConvPlan& plan, const float* w, std::vector<uint8_t>& dst);
void tc_pack_pair_weights(const TcConvPlan& plan, const float* w3, const float* w5, bool three_first,
                          std::vector<uint8_t>& dst);

struct TcJob {
  int in_coff;                 // channel coordinate of the job's first input channel in the tensor map
  const uint8_t* w;            // device pointer to the packed stream
  void* out;       int out_stride; int out_off;   // elements
  const void* res; int res_stride; int res_off;
  int outer_col;               // pair plans: accumulator column of the 5x5-only (outer) taps
  float descale = 1.f;         // TC_SPLIT16: 1 / TcConvPlan::scale of this job's weights (applied to the accumulator)
  float descale2 = 1.f;        // TC_SPLIT16, fused 1x1: 1 / scale of the 1x1 weights
  // fused 1x1 (TcLaunch::fuse): out2 = conv1x1(relu(this conv)) (+ res2), 64 channels of the activation type
  void* out2 = nullptr;       int out2_stride = 0; int out2_off = 0;
  const void* res2 = nullptr; int res2_stride = 0; int res2_off = 0;
  float2* pool;                // optional (64-column launches, 1-CTA kernel): per-pixel (max, sum) over this job's
                               // 64 output channels -> pool[pixel]; the CAC ChannelPool partial (CAC_module.py:78-81)
  // fused mode, optional: per-channel (sum, max) of out2 over the 32 pixels of every epilogue warp -- the partials of
  // CAC_channel's global average / max pool (CAC_module.py:43,47).  Layout: [frame][cell_y][cell_x][lane quarter 4]
  // [channel 64] float2, one cell = one 8 x 16-pixel sub-tile (cells_x = ceil(W / 8), cells_y = ceil(H / 16)).
  float2* cstat = nullptr;
};

struct TcLaunch {
  TcJob job[2];
  int njobs = 1;
  int B = 0, H = 0, W = 0;
  int relu = 0;
  int out_act = 0;             // ActType of out / res
  int nacc = 4;                // accumulators (128-pixel sub-tiles) per CTA tile: 1, 2 or 4
  size_t pool_stride = 0;      // fused mode: pixels between the two 32-channel-half pool maps of a job (0: B*H*W)
  // fused mode, row-band multi-GPU forward: only image rows [core_y0, core_y1) count towards TcJob::cstat (the other
  // rows are halo rows owned by a neighbouring band); core_y1 <= 0 means all rows
  int core_y0 = 0, core_y1 = 0;
  int fuse = 0;                // 1 (two_cta, 5x5 128->128 only): the following 1x1 128->64 convolution runs as a second
                               // GEMM out of TMEM inside the same kernel (job.out2 / res2, wmap); job.out is not written
  int y16_operand = TC_BF16;   // 16-bit type of the staged ReLU output and of the 1x1 weights (TC_F16 / TC_BF16)
  const CUtensorMap* wmap[2] = {nullptr, nullptr};   // per job: 2-D map of the packed 1x1 weight stream
  int two_cta = 0;             // 1: cluster-of-2 kernel (tcgen05 cta_group::2, M = 256); needs bmap
  const CUtensorMap* bmap[2] = {nullptr, nullptr};   // per job: 2-D map of the packed weight stream
};

// Tile geometry for `nacc` accumulators: NAX x NAY sub-tiles (1x1, 2x1, 2x2).
inline int tc_tile_w(int nacc) { return (nacc >= 2 ? 2 : 1) * kTcSubW; }
inline int tc_tile_h(int nacc) { return (nacc >= 4 ? 2 : 1) * kTcSubH; }
// Patch box (pixels, rows) the A tensor map must be encoded with for this plan / nacc.
inline int tc_box_w(const TcConvPlan& p, int nacc) { return tc_tile_w(nacc) + p.ks - 1; }
inline int tc_box_h(const TcConvPlan& p, int nacc) { return tc_tile_h(nacc) + p.ks - 1; }

// Encodes the 4-D NHWC tensor map {C, W, H, B} with box {slab_elems, box_w, box_h, 1}, SWIZZLE_128B.
cudaError_t tc_encode_tmap(CUtensorMap* map, const void* base, int act, int C, int W, int H, int B,
                           int slab_elems, int box_w, int box_h);

// 2-D tensor map over a packed weight stream (rows of 128 B, 32-row boxes) for the 2-CTA kernel.
cudaError_t tc_encode_bmap(CUtensorMap* map, const void* base, size_t bytes);

// Host-only consistency check (no GPU): every plan the engine builds must equal the compile-time tap schedule of
// the cluster kernel it is launched with, and packed weights must land where the K-major SWIZZLE_128B image says.
// Returns 0, or a bit per failed check.
int tc_selftest();

// tmap / tmapj1: activation tensor maps of job 0 / job 1 (the same map twice when both jobs read one buffer).
cudaError_t launch_conv_tc(const CUtensorMap& tmap, const CUtensorMap& tmapj1, const TcConvPlan& plan, const TcLaunch& L,
                           cudaStream_t st);

}  // namespace codon
