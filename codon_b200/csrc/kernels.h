// Host-side launchers of the hand-written kernels (internal to libcodon_b200).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace codon {

// One convolution "job" of a launch.  A launch runs up to two jobs (the depth and the colour
// branch of a stage) over the same B x H x W pixel grid.  Tensors are NHWC with an arbitrary
// pixel stride (channels of the whole buffer) and a channel offset, so producers write into
// channel slices of a shared buffer and torch.cat (CODON_x4.py:79,80,85,119,125) disappears.
struct ConvJob {
  const void* in;  int in_stride;  int in_off;    // elements
  const void* w;                                   // layout depends on the kernel
  void* out;       int out_stride; int out_off;
  const void* res; int res_stride; int res_off;   // optional residual added after (no) ReLU
};

// ---- fp32 CUDA-core direct convolution (parity mode) -------------------------------------
// w: fp32 [KS*KS][Cin][Cout].  Cin % 8 == 0, Cout % 64 == 0.
cudaError_t launch_conv_direct_f32(const ConvJob* jobs, int njobs, int B, int H, int W,
                                   int Cin, int Cout, int KS, bool relu, cudaStream_t st);

// ---- edge layers (all modes) ----------------------------------------------------------------
// input / input_c: 1 -> 64, 3x3, ReLU (CODON_x4.py:68,71).  x, y fp32 [B,H,W]; w_d, w_c fp32
// [9][64]; out NHWC 128 ch (depth | colour) of type act.
cudaError_t launch_conv_first(const float* x, const float* y, const float* w_d, const float* w_c,
                              void* out, int act, int B, int H, int W, cudaStream_t st, int rnd_tf32 = 0);
// output: 64 -> 1, 3x3, + global residual x (CODON_x4.py:130-131).  in NHWC 64 ch of type act
// (pixel stride in_stride); w fp32 [9][64]; x, out fp32 [B,H,W].
cudaError_t launch_conv_last(const void* in, int in_stride, int act, const float* w, const float* x,
                             float* out, int B, int H, int W, cudaStream_t st);

// ---- CAC cross-domain attention (CAC_module.py, CODON_x4.py:85-118) -------------------------
// F: NHWC 128 ch (depth | colour).  stats: per-pixel (max, mean) over the 128 channels ->
// pooled [B,H,W,2] fp32, and per-chunk partial per-channel (sum, max) -> part [B][chunks][2][128].
int cac_stats_chunks(int B, int H, int W);
cudaError_t launch_cac_stats(const void* F, int act, int B, int H, int W, float* pooled,
                             float* part, int chunks, cudaStream_t st);
// lean stats (tensor-core modes): per-chunk per-channel (sum, max) only; the per-pixel ChannelPool partials
// come from the 1x1 conv epilogues (two [B,H,W] float2 arrays, one per branch: (max, sum) over 64 channels).
cudaError_t launch_cac_chan_stats(const void* F, int act, int B, int H, int W, float* part, int chunks,
                                  cudaStream_t st);
// Spatial gate map gate[B][H][W] = sigmoid(conv5x5(ChannelPool)) (CAC_module.py:90-94).  pool_parts 1: pooled = final
// (max, mean) [B,H,W,2]; 2 / 4: (max, sum) partial maps `part_stride` pixels apart (0: B*H*W).  ws fp32 [2][25]: max map
// taps, then mean map taps.
struct CacGate {
  const float* pooled;
  const float* ws;
  float* gate;
  int H, W, pool_parts;
  size_t part_stride;
};
cudaError_t launch_cac_gate(const CacGate& gate, int B, cudaStream_t st);
// fused conv path: folds the per-cell channel partials the conv epilogue wrote (conv_tc.h, TcJob::cstat; `cells` cells
// of 8 x 16 pixels per frame and branch) into chunks of the `part` layout; chunks = cac_cell_chunks(cells).  With `gate`
// the same launch also computes the spatial gate map (extra blocks).
int cac_cell_chunks(int cells);
cudaError_t launch_cac_cell_reduce(const void* cstat_d, const void* cstat_c, int B, int cells, float* part, int chunks,
                                   cudaStream_t st, const CacGate* gate = nullptr);
// mlp: deterministic reduce of the partials, MLP 128->8->64 on avg and max, sigmoid -> sc [B,64].
// w1 [8][128] indexed by Fcat channel (colour | depth, CODON_x4.py:85), b1 [8], w2 [64][8], b2 [64].
cudaError_t launch_cac_mlp(const float* part, int chunks, int B, int HW, const float* w1,
                           const float* b1, const float* w2, const float* b2, float* sc,
                           cudaStream_t st);
// apply: F = F * sc[b, c % 64] * gate[b,h,w] + E   (in place on F).
cudaError_t launch_cac_apply(void* F, const void* E, int act, const float* gate, const float* sc, int B, int H, int W,
                             cudaStream_t st, int rnd_tf32 = 0);

// ---- utility -------------------------------------------------------------------------------------
cudaError_t launch_convert_to_f32(const void* src, int dtype, float* dst, size_t n, cudaStream_t st);
cudaError_t launch_convert_from_f32(const float* src, void* dst, int dtype, size_t n, cudaStream_t st);
// NHWC (type act, pixel stride `stride`, channel offset `off`, C channels) -> fp32 NCHW
cudaError_t launch_nhwc_to_nchw_f32(const void* src, int act, int stride, int off, int C, int B, int HW,
                                    float* dst, cudaStream_t st);

// ---- stand-alone NCHW fp32 CAC / CBAM pieces ----------------------------------------------------
// stats [4][B*C]: mean, max, lp(2), lse per (b,c) plane; pool_mask bit k selects pool k in the MLP sum
cudaError_t launch_nchw_channel_stats(const float* x, int B, int C, int HW, int pool_mask, float* stats,
                                      cudaStream_t st);
cudaError_t launch_gate_mlp(const float* stats, int B, int C, int pool_mask, const float* w1, const float* b1,
                            const float* w2, const float* b2, int hidden, int c_out, float* scale,
                            cudaStream_t st);
cudaError_t launch_nchw_channel_pool(const float* x, int B, int C, int HW, float* pooled, cudaStream_t st);
cudaError_t launch_nchw_spatial_scale(const float* pooled, const float* w, int B, int H, int W,
                                      float* scale, cudaStream_t st);
cudaError_t launch_nchw_apply(const float* x, const float* sc, const float* ss, const float* res, int B,
                              int C, int HW, int c_gate, float* y, cudaStream_t st);

cudaError_t launch_conv2d_nchw(const float* x, const float* w, const float* bias, int B, int Cin, int H, int W,
                               int Cout, int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int groups,
                               int relu, float* y, cudaStream_t st);

// ---- driver post-processing and metrics (metrics.cu) ---------------------------------------------
cudaError_t launch_quantise_u8(const float* src, uint8_t* dst, size_t n, int via_half, cudaStream_t st);
cudaError_t launch_masked_rmse(const uint8_t* label, const uint8_t* out, int B, int HW, double* rmse, cudaStream_t st);
size_t ssim_workspace_bytes(int B, int H, int W);
cudaError_t launch_ssim_gauss(const void* a, const void* b, int img_dtype, int B, int H, int W, double sd, double c1, double c2,
                              double* ssim, double* ws, cudaStream_t st);

// ---- driver pre-processing (preproc.cu) ------------------------------------------------------------
cudaError_t launch_bgr_to_gray(const uint8_t* bgr, uint8_t* gray, size_t n, int method, cudaStream_t st);
cudaError_t launch_u8_to_unit(const uint8_t* src, float* dst, size_t n, cudaStream_t st);
cudaError_t launch_bicubic_up(const float* src, float* dst, int B, int h, int w, int H, int W, cudaStream_t st);

// ---- tcgen05 implicit-GEMM convolution (BF16 / FP16 / TF32 modes), conv_tc.cu -------------------
struct TcConvPlan;   // defined in conv_tc.h

}  // namespace codon
