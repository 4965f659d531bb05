/*
 * codon_b200.h -- C ABI of libcodon_b200.so: the B200 (sm_100a) engine for the CODON guided
 * depth super-resolution forward pass.
 *
 * The reference (619862306/CODON) is pure Python/PyTorch and has no native interface to
 * mirror; each entry point below names the reference call it replaces (paths relative to the
 * reference checkout).  Plain pointers and sizes only: no torch types cross this boundary.
 * The Python host (codon_b200/engine.py) binds these with ctypes; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns CODON_OK (0) or a negative codon_status; the message of the
 *     last failure on a context is available from codon_last_error().
 *   - device pointers are borrowed for the duration of the stream-ordered call; the library
 *     owns only its re-laid-out weight copies.
 *   - frames are single-channel, row-major, B x H x W (== NCHW == NHWC for C = 1), values in
 *     normalised depth / gray [0,1] as produced by CODON_X4/test.py:122-123.
 *   - there is no CPU fallback: without a CUDA device every call fails with CODON_ERR_CUDA.
 */
#ifndef CODON_B200_H
#define CODON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct codon_ctx codon_ctx;

typedef enum codon_status {
  CODON_OK = 0,
  CODON_ERR_ARG = -1,       /* bad argument (null pointer, unknown name, wrong shape) */
  CODON_ERR_STATE = -2,     /* call order (e.g. forward before finalize_weights) */
  CODON_ERR_CUDA = -3,      /* CUDA runtime / driver failure, or no device */
  CODON_ERR_WORKSPACE = -4  /* workspace too small */
} codon_status;

/* Arithmetic mode of the conv trunk.
 *   FP32 : fp32 activations, fp32 FFMA direct convolution (parity mode, max-abs <= 1e-3).
 *   BF16 : bf16 NHWC activations, tcgen05 kind::f16 (bf16 operands), fp32 accumulation in TMEM.
 *   FP16 : as BF16 with fp16 operands (what the reference runs on GPU: model.half(),
 *          CODON_X4/test.py:52).
 *   TF32 : fp32 NHWC activations, tcgen05 kind::tf32, fp32 accumulation.
 *   F16X3: fp32-accurate tensor-core mode.  Activations and weights are carried as two fp16 planes
 *          (hi = fp16(v), lo = fp16(v - hi): ~22 mantissa bits; weights pre-scaled by a power of two
 *          into the fp16 normal range) and every K step issues hi*hi + lo*hi + hi*lo as three
 *          tcgen05 kind::f16 MMAs into one fp32 TMEM accumulator.  Same 4 bytes per activation
 *          element as TF32, one third of the FP16 tensor rate, products as accurate as fp32.
 * In every mode the depth input, the global residual add and the output stay fp32, and the
 * CAC statistics / gates are computed in fp32. */
typedef enum codon_mode {
  CODON_MODE_FP32 = 0,
  CODON_MODE_BF16 = 1,
  CODON_MODE_FP16 = 2,
  CODON_MODE_TF32 = 3,
  CODON_MODE_F16X3 = 4
} codon_mode;

/* dtype of the frames passed to codon_forward */
typedef enum codon_dtype {
  CODON_DTYPE_F32 = 0,
  CODON_DTYPE_F16 = 1,
  CODON_DTYPE_BF16 = 2
} codon_dtype;

/* Replaces: CODONNet() + .cuda().half() + .eval()  (CODON_X4/test.py:48,52,67;
 * CODON_X16/test.py:51-55).  scale is 4, 8 or 16 and only selects the expected parameter
 * set (x4/x8 carry the unused attention_c5/attention_s5 keys, CODON_X4/CODON_x4.py:64-65). */
int codon_create(codon_ctx** out, int device, int scale, int mode);
void codon_destroy(codon_ctx* ctx);
const char* codon_last_error(const codon_ctx* ctx);   /* ctx may be NULL: last global error */
const char* codon_version(void);
/* Host-only consistency check of the weight packing and of the compile-time tap schedules of the convolution kernels
 * (no GPU needed).  Returns 0 when everything agrees, otherwise one bit per failed check. */
int codon_selftest(void);

/* Replaces: model.load_state_dict(...)  (CODON_X4/test.py:59, CODON_X16/test.py:60).
 * name is the reference state_dict key ("conv3.weight", "attention_c0.mlp.1.bias", ...; a
 * leading "module." from DataParallel is accepted); data is HOST fp32 in the PyTorch layout
 * (OIHW for convolutions, [out,in] for Linear).  Keys of the never-executed
 * attention_c5 / attention_s5 are accepted and ignored.  codon_finalize_weights re-lays the
 * weights out for the kernels (tap-major, pre-swizzled K-major slabs) and uploads them; it
 * fails if a parameter the forward needs is missing. */
int codon_set_weight(codon_ctx* ctx, const char* name, const float* data,
                     const int64_t* shape, int ndim);
int codon_finalize_weights(codon_ctx* ctx);
/* Number of successful codon_finalize_weights calls on this context.  A re-finalize re-uses the device
 * allocations of the previous one (same sizes), so pointers captured in a CUDA graph stay valid, but a graph also
 * holds per-layer constants derived from the weights: holders of a captured forward compare this counter and
 * re-capture when it has moved (codon_b200.engine.GraphedForward does). */
unsigned long long codon_weights_generation(const codon_ctx* ctx);
/* Replaces, for hosts without Python: torch.load('X4.pth') + load_state_dict (CODON_X4/test.py:56-59).  Reads a flat
 * weight file written by codon_b200.checkpoint.export_flat (the importer turns the reference's pickled-module .pth
 * into it), calls codon_set_weight for every tensor and codon_finalize_weights.  Format, little endian, unpadded:
 *   "CODONW1\0" | uint32 n | n x { uint16 name_len | name | uint8 ndim | ndim x int64 dims | prod(dims) x float32 } */
int codon_load_weights_file(codon_ctx* ctx, const char* path);

/* Bytes of device scratch codon_forward needs for B frames of H x W (activations live here
 * so that the memory stays owned by and visible to the caller's allocator). */
size_t codon_workspace_bytes(const codon_ctx* ctx, int B, int H, int W);

/* Replaces: out = model(input_pic, gray_pic)  (CODON_X4/test.py:125, CODON_X16/test.py:132;
 * CODONNet.forward, CODON_X4/CODON_x4.py:66-132).  depth, guide, out: DEVICE pointers to
 * B*H*W elements of io_dtype.  Asynchronous on cuda_stream (a cudaStream_t, may be NULL for
 * the legacy default stream).  A context is not re-entrant: one in-flight forward per
 * ctx/workspace. */
int codon_forward(codon_ctx* ctx, const void* depth, const void* guide, void* out,
                  int B, int H, int W, int io_dtype,
                  void* workspace, size_t workspace_bytes, void* cuda_stream);

/* End-to-end variant with HOST frames (what test.py does around the model call,
 * CODON_X4/test.py:122-128): copies depth/guide from host, runs the forward, copies the
 * result back and synchronises.  Frames are fp32.  The context keeps pinned staging buffers
 * and its own workspace for this entry point. */
int codon_forward_host(codon_ctx* ctx, const float* depth, const float* guide, float* out,
                       int B, int H, int W);

/* Streaming form of codon_forward_host for a sequence of calls (the per-image loop of
 * CODON_X4/test.py:109-145, a video, a batch queue): submit enqueues H2D copy -> forward -> D2H copy
 * and returns at once; wait blocks until the OLDEST outstanding submit has delivered `out`.  With one
 * call submitted ahead, the copies of neighbouring calls run on their own streams under the kernels of
 * the current one (two device-side I/O slots, one workspace; the forwards themselves stay serialised,
 * so results are bit-identical to codon_forward_host).  depth, guide and out must be page-locked host
 * memory (CODON_ERR_ARG otherwise) and must stay valid and untouched until the matching wait returns.
 * At most two submits may be outstanding (CODON_ERR_STATE on a third); codon_forward_host refuses to run
 * while any are outstanding. */
int codon_forward_host_submit(codon_ctx* ctx, const float* depth, const float* guide, float* out,
                              int B, int H, int W);
int codon_forward_host_wait(codon_ctx* ctx);

/* Number of kernels the last codon_forward on this context launched. */
int codon_last_launch_count(const codon_ctx* ctx);

/* Per-kernel-class timing of the launches codon_forward makes, with CUDA events recorded on the
 * forward's own stream (what bench.py's roofline is computed from).  Categories 0..7:
 * conv5x5_128to128, pair_3x3_5x5_64to128, conv3x3, conv1x1_128to64, edge_1to64_64to1, cac_stats,
 * cac_mlp, cac_apply.  codon_profile_read waits for the recorded events and returns the totals
 * accumulated since the last reset: device milliseconds, algorithmic work (FLOP for the convs,
 * HBM bytes for edge / CAC kernels) and the number of launches. */
int codon_profile_enable(codon_ctx* ctx, int on);
int codon_profile_read(codon_ctx* ctx, int category, double* total_ms, double* work, long long* launches);
int codon_profile_reset(codon_ctx* ctx);
const char* codon_profile_category_name(int category);

/* Copies an intermediate activation of the last forward to dst as fp32 NCHW [B,C,H,W]
 * (DEVICE pointer, C returned through channels).  Names: "enc" (128: depth|colour encoder
 * outputs), "feat" (128: depth|colour stage outputs after the last stage), "ms" (128: the
 * last multi-scale pair of the depth / fusion branch),
 * "fuse" (64), "out_fuse" (64).  For layer-level parity tests. */
int codon_debug_tap(codon_ctx* ctx, const char* name, float* dst, int* channels,
                    void* cuda_stream);

/* ---- stand-alone CAC / CBAM pieces (unit-testable; NCHW fp32 DEVICE tensors) -------------
 * Replaces CAC_channel.forward (CODON_X4/CAC_module.py:38-63): x [B,C,H,W] -> scale [B,C_out]
 * (the reference returns it expanded to [B,C_out,H,W]); w1 [hidden,C], b1 [hidden],
 * w2 [C_out,hidden], b2 [C_out] are DEVICE fp32.  pool_mask selects the pool_types summed into
 * the gate (:40-61): bit 0 'avg', bit 1 'max', bit 2 'lp', bit 3 'lse'; the default
 * ['avg','max'] is 3.  Also serves attention/ResCBAM.py:38-61 (ChannelGate) with C_out == C. */
#define CODON_POOL_AVG 1
#define CODON_POOL_MAX 2
#define CODON_POOL_LP 4
#define CODON_POOL_LSE 8
int codon_cac_channel(const float* x, int B, int C, int H, int W,
                      const float* w1, const float* b1, const float* w2, const float* b2,
                      int hidden, int c_out, int pool_mask, float* scale, void* cuda_stream);
/* Replaces CAC_spatial.forward (CODON_X4/CAC_module.py:90-94) and ChannelPool (:78-81):
 * x [B,C,H,W] -> scale [B,1,H,W]; w [1,2,5,5] DEVICE fp32 (channel 0 weights the max map,
 * channel 1 the mean map).  pooled (may be NULL) receives the [B,2,H,W] ChannelPool output. */
int codon_cac_spatial(const float* x, int B, int C, int H, int W, const float* w,
                      float* scale, float* pooled, void* cuda_stream);
/* Gate + residual (CODON_X4/CODON_x4.py:88-91,117-118): y = x * sc[b,c % c_gate] * ss[b,h,w]
 * (+ res if res != NULL); sc may be NULL (ones) and ss may be NULL (ones), which gives
 * ChannelGate / SpatialGate (attention/ResCBAM.py:61,87). */
int codon_cac_apply(const float* x, const float* sc, const float* ss, const float* res,
                    int B, int C, int H, int W, int c_gate, float* y, void* cuda_stream);

/* Per-plane pooled statistics behind the pool_types of CAC_channel / ChannelGate
 * (CAC_module.py:43,47,50-55,71-76): stats [4][B*C] DEVICE fp32 = mean, max, lp (p = 2), lse. */
int codon_channel_stats(const float* x, int B, int C, int H, int W, float* stats, void* cuda_stream);
/* ChannelPool (CAC_module.py:78-81): x [B,C,H,W] -> pooled [B,2,H,W] = (max over C, mean over C). */
int codon_channel_pool(const float* x, int B, int C, int H, int W, float* pooled, void* cuda_stream);
/* BasicConv (CAC_module.py:6-20): generic NCHW fp32 cross-correlation with optional bias (may be
 * NULL) and ReLU; w [Cout, Cin/groups, kh, kw]; y [B, Cout, OH, OW] with the usual output size. */
int codon_conv2d_nchw(const float* x, const float* w, const float* bias, int B, int Cin, int H, int W,
                      int Cout, int kh, int kw, int stride_h, int stride_w, int pad_h, int pad_w,
                      int dil_h, int dil_w, int groups, int relu, float* y, void* cuda_stream);

/* ---- driver post-processing and evaluation metrics on the GPU (DEVICE pointers) ---------------
 * Replaces np.clip(out,0,1); (out*255).astype(np.uint8)  (CODON_X4/test.py:130,132): n fp32
 * values -> uint8 by truncation.  via_half != 0 reproduces the reference's float16 evaluation
 * (the model output there is a float16 array, test.py:52,127-128). */
int codon_quantise_u8(const float* src, uint8_t* dst, size_t n, int via_half, void* cuda_stream);
/* Replaces EvaluationResults (CODON_X4/test.py:148-164): label, out uint8 [B,H,W] (label already
 * cropped to the output's size, :150); rmse[b] (DEVICE double) = RMSE in grey levels over the
 * pixels with label != 0. */
int codon_masked_rmse(const uint8_t* label, const uint8_t* out, int B, int H, int W, double* rmse,
                      void* cuda_stream);
/* Replaces ssim_exact(img1, img2, sd, C1, C2)  (CODON_X4/ssim_2.py:36-52): images [B,H,W] either
 * uint8 (img_dtype 0; scaled by 1/255 as the driver does, test.py:139) or float64 (img_dtype 1,
 * used as given); float64 arithmetic, scipy gaussian_filter semantics; ssim[b] DEVICE double.
 * workspace: DEVICE scratch of at least (5*H*W + H) * B * 8 bytes. */
int codon_ssim_gauss(const void* img1, const void* img2, int img_dtype, int B, int H, int W, double sd,
                     double c1, double c2, double* ssim, void* workspace, size_t workspace_bytes,
                     void* cuda_stream);

/* ---- one frame over several GPUs (SURVEY.md 8e "single-frame latency mode") ----------------------
 * A group owns one finalized context per GPU (same scale / mode / weights, distinct devices with peer
 * access).  codon_group_forward_host splits ONE H x W frame into horizontal bands, one per GPU; after every
 * layer each GPU pulls 2 halo rows from its neighbours' bands over NVLink peer memory, and the per-band CAC
 * channel statistics are all-gathered so that every GPU evaluates the same gates.  No NCCL.  depth / guide /
 * out are HOST fp32 [H, W].  The result equals the single-GPU forward up to the summation order of the CAC
 * average pool. */
typedef struct codon_group codon_group;
int codon_group_create(codon_group** out, codon_ctx** ctxs, int n);
void codon_group_destroy(codon_group* group);
int codon_group_forward_host(codon_group* group, const float* depth, const float* guide, float* out, int H, int W);
const char* codon_group_last_error(const codon_group* group);
double codon_group_last_ms(const codon_group* group);   /* wall-clock milliseconds of the last forward_host */

/* ---- driver pre-processing on the GPU (DEVICE pointers) ------------------------------------------
 * Replaces the colour -> gray conversion of cv2.imread(path, 0) (CODON_X4/test.py:118) for an already decoded
 * 8-bit BGR image [n_pixels, 3].  method 0: what imread(.., 0) yields for a colour PNG (libpng's
 * rgb_to_gray, (R*9797 + G*19234 + B*3737) >> 15); method 1: cv2.cvtColor BGR2GRAY
 * ((B*1868 + G*9617 + R*4899 + 8192) >> 14). */
int codon_bgr_to_gray_u8(const uint8_t* bgr, uint8_t* gray, size_t n_pixels, int method, void* cuda_stream);
/* Replaces torch.from_numpy(img / 255).float() (test.py:122-123): dst[i] = float32(double(src[i]) / 255). */
int codon_u8_to_unit_f32(const uint8_t* src, float* dst, size_t n, void* cuda_stream);
/* The pre-upsampling of the low-resolution depth that the reference leaves to an unshipped offline step
 * (test.py:77 "Bicubic/X4"): src [B,h,w] -> dst [B,H,W] fp32 with the semantics of
 * cv2.resize(src, (W, H), interpolation=cv2.INTER_CUBIC) (a = -0.75, half-pixel centres, replicated border). */
int codon_bicubic_upsample_f32(const float* src, float* dst, int B, int h, int w, int H, int W, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* CODON_B200_H */
