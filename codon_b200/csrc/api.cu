// libcodon_b200: C ABI (include/codon_b200.h) and the forward-pass launch plan.
//
// The plan is the B200 restatement of CODONNet.forward (CODON_X4/CODON_x4.py:66-132 ==
// CODON_X8/CODON_x8.py; CODON_X16/CODON_x16.py:136-202).  Activations are NHWC in a caller
// provided workspace; the torch.cat calls of the reference (:79,80,85,119,125) are replaced by
// producers writing into channel slices of shared buffers:
//
//   E    128 ch  [enc_d | enc_c]      encoder outputs = residual carriers of the 5 stages (:70,73)
//   F    128 ch  [out_d | out_c]      stage outputs; conv7 reads it as cat(out, out_c) (:119)
//   MS   2 x 128 ch  ms_d, ms_c       multi-scale pairs: depth [3x3|5x5] (:79), colour [5x5|3x3] (:80); one
//                                     128-channel buffer per branch (depth first, colour P*128 elements later)
//   R2   2 x 128 ch  r2_d, r2_c       conv3 / conv6 outputs (:81-82), same split
//   FUSE  64 ch  conv7 output = residual carrier of the fusion stages (:120-121)
//   OF    64 ch  out_fuse (:127-128)
//
// One launch handles the depth and the colour branch of a layer as two "jobs".
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>
#include <chrono>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/codon_b200.h"
#include "common.cuh"
#include "conv_tc.h"
#include "kernels.h"

using namespace codon;

namespace {

thread_local std::string g_last_error;

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
};

struct ConvSpec { const char* name; int cout, cin, ks; };
// CODON_X4/CODON_x4.py:24-47
const ConvSpec kTrunk[] = {
    {"input", 64, 1, 3},    {"conv_input", 64, 64, 3}, {"conv1", 64, 64, 3},   {"conv2", 64, 64, 5},
    {"conv3", 128, 128, 5}, {"confuse", 64, 128, 1},   {"input_c", 64, 1, 3},  {"conv_input_c", 64, 64, 3},
    {"conv4", 64, 64, 5},   {"conv5", 64, 64, 3},      {"conv6", 128, 128, 5}, {"confuse_c", 64, 128, 1},
    {"conv7", 64, 128, 3},  {"conv8", 64, 64, 5},      {"conv9", 64, 64, 3},   {"conv10", 128, 128, 5},
    {"confuse_fuse", 64, 128, 1}, {"conv11", 64, 64, 3}, {"output", 1, 64, 3}};

struct TcLayer {
  TcConvPlan plan;
  uint8_t* dev = nullptr;
  CUtensorMap bmap;            // 2-D view of the packed stream (2-CTA kernel)
};

struct Buffers {
  // byte offsets into the workspace
  size_t xf = 0, yf = 0, of32 = 0, E = 0, F = 0, MS = 0, R2 = 0, FUSE = 0, OF = 0, pooled = 0, part = 0, sc = 0;
  size_t gate = 0;                    // spatial gate map s_s [B,H,W] fp32 (written next to the CAC MLP, read by cac_apply)
  size_t cstat = 0, cstat_half = 0;   // per-cell channel partials of the two branches (fused conv epilogue), bytes of one
  int cells = 0;                      // 8 x 16-pixel cells per frame
  size_t total = 0;
  size_t px = 0;      // pixels the layout was planned for (B*H*W; the band mode plans every GPU for the tallest band)
  int chunks = 0;
};

}  // namespace

struct codon_ctx {
  int device = 0, scale = 4, mode = 0, act = ACT_F32;
  bool finalized = false;
  std::map<std::string, HostTensor> host_w;
  std::string err;
  int launches = 0;
  std::vector<void*> dev_allocs;
  // Weight uploads of the previous finalize, in upload order: a re-finalize (load_state_dict after a forward) copies
  // into the same allocations when the sizes match, so device pointers and tensor maps baked into a captured CUDA
  // graph stay valid.  `generation` counts successful finalizes (GraphedForward re-captures when it moves: the
  // kernel parameters of a graph also hold per-layer weight scales).
  std::vector<size_t> alloc_bytes;
  size_t upload_cursor = 0;
  unsigned long long generation = 0;

  // fp32 direct-conv weights [T][Cin][Cout]
  std::map<std::string, float*> w_direct;
  // tcgen05 packed weights
  std::map<std::string, TcLayer> w_tc;
  float *w_in_d = nullptr, *w_in_c = nullptr, *w_out = nullptr;   // [9][64]
  float *cac_w1[5] = {}, *cac_b1[5] = {}, *cac_w2[5] = {}, *cac_b2[5] = {}, *cac_ws[5] = {};

  // tensor-map cache (valid while workspace / shape unchanged)
  struct MapKey {
    const void* base; int C, box_w, box_h, B, H, W;
    bool operator<(const MapKey& o) const {
      return std::tie(base, C, box_w, box_h, B, H, W) < std::tie(o.base, o.C, o.box_w, o.box_h, o.B, o.H, o.W);
    }
  };
  std::map<MapKey, CUtensorMap> tmaps;

  // per-category CUDA-event profiling of the launches (bench.py roofline)
  bool prof_on = false;
  struct ProfRec { cudaEvent_t a, b; int cat; double work; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[8] = {}, prof_work[8] = {};
  long long prof_n[8] = {};

  // last forward (debug taps)
  uint8_t* last_ws = nullptr;
  Buffers last_buf;
  int last_B = 0, last_H = 0, last_W = 0;

  // codon_forward_host state
  cudaStream_t host_stream = nullptr;
  void* host_ws = nullptr; size_t host_ws_bytes = 0;
  float *pin_in = nullptr, *pin_out = nullptr; size_t pin_elems = 0;
  float *dev_x = nullptr, *dev_y = nullptr, *dev_o = nullptr; size_t dev_elems = 0;

  // codon_forward_host_submit / _wait: two device-side I/O slots, copy streams beside the compute stream
  struct HostSlot {
    float *x = nullptr, *y = nullptr, *o = nullptr; size_t elems = 0;
    cudaEvent_t in_ready = nullptr, computed = nullptr, out_ready = nullptr;
  } slot[2];
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  unsigned long long submitted = 0, waited = 0;
};

namespace {

int fail(codon_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  if (ctx) ctx->err = buf;
  return code;
}

#define CU_TRY(ctx, expr)                                                                          \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(ctx, CODON_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Kernel-variant switches for perf experiments (CODON_TC_2CTA, _2CTA_ALL, _FUSE, _CSTAT, _NACC_CONV, _NACC_PAIR) exist
// only in builds with -DCODON_TC_EXPERIMENT; the product library takes every decision from the frame geometry.
#ifdef CODON_TC_EXPERIMENT
int exp_knob(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
#else
inline int exp_knob(const char*, int dflt) { return dflt; }
#endif

enum ProfCat { PC_CONV5_128 = 0, PC_PAIR = 1, PC_CONV3 = 2, PC_CONV1 = 3, PC_EDGE = 4, PC_CAC_STATS = 5, PC_CAC_MLP = 6, PC_CAC_APPLY = 7 };
const char* const kProfNames[8] = {"conv5x5_128to128", "pair_3x3_5x5_64to128", "conv3x3", "conv1x1_128to64", "edge_1to64_64to1",
                                   "cac_stats", "cac_mlp", "cac_apply"};

// Scoped event pair around one launch (no-op unless profiling is enabled on the context).
struct ProfScope {
  codon_ctx* ctx; cudaStream_t st; cudaEvent_t b = nullptr;
  ProfScope(codon_ctx* c, int cat, double work, cudaStream_t s) : ctx(c), st(s) {
    if (!ctx->prof_on) return;
    cudaEvent_t ev[2];
    for (auto& e : ev) {
      if (ctx->prof_pool.empty()) { if (cudaEventCreate(&e) != cudaSuccess) return; }
      else { e = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); }
    }
    cudaEventRecord(ev[0], st);
    b = ev[1];
    ctx->prof_recs.push_back({ev[0], ev[1], cat, work});
  }
  ~ProfScope() { if (b) cudaEventRecord(b, st); }
};

int pick_nacc(int B, int H, int W, int njobs, int prefer, const char* env);
int use_two_cta(int B, int H, int W, int nacc);

// True when every 5x5 128->128 + 1x1 pair of the forward runs as the fused cluster kernel for this geometry (the
// decisions of Runner::conv5_fused for the two-job CAC stages and the one-job fusion stages): the 128-channel
// intermediate then never exists in HBM and the R2 region only has to hold the 128-channel encoder scratch.
bool fused_everywhere(const codon_ctx* ctx, int B, int H, int W) {
  if (ctx->mode == CODON_MODE_FP32 || !exp_knob("CODON_TC_FUSE", 1)) return false;
  if (ctx->mode == CODON_MODE_F16X3) return true;
  for (int njobs = 1; njobs <= 2; ++njobs) {
    const int nacc = pick_nacc(B, H, W, njobs, 2, "CODON_TC_NACC_CONV");
    if (nacc > 2 || !use_two_cta(B, H, W, nacc)) return false;
  }
  return true;
}

Buffers plan_buffers(const codon_ctx* ctx, int B, int H, int W, int part_chunks = 0) {
  Buffers b;
  const size_t P = (size_t)B * H * W, e = act_bytes(ctx->act);
  b.px = P;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  b.xf = take(P * 4); b.yf = take(P * 4); b.of32 = take(P * 4);
  b.E = take(P * 128 * e); b.F = take(P * 128 * e);
  // R2: conv3 / conv6 outputs of the un-fused path (2 x 128 channels); with the fused kernels only its first 128
  // channels are used (scratch of the encoder's first layer)
  b.MS = take(P * 256 * e); b.R2 = take(P * (fused_everywhere(ctx, B, H, W) ? 128 : 256) * e);
  b.FUSE = take(P * 64 * e); b.OF = take(P * 64 * e);
  b.pooled = take(P * 2 * 4 * 4);   // (max, mean) map, or 2 / 4 (max, sum) partial maps from the 1x1 epilogues
  b.chunks = cac_stats_chunks(B, H, W);
  b.cells = cdiv(W, kTcSubW) * cdiv(H, kTcSubH);
  // `part` holds the chunk partials of whichever statistics path runs: 256-pixel chunks (stand-alone pass) or folded
  // 8 x 16-pixel cells (fused conv epilogue) -- on narrow frames the cell chunks outnumber the pixel chunks
  int max_chunks = b.chunks > part_chunks ? b.chunks : part_chunks;
  if (cac_cell_chunks(b.cells) > max_chunks) max_chunks = cac_cell_chunks(b.cells);
  b.part = take((size_t)B * max_chunks * 256 * 4);
  b.sc = take((size_t)B * 64 * 4);
  b.gate = take(P * 4);             // spatial gate map s_s
  b.cstat_half = align_up((size_t)B * b.cells * 256 * sizeof(float2), 1024);
  b.cstat = take(2 * b.cstat_half);
  b.total = off + 1024;   // slack for aligning the caller's pointer
  return b;
}

template <typename T>
int upload(codon_ctx* ctx, const std::vector<T>& host, T** dev) {
  const size_t bytes = host.size() * sizeof(T);
  void* p = nullptr;
  if (ctx->upload_cursor < ctx->dev_allocs.size() && ctx->alloc_bytes[ctx->upload_cursor] == bytes) {
    p = ctx->dev_allocs[ctx->upload_cursor];          // same upload sequence as last time: reuse in place
  } else {
    // first finalize, or the sequence changed: drop this and every later allocation of the old sequence
    for (size_t i = ctx->upload_cursor; i < ctx->dev_allocs.size(); ++i) cudaFree(ctx->dev_allocs[i]);
    ctx->dev_allocs.resize(ctx->upload_cursor);
    ctx->alloc_bytes.resize(ctx->upload_cursor);
    CU_TRY(ctx, cudaMalloc(&p, bytes));
    ctx->dev_allocs.push_back(p);
    ctx->alloc_bytes.push_back(bytes);
  }
  ctx->upload_cursor++;
  CU_TRY(ctx, cudaMemcpy(p, host.data(), bytes, cudaMemcpyHostToDevice));
  *dev = static_cast<T*>(p);
  return CODON_OK;
}

const HostTensor* find_w(const codon_ctx* ctx, const std::string& key) {
  auto it = ctx->host_w.find(key);
  return it == ctx->host_w.end() ? nullptr : &it->second;
}

// OIHW -> [T][Cin][Cout]
std::vector<float> to_tap_major(const HostTensor& t, int cout, int cin, int ks) {
  std::vector<float> r((size_t)ks * ks * cin * cout);
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int k = 0; k < ks * ks; ++k)
        r[((size_t)k * cin + ci) * cout + co] = t.data[((size_t)co * cin + ci) * ks * ks + k];
  return r;
}

int get_tmap(codon_ctx* ctx, const void* base, int C, int box_w, int box_h, int slab_elems, int B, int H, int W,
             const CUtensorMap** out) {
  codon_ctx::MapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.C = C; key.box_w = box_w; key.box_h = box_h; key.B = B; key.H = H; key.W = W;
  auto it = ctx->tmaps.find(key);
  if (it == ctx->tmaps.end()) {
    if (ctx->tmaps.size() > 256) ctx->tmaps.clear();
    CUtensorMap m;
    CU_TRY(ctx, tc_encode_tmap(&m, base, ctx->act, C, W, H, B, slab_elems, box_w, box_h));
    it = ctx->tmaps.emplace(key, m).first;
  }
  *out = &it->second;
  return CODON_OK;
}

// Accumulators (128-pixel sub-tiles) per CTA tile.  `prefer` is the size that measured fastest for the
// layer class when the frame is large enough: 2 for the 5x5 layers in cluster mode (two TMEM buffers of
// 2 accumulators -> the epilogue of a tile overlaps the MMAs of the next one), 4 for the HBM-bound
// 1x1 / 3x3 layers (fewer, larger TMA boxes).  Small frames step down until there are >= 2 waves.
int pick_nacc(int B, int H, int W, int njobs, int prefer, const char* env) {
  if (env) {                      // perf experiments: CODON_TC_NACC_PAIR / CODON_TC_NACC_CONV
    const int v = exp_knob(env, 0);
    if (v == 1 || v == 2 || v == 4) return v;
  }
  for (int nacc = prefer; nacc > 1; nacc >>= 1) {
    const long tiles = (long)B * cdiv(W, tc_tile_w(nacc)) * cdiv(H, tc_tile_h(nacc)) * njobs;
    if (tiles >= 2 * 148) return nacc;
  }
  return 1;
}

// The cluster-of-2 kernel pays off once a job has at least a couple of waves of tile pairs.
int use_two_cta(int B, int H, int W, int nacc) {
  const int env = exp_knob("CODON_TC_2CTA", -1);
  if (env == 0 || env == 1) return env;
  const long tiles = (long)B * cdiv(W, tc_tile_w(nacc)) * cdiv(H, tc_tile_h(nacc));
  return tiles >= 148 ? 1 : 0;
}

// Row-band mode (codon_group_*): one frame is split into horizontal bands, one per GPU.  Each band keeps
// kBandHalo extra rows on its interior sides; after every layer the hook refreshes those rows from the
// neighbouring GPUs' core rows over NVLink peer memory, and the per-band CAC channel partials are
// all-gathered so that every GPU evaluates the same gate MLP.
constexpr int kBandHalo = 2;
struct BandHook {
  int cy0 = 0, cy1 = 0;             // core rows [cy0, cy1) of the local image
  long global_hw = 0;               // pixels of the whole frame (the average pool divides by it)
  int chunk_off = 0, chunks_total = 0;   // this band's slot in / the size of the gathered partial buffer
  int my_chunks = 0;                // chunk partials this band contributes (cell chunks on the fused path)
  virtual int halo(const size_t* off, const size_t* row_bytes, int n) = 0;
  // all-gather of the bands' CAC chunk partials; optionally refreshes the halo rows of `n` buffers in the same exchange
  virtual int gather_parts(size_t part_off, const size_t* off = nullptr, const size_t* row_bytes = nullptr, int n = 0) = 0;
  virtual ~BandHook() {}
};

// One conv layer of the plan: up to two jobs reading channel slices of `in` (in_C channels).
struct LayerJob { const char* w; int in_off; size_t out; int out_stride, out_off; size_t res; int res_stride, res_off; bool has_res; size_t pool = 0; bool has_pool = false; size_t in_add = 0; };

struct Runner {
  codon_ctx* ctx; uint8_t* ws; int B, H, W; cudaStream_t st; int e;
  int Hdec;          // height the kernel-variant decisions are based on (band mode: identical on every GPU)
  size_t px;         // planned pixels (Buffers::px): stride of the per-branch halves and of the pool partial maps

  int conv(const char* plan_name, size_t in, int in_C, int cin, int cout, int ks, bool relu,
           const LayerJob* jobs, int njobs) {
    const double flops = 2.0 * B * H * W * (double)cin * cout * ks * ks * njobs;
    if (ctx->mode == CODON_MODE_FP32) {
      ConvJob cj[2];
      for (int i = 0; i < njobs; ++i) {
        cj[i].in = ws + in + jobs[i].in_add; cj[i].in_stride = in_C; cj[i].in_off = jobs[i].in_off;
        cj[i].w = ctx->w_direct.at(jobs[i].w);
        cj[i].out = ws + jobs[i].out; cj[i].out_stride = jobs[i].out_stride; cj[i].out_off = jobs[i].out_off;
        cj[i].res = jobs[i].has_res ? ws + jobs[i].res : nullptr;
        cj[i].res_stride = jobs[i].res_stride; cj[i].res_off = jobs[i].res_off;
      }
      ProfScope ps(ctx, cat_of(cin, cout, ks), flops, st);
      CU_TRY(ctx, launch_conv_direct_f32(cj, njobs, B, H, W, cin, cout, ks, relu, st));
      ctx->launches++;
      return CODON_OK;
    }
    (void)plan_name;
    const TcLayer& l0 = ctx->w_tc.at(jobs[0].w);
    TcLaunch L;
    L.njobs = njobs; L.B = B; L.H = H; L.W = W; L.relu = relu; L.out_act = ctx->act;
    const bool split = ctx->mode == CODON_MODE_F16X3;
    // split-fp16 operands: one accumulator pair (big, small) per tile, promoted chunk by chunk (conv_tc.cu)
    L.nacc = split ? 1 : pick_nacc(B, Hdec, W, njobs, ks == 5 ? 2 : 4, "CODON_TC_NACC_CONV");
    if (split && ks != 3) return fail(ctx, CODON_ERR_STATE, "f16x3 mode runs the 5x5 / 1x1 layers through the fused kernels only");
    for (int i = 0; i < njobs; ++i) {
      const TcLayer& l = ctx->w_tc.at(jobs[i].w);
      L.job[i].in_coff = jobs[i].in_off;
      L.job[i].w = l.dev;
      L.job[i].descale = 1.f / l.plan.scale;
      L.job[i].out = ws + jobs[i].out; L.job[i].out_stride = jobs[i].out_stride; L.job[i].out_off = jobs[i].out_off;
      L.job[i].res = jobs[i].has_res ? ws + jobs[i].res : nullptr;
      L.job[i].res_stride = jobs[i].res_stride; L.job[i].res_off = jobs[i].res_off;
      L.job[i].outer_col = 0;
      L.job[i].pool = jobs[i].has_pool ? reinterpret_cast<float2*>(ws + jobs[i].pool) : nullptr;
      L.bmap[i] = &l.bmap;
    }
    // the HBM-bound 1x1 / 3x3 layers measured slower in cluster mode; the 5x5 layers gain 15-20 %
    {
      const int all = exp_knob("CODON_TC_2CTA_ALL", 0);
      // cluster mode: the 5x5 layers (+15-20 %) and, with one patch per slab, the 3x3 layers (+30 %); the
      // stand-alone 1x1 (fallback path only) stays single-CTA: its epilogue emits the ChannelPool partials
      L.two_cta = (ks == 5 || ks == 3 || (all && !jobs[0].has_pool)) ? use_two_cta(B, Hdec, W, L.nacc) : 0;
      if (split) L.two_cta = 1;     // the split path exists in the cluster kernel only (any frame size)
    }
    const CUtensorMap* tm[2] = {nullptr, nullptr};
    for (int i = 0; i < njobs; ++i) {
      int rc = get_tmap(ctx, ws + in + jobs[i].in_add, in_C, tc_box_w(l0.plan, L.nacc), tc_box_h(l0.plan, L.nacc),
                        l0.plan.slab_elems, B, H, W, &tm[i]);
      if (rc) return rc;
    }
    ProfScope ps(ctx, cat_of(cin, cout, ks), flops, st);
    CU_TRY(ctx, launch_conv_tc(*tm[0], *tm[njobs - 1], l0.plan, L, st));
    ctx->launches++;
    return CODON_OK;
  }

  // conv5x5 128->128 (+ReLU) immediately followed by conv1x1 128->64 (no ReLU, optional residual) as one
  // cluster kernel (conv_tc.cu, FUSE).  Returns 1 if the fused path is not applicable (caller falls back).
  struct FusedJob { const char* w5; const char* w1; size_t in_add; size_t out2; int out2_stride, out2_off;
                    size_t res2; int res2_stride, res2_off; bool has_res; size_t pool; bool has_pool;
                    size_t cstat = 0; bool has_cstat = false; };
  int core_y0 = 0, core_y1 = 0;      // row-band mode: the rows of the local image this GPU owns (cstat masking)
  int conv5_fused(size_t in, const FusedJob* jobs, int njobs) {
    const int env = exp_knob("CODON_TC_FUSE", 1);
    if (ctx->mode == CODON_MODE_FP32 || !env) return 1;
    const bool split = ctx->mode == CODON_MODE_F16X3;
    // split-fp16 operands: one accumulator per tile (four patch stages + the two-plane Y tile fill shared memory),
    // always through this kernel
    const int nacc = split ? 1 : pick_nacc(B, Hdec, W, njobs, 2, "CODON_TC_NACC_CONV");
    if (!split && (nacc > 2 || !use_two_cta(B, Hdec, W, nacc))) return 1;
    const TcLayer& l0 = ctx->w_tc.at(jobs[0].w5);
    TcLaunch L;
    L.njobs = njobs; L.B = B; L.H = H; L.W = W; L.relu = 1; L.out_act = ctx->act;
    L.nacc = nacc; L.two_cta = 1; L.fuse = 1; L.pool_stride = px;
    L.core_y0 = core_y0; L.core_y1 = core_y1;
    L.y16_operand = ctx->mode == CODON_MODE_BF16 ? TC_BF16 : TC_F16;
    const CUtensorMap* tm[2] = {nullptr, nullptr};
    for (int i = 0; i < njobs; ++i) {
      const TcLayer& l = ctx->w_tc.at(jobs[i].w5);
      const TcLayer& l1 = ctx->w_tc.at(std::string(jobs[i].w1) + "@y16");
      L.job[i].in_coff = 0;
      L.job[i].w = l.dev;
      L.job[i].descale = 1.f / l.plan.scale; L.job[i].descale2 = 1.f / l1.plan.scale;
      L.job[i].out = nullptr; L.job[i].out_stride = 0; L.job[i].out_off = 0;
      L.job[i].res = nullptr; L.job[i].res_stride = 0; L.job[i].res_off = 0;
      L.job[i].outer_col = 0;
      L.job[i].out2 = ws + jobs[i].out2; L.job[i].out2_stride = jobs[i].out2_stride; L.job[i].out2_off = jobs[i].out2_off;
      L.job[i].res2 = jobs[i].has_res ? ws + jobs[i].res2 : nullptr;
      L.job[i].res2_stride = jobs[i].res2_stride; L.job[i].res2_off = jobs[i].res2_off;
      L.job[i].pool = jobs[i].has_pool ? reinterpret_cast<float2*>(ws + jobs[i].pool) : nullptr;
      L.job[i].cstat = jobs[i].has_cstat ? reinterpret_cast<float2*>(ws + jobs[i].cstat) : nullptr;
      L.bmap[i] = &l.bmap;
      L.wmap[i] = &l1.bmap;
      int rc = get_tmap(ctx, ws + in + jobs[i].in_add, 128, tc_box_w(l0.plan, nacc), tc_box_h(l0.plan, nacc),
                        l0.plan.slab_elems, B, H, W, &tm[i]);
      if (rc) return rc;
    }
    const double px = (double)B * H * W;
    {
      ProfScope ps(ctx, PC_CONV5_128, 2.0 * px * 128.0 * 128.0 * 25.0 * njobs, st);
      CU_TRY(ctx, launch_conv_tc(*tm[0], *tm[njobs - 1], l0.plan, L, st));
    }
    // the fused 1x1 is booked with zero time: its FLOP stay in the trunk total, its launches are gone
    ctx->launches++;
    return CODON_OK;
  }

  static int cat_of(int cin, int cout, int ks) {
    if (ks == 5 && cin == 128) return PC_CONV5_128;
    if (ks == 1) return PC_CONV1;
    if (ks == 5) return PC_PAIR;      // fp32 mode runs the pair as separate 3x3 / 5x5 launches
    return PC_CONV3;
  }

  // 3x3 || 5x5 multi-scale pair on a 64-channel input slice -> 128 output channels at out_off.
  // three_first[i]: job i writes [3x3 | 5x5] (depth branch) else [5x5 | 3x3].
  int pair(size_t in, int in_C, const int* in_off, const char* const* w3, const char* const* w5,
           const char* const* wpair, const bool* three_first, size_t out, int out_stride, const int* out_off,
           const size_t* out_add, int njobs) {
    if (ctx->mode == CODON_MODE_FP32) {
      LayerJob j3[2], j5[2];
      for (int i = 0; i < njobs; ++i) {
        j3[i] = {w3[i], in_off[i], out + out_add[i], out_stride, out_off[i] + (three_first[i] ? 0 : 64), 0, 0, 0, false};
        j5[i] = {w5[i], in_off[i], out + out_add[i], out_stride, out_off[i] + (three_first[i] ? 64 : 0), 0, 0, 0, false};
      }
      int rc = conv("", in, in_C, 64, 64, 3, true, j3, njobs);
      if (rc) return rc;
      return conv("", in, in_C, 64, 64, 5, true, j5, njobs);
    }
    const TcLayer& l0 = ctx->w_tc.at(wpair[0]);
    TcLaunch L;
    L.njobs = njobs; L.B = B; L.H = H; L.W = W; L.relu = 1; L.out_act = ctx->act;
    const bool split = ctx->mode == CODON_MODE_F16X3;
    L.nacc = split ? 1 : pick_nacc(B, Hdec, W, njobs, 2, "CODON_TC_NACC_PAIR");
    for (int i = 0; i < njobs; ++i) {
      const TcLayer& l = ctx->w_tc.at(wpair[i]);
      L.job[i].in_coff = in_off[i];
      L.job[i].w = l.dev;
      L.job[i].descale = 1.f / l.plan.scale;
      L.job[i].out = ws + out + out_add[i]; L.job[i].out_stride = out_stride; L.job[i].out_off = out_off[i];
      L.job[i].res = nullptr; L.job[i].res_stride = 0; L.job[i].res_off = 0;
      L.job[i].outer_col = three_first[i] ? 64 : 0;
      L.job[i].pool = nullptr;
      L.bmap[i] = &l.bmap;
    }
    L.two_cta = split ? 1 : use_two_cta(B, Hdec, W, L.nacc);
    const CUtensorMap* tm = nullptr;
    int rc = get_tmap(ctx, ws + in, in_C, tc_box_w(l0.plan, L.nacc), tc_box_h(l0.plan, L.nacc), l0.plan.slab_elems, B, H, W, &tm);
    if (rc) return rc;
    ProfScope ps(ctx, PC_PAIR, 2.0 * B * H * W * 64.0 * 64.0 * 34.0 * njobs, st);
    CU_TRY(ctx, launch_conv_tc(*tm, *tm, l0.plan, L, st));
    ctx->launches++;
    return CODON_OK;
  }
};

int run_forward(codon_ctx* ctx, const float* x, const float* y, float* out, int B, int H, int W, uint8_t* ws,
                const Buffers& bf, cudaStream_t st, BandHook* hook = nullptr) {
  Runner r{ctx, ws, B, H, W, st, act_bytes(ctx->act), hook ? (int)(bf.px / (size_t)W) : H, bf.px};
  if (hook) { r.core_y0 = hook->cy0; r.core_y1 = hook->cy1; }
  // band mode: refresh the halo rows of the buffers a layer wrote (no-op otherwise)
  auto halo = [&](std::initializer_list<size_t> offs, int channels_or_bytes, bool raw_bytes = false) -> int {
    if (!hook) return CODON_OK;
    size_t o[4], rb[4];
    int n = 0;
    for (size_t v : offs) { o[n] = v; rb[n] = raw_bytes ? (size_t)W * channels_or_bytes : (size_t)W * channels_or_bytes * r.e; ++n; }
    return hook->halo(o, rb, n);
  };
  const bool tc_mode = ctx->mode != CODON_MODE_FP32;
  // MS and R2 hold the depth branch in their first half and the colour branch in the second (128 channels
  // each, pixel-contiguous per branch: a branch's 256/512-byte pixel rows are read and written whole)
  const size_t half = bf.px * 128 * r.e;
  int rc;
  // encoders (CODON_x4.py:68-73): input/input_c 1->64 (+ReLU) into the R2 region viewed as 128 ch
  const size_t T0 = bf.R2;
  const double P = (double)B * H * W;
  {
    ProfScope ps(ctx, PC_EDGE, P * (8 + 128 * r.e), st);
    CU_TRY(ctx, launch_conv_first(x, y, ctx->w_in_d, ctx->w_in_c, ws + T0, ctx->act, B, H, W, st, ctx->mode == CODON_MODE_TF32));
  }
  ctx->launches++;
  if ((rc = halo({T0}, 128))) return rc;
  {
    LayerJob j[2] = {{"conv_input", 0, bf.E, 128, 0, 0, 0, 0, false}, {"conv_input_c", 64, bf.E, 128, 64, 0, 0, 0, false}};
    if ((rc = r.conv("", T0, 128, 64, 64, 3, true, j, 2))) return rc;
  }
  if ((rc = halo({bf.E}, 128))) return rc;
  // five multi-scale + CAC stages (:74-118)
  for (int s = 0; s < 5; ++s) {
    const size_t src = s == 0 ? bf.E : bf.F;
    {
      const int in_off[2] = {0, 64}, out_off[2] = {0, 0};
      const size_t out_add[2] = {0, half};
      const char* w3[2] = {"conv1", "conv5"}; const char* w5[2] = {"conv2", "conv4"};
      const char* wp[2] = {"pair_d", "pair_c"};
      const bool tf[2] = {true, false};
      if ((rc = r.pair(src, 128, in_off, w3, w5, wp, tf, bf.MS, 128, out_off, out_add, 2))) return rc;
    }
    if ((rc = halo({bf.MS, bf.MS + half}, 128))) return rc;
    const size_t pmap = bf.px * 8;                 // bytes of one per-pixel float2 partial map
    int pool_parts = 1;
    bool fj_cstat = false;
    {
      Runner::FusedJob fj[2] = {{"conv3", "confuse", 0, bf.F, 128, 0, 0, 0, 0, false, bf.pooled, true},
                                {"conv6", "confuse_c", half, bf.F, 128, 64, 0, 0, 0, false, bf.pooled + 2 * pmap, true}};
      const int use_cstat = exp_knob("CODON_TC_CSTAT", 1);   // 0: perf experiments, stand-alone statistics pass instead
      if (use_cstat) {   // the epilogue also leaves the per-cell channel partials of the global pools (band mode: core rows only)
        fj[0].cstat = bf.cstat; fj[0].has_cstat = true;
        fj[1].cstat = bf.cstat + bf.cstat_half; fj[1].has_cstat = true;
        fj_cstat = true;
      }
      rc = r.conv5_fused(bf.MS, fj, 2);
      if (rc < 0) return rc;
      if (rc == 0) pool_parts = 4;
    }
    if (pool_parts == 1) {
      {
        LayerJob j[2] = {{"conv3", 0, bf.R2, 128, 0, 0, 0, 0, false}, {"conv6", 0, bf.R2 + half, 128, 0, 0, 0, 0, false}};
        j[1].in_add = half;
        if ((rc = r.conv("", bf.MS, 128, 128, 128, 5, true, j, 2))) return rc;
      }
      if ((rc = halo({bf.R2, bf.R2 + half}, 128))) return rc;
      LayerJob j[2] = {{"confuse", 0, bf.F, 128, 0, 0, 0, 0, false}, {"confuse_c", 0, bf.F, 128, 64, 0, 0, 0, false}};
      j[1].in_add = half;
      if (tc_mode) {   // the 1x1 epilogues emit the per-branch ChannelPool partials (max, sum) per pixel
        j[0].pool = bf.pooled; j[0].has_pool = true;
        j[1].pool = bf.pooled + pmap; j[1].has_pool = true;
        pool_parts = 2;
      }
      if ((rc = r.conv("", bf.R2, 128, 128, 64, 1, false, j, 2))) return rc;
    }
    // CAC gates (:85-118, CAC_module.py:38-63, 78-94)
    float* pooled = reinterpret_cast<float*>(ws + bf.pooled);
    float* part = reinterpret_cast<float*>(ws + bf.part);
    float* sc = reinterpret_cast<float*>(ws + bf.sc);
    float* gate = reinterpret_cast<float*>(ws + bf.gate);
    const CacGate cg = {pooled, ctx->cac_ws[s], gate, H, W, pool_parts, bf.px};
    // algorithmic HBM bytes: stats reads F (128e B/px); apply reads F and E, writes F (384e B/px)
    if (!hook) {
      int chunks = bf.chunks;
      if (pool_parts == 4 && fj_cstat) {
        // fused conv path: fold the epilogue's per-cell partials (32 B per pixel) instead of re-reading F
        chunks = cac_cell_chunks(bf.cells);
        // + the gate map computed by the same launch: pool_parts partial maps of 8 B/px in, 4 B/px out
        ProfScope ps(ctx, PC_CAC_STATS, P * (32 + pool_parts * 8 + 4), st);
        CU_TRY(ctx, launch_cac_cell_reduce(ws + bf.cstat, ws + bf.cstat + bf.cstat_half, B, bf.cells, part, chunks, st, &cg));
      } else {
        ProfScope ps(ctx, PC_CAC_STATS, P * (128 * r.e + pool_parts * 8 + 4), st);
        if (tc_mode) CU_TRY(ctx, launch_cac_chan_stats(ws + bf.F, ctx->act, B, H, W, part, bf.chunks, st));
        else CU_TRY(ctx, launch_cac_stats(ws + bf.F, ctx->act, B, H, W, pooled, part, bf.chunks, st));
        CU_TRY(ctx, launch_cac_gate(cg, B, st));
        ctx->launches++;
      }
      ProfScope ps(ctx, PC_CAC_MLP, 0.0, st);
      CU_TRY(ctx, launch_cac_mlp(part, chunks, B, H * W, ctx->cac_w1[s], ctx->cac_b1[s], ctx->cac_w2[s],
                                 ctx->cac_b2[s], sc, st));
    } else {
      // band mode (B == 1): statistics over this band's core rows only, into its slot of the gathered buffer;
      // the ChannelPool maps of the halo rows come from the neighbours; every GPU then reduces ALL bands'
      // partials in the same fixed order, with the whole frame's pixel count
      const int core_h = hook->cy1 - hook->cy0;
      const size_t row0 = (size_t)hook->cy0 * W;
      float* my_part = part + (size_t)hook->chunk_off * 256;
      const int my_chunks = hook->my_chunks;
      if (tc_mode) {
        if (pool_parts == 4 && fj_cstat) {
          // fused path: the epilogue's per-cell partials already count the core rows only
          const int cells = cdiv(W, kTcSubW) * cdiv(H, kTcSubH);
          if (my_chunks != cac_cell_chunks(cells)) return fail(ctx, CODON_ERR_STATE, "band mode: chunk plan does not match the fused path");
          CU_TRY(ctx, launch_cac_cell_reduce(ws + bf.cstat, ws + bf.cstat + bf.cstat_half, 1, cells, my_part, my_chunks, st));
        } else {
          if (my_chunks != cac_stats_chunks(1, core_h, W)) return fail(ctx, CODON_ERR_STATE, "band mode: chunk plan does not match the stand-alone statistics pass");
          CU_TRY(ctx, launch_cac_chan_stats(ws + bf.F + row0 * 128 * r.e, ctx->act, 1, core_h, W, my_part, my_chunks, st));
        }
        // the ChannelPool maps of the halo rows travel in the same exchange as the channel partials
        size_t o[4]; for (int k = 0; k < pool_parts; ++k) o[k] = bf.pooled + k * pmap;
        size_t rb[4] = {(size_t)W * 8, (size_t)W * 8, (size_t)W * 8, (size_t)W * 8};
        if ((rc = hook->gather_parts(bf.part, o, rb, pool_parts))) return rc;
      } else {
        if (my_chunks != cac_stats_chunks(1, core_h, W)) return fail(ctx, CODON_ERR_STATE, "band mode: chunk plan does not match the statistics pass");
        CU_TRY(ctx, launch_cac_stats(ws + bf.F + row0 * 128 * r.e, ctx->act, 1, core_h, W, pooled + row0 * 2, my_part, my_chunks, st));
        if ((rc = halo({bf.pooled}, 8, true))) return rc;
        if ((rc = hook->gather_parts(bf.part))) return rc;
      }
      // the gate map needs the neighbours' ChannelPool halo rows: after the exchange
      CU_TRY(ctx, launch_cac_gate(cg, 1, st));
      ctx->launches++;
      CU_TRY(ctx, launch_cac_mlp(part, hook->chunks_total, 1, (int)hook->global_hw, ctx->cac_w1[s], ctx->cac_b1[s],
                                 ctx->cac_w2[s], ctx->cac_b2[s], sc, st));
    }
    {
      ProfScope ps(ctx, PC_CAC_APPLY, P * 384 * r.e, st);
      CU_TRY(ctx, launch_cac_apply(ws + bf.F, ws + bf.E, ctx->act, gate, sc, B, H, W, st, ctx->mode == CODON_MODE_TF32));
    }
    ctx->launches += 3;
    if ((rc = halo({bf.F}, 128))) return rc;
  }
  // fusion head (:119-121): cat(out, out_c) is F itself
  {
    LayerJob j[1] = {{"conv7", 0, bf.FUSE, 64, 0, 0, 0, 0, false}};
    if ((rc = r.conv("", bf.F, 128, 128, 64, 3, true, j, 1))) return rc;
  }
  if ((rc = halo({bf.FUSE}, 64))) return rc;
  // three fusion stages (:122-128)
  for (int k = 0; k < 3; ++k) {
    const size_t src = k == 0 ? bf.FUSE : bf.OF;
    {
      const int in_off[1] = {0}, out_off[1] = {0};
      const size_t out_add[1] = {0};
      const char* w3[1] = {"conv9"}; const char* w5[1] = {"conv8"}; const char* wp[1] = {"pair_f"};
      const bool tf[1] = {false};
      if ((rc = r.pair(src, 64, in_off, w3, w5, wp, tf, bf.MS, 128, out_off, out_add, 1))) return rc;
    }
    if ((rc = halo({bf.MS}, 128))) return rc;
    {
      Runner::FusedJob fj[1] = {{"conv10", "confuse_fuse", 0, bf.OF, 64, 0, bf.FUSE, 64, 0, true, 0, false}};
      rc = r.conv5_fused(bf.MS, fj, 1);
      if (rc < 0) return rc;
    }
    if (rc == 1) {
      {
        LayerJob j[1] = {{"conv10", 0, bf.R2, 128, 0, 0, 0, 0, false}};
        if ((rc = r.conv("", bf.MS, 128, 128, 128, 5, true, j, 1))) return rc;
      }
      if ((rc = halo({bf.R2}, 128))) return rc;
      LayerJob j[1] = {{"confuse_fuse", 0, bf.OF, 64, 0, bf.FUSE, 64, 0, true}};
      if ((rc = r.conv("", bf.R2, 128, 128, 64, 1, false, j, 1))) return rc;
    }
    if ((rc = halo({bf.OF}, 64))) return rc;
  }
  // reconstruction (:129-131): conv11 into the MS region viewed as 64 ch, then output + x
  {
    LayerJob j[1] = {{"conv11", 0, bf.MS, 64, 0, 0, 0, 0, false}};
    if ((rc = r.conv("", bf.OF, 64, 64, 64, 3, true, j, 1))) return rc;
  }
  if ((rc = halo({bf.MS}, 64))) return rc;
  {
    ProfScope ps(ctx, PC_EDGE, P * (8 + 64 * r.e), st);
    CU_TRY(ctx, launch_conv_last(ws + bf.MS, 64, ctx->act, ctx->w_out, x, out, B, H, W, st));
  }
  ctx->launches++;
  return CODON_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// codon_group: one frame over several GPUs of one process (row bands + NVLink peer-memory halo exchange)
struct codon_group {
  int n = 0;
  std::vector<codon_ctx*> ctx;
  std::vector<int> dev;
  std::vector<cudaStream_t> st;
  std::vector<uint8_t*> ws;          // per GPU workspace (identical layout on every GPU)
  std::vector<float*> din, dout;     // per GPU band inputs (x | y) and output
  std::vector<float*> pin;           // per GPU pinned staging (x | y | out)
  size_t cap_px = 0, ws_cap = 0;     // capacity (pixels of the tallest band, workspace bytes)
  // Cross-GPU ordering without host rendezvous: flags[g] is an array of n counters in GPU g's memory; counter h is
  // advanced by GPU h (a store through NVLink peer memory from h's exchange kernel, stream-ordered behind h's layer) and
  // polled by GPU g's exchange kernel before it pulls rows from h.  seq_out[g][h] / seq_in[g][h] are the host-side
  // mirrors of "how often g has signalled h" / "how many signals g has consumed from h"; every GPU enqueues the same
  // exchange sequence, so the two sides count alike.
  std::vector<uint32_t*> flags;
  std::vector<std::vector<uint32_t>> seq_out, seq_in;
  bool poisoned = false;             // a forward failed: streams are drained and the counters reset before the next one
  // geometry of the current forward
  int H = 0, W = 0;
  std::vector<int> r0, r1, mt, mb, hloc, chunk_off, chunks;
  int chunks_total = 0;
  Buffers bf;
  std::atomic<bool> failed{false};
  std::mutex err_mu;
  std::string err;
  double last_ms = 0.0;

  void set_error(const std::string& m) {
    {
      std::lock_guard<std::mutex> lk(err_mu);
      if (err.empty()) err = m;
    }
    if (!failed.exchange(true, std::memory_order_acq_rel)) {
      // release every exchange kernel that may be polling a counter this thread will never advance: all counters far
      // ahead (cyclic comparison: 0x7f7f7f7f - want >= 0); the results of this forward are discarded anyway
      for (int h = 0; h < n; ++h) {
        if (!flags.empty() && flags[h] && cudaSetDevice(dev[h]) == cudaSuccess) cudaMemset(flags[h], 0x7F, (size_t)n * sizeof(uint32_t));
      }
    }
  }
};

namespace {

// One exchange = one kernel on the exchanging GPU's stream (stream-ordered behind the layer whose rows it publishes):
//   1. signal: one thread stores this GPU's new counter value into each peer's flag array (system-scope fence first:
//      the layer's results are visible before the counter is);
//   2. wait:   every CTA polls this GPU's own flag array until each peer's counter has reached the expected value
//      (acquire loads; a watchdog traps after ~2 s instead of hanging the GPU on a protocol bug);
//   3. pull:   rows are copied out of peer memory with plain 16-byte loads over NVLink.
// No DMA peer copies: the runtime may serialise those against other work of the source device, and every GPU must be
// free to run ahead of its neighbours here; no host thread ever waits for another one.
constexpr int kPullMax = 24, kPeerMax = 16;
struct ExchangeArgs {
  const uint8_t* src[kPullMax];
  uint8_t* dst[kPullMax];
  uint32_t bytes[kPullMax];      // multiples of 8 (rows of float2 maps, or of >= 64-channel pixels)
  int nseg;
  uint32_t* peer_flag[kPeerMax]; // where to publish my counter (peer memory)
  uint32_t peer_val[kPeerMax];
  const uint32_t* my_flag[kPeerMax];   // where the peers publish theirs (my memory)
  uint32_t want[kPeerMax];
  int npeer;
  void add(const void* s, void* d, size_t n) { src[nseg] = static_cast<const uint8_t*>(s); dst[nseg] = static_cast<uint8_t*>(d); bytes[nseg] = (uint32_t)n; ++nseg; }
};
__global__ void __launch_bounds__(256) exchange_kernel(const ExchangeArgs a) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && (int)threadIdx.x < a.npeer) {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(a.peer_flag[threadIdx.x]) = a.peer_val[threadIdx.x];
  }
  if ((int)threadIdx.x < a.npeer) {
    const uint32_t* f = a.my_flag[threadIdx.x];
    const uint32_t want = a.want[threadIdx.x];
    const long long t0 = clock64();
    while (true) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if ((int32_t)(v - want) >= 0) break;          // cyclic comparison
      if (clock64() - t0 > 4000000000LL) {
        printf("codon_group exchange: peer counter %d stuck at %u, expected %u\n", (int)threadIdx.x, v, want);
        __trap();
      }
    }
  }
  __syncthreads();
  for (int sgm = blockIdx.y; sgm < a.nseg; sgm += gridDim.y) {
    const uint8_t* s = a.src[sgm];
    uint8_t* d = a.dst[sgm];
    const uint32_t n = a.bytes[sgm];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if ((((uintptr_t)s | (uintptr_t)d | n) & 15) == 0) {
      for (uint32_t i = tid; i < (n >> 4); i += nth) reinterpret_cast<uint4*>(d)[i] = reinterpret_cast<const uint4*>(s)[i];
    } else {
      for (uint32_t i = tid; i < (n >> 3); i += nth) reinterpret_cast<uint2*>(d)[i] = reinterpret_cast<const uint2*>(s)[i];
    }
  }
}

struct GroupHook : BandHook {
  codon_group* G; int g;
  GroupHook(codon_group* G_, int g_) : G(G_), g(g_) {}

  // fills the signal / wait part for the given peers (advances the host-side mirrors of the counters) and launches
  int launch(ExchangeArgs& a, const int* peers, int npeers) {
    if (G->failed.load(std::memory_order_acquire)) return CODON_ERR_CUDA;
    a.npeer = npeers;
    for (int i = 0; i < npeers; ++i) {
      const int h = peers[i];
      a.peer_flag[i] = G->flags[h] + g;
      a.peer_val[i] = ++G->seq_out[g][h];
      a.my_flag[i] = G->flags[g] + h;
      a.want[i] = ++G->seq_in[g][h];
    }
    uint32_t mx = 0;
    for (int i = 0; i < a.nseg; ++i) mx = a.bytes[i] > mx ? a.bytes[i] : mx;
    int bx = (int)(((mx >> 4) + 255) / 256);
    if (bx > 32) bx = 32;
    if (bx < 1) bx = 1;
    exchange_kernel<<<dim3(bx, a.nseg > 0 ? a.nseg : 1), 256, 0, G->st[g]>>>(a);
    if (cudaGetLastError() != cudaSuccess) { G->set_error("exchange_kernel launch failed"); return CODON_ERR_CUDA; }
    return CODON_OK;
  }

  int add_halo(ExchangeArgs& a, const size_t* off, const size_t* rb, int nbuf) {
    if (a.nseg + nbuf * 2 > kPullMax) { G->set_error("halo: too many buffers"); return CODON_ERR_STATE; }
    for (int i = 0; i < nbuf; ++i)
      if ((rb[i] & 7) || (off[i] & 7)) { G->set_error("halo rows are not 8-byte granular"); return CODON_ERR_STATE; }
    auto add = [&](int h, size_t src_row, size_t dst_row) {
      for (int i = 0; i < nbuf; ++i)
        a.add(G->ws[h] + off[i] + src_row * rb[i], G->ws[g] + off[i] + dst_row * rb[i], (size_t)kBandHalo * rb[i]);
    };
    if (g > 0)            // my top halo rows <- the last core rows of the band above
      add(g - 1, (size_t)(G->hloc[g - 1] - G->mb[g - 1] - kBandHalo), 0);
    if (g < G->n - 1)     // my bottom halo rows <- the first core rows of the band below
      add(g + 1, (size_t)G->mt[g + 1], (size_t)(G->hloc[g] - G->mb[g]));
    return CODON_OK;
  }

  int halo(const size_t* off, const size_t* rb, int nbuf) override {
    int peers[2], np = 0;
    if (g > 0) peers[np++] = g - 1;
    if (g < G->n - 1) peers[np++] = g + 1;
    ExchangeArgs a;
    a.nseg = 0;
    if (add_halo(a, off, rb, nbuf)) return CODON_ERR_STATE;
    return launch(a, peers, np);
  }

  int gather_parts(size_t part_off, const size_t* off, const size_t* rb, int nbuf) override {
    int peers[kPeerMax], np = 0;
    ExchangeArgs a;
    a.nseg = 0;
    if (nbuf > 0 && add_halo(a, off, rb, nbuf)) return CODON_ERR_STATE;
    if (a.nseg + G->n - 1 > kPullMax) { G->set_error("gather: too many segments"); return CODON_ERR_STATE; }
    for (int h = 0; h < G->n; ++h) {
      if (h == g) continue;
      peers[np++] = h;
      const size_t o = part_off + (size_t)G->chunk_off[h] * 256 * sizeof(float);
      a.add(G->ws[h] + o, G->ws[g] + o, (size_t)G->chunks[h] * 256 * sizeof(float));
    }
    return launch(a, peers, np);
  }
};

void group_free_buffers(codon_group* G) {
  for (int g = 0; g < G->n; ++g) {
    cudaSetDevice(G->dev[g]);
    if (G->ws[g]) cudaFree(G->ws[g]);
    if (G->din[g]) cudaFree(G->din[g]);
    if (G->dout[g]) cudaFree(G->dout[g]);
    if (G->pin[g]) cudaFreeHost(G->pin[g]);
    G->ws[g] = nullptr; G->din[g] = nullptr; G->dout[g] = nullptr; G->pin[g] = nullptr;
  }
  G->cap_px = 0; G->ws_cap = 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" {

const char* codon_version(void) { return "codon_b200 0.1 (sm_100a)"; }

int codon_selftest(void) { return tc_selftest(); }

const char* codon_last_error(const codon_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int codon_create(codon_ctx** out, int device, int scale, int mode) {
  if (!out) return fail(nullptr, CODON_ERR_ARG, "codon_create: out is NULL");
  *out = nullptr;
  if (scale != 4 && scale != 8 && scale != 16) return fail(nullptr, CODON_ERR_ARG, "codon_create: scale must be 4, 8 or 16 (got %d)", scale);
  if (mode < CODON_MODE_FP32 || mode > CODON_MODE_F16X3) return fail(nullptr, CODON_ERR_ARG, "codon_create: unknown mode %d", mode);
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, CODON_ERR_CUDA, "codon_create: no CUDA device (%s); there is no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= n) return fail(nullptr, CODON_ERR_ARG, "codon_create: device %d out of range [0,%d)", device, n);
  cudaDeviceProp prop;
  CU_TRY(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(nullptr, CODON_ERR_CUDA, "codon_create: device %d is sm_%d%d; this library is built for sm_100a only",
                device, prop.major, prop.minor);
  CU_TRY(nullptr, cudaSetDevice(device));
  codon_ctx* c = new codon_ctx();
  c->device = device; c->scale = scale; c->mode = mode;
  c->act = mode == CODON_MODE_BF16 ? ACT_BF16 : mode == CODON_MODE_FP16 ? ACT_F16 : mode == CODON_MODE_F16X3 ? ACT_SPLIT16 : ACT_F32;
  *out = c;
  return CODON_OK;
}

void codon_destroy(codon_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (void* p : ctx->dev_allocs) cudaFree(p);
  if (ctx->host_ws) cudaFree(ctx->host_ws);
  if (ctx->dev_x) cudaFree(ctx->dev_x);
  if (ctx->dev_y) cudaFree(ctx->dev_y);
  if (ctx->dev_o) cudaFree(ctx->dev_o);
  if (ctx->pin_in) cudaFreeHost(ctx->pin_in);
  if (ctx->pin_out) cudaFreeHost(ctx->pin_out);
  for (auto& sl : ctx->slot) {
    for (float* p : {sl.x, sl.y, sl.o}) if (p) cudaFree(p);
    for (cudaEvent_t e : {sl.in_ready, sl.computed, sl.out_ready}) if (e) cudaEventDestroy(e);
  }
  if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
  if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
  if (ctx->host_stream) cudaStreamDestroy(ctx->host_stream);
  for (auto& r : ctx->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
  delete ctx;
}

int codon_set_weight(codon_ctx* ctx, const char* name, const float* data, const int64_t* shape, int ndim) {
  if (!ctx || !name || !data || !shape || ndim < 1 || ndim > 4) return fail(ctx, CODON_ERR_ARG, "codon_set_weight: bad argument");
  std::string key(name);
  if (key.rfind("module.", 0) == 0) key = key.substr(7);   // DataParallel prefix (CODON_X16/test.py:52,60)
  // validate against the parameter inventory
  std::vector<int64_t> shp(shape, shape + ndim), want;
  const std::string base = key.substr(0, key.find('.'));
  bool known = false;
  for (const ConvSpec& s : kTrunk)
    if (key == std::string(s.name) + ".weight") { want = {s.cout, s.cin, s.ks, s.ks}; known = true; }
  if (!known && base.rfind("attention_c", 0) == 0 && base.size() == 12) {
    const int k = base[11] - '0';
    const int C = k == 5 ? 64 : 128, hid = k == 5 ? 4 : 8;
    if (k >= 0 && k <= 5) {
      const std::string rest = key.substr(base.size());
      if (rest == ".mlp.1.weight") { want = {hid, C}; known = true; }
      else if (rest == ".mlp.1.bias") { want = {hid}; known = true; }
      else if (rest == ".mlp.3.weight") { want = {64, hid}; known = true; }
      else if (rest == ".mlp.3.bias") { want = {64}; known = true; }
    }
  }
  if (!known && base.rfind("attention_s", 0) == 0 && base.size() == 12 && base[11] >= '0' && base[11] <= '5' &&
      key.substr(base.size()) == ".spatial.conv.weight") { want = {1, 2, 5, 5}; known = true; }
  if (!known) return fail(ctx, CODON_ERR_ARG, "codon_set_weight: unexpected key '%s'", name);
  if (shp != want) return fail(ctx, CODON_ERR_ARG, "codon_set_weight: shape mismatch for '%s'", name);
  size_t n = 1;
  for (int64_t d : shp) n *= (size_t)d;
  HostTensor t;
  t.shape = shp;
  t.data.assign(data, data + n);
  ctx->host_w[key] = std::move(t);
  ctx->finalized = false;
  return CODON_OK;
}

int codon_finalize_weights(codon_ctx* ctx) {
  if (!ctx) return fail(nullptr, CODON_ERR_ARG, "codon_finalize_weights: ctx is NULL");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  for (const ConvSpec& s : kTrunk)
    if (!find_w(ctx, std::string(s.name) + ".weight"))
      return fail(ctx, CODON_ERR_STATE, "codon_finalize_weights: missing parameter '%s.weight'", s.name);
  for (int k = 0; k < 5; ++k)
    for (const char* suf : {"c%d.mlp.1.weight", "c%d.mlp.1.bias", "c%d.mlp.3.weight", "c%d.mlp.3.bias", "s%d.spatial.conv.weight"}) {
      char buf[64], key[96];
      snprintf(buf, sizeof(buf), suf, k);
      snprintf(key, sizeof(key), "attention_%s", buf);
      if (!find_w(ctx, key)) return fail(ctx, CODON_ERR_STATE, "codon_finalize_weights: missing parameter '%s'", key);
    }
  // earlier uploads are overwritten in place (same sizes in the same order); nothing in flight may read them
  CU_TRY(ctx, cudaDeviceSynchronize());
  ctx->upload_cursor = 0;
  ctx->w_direct.clear(); ctx->w_tc.clear();

  int rc;
  auto W = [&](const char* n) -> const HostTensor& { return *find_w(ctx, std::string(n) + ".weight"); };
  // edge layers: [9][64]
  {
    std::vector<float> a(576), b(576), c(576);
    for (int t = 0; t < 9; ++t)
      for (int ch = 0; ch < 64; ++ch) {
        a[t * 64 + ch] = W("input").data[ch * 9 + t];
        b[t * 64 + ch] = W("input_c").data[ch * 9 + t];
        c[t * 64 + ch] = W("output").data[ch * 9 + t];
      }
    if ((rc = upload(ctx, a, &ctx->w_in_d)) || (rc = upload(ctx, b, &ctx->w_in_c)) || (rc = upload(ctx, c, &ctx->w_out))) return rc;
  }
  for (int k = 0; k < 5; ++k) {
    char key[96];
    auto G = [&](const char* fmt) -> const std::vector<float>& { snprintf(key, sizeof(key), fmt, k); return find_w(ctx, key)->data; };
    if ((rc = upload(ctx, G("attention_c%d.mlp.1.weight"), &ctx->cac_w1[k])) ||
        (rc = upload(ctx, G("attention_c%d.mlp.1.bias"), &ctx->cac_b1[k])) ||
        (rc = upload(ctx, G("attention_c%d.mlp.3.weight"), &ctx->cac_w2[k])) ||
        (rc = upload(ctx, G("attention_c%d.mlp.3.bias"), &ctx->cac_b2[k])) ||
        (rc = upload(ctx, G("attention_s%d.spatial.conv.weight"), &ctx->cac_ws[k]))) return rc;
  }
  if (ctx->mode == CODON_MODE_FP32) {
    for (const ConvSpec& s : kTrunk) {
      if (s.cin == 1 || s.cout == 1) continue;
      float* d = nullptr;
      if ((rc = upload(ctx, to_tap_major(W(s.name), s.cout, s.cin, s.ks), &d))) return rc;
      ctx->w_direct[s.name] = d;
    }
  } else {
    const int operand = ctx->mode == CODON_MODE_BF16 ? TC_BF16 : ctx->mode == CODON_MODE_FP16 ? TC_F16
                        : ctx->mode == CODON_MODE_F16X3 ? TC_SPLIT16 : TC_TF32;
    // split-fp16 operands: per-layer power-of-two weight scale (TcConvPlan::scale), undone in the epilogue
    auto scale_of = [&](const char* a, const char* b = nullptr) {
      if (operand != TC_SPLIT16) return 1.f;
      const std::vector<float>& wa = W(a).data;
      return b ? tc_pick_scale(wa.data(), wa.size(), W(b).data.data(), W(b).data.size()) : tc_pick_scale(wa.data(), wa.size());
    };
    std::vector<uint8_t> packed;
    for (const ConvSpec& s : kTrunk) {
      if (s.cin == 1 || s.cout == 1) continue;
      TcLayer l;
      l.plan = tc_make_plan(s.ks, s.cin, s.cout, operand, scale_of(s.name));
      tc_pack_weights(l.plan, W(s.name).data.data(), packed);
      if ((rc = upload(ctx, packed, &l.dev))) return rc;
      CU_TRY(ctx, tc_encode_bmap(&l.bmap, l.dev, packed.size()));
      ctx->w_tc[s.name] = l;
    }
    // 16-bit copies of the 1x1 weights for the fused conv5+1x1 kernel (fp16 in tf32 mode: same mantissa width)
    const int y16 = operand == TC_BF16 ? TC_BF16 : operand == TC_SPLIT16 ? TC_SPLIT16 : TC_F16;
    for (const char* n : {"confuse", "confuse_c", "confuse_fuse"}) {
      TcLayer l;
      l.plan = tc_make_plan(1, 128, 64, y16, scale_of(n));
      tc_pack_weights(l.plan, W(n).data.data(), packed);
      if ((rc = upload(ctx, packed, &l.dev))) return rc;
      CU_TRY(ctx, tc_encode_bmap(&l.bmap, l.dev, packed.size()));
      ctx->w_tc[std::string(n) + "@y16"] = l;
    }
    struct PairSpec { const char* name; const char* w3; const char* w5; bool three_first; };
    // depth [3x3|5x5] (:75,77,79); colour [5x5|3x3] (:76,78,80); fusion [5x5|3x3] (:123-125)
    for (const PairSpec& p : {PairSpec{"pair_d", "conv1", "conv2", true}, PairSpec{"pair_c", "conv5", "conv4", false},
                              PairSpec{"pair_f", "conv9", "conv8", false}}) {
      TcLayer l;
      l.plan = tc_make_pair_plan(64, operand, scale_of(p.w3, p.w5));
      tc_pack_pair_weights(l.plan, W(p.w3).data.data(), W(p.w5).data.data(), p.three_first, packed);
      if ((rc = upload(ctx, packed, &l.dev))) return rc;
      CU_TRY(ctx, tc_encode_bmap(&l.bmap, l.dev, packed.size()));
      ctx->w_tc[p.name] = l;
    }
  }
  // allocations of a longer previous sequence (never happens for one mode; kept for safety)
  for (size_t i = ctx->upload_cursor; i < ctx->dev_allocs.size(); ++i) cudaFree(ctx->dev_allocs[i]);
  ctx->dev_allocs.resize(ctx->upload_cursor);
  ctx->alloc_bytes.resize(ctx->upload_cursor);
  ctx->finalized = true;
  ctx->generation++;
  return CODON_OK;
}

unsigned long long codon_weights_generation(const codon_ctx* ctx) { return ctx ? ctx->generation : 0; }

// Flat weight file (codon_b200.checkpoint.export_flat): "CODONW1\0", uint32 count, then per tensor
// uint16 name length, name bytes, uint8 ndim, ndim x int64 dims, prod(dims) x float32 -- little endian, no padding.
int codon_load_weights_file(codon_ctx* ctx, const char* path) {
  if (!ctx || !path) return fail(ctx, CODON_ERR_ARG, "codon_load_weights_file: bad argument");
  FILE* f = fopen(path, "rb");
  if (!f) return fail(ctx, CODON_ERR_ARG, "codon_load_weights_file: cannot open '%s'", path);
  struct Closer { FILE* f; ~Closer() { fclose(f); } } closer{f};
  char magic[8];
  uint32_t count = 0;
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "CODONW1\0", 8) != 0 || fread(&count, 4, 1, f) != 1 || count > 4096)
    return fail(ctx, CODON_ERR_ARG, "codon_load_weights_file: '%s' is not a CODONW1 weight file", path);
  std::vector<float> data;
  for (uint32_t i = 0; i < count; ++i) {
    uint16_t nlen = 0;
    uint8_t ndim = 0;
    char name[256];
    int64_t dims[4];
    if (fread(&nlen, 2, 1, f) != 1 || nlen == 0 || nlen >= sizeof(name) || fread(name, 1, nlen, f) != nlen ||
        fread(&ndim, 1, 1, f) != 1 || ndim < 1 || ndim > 4 || fread(dims, 8, ndim, f) != ndim)
      return fail(ctx, CODON_ERR_ARG, "codon_load_weights_file: truncated or malformed entry %u", i);
    name[nlen] = 0;
    size_t n = 1;
    for (int d = 0; d < ndim; ++d) {
      if (dims[d] < 1 || dims[d] > (1 << 20)) return fail(ctx, CODON_ERR_ARG, "codon_load_weights_file: bad shape for '%s'", name);
      n *= (size_t)dims[d];
    }
    if (n > ((size_t)1 << 26)) return fail(ctx, CODON_ERR_ARG, "codon_load_weights_file: '%s' is too large", name);
    data.resize(n);
    if (fread(data.data(), 4, n, f) != n) return fail(ctx, CODON_ERR_ARG, "codon_load_weights_file: truncated data of '%s'", name);
    int rc = codon_set_weight(ctx, name, data.data(), dims, ndim);
    if (rc) return rc;
  }
  return codon_finalize_weights(ctx);
}

size_t codon_workspace_bytes(const codon_ctx* ctx, int B, int H, int W) {
  if (!ctx || B < 1 || H < 1 || W < 1) return 0;
  return plan_buffers(ctx, B, H, W).total;
}

int codon_forward(codon_ctx* ctx, const void* depth, const void* guide, void* out, int B, int H, int W,
                  int io_dtype, void* workspace, size_t workspace_bytes, void* cuda_stream) {
  if (!ctx) return fail(nullptr, CODON_ERR_ARG, "codon_forward: ctx is NULL");
  if (!ctx->finalized) return fail(ctx, CODON_ERR_STATE, "codon_forward: weights not finalized");
  if (!depth || !guide || !out || !workspace) return fail(ctx, CODON_ERR_ARG, "codon_forward: NULL pointer");
  if (B < 1 || H < 1 || W < 1) return fail(ctx, CODON_ERR_ARG, "codon_forward: bad shape %dx%dx%d", B, H, W);
  if ((size_t)B * H * W >= (1u << 30)) return fail(ctx, CODON_ERR_ARG, "codon_forward: more than 2^30 pixels per call");
  if (io_dtype < CODON_DTYPE_F32 || io_dtype > CODON_DTYPE_BF16) return fail(ctx, CODON_ERR_ARG, "codon_forward: unknown io_dtype %d", io_dtype);
  const Buffers bf = plan_buffers(ctx, B, H, W);
  if (workspace_bytes < bf.total)
    return fail(ctx, CODON_ERR_WORKSPACE, "codon_forward: workspace %zu B < required %zu B", workspace_bytes, bf.total);
  for (const void* p : {depth, guide, (const void*)out, (const void*)workspace}) {
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess || at.type != cudaMemoryTypeDevice) {
      cudaGetLastError();
      return fail(ctx, CODON_ERR_ARG, "codon_forward: pointer %p is not device memory (no CPU path)", p);
    }
  }
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 1024));
  const size_t P = (size_t)B * H * W;
  ctx->launches = 0;
  const float *x = static_cast<const float*>(depth), *y = static_cast<const float*>(guide);
  float* o = static_cast<float*>(out);
  if (io_dtype != CODON_DTYPE_F32) {
    float* xf = reinterpret_cast<float*>(ws + bf.xf);
    float* yf = reinterpret_cast<float*>(ws + bf.yf);
    CU_TRY(ctx, launch_convert_to_f32(depth, io_dtype, xf, P, st));
    CU_TRY(ctx, launch_convert_to_f32(guide, io_dtype, yf, P, st));
    ctx->launches += 2;
    x = xf; y = yf; o = reinterpret_cast<float*>(ws + bf.of32);
  }
  int rc = run_forward(ctx, x, y, o, B, H, W, ws, bf, st);
  if (rc) return rc;
  if (io_dtype != CODON_DTYPE_F32) {
    CU_TRY(ctx, launch_convert_from_f32(o, out, io_dtype, P, st));
    ctx->launches++;
  }
  ctx->last_ws = ws; ctx->last_buf = bf; ctx->last_B = B; ctx->last_H = H; ctx->last_W = W;
  return CODON_OK;
}

int codon_forward_host(codon_ctx* ctx, const float* depth, const float* guide, float* out, int B, int H, int W) {
  if (!ctx) return fail(nullptr, CODON_ERR_ARG, "codon_forward_host: ctx is NULL");
  if (!depth || !guide || !out) return fail(ctx, CODON_ERR_ARG, "codon_forward_host: NULL pointer");
  if (B < 1 || H < 1 || W < 1) return fail(ctx, CODON_ERR_ARG, "codon_forward_host: bad shape");
  if (ctx->submitted != ctx->waited)
    return fail(ctx, CODON_ERR_STATE, "codon_forward_host: %llu submitted call(s) not waited for", ctx->submitted - ctx->waited);
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t P = (size_t)B * H * W;
  if (!ctx->host_stream) CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->host_stream, cudaStreamNonBlocking));
  // Buffers the caller already page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory) are used directly;
  // pageable ones are staged through the context's pinned buffers (one extra host memcpy each way).
  auto is_pinned = [](const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
  };
  const bool pin_d = is_pinned(depth), pin_g = is_pinned(guide), pin_o = is_pinned(out);
  if ((!pin_d || !pin_g || !pin_o) && ctx->pin_elems < P) {
    if (ctx->pin_in) cudaFreeHost(ctx->pin_in);
    if (ctx->pin_out) cudaFreeHost(ctx->pin_out);
    ctx->pin_in = ctx->pin_out = nullptr; ctx->pin_elems = 0;
    CU_TRY(ctx, cudaMallocHost(reinterpret_cast<void**>(&ctx->pin_in), 2 * P * sizeof(float)));
    CU_TRY(ctx, cudaMallocHost(reinterpret_cast<void**>(&ctx->pin_out), P * sizeof(float)));
    ctx->pin_elems = P;
  }
  if (ctx->dev_elems < P) {
    for (float** p : {&ctx->dev_x, &ctx->dev_y, &ctx->dev_o}) { if (*p) cudaFree(*p); *p = nullptr; }
    ctx->dev_elems = 0;
    CU_TRY(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->dev_x), P * sizeof(float)));
    CU_TRY(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->dev_y), P * sizeof(float)));
    CU_TRY(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->dev_o), P * sizeof(float)));
    ctx->dev_elems = P;
  }
  const size_t need = codon_workspace_bytes(ctx, B, H, W);
  if (ctx->host_ws_bytes < need) {
    if (ctx->host_ws) cudaFree(ctx->host_ws);
    ctx->host_ws = nullptr; ctx->host_ws_bytes = 0;
    CU_TRY(ctx, cudaMalloc(&ctx->host_ws, need));
    ctx->host_ws_bytes = need;
  }
  cudaStream_t st = ctx->host_stream;
  const float* src_d = depth;
  const float* src_g = guide;
  if (!pin_d) { memcpy(ctx->pin_in, depth, P * sizeof(float)); src_d = ctx->pin_in; }
  if (!pin_g) { memcpy(ctx->pin_in + P, guide, P * sizeof(float)); src_g = ctx->pin_in + P; }
  CU_TRY(ctx, cudaMemcpyAsync(ctx->dev_x, src_d, P * sizeof(float), cudaMemcpyHostToDevice, st));
  CU_TRY(ctx, cudaMemcpyAsync(ctx->dev_y, src_g, P * sizeof(float), cudaMemcpyHostToDevice, st));
  int rc = codon_forward(ctx, ctx->dev_x, ctx->dev_y, ctx->dev_o, B, H, W, CODON_DTYPE_F32, ctx->host_ws,
                         ctx->host_ws_bytes, st);
  if (rc) return rc;
  CU_TRY(ctx, cudaMemcpyAsync(pin_o ? out : ctx->pin_out, ctx->dev_o, P * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU_TRY(ctx, cudaStreamSynchronize(st));
  if (!pin_o) memcpy(out, ctx->pin_out, P * sizeof(float));
  return CODON_OK;
}

int codon_forward_host_submit(codon_ctx* ctx, const float* depth, const float* guide, float* out, int B, int H, int W) {
  if (!ctx) return fail(nullptr, CODON_ERR_ARG, "codon_forward_host_submit: ctx is NULL");
  if (!depth || !guide || !out) return fail(ctx, CODON_ERR_ARG, "codon_forward_host_submit: NULL pointer");
  if (B < 1 || H < 1 || W < 1) return fail(ctx, CODON_ERR_ARG, "codon_forward_host_submit: bad shape");
  if (ctx->submitted - ctx->waited >= 2)
    return fail(ctx, CODON_ERR_STATE, "codon_forward_host_submit: two calls already in flight (call codon_forward_host_wait)");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  for (const void* p : {(const void*)depth, (const void*)guide, (const void*)out}) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess || a.type != cudaMemoryTypeHost) {
      cudaGetLastError();
      return fail(ctx, CODON_ERR_ARG, "codon_forward_host_submit: %p is not page-locked host memory (asynchronous copies need "
                                      "cudaHostAlloc / cudaHostRegister / pin_memory buffers; codon_forward_host stages pageable ones)", p);
    }
  }
  const size_t P = (size_t)B * H * W;
  if (!ctx->host_stream) CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->host_stream, cudaStreamNonBlocking));
  if (!ctx->h2d_stream) CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
  if (!ctx->d2h_stream) CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
  codon_ctx::HostSlot& sl = ctx->slot[ctx->submitted & 1];
  // the slot's previous call (submitted - 2) has been waited for (at most two in flight): its buffers are idle
  if (!sl.in_ready) {
    for (cudaEvent_t* e : {&sl.in_ready, &sl.computed, &sl.out_ready}) CU_TRY(ctx, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }
  if (sl.elems < P) {
    for (float** p : {&sl.x, &sl.y, &sl.o}) { if (*p) cudaFree(*p); *p = nullptr; }
    sl.elems = 0;
    for (float** p : {&sl.x, &sl.y, &sl.o}) CU_TRY(ctx, cudaMalloc(reinterpret_cast<void**>(p), P * sizeof(float)));
    sl.elems = P;
  }
  const size_t need = codon_workspace_bytes(ctx, B, H, W);
  if (ctx->host_ws_bytes < need) {
    CU_TRY(ctx, cudaStreamSynchronize(ctx->host_stream));   // the call in flight still computes in the old workspace
    if (ctx->host_ws) cudaFree(ctx->host_ws);
    ctx->host_ws = nullptr; ctx->host_ws_bytes = 0;
    CU_TRY(ctx, cudaMalloc(&ctx->host_ws, need));
    ctx->host_ws_bytes = need;
  }
  // H2D on the copy-in stream (overlaps the kernels of the call before), kernels on the compute stream, D2H on the
  // copy-out stream (overlaps the kernels of the call after); events order the three per call
  CU_TRY(ctx, cudaMemcpyAsync(sl.x, depth, P * sizeof(float), cudaMemcpyHostToDevice, ctx->h2d_stream));
  CU_TRY(ctx, cudaMemcpyAsync(sl.y, guide, P * sizeof(float), cudaMemcpyHostToDevice, ctx->h2d_stream));
  CU_TRY(ctx, cudaEventRecord(sl.in_ready, ctx->h2d_stream));
  CU_TRY(ctx, cudaStreamWaitEvent(ctx->host_stream, sl.in_ready, 0));
  int rc = codon_forward(ctx, sl.x, sl.y, sl.o, B, H, W, CODON_DTYPE_F32, ctx->host_ws, ctx->host_ws_bytes, ctx->host_stream);
  if (rc) return rc;
  CU_TRY(ctx, cudaEventRecord(sl.computed, ctx->host_stream));
  CU_TRY(ctx, cudaStreamWaitEvent(ctx->d2h_stream, sl.computed, 0));
  CU_TRY(ctx, cudaMemcpyAsync(out, sl.o, P * sizeof(float), cudaMemcpyDeviceToHost, ctx->d2h_stream));
  CU_TRY(ctx, cudaEventRecord(sl.out_ready, ctx->d2h_stream));
  ctx->submitted++;
  return CODON_OK;
}

int codon_forward_host_wait(codon_ctx* ctx) {
  if (!ctx) return fail(nullptr, CODON_ERR_ARG, "codon_forward_host_wait: ctx is NULL");
  if (ctx->submitted == ctx->waited) return fail(ctx, CODON_ERR_STATE, "codon_forward_host_wait: nothing submitted");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  CU_TRY(ctx, cudaEventSynchronize(ctx->slot[ctx->waited & 1].out_ready));
  ctx->waited++;
  return CODON_OK;
}

int codon_last_launch_count(const codon_ctx* ctx) { return ctx ? ctx->launches : 0; }

int codon_profile_enable(codon_ctx* ctx, int on) {
  if (!ctx) return fail(nullptr, CODON_ERR_ARG, "codon_profile_enable: ctx is NULL");
  ctx->prof_on = on != 0;
  return CODON_OK;
}

static int prof_collect(codon_ctx* ctx) {
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  for (auto& r : ctx->prof_recs) {
    CU_TRY(ctx, cudaEventSynchronize(r.b));
    float ms = 0.f;
    CU_TRY(ctx, cudaEventElapsedTime(&ms, r.a, r.b));
    ctx->prof_ms[r.cat] += ms; ctx->prof_work[r.cat] += r.work; ctx->prof_n[r.cat]++;
    ctx->prof_pool.push_back(r.a); ctx->prof_pool.push_back(r.b);
  }
  ctx->prof_recs.clear();
  return CODON_OK;
}

int codon_profile_read(codon_ctx* ctx, int category, double* total_ms, double* work, long long* launches) {
  if (!ctx || category < 0 || category >= 8) return fail(ctx, CODON_ERR_ARG, "codon_profile_read: bad argument");
  int rc = prof_collect(ctx);
  if (rc) return rc;
  if (total_ms) *total_ms = ctx->prof_ms[category];
  if (work) *work = ctx->prof_work[category];
  if (launches) *launches = ctx->prof_n[category];
  return CODON_OK;
}

int codon_profile_reset(codon_ctx* ctx) {
  if (!ctx) return fail(nullptr, CODON_ERR_ARG, "codon_profile_reset: ctx is NULL");
  int rc = prof_collect(ctx);
  if (rc) return rc;
  for (int i = 0; i < 8; ++i) { ctx->prof_ms[i] = 0; ctx->prof_work[i] = 0; ctx->prof_n[i] = 0; }
  return CODON_OK;
}

const char* codon_profile_category_name(int category) { return category >= 0 && category < 8 ? kProfNames[category] : ""; }

int codon_debug_tap(codon_ctx* ctx, const char* name, float* dst, int* channels, void* cuda_stream) {
  if (!ctx || !name || !dst || !channels) return fail(ctx, CODON_ERR_ARG, "codon_debug_tap: bad argument");
  if (!ctx->last_ws) return fail(ctx, CODON_ERR_STATE, "codon_debug_tap: no forward has run");
  const Buffers& bf = ctx->last_buf;
  size_t off; int C, stride;
  const std::string n(name);
  if (n == "enc") { off = bf.E; C = 128; stride = 128; }
  else if (n == "feat") { off = bf.F; C = 128; stride = 128; }
  else if (n == "ms") { off = bf.MS; C = 128; stride = 128; }
  else if (n == "fuse") { off = bf.FUSE; C = 64; stride = 64; }
  else if (n == "out_fuse") { off = bf.OF; C = 64; stride = 64; }
  else return fail(ctx, CODON_ERR_ARG, "codon_debug_tap: unknown tap '%s'", name);
  *channels = C;
  CU_TRY(ctx, launch_nhwc_to_nchw_f32(ctx->last_ws + off, ctx->act, stride, 0, C, ctx->last_B,
                                      ctx->last_H * ctx->last_W, dst, static_cast<cudaStream_t>(cuda_stream)));
  return CODON_OK;
}

// ---- stand-alone pieces --------------------------------------------------------------------------
int codon_cac_channel(const float* x, int B, int C, int H, int W, const float* w1, const float* b1,
                      const float* w2, const float* b2, int hidden, int c_out, int pool_mask, float* scale,
                      void* cuda_stream) {
  if (!x || !w1 || !b1 || !w2 || !b2 || !scale || B < 1 || C < 1 || H < 1 || W < 1 || hidden < 1 || c_out < 1 ||
      pool_mask < 1 || pool_mask > 15)
    return fail(nullptr, CODON_ERR_ARG, "codon_cac_channel: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  float* tmp = nullptr;
  CU_TRY(nullptr, cudaMallocAsync(reinterpret_cast<void**>(&tmp), (size_t)4 * B * C * sizeof(float), st));
  // the temporary is released on every path (stream-ordered free), also when a launch fails
  cudaError_t e = launch_nchw_channel_stats(x, B, C, H * W, pool_mask, tmp, st);
  if (e == cudaSuccess) e = launch_gate_mlp(tmp, B, C, pool_mask, w1, b1, w2, b2, hidden, c_out, scale, st);
  const cudaError_t ef = cudaFreeAsync(tmp, st);
  if (e != cudaSuccess) return fail(nullptr, CODON_ERR_CUDA, "codon_cac_channel: %s", cudaGetErrorString(e));
  CU_TRY(nullptr, ef);
  return CODON_OK;
}

int codon_cac_spatial(const float* x, int B, int C, int H, int W, const float* w, float* scale, float* pooled,
                      void* cuda_stream) {
  if (!x || !w || !scale || B < 1 || C < 1 || H < 1 || W < 1) return fail(nullptr, CODON_ERR_ARG, "codon_cac_spatial: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  float* tmp = pooled;
  if (!tmp) CU_TRY(nullptr, cudaMallocAsync(reinterpret_cast<void**>(&tmp), (size_t)2 * B * H * W * sizeof(float), st));
  cudaError_t e = launch_nchw_channel_pool(x, B, C, H * W, tmp, st);
  if (e == cudaSuccess) e = launch_nchw_spatial_scale(tmp, w, B, H, W, scale, st);
  const cudaError_t ef = pooled ? cudaSuccess : cudaFreeAsync(tmp, st);
  if (e != cudaSuccess) return fail(nullptr, CODON_ERR_CUDA, "codon_cac_spatial: %s", cudaGetErrorString(e));
  CU_TRY(nullptr, ef);
  return CODON_OK;
}

int codon_cac_apply(const float* x, const float* sc, const float* ss, const float* res, int B, int C, int H, int W,
                    int c_gate, float* y, void* cuda_stream) {
  if (!x || !y || B < 1 || C < 1 || H < 1 || W < 1 || (sc && c_gate < 1)) return fail(nullptr, CODON_ERR_ARG, "codon_cac_apply: bad argument");
  CU_TRY(nullptr, launch_nchw_apply(x, sc, ss, res, B, C, H * W, c_gate > 0 ? c_gate : 1, y, static_cast<cudaStream_t>(cuda_stream)));
  return CODON_OK;
}

int codon_channel_stats(const float* x, int B, int C, int H, int W, float* stats, void* cuda_stream) {
  if (!x || !stats || B < 1 || C < 1 || H < 1 || W < 1) return fail(nullptr, CODON_ERR_ARG, "codon_channel_stats: bad argument");
  CU_TRY(nullptr, launch_nchw_channel_stats(x, B, C, H * W, 15, stats, static_cast<cudaStream_t>(cuda_stream)));
  return CODON_OK;
}

int codon_channel_pool(const float* x, int B, int C, int H, int W, float* pooled, void* cuda_stream) {
  if (!x || !pooled || B < 1 || C < 1 || H < 1 || W < 1) return fail(nullptr, CODON_ERR_ARG, "codon_channel_pool: bad argument");
  CU_TRY(nullptr, launch_nchw_channel_pool(x, B, C, H * W, pooled, static_cast<cudaStream_t>(cuda_stream)));
  return CODON_OK;
}

int codon_conv2d_nchw(const float* x, const float* w, const float* bias, int B, int Cin, int H, int W, int Cout,
                      int kh, int kw, int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                      int groups, int relu, float* y, void* cuda_stream) {
  if (!x || !w || !y || B < 1 || Cin < 1 || H < 1 || W < 1 || Cout < 1 || kh < 1 || kw < 1 || stride_h < 1 ||
      stride_w < 1 || pad_h < 0 || pad_w < 0 || dil_h < 1 || dil_w < 1 || groups < 1 || Cin % groups || Cout % groups)
    return fail(nullptr, CODON_ERR_ARG, "codon_conv2d_nchw: bad argument");
  CU_TRY(nullptr, launch_conv2d_nchw(x, w, bias, B, Cin, H, W, Cout, kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h,
                                     dil_w, groups, relu, y, static_cast<cudaStream_t>(cuda_stream)));
  return CODON_OK;
}

int codon_quantise_u8(const float* src, uint8_t* dst, size_t n, int via_half, void* cuda_stream) {
  if (!src || !dst) return fail(nullptr, CODON_ERR_ARG, "codon_quantise_u8: NULL pointer");
  if (n == 0) return CODON_OK;
  CU_TRY(nullptr, launch_quantise_u8(src, dst, n, via_half, static_cast<cudaStream_t>(cuda_stream)));
  return CODON_OK;
}

int codon_masked_rmse(const uint8_t* label, const uint8_t* out, int B, int H, int W, double* rmse, void* cuda_stream) {
  if (!label || !out || !rmse || B < 1 || H < 1 || W < 1) return fail(nullptr, CODON_ERR_ARG, "codon_masked_rmse: bad argument");
  CU_TRY(nullptr, launch_masked_rmse(label, out, B, H * W, rmse, static_cast<cudaStream_t>(cuda_stream)));
  return CODON_OK;
}

int codon_ssim_gauss(const void* img1, const void* img2, int img_dtype, int B, int H, int W, double sd, double c1,
                     double c2, double* ssim, void* workspace, size_t workspace_bytes, void* cuda_stream) {
  if (!img1 || !img2 || !ssim || !workspace || B < 1 || H < 1 || W < 1 || !(sd > 0) || img_dtype < 0 || img_dtype > 1)
    return fail(nullptr, CODON_ERR_ARG, "codon_ssim_gauss: bad argument");
  if (workspace_bytes < ssim_workspace_bytes(B, H, W))
    return fail(nullptr, CODON_ERR_WORKSPACE, "codon_ssim_gauss: workspace %zu B < required %zu B", workspace_bytes,
                ssim_workspace_bytes(B, H, W));
  CU_TRY(nullptr, launch_ssim_gauss(img1, img2, img_dtype, B, H, W, sd, c1, c2, ssim, static_cast<double*>(workspace),
                                    static_cast<cudaStream_t>(cuda_stream)));
  return CODON_OK;
}

int codon_bgr_to_gray_u8(const uint8_t* bgr, uint8_t* gray, size_t n_pixels, int method, void* cuda_stream) {
  if (!bgr || !gray || method < 0 || method > 1) return fail(nullptr, CODON_ERR_ARG, "codon_bgr_to_gray_u8: bad argument");
  if (n_pixels == 0) return CODON_OK;
  CU_TRY(nullptr, launch_bgr_to_gray(bgr, gray, n_pixels, method, static_cast<cudaStream_t>(cuda_stream)));
  return CODON_OK;
}

int codon_u8_to_unit_f32(const uint8_t* src, float* dst, size_t n, void* cuda_stream) {
  if (!src || !dst) return fail(nullptr, CODON_ERR_ARG, "codon_u8_to_unit_f32: NULL pointer");
  if (n == 0) return CODON_OK;
  CU_TRY(nullptr, launch_u8_to_unit(src, dst, n, static_cast<cudaStream_t>(cuda_stream)));
  return CODON_OK;
}

int codon_bicubic_upsample_f32(const float* src, float* dst, int B, int h, int w, int H, int W, void* cuda_stream) {
  if (!src || !dst || B < 1 || h < 1 || w < 1 || H < 1 || W < 1)
    return fail(nullptr, CODON_ERR_ARG, "codon_bicubic_upsample_f32: bad argument");
  CU_TRY(nullptr, launch_bicubic_up(src, dst, B, h, w, H, W, static_cast<cudaStream_t>(cuda_stream)));
  return CODON_OK;
}

int codon_group_create(codon_group** out, codon_ctx** ctxs, int n) {
  if (!out || !ctxs || n < 1 || n > 16) return fail(nullptr, CODON_ERR_ARG, "codon_group_create: bad argument");
  *out = nullptr;
  for (int i = 0; i < n; ++i) {
    if (!ctxs[i] || !ctxs[i]->finalized) return fail(nullptr, CODON_ERR_STATE, "codon_group_create: context %d has no finalized weights", i);
    if (ctxs[i]->mode != ctxs[0]->mode || ctxs[i]->scale != ctxs[0]->scale) return fail(nullptr, CODON_ERR_ARG, "codon_group_create: contexts differ in mode / scale");
    for (int j = 0; j < i; ++j)
      if (ctxs[j]->device == ctxs[i]->device) return fail(nullptr, CODON_ERR_ARG, "codon_group_create: two contexts on device %d", ctxs[i]->device);
  }
  if (n > 16) return fail(nullptr, CODON_ERR_ARG, "codon_group_create: at most 16 GPUs");
  codon_group* G = new codon_group();
  G->n = n;
  G->ctx.assign(ctxs, ctxs + n);
  G->dev.resize(n); G->st.resize(n); G->ws.assign(n, nullptr); G->din.assign(n, nullptr); G->dout.assign(n, nullptr);
  G->pin.assign(n, nullptr); G->flags.assign(n, nullptr);
  G->seq_out.assign(n, std::vector<uint32_t>(n, 0)); G->seq_in.assign(n, std::vector<uint32_t>(n, 0));
  G->r0.resize(n); G->r1.resize(n); G->mt.resize(n); G->mb.resize(n); G->hloc.resize(n); G->chunk_off.resize(n); G->chunks.resize(n);
  for (int g = 0; g < n; ++g) G->dev[g] = ctxs[g]->device;
  for (int g = 0; g < n; ++g) {
    CU_TRY(nullptr, cudaSetDevice(G->dev[g]));
    for (int h = 0; h < n; ++h) {
      if (h == g) continue;
      int can = 0;
      CU_TRY(nullptr, cudaDeviceCanAccessPeer(&can, G->dev[g], G->dev[h]));
      if (!can) { delete G; return fail(nullptr, CODON_ERR_CUDA, "codon_group_create: device %d cannot access device %d", G->dev[g], G->dev[h]); }
      cudaError_t e = cudaDeviceEnablePeerAccess(G->dev[h], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { delete G; return fail(nullptr, CODON_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); }
      cudaGetLastError();
    }
    CU_TRY(nullptr, cudaStreamCreateWithFlags(&G->st[g], cudaStreamNonBlocking));
    void* fl = nullptr;
    CU_TRY(nullptr, cudaMalloc(&fl, (size_t)n * sizeof(uint32_t)));
    CU_TRY(nullptr, cudaMemset(fl, 0, (size_t)n * sizeof(uint32_t)));
    G->flags[g] = static_cast<uint32_t*>(fl);
  }
  *out = G;
  return CODON_OK;
}

void codon_group_destroy(codon_group* G) {
  if (!G) return;
  group_free_buffers(G);
  for (int g = 0; g < G->n; ++g) {
    cudaSetDevice(G->dev[g]);
    if (G->st[g]) { cudaStreamSynchronize(G->st[g]); cudaStreamDestroy(G->st[g]); }
    if (G->flags[g]) cudaFree(G->flags[g]);
  }
  delete G;
}

const char* codon_group_last_error(const codon_group* G) { return G ? G->err.c_str() : g_last_error.c_str(); }

double codon_group_last_ms(const codon_group* G) { return G ? G->last_ms : 0.0; }

int codon_group_forward_host(codon_group* G, const float* depth, const float* guide, float* out, int H, int W) {
  if (!G || !depth || !guide || !out || H < 1 || W < 1) return fail(nullptr, CODON_ERR_ARG, "codon_group_forward_host: bad argument");
  const int n = G->n;
  if (H < n * 2 * kBandHalo) return fail(nullptr, CODON_ERR_ARG, "codon_group_forward_host: frame of %d rows is too small for %d bands", H, n);
  // ---- band geometry: rows split evenly; interior sides carry kBandHalo halo rows ----------------
  if (G->poisoned) {
    // a previous forward failed half-way: drain every stream, then restart the cross-GPU counters from zero
    for (int g = 0; g < n; ++g) {
      CU_TRY(nullptr, cudaSetDevice(G->dev[g]));
      cudaDeviceSynchronize();
      cudaGetLastError();
      CU_TRY(nullptr, cudaMemset(G->flags[g], 0, (size_t)n * sizeof(uint32_t)));
      std::fill(G->seq_out[g].begin(), G->seq_out[g].end(), 0u);
      std::fill(G->seq_in[g].begin(), G->seq_in[g].end(), 0u);
    }
    G->poisoned = false;
  }
  int hmax = 0;
  G->H = H; G->W = W;
  for (int g = 0; g < n; ++g) {
    G->r0[g] = (int)((long)H * g / n); G->r1[g] = (int)((long)H * (g + 1) / n);
    G->mt[g] = g > 0 ? kBandHalo : 0; G->mb[g] = g < n - 1 ? kBandHalo : 0;
    G->hloc[g] = G->r1[g] - G->r0[g] + G->mt[g] + G->mb[g];
    if (G->hloc[g] > hmax) hmax = G->hloc[g];
  }
  // CAC chunk partials per band: folded 8 x 16-pixel cells of the local image when the fused conv epilogue produces the
  // statistics (tensor-core modes, every GPU takes that decision from the tallest band), 256-pixel chunks of the core
  // rows otherwise
  const bool fused_stats = fused_everywhere(G->ctx[0], 1, hmax, W) && exp_knob("CODON_TC_CSTAT", 1);
  G->chunks_total = 0;
  for (int g = 0; g < n; ++g) {
    G->chunk_off[g] = G->chunks_total;
    G->chunks[g] = fused_stats ? cac_cell_chunks(cdiv(W, kTcSubW) * cdiv(G->hloc[g], kTcSubH))
                               : cac_stats_chunks(1, G->r1[g] - G->r0[g], W);
    G->chunks_total += G->chunks[g];
  }
  G->bf = plan_buffers(G->ctx[0], 1, hmax, W, G->chunks_total);
  const size_t px = (size_t)hmax * W;
  if (px > G->cap_px || G->bf.total > G->ws_cap) {
    group_free_buffers(G);
    for (int g = 0; g < n; ++g) {
      CU_TRY(nullptr, cudaSetDevice(G->dev[g]));
      void* p = nullptr;
      CU_TRY(nullptr, cudaMalloc(&p, G->bf.total)); G->ws[g] = static_cast<uint8_t*>(p);
      CU_TRY(nullptr, cudaMalloc(&p, 2 * px * sizeof(float))); G->din[g] = static_cast<float*>(p);
      CU_TRY(nullptr, cudaMalloc(&p, px * sizeof(float))); G->dout[g] = static_cast<float*>(p);
      CU_TRY(nullptr, cudaMallocHost(&p, 3 * px * sizeof(float))); G->pin[g] = static_cast<float*>(p);
    }
    G->cap_px = px; G->ws_cap = G->bf.total;
  }
  G->failed.store(false); G->err.clear();

  auto worker = [&](int g) {
    codon_ctx* ctx = G->ctx[g];
    if (cudaSetDevice(G->dev[g]) != cudaSuccess) { G->set_error("cudaSetDevice failed"); return; }
    cudaStream_t st = G->st[g];
    const int y0 = G->r0[g] - G->mt[g], hl = G->hloc[g];
    const size_t npx = (size_t)hl * W;
    float* pin = G->pin[g];
    memcpy(pin, depth + (size_t)y0 * W, npx * sizeof(float));
    memcpy(pin + px, guide + (size_t)y0 * W, npx * sizeof(float));
    bool ok = cudaMemcpyAsync(G->din[g], pin, npx * sizeof(float), cudaMemcpyHostToDevice, st) == cudaSuccess &&
              cudaMemcpyAsync(G->din[g] + px, pin + px, npx * sizeof(float), cudaMemcpyHostToDevice, st) == cudaSuccess;
    if (!ok) { G->set_error("band upload failed"); return; }
    GroupHook hook(G, g);
    hook.cy0 = G->mt[g]; hook.cy1 = hl - G->mb[g];
    hook.global_hw = (long)H * W;
    hook.chunk_off = G->chunk_off[g]; hook.chunks_total = G->chunks_total; hook.my_chunks = G->chunks[g];
    uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(G->ws[g]), 1024));
    ctx->launches = 0;
    int rc = run_forward(ctx, G->din[g], G->din[g] + px, G->dout[g], 1, hl, W, ws, G->bf, st, &hook);
    if (rc) { G->set_error(std::string("band ") + std::to_string(g) + ": " + ctx->err); return; }
    const size_t core_px = (size_t)(G->r1[g] - G->r0[g]) * W;
    if (cudaMemcpyAsync(pin + 2 * px, G->dout[g] + (size_t)G->mt[g] * W, core_px * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) { G->set_error(std::string("band ") + std::to_string(g) + ": " + cudaGetErrorString(cudaGetLastError())); return; }
    memcpy(out + (size_t)G->r0[g] * W, pin + 2 * px, core_px * sizeof(float));
  };
  // the aligned workspace base must be the same offset on every GPU (cudaMalloc returns >= 256-B aligned
  // pointers; the layout offsets are relative to the 1024-aligned base, and peers address ws[h] + offset)
  for (int g = 0; g < n; ++g)
    if (reinterpret_cast<uintptr_t>(G->ws[g]) % 1024 != 0) return fail(nullptr, CODON_ERR_CUDA, "codon_group: workspace not 1024-byte aligned");
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int g = 1; g < n; ++g) th.emplace_back(worker, g);
  worker(0);
  for (auto& t : th) t.join();
  G->last_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (G->failed.load()) {
    G->poisoned = true;
    return fail(nullptr, CODON_ERR_CUDA, "codon_group_forward_host: %s", G->err.c_str());
  }
  return CODON_OK;
}

}  // extern "C"
