// libspfe: context, buffers, launch plan and C ABI of the B200 SuperPoint
// front-end.  See include/spfe.h for the boundary and DESIGN.md for the plan.
#include "../../include/spfe.h"

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "conv1ab.cuh"
#include "conv1ab_mma.cuh"
#include "conv_tc.cuh"
#include "cov.cuh"
#include "dustpose.cuh"
#include "guided.cuh"
#include "kernels_misc.cuh"
#include "match.cuh"
#include "weights.h"

using namespace spfe;

static constexpr int GUIDED_SMEM_MAX = 160 * 1024;  // guided_kernel: taken / claim arrays of the resolve phase (n * 5 bytes)
static constexpr int DUST_SMEM_MAX = 200 * 1024;  // dust_pose_kernel stages the dust map in shared memory up to this size

namespace {

// Message of the last failure on the calling thread (like errno): the matcher / guided / dust entries run on the
// tracking, local-mapping and loop-closing threads at once, so the message must not live in the shared context.
thread_local std::string g_last_error;
#define g_create_error g_last_error

std::string fmt(const char *f, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof buf, f, ap);
  va_end(ap);
  return buf;
}

// ----------------------------------------------------------------------------
// kernel configurations (see conv_tc.cuh)
//                         TAPS CB  N   epilogue        resident-W  A-ring B-ring
//                                                                                     tile pair (HALVES = 2)
using CfgC64 = ConvCfg<9, 1, 64, EPI_RELU, true, 2, 1, 2>;          // conv2a          slab 54 KB x2 + 72 KB weights
using CfgC64P = ConvCfg<9, 1, 64, EPI_RELU_POOL, true, 2, 1, 2>;    // conv1b, conv2b
using CfgC64X = ConvCfg<9, 1, 64, EPI_RELU, true, 3, 1, 2, 2, true>;        // the same as CTA pairs (cta_group::2): default
using CfgC64PX = ConvCfg<9, 1, 64, EPI_RELU_POOL, true, 3, 1, 2, 1, true>;
using CfgC3a = ConvCfg<9, 1, 128, EPI_RELU, true, 2, 1, 1>;         // conv3a          144 KB weights: single tiles
using CfgC128 = ConvCfg<9, 2, 128, EPI_RELU, false, 2, 6, 2>;       // conv4a, conv4b  slab 54 KB x2 + 6 x 16 KB weights
using CfgC128P = ConvCfg<9, 2, 128, EPI_RELU_POOL, false, 2, 6, 2>; // conv3b
using CfgHeads = ConvCfg<9, 2, 256, EPI_RELU, false, 2, 4, 1>;      // convPa || convDa (NB = 2), N = 256: single tiles
// CTA-pair variants of the streamed-weight layers: each rank streams half of every weight block (half the L2 -> SM
// traffic that bounds these layers) and conv3a keeps half of its weights, which makes room for tile pairs
using CfgC3aX = ConvCfg<9, 1, 128, EPI_RELU, true, 2, 1, 2, 1, true>;
using CfgC128X = ConvCfg<9, 2, 128, EPI_RELU, false, 2, 9, 2, 1, true>;
using CfgC128PX = ConvCfg<9, 2, 128, EPI_RELU_POOL, false, 2, 9, 2, 1, true>;
using CfgHeadsX = ConvCfg<9, 2, 256, EPI_RELU, false, 2, 8, 1, 1, true>;
using CfgPb = ConvCfg<1, 4, 80, EPI_DETECT, true, 4, 1, 1, 2>;  // convPb + detector head   (epilogue-bound: two epilogue groups)
using CfgDb = ConvCfg<1, 4, 256, EPI_L2NORM, true, 4, 1, 1, 2>; // convDb + L2 norm
using CfgMatch = ConvCfg<1, 4, 256, EPI_TOP2, false, 4, 4, 1, 2>;  // descriptor matching: Q.T^T + top-2 per 256-column block
using CfgMatch3 = ConvCfg<1, 4, 256, EPI_TOP3, false, 4, 4, 1, 2>; // descriptor-set matching (spfe_match_*): top-3 per block

// "exact" mode (SPFE_EXACT): hi/lo-split operands, three MMAs per product (ConvCfg::XP); CB counts slabs = 2 x real
// 64-channel blocks.  Weights are streamed as CTA pairs for every 3x3 layer, resident for the 1x1 detector head.
using CfgX64 = ConvCfg<9, 2, 64, EPI_RELU, false, 3, 9, 2, 1, true, true>;          // conv2a
using CfgX64P = ConvCfg<9, 2, 64, EPI_RELU_POOL, false, 3, 9, 2, 1, true, true>;    // conv1b, conv2b
using CfgX3a = ConvCfg<9, 2, 128, EPI_RELU, false, 2, 9, 2, 1, true, true>;         // conv3a
using CfgX128 = ConvCfg<9, 4, 128, EPI_RELU, false, 2, 9, 2, 1, true, true>;        // conv4a, conv4b
using CfgX128P = ConvCfg<9, 4, 128, EPI_RELU_POOL, false, 2, 9, 2, 1, true, true>;  // conv3b
using CfgXHeads = ConvCfg<9, 4, 256, EPI_RELU, false, 2, 8, 1, 1, true, true>;      // convPa || convDa
using CfgXPb = ConvCfg<1, 8, 80, EPI_DETECT, true, 8, 1, 1, 2, false, true>;        // convPb + detector head (Wh | Wl resident)

enum { L1B = 0, L2A, L2B, L3A, L3B, L4A, L4B, LHEADS, LPB, LDB, NLAYERS };
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Layer {
  __half *w = nullptr;  // packed [tap][cblock][cout_total][64]
  float *bias = nullptr;
  CUtensorMap tm;
  CUtensorMap tm_half;  // box of n_tile / 2 rows: what one member of a CTA pair loads
  int taps = 0, cb = 0, cout_total = 0, n_tile = 0;
  double flop_per_px = 0;  // 2 * taps * cin * cout (real couts)
};

struct Slot {
  // kernels / H2D / D2H queues of this slot.  By default all slots share the context's three streams (one in-order
  // compute queue: the persistent kernels of consecutive batches never fight for SMs; copies overlap on their own
  // engines); SPFE_SLOT_STREAMS=1 gives every slot a private stream for all three instead.
  cudaStream_t stream = nullptr, in_stream = nullptr, out_stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_out = nullptr, ev_fork = nullptr, ev_join = nullptr;
  int batch = 0;
  bool pending = false, on_host = false;
  // device activations (NHWC fp16)
  uint8_t *d_gray = nullptr;
  __half *a1a = nullptr, *a1b = nullptr, *a2a = nullptr, *a2b = nullptr, *a3a = nullptr, *a3b = nullptr, *a4a = nullptr,
         *a4b = nullptr, *heads = nullptr, *coarse = nullptr;
  float *score = nullptr, *semi_dust = nullptr, *dense_dust = nullptr, *heat_log = nullptr, *heat = nullptr,
        *heat_inv = nullptr, *heat_mm_f = nullptr;
  unsigned *heat_mm = nullptr;
  uint8_t *argmax = nullptr;
  int *count = nullptr;
  float *kp_xy = nullptr, *kp_score = nullptr, *desc = nullptr;
  int16_t *occ = nullptr;
  unsigned long long *scratch = nullptr;
  CUtensorMap tmA[NLAYERS];
  CUtensorMap tmXA[NLAYERS];  // exact mode: maps over the hi/lo-interleaved activation tensors
  CUtensorMap tmA3aX;  // conv3a as a CTA pair works on tile pairs (slab of 24 pixels per row instead of 16)
  MatchScratch match;
  // SPFE_MATCH_PREV: descriptor "slots": slot 0 = last frame of the previous batch (carry), slot z+1 = frame z
  float *desc_all = nullptr;    // [Bm+1][cap][256]  (desc = desc_all + cap*256)
  int *count_all = nullptr;     // [Bm+1]            (count = count_all + 1)
  __half *x16 = nullptr;        // [Bm+1][rows_pad][256] fp16 copies for the tensor-core candidate GEMM
  float2 *cand = nullptr;       // [2][Bm][rows_pad][NB][2]
  CUtensorMap tmQ, tmT;
  Layer match_layer;            // dummy layer record for launch_conv
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // pinned host mirrors
  uint8_t *h_gray = nullptr;
  int *h_count = nullptr;
  float *h_kp_xy = nullptr, *h_kp_score = nullptr, *h_desc = nullptr, *h_dense = nullptr, *h_semi = nullptr,
        *h_heat = nullptr, *h_heat_inv = nullptr;
  int16_t *h_occ = nullptr;
  uint16_t *h_desc16 = nullptr;  // DESC_F16: [Bm][cap][256] binary16 instead of h_desc
  long long d2h_bytes = 0;       // what the last spfe_wait moved device -> host
  // spfe_extract's single-frame launch plan as a CUDA graph (slot 0): H2D, every kernel, all D2H copies in one launch
  cudaGraphExec_t graph = nullptr;
  cudaEvent_t ev_graph = nullptr, ev_heat = nullptr, ev_nms = nullptr;
  bool capturing = false, graph_inflight = false, eager_desc = false;
  long long graph_launches = 0, graph_d2h = 0;
  float graph_thresh = 0.f;
  int extract_calls = 0;
  int *h_match = nullptr, *h_nprev = nullptr;
  float *h_mdist = nullptr;
  // covariance (device)
  int *cov_owner = nullptr, *cov_qlen = nullptr, *cov_overflow = nullptr, *cov_frame_flag = nullptr, *cov_n_replay = nullptr;
  unsigned long long cov_epoch = 0;  // batches run on this slot (see cov_tag)
  int *cov_done = nullptr, *cov_ctr = nullptr, *cov_big = nullptr, *cov_pend = nullptr, *cov_isbig = nullptr;
  uint32_t *cov_visited = nullptr;
  uint32_t *cov_queue = nullptr;
  float *resp = nullptr, *cov2 = nullptr, *cov2_inv = nullptr;
  float *h_resp = nullptr, *h_cov2 = nullptr, *h_cov2_inv = nullptr;
  int *h_cov_overflow = nullptr;
};

}  // namespace

// Device-resident descriptor rows: fp32 (exact re-rank, guided searches) + the fp16 copy the tensor-core nomination
// reads, with their TMA views.  Guarded by the context's match_mu like every matcher entry.
struct spfe_desc_set {
  spfe_ctx *ctx = nullptr;
  int cap = 0, rows_pad = 0, n = 0;
  bool unit = true;       // every row is a unit vector (what the tensor-core score bound needs)
  float *d32 = nullptr;   // [rows_pad][256]
  __half *d16 = nullptr;  // [rows_pad][256]
  CUtensorMap tmA, tmB;   // d16 as the A operand (128-row tiles) / the B operand (256-row blocks)
};

struct spfe_ctx {
  spfe_config cfg;
  std::string weights_path;
  int H = 0, W = 0, hc = 0, wc = 0, cells = 0, cap = 0, num_sms = 0;
  int rows_pad = 0, match_nb = 0, match_tiles = 0;  // tensor-core matcher geometry
  bool heat = false, cov = false, match_prev = false;  // heat: heat maps computed on the device (EMIT_HEAT or EMIT_COV)
  bool heat_host = false, heat_inv_host = false;       // EMIT_HEAT / EMIT_HEAT_INV: heat_ / heat_inv_ are also copied to the host
  bool exact = false;       // EXACT: hi/lo-split operands, 3 MMAs per product, conv1a in fp32 (unfused)
  bool lazy_heat = false;   // LAZY_HEAT: heat_log + min / max stay on the device for spfe_fetch_heat
  bool desc_f16 = false;    // DESC_F16: descriptors cross PCIe as fp16
  float *fetch_heat = nullptr, *fetch_heat_inv = nullptr;  // [H*W] device staging of spfe_fetch_heat (guarded by match_mu)
  // conv1a + conv1b: 2 = one kernel, both layers on the tensor core (default); 1 = one kernel, conv1a on the CUDA cores
  // (SPFE_CONV1=ffma); 0 = two kernels (SPFE_CONV1=unfused or SPFE_FUSED_CONV1=0; materialises conv1a for inspection)
  bool pair = true;   // SPFE_PAIR=0: single-CTA MMAs for the 64 -> 64 layers as well
  bool pair_conv2a = true;   // SPFE_PAIR_CONV2A=0: conv2a single-CTA
  bool pair_stream = true;   // SPFE_PAIR_STREAM=0: conv3a .. convPa|Da single-CTA
  bool pair_conv1 = true;    // SPFE_PAIR_CONV1=0: the fused conv1a+1b kernel single-CTA
  int nms_smem = 0, nms_list_smem = 0;  // dynamic shared memory of nms_kernel / whether its key list fits in it
  int cov_force = 0;  // SPFE_COV_FORCE (test hook): push floods down the big / sequential fallback paths
  // spfe_dom_timing: CUDA events around the dominant kernel (conv1a+1b) of every batch, the last 64 kept -- its launch
  // duration INSIDE a long run, which is what bench.py's roofline block quotes against the sustained peak
  bool dom_timing = false;
  std::vector<cudaEvent_t> dom_ev;
  long long dom_n = 0;
  bool use_graph = true;  // SPFE_GRAPH=0: spfe_extract enqueues its launch plan call by call instead of replaying a CUDA graph
  bool pdl = false;  // SPFE_PDL=1: programmatic dependent launch of the tensor-core kernels (measured: no gain, the board is power-capped)
  int conv1_mode = 2;
  bool fused_conv1 = true;
  void *w1m = nullptr;  // conv1a weights as hi | lo fp16 UMMA operands (conv1ab_mma.cuh)
  EncodeTiledFn encode = nullptr;
  float *w1a = nullptr, *b1a = nullptr;  // conv1a fp32 [9][64], [64]
  Layer layers[NLAYERS];
  Layer xlayers[NLAYERS];  // exact mode: [tap][Wh_0 | Wl_0 | Wh_1 | Wl_1 ...][cout][64]
  std::vector<Slot> slots;
  std::vector<void *> dev_allocs, host_allocs;
  std::atomic<long long> launches{0};
  std::mutex match_mu;
  cudaStream_t match_stream = nullptr;
  cudaStream_t compute = nullptr, copy_in = nullptr, copy_out = nullptr, aux = nullptr;  // aux: covariance beside the matcher
  cudaStream_t copy_late = nullptr;  // the n[b]-row descriptor copies spfe_wait issues once the counts are on the host
  bool slot_streams = false;
  MatchScratch match;  // for spfe_match_* (descriptor sets)
  spfe_desc_set *tmp_q = nullptr, *tmp_t = nullptr;  // the host-pointer entries upload into these
  std::vector<spfe_desc_set *> sets;                 // every live descriptor set (leftovers are freed with the context)
  float2 *set_cand[2] = {nullptr, nullptr};            // [match_cap rows][match_cap / 256][3] nominees per direction
  unsigned long long *set_second = nullptr;            // [match_cap] second-best keys (k-NN 2)
  int *set_flag = nullptr, *h_set_flag = nullptr;      // "a row is not a unit vector" (device / pinned host)
  int *set_rows = nullptr, *h_set_rows = nullptr;      // gather lists of spfe_desc_set_from_frame
  Layer match_layer;                                   // dummy layer record for launch_conv<CfgMatch3>
  float *h_match_q = nullptr;
  int *h_match_idx = nullptr;
  float *h_match_dist = nullptr;
  int match_cap = 0;
  // spfe_search_guided scratch (device, grown on demand; guarded by match_mu)
  void *guided_buf = nullptr;
  size_t guided_bytes = 0;
  int guided_rounds = 0;         // parallel rounds of the last guided search's resolve phase (spfe_guided_last_rounds)
  void *guided_stage = nullptr;  // page-locked staging of the per-call arrays and results of spfe_search_guided*
  size_t guided_stage_bytes = 0;
  // spfe_dust_pose_* page-locked staging block (grown on demand; guarded by match_mu)
  void *dust_stage = nullptr;
  size_t dust_stage_bytes = 0;

  int fail(int code, const std::string &msg) const {
    g_last_error = msg;
    return code;
  }
};

namespace {

#define CU_OK(ctx, call)                                                                         \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      return (ctx)->fail(SPFE_ERR_CUDA, fmt("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); \
  } while (0)

template <class T>
int dev_alloc(spfe_ctx *c, T **p, size_t n) {
  void *q = nullptr;
  CU_OK(c, cudaMalloc(&q, n * sizeof(T) + 256));
  c->dev_allocs.push_back(q);
  *p = static_cast<T *>(q);
  return SPFE_OK;
}
template <class T>
int host_alloc(spfe_ctx *c, T **p, size_t n) {
  void *q = nullptr;
  CU_OK(c, cudaMallocHost(&q, n * sizeof(T) + 64));
  c->host_allocs.push_back(q);
  *p = static_cast<T *>(q);
  return SPFE_OK;
}

// 4-D NHWC fp16 activation tensor -> TMA map with box {64 ch, 8 px, box_rows, 1 frame}, 128-B swizzle.
int make_act_map(spfe_ctx *c, CUtensorMap *tm, const void *ptr, int C, int W, int H, int B, int box_rows, int box_w = 8) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = c->encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return c->fail(SPFE_ERR_CUDA, fmt("cuTensorMapEncodeTiled(act C=%d W=%d H=%d B=%d) -> %d", C, W, H, B, (int)r));
  return SPFE_OK;
}
// 2-D [rows][cols] fp16 matrix -> TMA map with box {64, box_rows}, 128-B swizzle.
int make_mat_map(spfe_ctx *c, CUtensorMap *tm, const void *ptr, int cols, int rows, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = c->encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return c->fail(SPFE_ERR_CUDA, fmt("cuTensorMapEncodeTiled(mat %dx%d) -> %d", rows, cols, (int)r));
  return SPFE_OK;
}

// Pack OIHW fp32 conv weights of one or two layers (concatenated along cout)
// into [tap][cblock][cout_total][64] fp16; couts beyond the real ones are zero.
int upload_layer(spfe_ctx *c, Layer &L, const std::vector<const HostTensor *> &ws, const std::vector<const HostTensor *> &bs,
                 int n_tile, int cout_total, bool xp = false) {
  const int ci = ws[0]->dims[1], k = ws[0]->dims[2];
  L.taps = k * k;
  L.cb = ci / 64 * (xp ? 2 : 1);  // exact mode: every 64-channel block is a (hi, lo) pair of weight blocks
  L.cout_total = cout_total;
  L.n_tile = n_tile;
  std::vector<__half> packed(static_cast<size_t>(L.taps) * L.cb * cout_total * 64, __float2half(0.f));
  std::vector<float> bias(cout_total, 0.f);
  int o0 = 0;
  double real_couts = 0;
  for (size_t li = 0; li < ws.size(); li++) {
    const HostTensor &w = *ws[li];
    const int co = w.dims[0];
    real_couts += co;
    for (int o = 0; o < co; o++) {
      bias[o0 + o] = bs[li]->data[o];
      for (int ch = 0; ch < ci; ch++)
        for (int t = 0; t < L.taps; t++) {
          const float v = w.data[(static_cast<size_t>(o) * ci + ch) * L.taps + t];
          if (xp) {
            const size_t wb = static_cast<size_t>(t) * L.cb + 2 * (ch / 64);
            const __half hi = __float2half_rn(v);
            packed[(wb * cout_total + o0 + o) * 64 + (ch % 64)] = hi;
            packed[((wb + 1) * cout_total + o0 + o) * 64 + (ch % 64)] = __float2half_rn(v - __half2float(hi));
            continue;
          }
          const size_t wb = static_cast<size_t>(t) * L.cb + ch / 64;
          packed[(wb * cout_total + o0 + o) * 64 + (ch % 64)] = __float2half_rn(v);
        }
    }
    o0 += co;
  }
  L.flop_per_px = 2.0 * L.taps * ci * real_couts;
  int rc;
  if ((rc = dev_alloc(c, &L.w, packed.size()))) return rc;
  if ((rc = dev_alloc(c, &L.bias, bias.size()))) return rc;
  CU_OK(c, cudaMemcpy(L.w, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice));
  CU_OK(c, cudaMemcpy(L.bias, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice));
  if (int rc2 = make_mat_map(c, &L.tm_half, L.w, 64, L.taps * L.cb * cout_total, n_tile / 2)) return rc2;
  return make_mat_map(c, &L.tm, L.w, 64, L.taps * L.cb * cout_total, n_tile);
}

template <class Cfg>
int launch_conv(spfe_ctx *c, cudaStream_t st, const CUtensorMap &tmA, const Layer &L, ConvArgs a, const CUtensorMap *tmB = nullptr) {
  static_assert(Cfg::TAPS == 9 || Cfg::TAPS == 1, "taps");
  if (L.taps != Cfg::TAPS || L.cb != Cfg::CB || L.n_tile != Cfg::N || (Cfg::WRES && a.NB != 1))
    return c->fail(SPFE_ERR_INVALID, "conv launch: layer / kernel configuration mismatch");
  const int smem = Cfg::smem_bytes(a.NB);
  {  // function attributes are per device: remember what this instantiation was granted on the context's device
    static std::mutex mu;
    static int granted[64] = {0};
    std::lock_guard<std::mutex> lock(mu);
    int &g = granted[c->cfg.device_id & 63];
    if (g < smem) {
      CU_OK(c, cudaFuncSetAttribute(conv_tc_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      g = smem;
    }
  }
  a.bias = L.bias;
  a.tiles_x = (a.W + Cfg::TILE_W - 1) / Cfg::TILE_W;
  a.tiles_y = (a.H + 15) / 16;
  a.n_items = a.B * a.tiles_x * a.tiles_y * a.NB;
  if (Cfg::MATCH) a.n_items = (a.m_set ? 1 : a.B * 2) * a.m_tiles * a.NB;
  int grid = a.n_items < c->num_sms ? a.n_items : c->num_sms;
  if (Cfg::PAIR) {  // clusters of two CTAs; every pair works on two items at a time
    const int pair_items = (a.n_items / a.NB + 1) / 2 * a.NB;
    const int pairs = pair_items < c->num_sms / 2 ? pair_items : c->num_sms / 2;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(2 * pairs); lc.blockDim = dim3(Cfg::THREADS); lc.dynamicSmemBytes = smem; lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    CU_OK(c, cudaLaunchKernelEx(&lc, conv_tc_kernel<Cfg>, tmA, L.tm_half, a));
  } else if (c->pdl) {  // programmatic dependent launch: overlap this kernel's prologue with the previous kernel's tail
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(Cfg::THREADS); lc.dynamicSmemBytes = smem; lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = 1;
    CU_OK(c, cudaLaunchKernelEx(&lc, conv_tc_kernel<Cfg>, tmA, tmB ? *tmB : L.tm, a));
  } else {
    conv_tc_kernel<Cfg><<<grid, Cfg::THREADS, smem, st>>>(tmA, tmB ? *tmB : L.tm, a);
  }
  c->launches++;
  CU_OK(c, cudaGetLastError());
  return SPFE_OK;
}

struct StageTimer {
  std::vector<cudaEvent_t> ev;
  std::vector<std::string> names;
  std::vector<double> flop, bytes;
  cudaStream_t st;
  bool on;
  void mark(const char *name, double f = 0, double b = 0) {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ev.push_back(e);
    names.push_back(name);
    flop.push_back(f);
    bytes.push_back(b);
  }
};

int run_match(spfe_ctx *c, cudaStream_t st, const MatchArgs &a, int Z, int rows);
int set_alloc(spfe_ctx *c, int capacity, spfe_desc_set **out);
void set_free(spfe_desc_set *s);

// The convolution stack in "exact" mode (SPFE_EXACT): conv1a in fp32 on the CUDA cores, stored as hi + lo fp16; every
// other layer as three MMAs per product on hi/lo-split operands (ConvCfg::XP).  Same stage names / algorithmic FLOPs as
// the default path, so the bench's roofline block reads the same way.
int run_convs_exact(spfe_ctx *c, Slot &s, int B, StageTimer *tm) {
  const int H = c->H, W = c->W, hc = c->hc, wc = c->wc;
  cudaStream_t st = s.stream;
  auto mark = [&](const char *n, double f = 0, double b = 0) { if (tm) tm->mark(n, f, b); };
  int rc;
  {
    dim3 grid((W + C1A_TW - 1) / C1A_TW, (H + C1A_TH - 1) / C1A_TH, B);
    conv1a_kernel<true><<<grid, 256, 0, st>>>(s.d_gray, s.a1a, c->w1a, c->b1a, B, H, W);
    c->launches++;
    CU_OK(c, cudaGetLastError());
    mark("conv1a", 2.0 * 9 * 64 * H * W * B, (1.0 + 256.0) * H * W * B);
  }
  auto args = [&](int h, int w, int nb, int cout_stride, __half *out) {
    ConvArgs a;
    memset(&a, 0, sizeof a);
    a.B = B; a.H = h; a.W = w; a.NB = nb; a.cout_stride = cout_stride; a.out = out;
    return a;
  };
  auto flop = [&](int l, int h, int w) { return c->layers[l].flop_per_px * h * w * B; };
  if ((rc = launch_conv<CfgX64P>(c, st, s.tmXA[L1B], c->xlayers[L1B], args(H, W, 1, 128, s.a1b)))) return rc;
  mark("conv1b", flop(L1B, H, W), (256.0 + 64.0) * H * W * B);
  if ((rc = launch_conv<CfgX64>(c, st, s.tmXA[L2A], c->xlayers[L2A], args(H / 2, W / 2, 1, 128, s.a2a)))) return rc;
  mark("conv2a", flop(L2A, H / 2, W / 2), 512.0 * (H / 2) * (W / 2) * B);
  if ((rc = launch_conv<CfgX64P>(c, st, s.tmXA[L2B], c->xlayers[L2B], args(H / 2, W / 2, 1, 128, s.a2b)))) return rc;
  mark("conv2b", flop(L2B, H / 2, W / 2), 320.0 * (H / 2) * (W / 2) * B);
  if ((rc = launch_conv<CfgX3a>(c, st, s.tmXA[L3A], c->xlayers[L3A], args(H / 4, W / 4, 1, 256, s.a3a)))) return rc;
  mark("conv3a", flop(L3A, H / 4, W / 4), 768.0 * (H / 4) * (W / 4) * B);
  if ((rc = launch_conv<CfgX128P>(c, st, s.tmXA[L3B], c->xlayers[L3B], args(H / 4, W / 4, 1, 256, s.a3b)))) return rc;
  mark("conv3b", flop(L3B, H / 4, W / 4), 640.0 * (H / 4) * (W / 4) * B);
  if ((rc = launch_conv<CfgX128>(c, st, s.tmXA[L4A], c->xlayers[L4A], args(hc, wc, 1, 256, s.a4a)))) return rc;
  mark("conv4a", flop(L4A, hc, wc), 1024.0 * hc * wc * B);
  if ((rc = launch_conv<CfgX128>(c, st, s.tmXA[L4B], c->xlayers[L4B], args(hc, wc, 1, 256, s.a4b)))) return rc;
  mark("conv4b", flop(L4B, hc, wc), 1024.0 * hc * wc * B);
  if ((rc = launch_conv<CfgXHeads>(c, st, s.tmXA[LHEADS], c->xlayers[LHEADS], args(hc, wc, 2, 1024, s.heads)))) return rc;
  mark("convPa|Da", flop(LHEADS, hc, wc), (512.0 + 2048.0) * hc * wc * B);
  {
    ConvArgs a = args(hc, wc, 1, 0, nullptr);
    a.score = s.score; a.argmax = s.argmax; a.semi_dust = s.semi_dust; a.dense_dust = s.dense_dust;
    a.heat_log = c->heat ? s.heat_log : nullptr;
    a.heat_minmax = c->heat ? s.heat_mm : nullptr;
    if ((rc = launch_conv<CfgXPb>(c, st, s.tmXA[LPB], c->xlayers[LPB], a))) return rc;
    mark("convPb+det", flop(LPB, hc, wc), (1024.0 + 13.0 + (c->heat ? 256.0 : 0.0)) * hc * wc * B);
  }
  {  // descriptor head: the parity bar is 1e-3 cosine, so convDb stays a single product on the hi halves of convDa
    ConvArgs a = args(hc, wc, 1, 256, s.coarse);
    a.cin_off = 512;         // convDa's (hi, lo) blocks follow convPa's four pairs
    a.cin_blk_stride = 2;    // hi blocks only
    if ((rc = launch_conv<CfgDb>(c, st, s.tmXA[LDB], c->layers[LDB], a))) return rc;
    mark("convDb+norm", flop(LDB, hc, wc), (512.0 + 512.0) * hc * wc * B);
  }
  return SPFE_OK;
}

// Enqueue the whole per-batch launch plan on the slot's stream.
int run_pipeline(spfe_ctx *c, Slot &s, int B, StageTimer *tm) {
  const int H = c->H, W = c->W, hc = c->hc, wc = c->wc;
  cudaStream_t st = s.stream;
  auto mark = [&](const char *n, double f = 0, double b = 0) { if (tm) tm->mark(n, f, b); };
  mark("start");
  if (c->heat) {
    init_minmax_kernel<<<(B + 127) / 128, 128, 0, st>>>(s.heat_mm, B);
    c->launches++;
  }
  int rc;
  const bool dom = c->dom_timing && !tm && !s.capturing && !c->exact && c->fused_conv1 && !c->dom_ev.empty();
  if (dom) CU_OK(c, cudaEventRecord(c->dom_ev[2 * (c->dom_n % 64)], st));
  if (c->exact) {
    if ((rc = run_convs_exact(c, s, B, tm))) return rc;
  } else {
  if (c->fused_conv1) {  // conv1a + conv1b + pool in one kernel: u8 image -> fp16 [H/2][W/2][64]
    Conv1abArgs a;
    a.img = s.d_gray; a.w1a = c->w1a; a.b1a = c->b1a; a.w1m = c->w1m; a.b1b = c->layers[L1B].bias; a.out = s.a1b;
    a.B = B; a.H = H; a.W = W; a.tiles_x = (W + 15) / 16; a.tiles_y = (H + 15) / 16;
    a.n_items = B * a.tiles_x * a.tiles_y;
    const int grid = a.n_items < c->num_sms ? a.n_items : c->num_sms;
    if (c->conv1_mode == 2 && c->pair_conv1) {  // CTA pairs: every pair works on two items at a time
      const int pairs = (a.n_items + 1) / 2 < c->num_sms / 2 ? (a.n_items + 1) / 2 : c->num_sms / 2;
      cudaLaunchConfig_t lc = {};
      lc.gridDim = dim3(2 * pairs); lc.blockDim = dim3(c1m::THREADS); lc.dynamicSmemBytes = c1m::SMEM; lc.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      lc.attrs = at; lc.numAttrs = 1;
      CU_OK(c, cudaLaunchKernelEx(&lc, conv1ab_mma_kernel<true>, c->layers[L1B].tm_half, a));
    } else if (c->conv1_mode == 2) conv1ab_mma_kernel<false><<<grid, c1m::THREADS, c1m::SMEM, st>>>(c->layers[L1B].tm, a);
    else conv1ab_kernel<<<grid, c1ab::THREADS, c1ab::SMEM, st>>>(c->layers[L1B].tm, a);
    c->launches++;
    CU_OK(c, cudaGetLastError());
    mark("conv1a+1b", 2.0 * 9 * 64 * H * W * B + c->layers[L1B].flop_per_px * H * W * B, (1.0 + 32.0) * H * W * B);
    if (dom) {
      CU_OK(c, cudaEventRecord(c->dom_ev[2 * (c->dom_n % 64) + 1], st));
      c->dom_n++;
    }
  } else {  // unfused path (debugging aid: materialises the conv1a activation)
    dim3 grid((W + C1A_TW - 1) / C1A_TW, (H + C1A_TH - 1) / C1A_TH, B);
    conv1a_kernel<false><<<grid, 256, 0, st>>>(s.d_gray, s.a1a, c->w1a, c->b1a, B, H, W);
    c->launches++;
    CU_OK(c, cudaGetLastError());
    mark("conv1a", 2.0 * 9 * 64 * H * W * B, (1.0 + 128.0) * H * W * B);
  }
  auto conv_args = [&](int h, int w, int nb, int cout_stride, __half *out) {
    ConvArgs a;
    memset(&a, 0, sizeof a);
    a.B = B; a.H = h; a.W = w; a.NB = nb; a.cin_off = 0; a.cout_stride = cout_stride; a.out = out;
    return a;
  };
  auto stage_flop = [&](int l, int h, int w) { return c->layers[l].flop_per_px * h * w * B; };
  if (!c->fused_conv1) {
    if ((rc = launch_conv<CfgC64P>(c, st, s.tmA[L1B], c->layers[L1B], conv_args(H, W, 1, 64, s.a1b)))) return rc;
    mark("conv1b", stage_flop(L1B, H, W), (128.0 + 32.0) * H * W * B);
  }
  if (c->pair_conv2a) rc = launch_conv<CfgC64X>(c, st, s.tmA[L2A], c->layers[L2A], conv_args(H / 2, W / 2, 1, 64, s.a2a));
  else rc = launch_conv<CfgC64>(c, st, s.tmA[L2A], c->layers[L2A], conv_args(H / 2, W / 2, 1, 64, s.a2a));
  if (rc) return rc;
  mark("conv2a", stage_flop(L2A, H / 2, W / 2), 256.0 * (H / 2) * (W / 2) * B);
  if (c->pair) rc = launch_conv<CfgC64PX>(c, st, s.tmA[L2B], c->layers[L2B], conv_args(H / 2, W / 2, 1, 64, s.a2b));
  else rc = launch_conv<CfgC64P>(c, st, s.tmA[L2B], c->layers[L2B], conv_args(H / 2, W / 2, 1, 64, s.a2b));
  if (rc) return rc;
  mark("conv2b", stage_flop(L2B, H / 2, W / 2), 160.0 * (H / 2) * (W / 2) * B);
  if (c->pair_stream) rc = launch_conv<CfgC3aX>(c, st, s.tmA3aX, c->layers[L3A], conv_args(H / 4, W / 4, 1, 128, s.a3a));
  else rc = launch_conv<CfgC3a>(c, st, s.tmA[L3A], c->layers[L3A], conv_args(H / 4, W / 4, 1, 128, s.a3a));
  if (rc) return rc;
  mark("conv3a", stage_flop(L3A, H / 4, W / 4), 384.0 * (H / 4) * (W / 4) * B);
  if (c->pair_stream) rc = launch_conv<CfgC128PX>(c, st, s.tmA[L3B], c->layers[L3B], conv_args(H / 4, W / 4, 1, 128, s.a3b));
  else rc = launch_conv<CfgC128P>(c, st, s.tmA[L3B], c->layers[L3B], conv_args(H / 4, W / 4, 1, 128, s.a3b));
  if (rc) return rc;
  mark("conv3b", stage_flop(L3B, H / 4, W / 4), 320.0 * (H / 4) * (W / 4) * B);
  if (c->pair_stream) rc = launch_conv<CfgC128X>(c, st, s.tmA[L4A], c->layers[L4A], conv_args(hc, wc, 1, 128, s.a4a));
  else rc = launch_conv<CfgC128>(c, st, s.tmA[L4A], c->layers[L4A], conv_args(hc, wc, 1, 128, s.a4a));
  if (rc) return rc;
  mark("conv4a", stage_flop(L4A, hc, wc), 512.0 * hc * wc * B);
  if (c->pair_stream) rc = launch_conv<CfgC128X>(c, st, s.tmA[L4B], c->layers[L4B], conv_args(hc, wc, 1, 128, s.a4b));
  else rc = launch_conv<CfgC128>(c, st, s.tmA[L4B], c->layers[L4B], conv_args(hc, wc, 1, 128, s.a4b));
  if (rc) return rc;
  mark("conv4b", stage_flop(L4B, hc, wc), 512.0 * hc * wc * B);
  if (c->pair_stream) rc = launch_conv<CfgHeadsX>(c, st, s.tmA[LHEADS], c->layers[LHEADS], conv_args(hc, wc, 2, 512, s.heads));
  else rc = launch_conv<CfgHeads>(c, st, s.tmA[LHEADS], c->layers[LHEADS], conv_args(hc, wc, 2, 512, s.heads));
  if (rc) return rc;
  mark("convPa|Da", stage_flop(LHEADS, hc, wc), (256.0 + 1024.0) * hc * wc * B);
  {
    ConvArgs a = conv_args(hc, wc, 1, 0, nullptr);
    a.cin_off = 0;
    a.score = s.score; a.argmax = s.argmax; a.semi_dust = s.semi_dust; a.dense_dust = s.dense_dust;
    a.heat_log = c->heat ? s.heat_log : nullptr;
    a.heat_minmax = c->heat ? s.heat_mm : nullptr;
    if ((rc = launch_conv<CfgPb>(c, st, s.tmA[LPB], c->layers[LPB], a))) return rc;
    mark("convPb+det", stage_flop(LPB, hc, wc), (512.0 + 13.0 + (c->heat ? 256.0 : 0.0)) * hc * wc * B);
  }
  {
    ConvArgs a = conv_args(hc, wc, 1, 256, s.coarse);
    a.cin_off = 256;
    if ((rc = launch_conv<CfgDb>(c, st, s.tmA[LDB], c->layers[LDB], a))) return rc;
    mark("convDb+norm", stage_flop(LDB, hc, wc), (512.0 + 512.0) * hc * wc * B);
  }
  }  // default (fp16) convolutions
  // computeCovariance (with to_heat) runs beside the rest of the tail: outside profiling runs it goes to the auxiliary
  // queue.  to_heat depends only on the detector head, so it starts at once (next to the NMS, which occupies 64 SMs);
  // the floods wait for the NMS (they need the key points) and overlap the descriptor sampling and the matcher; the two
  // queues join again at the end of the plan.
  const bool fork = tm == nullptr && c->cov && c->match_prev && c->aux != nullptr && !c->slot_streams;
  cudaStream_t st_main = st;
  if (fork) {
    CU_OK(c, cudaEventRecord(s.ev_fork, st_main));
    CU_OK(c, cudaStreamWaitEvent(c->aux, s.ev_fork, 0));
    st = c->aux;
  }
  if (c->cov || c->heat_host || c->heat_inv_host) {  // (LAZY_HEAT alone: heat_log + min / max stay put for spfe_fetch_heat)
    dim3 grid(64, B);
    heat_norm_kernel<<<grid, 256, 0, st>>>(s.heat_log, s.heat_mm, c->heat_host ? s.heat : nullptr, s.heat_inv, s.heat_mm_f, H * W);
    c->launches++;
    CU_OK(c, cudaGetLastError());
    mark("heat_norm", 0, (c->heat_host ? 12.0 : 8.0) * H * W * B);
    if (!tm && (c->heat_host || c->heat_inv_host)) CU_OK(c, cudaEventRecord(s.ev_heat, st));  // the heat maps may leave while the floods run
  }
  st = st_main;
  {
    NmsArgs n;
    n.score = s.score; n.argmax = s.argmax; n.hc = hc; n.wc = wc; n.thresh = c->cfg.score_thresh;
    n.radius = c->cfg.nms_radius; n.border = c->cfg.border; n.cap = c->cap;
    n.count = s.count; n.kp_xy = s.kp_xy; n.kp_score = s.kp_score; n.occ = s.occ; n.scratch = s.scratch;
    n.list_smem = c->nms_list_smem;
    nms_kernel<<<B, 1024, c->nms_smem, st>>>(n);
    c->launches++;
    CU_OK(c, cudaGetLastError());
    mark("nms", 0, 7.0 * c->cells * B);
  }
  if (fork) {
    CU_OK(c, cudaEventRecord(s.ev_nms, st_main));
    CU_OK(c, cudaStreamWaitEvent(c->aux, s.ev_nms, 0));
  }
  {
    dim3 grid((c->cap + 7) / 8, B);
    sample_desc_kernel<<<grid, 256, 0, st>>>(s.coarse, s.kp_xy, s.count, s.desc, hc, wc, c->cap,
                                             s.x16 ? s.x16 + static_cast<size_t>(c->rows_pad) * 256 : nullptr, c->rows_pad);
    c->launches++;
    CU_OK(c, cudaGetLastError());
    mark("sample_desc", 0, 3072.0 * c->cap * B);
  }
  if (fork) st = c->aux;
  if (c->cov) {  // computeCovariance on the device (cov.cuh): parallel floods, then sequential replay of the conflicted few
    const size_t px = static_cast<size_t>(H) * W;
    if (s.capturing)  // a replayed graph cannot count epochs: it clears its own frames and claims with the largest tag
      CU_OK(c, cudaMemsetAsync(s.cov_owner, 0x7F, static_cast<size_t>(B) * px * sizeof(int), st));
    else if (s.cov_epoch % 509 == 0)  // claims carry a falling batch tag (cov_tag): the 4-byte-per-pixel map needs no clearing in between
      CU_OK(c, cudaMemsetAsync(s.cov_owner, 0x7F, static_cast<size_t>(c->cfg.max_batch) * px * sizeof(int), st));
    const size_t vis_words = (px + 31) / 32;
    CU_OK(c, cudaMemsetAsync(s.cov_visited, 0, B * vis_words * sizeof(uint32_t), st));
    CU_OK(c, cudaMemsetAsync(s.cov_frame_flag, 0, B * sizeof(int), st));
    CU_OK(c, cudaMemsetAsync(s.cov_ctr, 0, COV_NCTR * sizeof(int), st));
    CU_OK(c, cudaMemsetAsync(s.cov_overflow, 0, sizeof(int), st));  // per batch: one overflow must not fail every later batch
    CovArgs a;
    a.heat_inv = s.heat_inv; a.kp_xy = s.kp_xy; a.count = s.count; a.owner = s.cov_owner; a.visited = s.cov_visited;
    a.queue = s.cov_queue; a.qlen = s.cov_qlen; a.response = s.resp; a.cov2 = s.cov2; a.cov2_inv = s.cov2_inv;
    a.overflow = s.cov_overflow; a.H = H; a.W = W; a.cap = c->cap; a.B = B; a.round = 0;
    a.force = c->cov_force;
    a.epoch_tag = (508 - static_cast<int>(s.capturing ? 0 : s.cov_epoch % 509)) << 22;  // every tag stays below the cleared value 0x7F7F7F7F
    if (!s.capturing) s.cov_epoch++;
    a.done = s.cov_done; a.isbig = s.cov_isbig; a.ctr = s.cov_ctr; a.big = s.cov_big; a.pend = s.cov_pend;
    a.frame_flag = s.cov_frame_flag; a.n_replay = s.cov_n_replay; a.vis_words = static_cast<int>(vis_words);
    mark("cov_memset", 0, 5.0 * px * B);
    cov_flood_kernel<COV_S_WIN, COV_S_QCAP, COV_S_WARPS, false><<<c->num_sms, COV_S_WARPS * 32, COV_S_SMEM, st>>>(a);
    cov_flood_kernel<COV_WIN, COV_QCAP, COV_B_WARPS, true><<<32, COV_B_WARPS * 32, COV_B_SMEM, st>>>(a);
    mark("cov_flood", 0, 0);
    cov_resolve0_kernel<<<(B * c->cap + 7) / 8, 256, 0, st>>>(a);
    mark("cov_resolve0", 0, 0);
    for (int r = 1; r <= COV_ROUNDS; r++) {
      a.round = r;
      cov_claim_kernel<<<c->num_sms, 256, 0, st>>>(a);
      cov_resolve_kernel<<<c->num_sms > 2 * COV_R_BIG_BLOCKS ? c->num_sms : 2 * COV_R_BIG_BLOCKS, COV_S_WARPS * 32, COV_S_SMEM, st>>>(a);
    }
    mark("cov_rounds", 0, 0);
    cov_replay_kernel<<<B, 32, (COV_SEQ_QCAP + COV_SEQ_BITMAP_WORDS) * sizeof(uint32_t), st>>>(a);
    c->launches += 4 + 2 * COV_ROUNDS;
    CU_OK(c, cudaGetLastError());
    mark("cov_replay", 0, 0);
  }
  if (fork) {
    CU_OK(c, cudaEventRecord(s.ev_join, c->aux));
    st = st_main;
  }
  if (c->match_prev) {
    // frame z vs frame z-1 (frame 0 vs the carry in slot 0): fp16 candidate GEMM on the tensor core for both
    // directions, exact fp32 re-rank of the candidates, cross-check; then the last frame becomes the carry.
    ConvArgs a;
    memset(&a, 0, sizeof a);
    a.B = B; a.H = c->rows_pad / 8; a.W = 8; a.NB = c->match_nb;
    a.m_count = s.count_all; a.m_cand = s.cand; a.m_rows_pad = c->rows_pad; a.m_tiles = c->match_tiles;
    if ((rc = launch_conv<CfgMatch>(c, st, s.tmQ, s.match_layer, a, &s.tmT))) return rc;
    mark("match_gemm", 2.0 * 2 * 256 * c->cap * c->cap * B, 0);
    RerankArgs r;
    r.desc_all = s.desc_all; r.count_all = s.count_all; r.cand = s.cand; r.rowbest = s.match.rowbest; r.colbest = s.match.colbest;
    r.cap = c->cap; r.rows_pad = c->rows_pad; r.NB = c->match_nb; r.Z = B;
    match_rerank_kernel<<<dim3((c->cap + 7) / 8, B, 2), 256, 0, st>>>(r);
    mark("match_rerank", 0, 0);
    MatchArgs m;
    memset(&m, 0, sizeof m);
    m.nq = s.count; m.rowbest = s.match.rowbest; m.colbest = s.match.colbest; m.q2t = s.match.q2t; m.dist = s.match.dist; m.cap = c->cap;
    match_final_kernel<<<dim3((c->cap + 255) / 256, B), 256, 0, st>>>(m);
    c->launches += 2;
    CU_OK(c, cudaGetLastError());
    CU_OK(c, cudaMemcpyAsync(s.match.dn, s.count_all, sizeof(int), cudaMemcpyDeviceToDevice, st));  // n_prev of frame 0
    CU_OK(c, cudaMemcpyAsync(s.desc_all, s.desc_all + static_cast<size_t>(B) * c->cap * 256, static_cast<size_t>(c->cap) * 256 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CU_OK(c, cudaMemcpyAsync(s.x16, s.x16 + static_cast<size_t>(B) * c->rows_pad * 256, static_cast<size_t>(c->rows_pad) * 256 * sizeof(__half), cudaMemcpyDeviceToDevice, st));
    CU_OK(c, cudaMemcpyAsync(s.count_all, s.count_all + B, sizeof(int), cudaMemcpyDeviceToDevice, st));
    mark("match_final", 0, 2048.0 * c->cap * B);
  }
  if (fork) CU_OK(c, cudaStreamWaitEvent(st_main, s.ev_join, 0));
  s.batch = B;
  return SPFE_OK;
}

}  // namespace

// ============================================================================
// C ABI
// ============================================================================
extern "C" {
#ifdef SPFE_C1M_TRACE
void *spfe_c1m_trace_ptr = nullptr;  // device address of g_c1m_trace (tools/c1m_trace.py)
#endif

void spfe_default_config(spfe_config *cfg, int32_t height, int32_t width, int32_t max_keypoints) {
  if (!cfg) return;
  memset(cfg, 0, sizeof *cfg);
  cfg->struct_size = sizeof *cfg;
  cfg->height = height;
  cfg->width = width;
  cfg->max_keypoints = max_keypoints;
  cfg->score_thresh = 0.007f;  // sp_extractor.cpp:122
  cfg->nms_radius = 4;         // sp_extractor.cpp:502
  cfg->border = 8;             // sp_extractor.cpp:502
  cfg->device_id = 0;
  cfg->max_batch = 1;
  cfg->num_slots = 1;
  cfg->flags = SPFE_EMIT_HEAT | SPFE_EMIT_HEAT_INV | SPFE_EMIT_COV;  // what SPExtractor::operator() always produces
  cfg->weights_path = nullptr;
}

const char *spfe_last_error(const spfe_ctx *) { return g_last_error.c_str(); }

static int create_impl(spfe_ctx *c) {
  const spfe_config &cfg = c->cfg;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return c->fail(SPFE_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU fallback)");
  if (cfg.device_id < 0 || cfg.device_id >= ndev) return c->fail(SPFE_ERR_INVALID, "device_id out of range");
  CU_OK(c, cudaSetDevice(cfg.device_id));
  cudaDeviceProp prop;
  CU_OK(c, cudaGetDeviceProperties(&prop, cfg.device_id));
  if (prop.major != 10)
    return c->fail(SPFE_ERR_NO_DEVICE, fmt("device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg.device_id, prop.major, prop.minor));
  c->num_sms = prop.multiProcessorCount;
  {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU_OK(c, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return c->fail(SPFE_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    c->encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const int H = c->H, W = c->W, hc = c->hc, wc = c->wc, Bm = cfg.max_batch;
  int rc;

  // ---- weights
  WeightMap wm;
  std::string err;
  if (!load_weights(c->weights_path, wm, err)) return c->fail(SPFE_ERR_WEIGHTS, err);
  {
    const HostTensor &w = wm["conv1a.weight"], &b = wm["conv1a.bias"];
    std::vector<float> w9(9 * 64);
    for (int o = 0; o < 64; o++)
      for (int t = 0; t < 9; t++) w9[t * 64 + o] = w.data[o * 9 + t];
    if ((rc = dev_alloc(c, &c->w1a, w9.size()))) return rc;
    if ((rc = dev_alloc(c, &c->b1a, 64))) return rc;
    CU_OK(c, cudaMemcpy(c->w1a, w9.data(), w9.size() * 4, cudaMemcpyHostToDevice));
    CU_OK(c, cudaMemcpy(c->b1a, b.data.data(), 64 * 4, cudaMemcpyHostToDevice));
    // conv1ab_mma.cuh: W1[part][64 couts][16 k] fp16, part 0 = hi, 1 = lo = fp16(v - hi); k = tap 0..8, k = 9 is the bias
    // (its im2col column holds 255, the conv1a epilogue scales by 1/255).  Un-swizzled K-major canonical layout:
    // 8-row x 16-byte core matrices, K chunks 128 B apart, 8-row groups 256 B apart.
    std::vector<__half> img(2 * 64 * 16, __float2half(0.f));
    for (int o = 0; o < 64; o++)
      for (int k = 0; k < 10; k++) {
        const float v = k < 9 ? w.data[o * 9 + k] : b.data[o];
        if (!(std::fabs(v) < 60000.f)) return c->fail(SPFE_ERR_WEIGHTS, "conv1a weight / bias outside the fp16 range");
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        const size_t off = static_cast<size_t>(o >> 3) * 128 + (k >> 3) * 64 + (o & 7) * 8 + (k & 7);  // in fp16 elements
        img[off] = hi;
        img[64 * 16 + off] = lo;
      }
    if ((rc = dev_alloc(c, reinterpret_cast<__half **>(&c->w1m), img.size()))) return rc;
    CU_OK(c, cudaMemcpy(c->w1m, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice));
  }
  auto up = [&](int l, std::vector<const char *> names, int n_tile, int cout_total) {
    std::vector<const HostTensor *> ws, bs;
    for (const char *n : names) {
      ws.push_back(&wm[std::string(n) + ".weight"]);
      bs.push_back(&wm[std::string(n) + ".bias"]);
    }
    return upload_layer(c, c->layers[l], ws, bs, n_tile, cout_total);
  };
  if ((rc = up(L1B, {"conv1b"}, 64, 64))) return rc;
  if ((rc = up(L2A, {"conv2a"}, 64, 64))) return rc;
  if ((rc = up(L2B, {"conv2b"}, 64, 64))) return rc;
  if ((rc = up(L3A, {"conv3a"}, 128, 128))) return rc;
  if ((rc = up(L3B, {"conv3b"}, 128, 128))) return rc;
  if ((rc = up(L4A, {"conv4a"}, 128, 128))) return rc;
  if ((rc = up(L4B, {"conv4b"}, 128, 128))) return rc;
  if ((rc = up(LHEADS, {"convPa", "convDa"}, 256, 512))) return rc;
  if ((rc = up(LPB, {"convPb"}, 80, 80))) return rc;
  if ((rc = up(LDB, {"convDb"}, 256, 256))) return rc;
  if (c->exact) {
    auto upx = [&](int l, std::vector<const char *> names, int n_tile, int cout_total) {
      std::vector<const HostTensor *> ws, bs;
      for (const char *n : names) {
        ws.push_back(&wm[std::string(n) + ".weight"]);
        bs.push_back(&wm[std::string(n) + ".bias"]);
      }
      return upload_layer(c, c->xlayers[l], ws, bs, n_tile, cout_total, true);
    };
    if ((rc = upx(L1B, {"conv1b"}, 64, 64))) return rc;
    if ((rc = upx(L2A, {"conv2a"}, 64, 64))) return rc;
    if ((rc = upx(L2B, {"conv2b"}, 64, 64))) return rc;
    if ((rc = upx(L3A, {"conv3a"}, 128, 128))) return rc;
    if ((rc = upx(L3B, {"conv3b"}, 128, 128))) return rc;
    if ((rc = upx(L4A, {"conv4a"}, 128, 128))) return rc;
    if ((rc = upx(L4B, {"conv4b"}, 128, 128))) return rc;
    if ((rc = upx(LHEADS, {"convPa", "convDa"}, 256, 512))) return rc;
    if ((rc = upx(LPB, {"convPb"}, 80, 80))) return rc;
  }

  // ---- per-slot buffers
  c->slots.resize(cfg.num_slots);
  {
    const char *e = getenv("SPFE_SLOT_STREAMS");
    c->slot_streams = e && e[0] == '1';
    if (!c->slot_streams) {
      CU_OK(c, cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
      CU_OK(c, cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
      CU_OK(c, cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
      const char *ax = getenv("SPFE_AUX_STREAM");
      if (!(ax && ax[0] == '0')) CU_OK(c, cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
    }
  }
  CU_OK(c, cudaStreamCreateWithFlags(&c->copy_late, cudaStreamNonBlocking));
  const size_t px = static_cast<size_t>(H) * W, cells = c->cells, cap = c->cap;
  if (c->heat) {
    if ((rc = dev_alloc(c, &c->fetch_heat, px))) return rc;
    if ((rc = dev_alloc(c, &c->fetch_heat_inv, px))) return rc;
  }
  for (Slot &s : c->slots) {
    if (c->slot_streams) {
      CU_OK(c, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
      s.in_stream = s.out_stream = s.stream;
    } else {
      s.stream = c->compute; s.in_stream = c->copy_in; s.out_stream = c->copy_out;
    }
    CU_OK(c, cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
    CU_OK(c, cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
    CU_OK(c, cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming));
    CU_OK(c, cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming));
    CU_OK(c, cudaEventCreateWithFlags(&s.ev_join, cudaEventDisableTiming));
    CU_OK(c, cudaEventCreateWithFlags(&s.ev_graph, cudaEventDisableTiming));
    CU_OK(c, cudaEventCreateWithFlags(&s.ev_heat, cudaEventDisableTiming));
    CU_OK(c, cudaEventCreateWithFlags(&s.ev_nms, cudaEventDisableTiming));
    const size_t xm = c->exact ? 2 : 1;  // exact mode: every activation is a (hi, lo) pair
    if ((rc = dev_alloc(c, &s.d_gray, Bm * px))) return rc;
    if (!c->fused_conv1 && (rc = dev_alloc(c, &s.a1a, Bm * px * 64 * xm))) return rc;
    if ((rc = dev_alloc(c, &s.a1b, Bm * px / 4 * 64 * xm))) return rc;
    if ((rc = dev_alloc(c, &s.a2a, Bm * px / 4 * 64 * xm))) return rc;
    if ((rc = dev_alloc(c, &s.a2b, Bm * px / 16 * 64 * xm))) return rc;
    if ((rc = dev_alloc(c, &s.a3a, Bm * px / 16 * 128 * xm))) return rc;
    if ((rc = dev_alloc(c, &s.a3b, Bm * cells * 128 * xm))) return rc;
    if ((rc = dev_alloc(c, &s.a4a, Bm * cells * 128 * xm))) return rc;
    if ((rc = dev_alloc(c, &s.a4b, Bm * cells * 128 * xm))) return rc;
    if ((rc = dev_alloc(c, &s.heads, Bm * cells * 512 * xm))) return rc;
    if ((rc = dev_alloc(c, &s.coarse, Bm * cells * 256))) return rc;
    if ((rc = dev_alloc(c, &s.score, Bm * cells))) return rc;
    if ((rc = dev_alloc(c, &s.argmax, Bm * cells))) return rc;
    if ((rc = dev_alloc(c, &s.semi_dust, Bm * cells))) return rc;
    if ((rc = dev_alloc(c, &s.dense_dust, Bm * cells))) return rc;
    if ((rc = dev_alloc(c, &s.count_all, Bm + 1))) return rc;
    s.count = s.count_all + 1;
    if ((rc = dev_alloc(c, &s.kp_xy, Bm * cap * 2))) return rc;
    if ((rc = dev_alloc(c, &s.kp_score, Bm * cap))) return rc;
    if ((rc = dev_alloc(c, &s.desc_all, (Bm + 1) * cap * 256))) return rc;
    s.desc = s.desc_all + cap * 256;
    if ((rc = dev_alloc(c, &s.occ, Bm * cells))) return rc;
    if ((rc = dev_alloc(c, &s.scratch, Bm * cells))) return rc;
    CU_OK(c, cudaMemset(s.count_all, 0, (Bm + 1) * sizeof(int)));
    CU_OK(c, cudaMemset(s.kp_xy, 0, Bm * cap * 2 * sizeof(float)));
    CU_OK(c, cudaMemset(s.desc_all, 0, (Bm + 1) * cap * 256 * sizeof(float)));
    if (c->heat) {
      if ((rc = dev_alloc(c, &s.heat_log, Bm * px))) return rc;
      if ((rc = dev_alloc(c, &s.heat, Bm * px))) return rc;
      if ((rc = dev_alloc(c, &s.heat_inv, Bm * px))) return rc;
      if ((rc = dev_alloc(c, &s.heat_mm, Bm * 2))) return rc;
      if ((rc = dev_alloc(c, &s.heat_mm_f, Bm * 2))) return rc;
      if (c->heat_host && (rc = host_alloc(c, &s.h_heat, Bm * px))) return rc;
      if (c->heat_inv_host && (rc = host_alloc(c, &s.h_heat_inv, Bm * px))) return rc;
    }
    if (c->cov) {
      if ((rc = dev_alloc(c, &s.cov_owner, Bm * px))) return rc;
      if ((rc = dev_alloc(c, &s.cov_visited, Bm * ((px + 31) / 32)))) return rc;
      if ((rc = dev_alloc(c, &s.cov_queue, Bm * cap * COV_QCAP))) return rc;
      if ((rc = dev_alloc(c, &s.cov_qlen, Bm * cap))) return rc;
      if ((rc = dev_alloc(c, &s.cov_overflow, 1))) return rc;
      if ((rc = dev_alloc(c, &s.cov_frame_flag, Bm))) return rc;
      if ((rc = dev_alloc(c, &s.cov_n_replay, Bm * 2))) return rc;
      if ((rc = dev_alloc(c, &s.cov_done, Bm * cap))) return rc;
      if ((rc = dev_alloc(c, &s.cov_ctr, COV_NCTR))) return rc;
      if ((rc = dev_alloc(c, &s.cov_isbig, Bm * cap))) return rc;
      if ((rc = dev_alloc(c, &s.cov_big, Bm * cap))) return rc;
      if ((rc = dev_alloc(c, &s.cov_pend, Bm * cap))) return rc;
      CU_OK(c, cudaMemset(s.cov_overflow, 0, sizeof(int)));
      if ((rc = dev_alloc(c, &s.resp, Bm * cap))) return rc;
      if ((rc = dev_alloc(c, &s.cov2, Bm * cap * 2))) return rc;
      if ((rc = dev_alloc(c, &s.cov2_inv, Bm * cap * 2))) return rc;
      if ((rc = host_alloc(c, &s.h_resp, Bm * cap))) return rc;
      if ((rc = host_alloc(c, &s.h_cov2, Bm * cap * 2))) return rc;
      if ((rc = host_alloc(c, &s.h_cov2_inv, Bm * cap * 2))) return rc;
      if ((rc = host_alloc(c, &s.h_cov_overflow, 1))) return rc;
    }
    // matcher scratch (device-resident frame-vs-frame matching inside a slot)
    if ((rc = dev_alloc(c, &s.match.rowbest, Bm * cap))) return rc;
    if ((rc = dev_alloc(c, &s.match.colbest, Bm * cap))) return rc;
    if ((rc = dev_alloc(c, &s.match.q2t, Bm * cap))) return rc;
    if ((rc = dev_alloc(c, &s.match.dist, Bm * cap))) return rc;
    if ((rc = dev_alloc(c, &s.match.dn, 2))) return rc;
    s.match.cap = static_cast<int>(cap);
    if (c->match_prev || c->desc_f16) {
      const size_t rp = c->rows_pad;
      if ((rc = dev_alloc(c, &s.x16, (Bm + 1) * rp * 256))) return rc;
      CU_OK(c, cudaMemset(s.x16, 0, (Bm + 1) * rp * 256 * sizeof(__half)));
    }
    if (c->match_prev) {
      const size_t rp = c->rows_pad, nbk = c->match_nb;
      if ((rc = dev_alloc(c, &s.cand, 2 * Bm * rp * nbk * 4 * 2))) return rc;  // [dir][frame][row][block][residue][top-2]
      if ((rc = make_act_map(c, &s.tmQ, s.x16, 256, 8, static_cast<int>(rp / 8), Bm + 1, 16))) return rc;
      if ((rc = make_mat_map(c, &s.tmT, s.x16, 256, static_cast<int>((Bm + 1) * rp), 256))) return rc;
      s.match_layer.taps = 1; s.match_layer.cb = 4; s.match_layer.n_tile = 256; s.match_layer.cout_total = 256;
    }
    CU_OK(c, cudaEventCreate(&s.ev0));
    CU_OK(c, cudaEventCreate(&s.ev1));
    if ((rc = host_alloc(c, &s.h_match, Bm * cap))) return rc;
    if ((rc = host_alloc(c, &s.h_mdist, Bm * cap))) return rc;
    if ((rc = host_alloc(c, &s.h_nprev, 1))) return rc;
    if ((rc = host_alloc(c, &s.h_gray, Bm * px))) return rc;
    if ((rc = host_alloc(c, &s.h_count, Bm))) return rc;
    if ((rc = host_alloc(c, &s.h_kp_xy, Bm * cap * 2))) return rc;
    if ((rc = host_alloc(c, &s.h_kp_score, Bm * cap))) return rc;
    if (c->desc_f16) { if ((rc = host_alloc(c, &s.h_desc16, Bm * cap * 256))) return rc; }
    else if ((rc = host_alloc(c, &s.h_desc, Bm * cap * 256))) return rc;
    if ((rc = host_alloc(c, &s.h_dense, Bm * cells))) return rc;
    if ((rc = host_alloc(c, &s.h_semi, Bm * cells))) return rc;
    if ((rc = host_alloc(c, &s.h_occ, Bm * cells))) return rc;
    if (c->exact) {  // maps over the hi/lo-interleaved tensors (twice the channels)
      if ((rc = make_act_map(c, &s.tmXA[L1B], s.a1a, 128, W, H, Bm, 18, CfgX64P::PW))) return rc;
      if ((rc = make_act_map(c, &s.tmXA[L2A], s.a1b, 128, W / 2, H / 2, Bm, 18, CfgX64::PW))) return rc;
      if ((rc = make_act_map(c, &s.tmXA[L2B], s.a2a, 128, W / 2, H / 2, Bm, 18, CfgX64P::PW))) return rc;
      if ((rc = make_act_map(c, &s.tmXA[L3A], s.a2b, 128, W / 4, H / 4, Bm, 18, CfgX3a::PW))) return rc;
      if ((rc = make_act_map(c, &s.tmXA[L3B], s.a3a, 256, W / 4, H / 4, Bm, 18, CfgX128P::PW))) return rc;
      if ((rc = make_act_map(c, &s.tmXA[L4A], s.a3b, 256, wc, hc, Bm, 18, CfgX128::PW))) return rc;
      if ((rc = make_act_map(c, &s.tmXA[L4B], s.a4a, 256, wc, hc, Bm, 18, CfgX128::PW))) return rc;
      if ((rc = make_act_map(c, &s.tmXA[LHEADS], s.a4b, 256, wc, hc, Bm, 18, CfgXHeads::PW))) return rc;
      if ((rc = make_act_map(c, &s.tmXA[LPB], s.heads, 1024, wc, hc, Bm, 16))) return rc;
      if ((rc = make_act_map(c, &s.tmXA[LDB], s.heads, 1024, wc, hc, Bm, 16))) return rc;
      continue;
    }
    // TMA maps of every layer's input tensor
    // 3x3 layers: one slab of 18 rows x PW pixels per item (PW = 24 for tile pairs, 16 for single tiles)
    if (!c->fused_conv1 && (rc = make_act_map(c, &s.tmA[L1B], s.a1a, 64, W, H, Bm, 18, CfgC64P::PW))) return rc;
    if ((rc = make_act_map(c, &s.tmA[L2A], s.a1b, 64, W / 2, H / 2, Bm, 18, CfgC64::PW))) return rc;
    if ((rc = make_act_map(c, &s.tmA[L2B], s.a2a, 64, W / 2, H / 2, Bm, 18, CfgC64P::PW))) return rc;
    if ((rc = make_act_map(c, &s.tmA[L3A], s.a2b, 64, W / 4, H / 4, Bm, 18, CfgC3a::PW))) return rc;
    if ((rc = make_act_map(c, &s.tmA3aX, s.a2b, 64, W / 4, H / 4, Bm, 18, CfgC3aX::PW))) return rc;
    if ((rc = make_act_map(c, &s.tmA[L3B], s.a3a, 128, W / 4, H / 4, Bm, 18, CfgC128P::PW))) return rc;
    if ((rc = make_act_map(c, &s.tmA[L4A], s.a3b, 128, wc, hc, Bm, 18, CfgC128::PW))) return rc;
    if ((rc = make_act_map(c, &s.tmA[L4B], s.a4a, 128, wc, hc, Bm, 18, CfgC128::PW))) return rc;
    if ((rc = make_act_map(c, &s.tmA[LHEADS], s.a4b, 128, wc, hc, Bm, 18, CfgHeads::PW))) return rc;
    if ((rc = make_act_map(c, &s.tmA[LPB], s.heads, 512, wc, hc, Bm, 16))) return rc;
    if ((rc = make_act_map(c, &s.tmA[LDB], s.heads, 512, wc, hc, Bm, 16))) return rc;
  }
  // ---- host-pointer matcher scratch
  c->match_cap = ((static_cast<int>(cap) > 4096 ? static_cast<int>(cap) : 4096) + 255) / 256 * 256;
  CU_OK(c, cudaStreamCreateWithFlags(&c->match_stream, cudaStreamNonBlocking));
  if ((rc = dev_alloc(c, &c->match.rowbest, c->match_cap))) return rc;
  if ((rc = dev_alloc(c, &c->match.colbest, c->match_cap))) return rc;
  if ((rc = dev_alloc(c, &c->match.q2t, c->match_cap))) return rc;
  if ((rc = dev_alloc(c, &c->match.dist, c->match_cap))) return rc;
  for (int d = 0; d < 2; d++)
    if ((rc = dev_alloc(c, &c->set_cand[d], static_cast<size_t>(c->match_cap) * (c->match_cap / 256) * 4 * 3))) return rc;
  if ((rc = dev_alloc(c, &c->set_second, c->match_cap))) return rc;
  if ((rc = dev_alloc(c, &c->set_flag, 2))) return rc;
  if ((rc = host_alloc(c, &c->h_set_flag, 2))) return rc;
  if ((rc = dev_alloc(c, &c->set_rows, c->match_cap))) return rc;
  if ((rc = host_alloc(c, &c->h_set_rows, c->match_cap))) return rc;
  c->match_layer.taps = 1; c->match_layer.cb = 4; c->match_layer.n_tile = 256; c->match_layer.cout_total = 256;
  if ((rc = set_alloc(c, c->match_cap, &c->tmp_q))) return rc;
  if ((rc = set_alloc(c, c->match_cap, &c->tmp_t))) return rc;
  if ((rc = dev_alloc(c, &c->match.dn, 2))) return rc;
  c->match.cap = c->match_cap;
  if ((rc = host_alloc(c, &c->h_match_q, static_cast<size_t>(c->match_cap) * 4))) return rc;  // k-NN results: [nq][2] idx + [nq][2] dist
  if ((rc = host_alloc(c, &c->h_match_idx, c->match_cap + 2))) return rc;
  if ((rc = host_alloc(c, &c->h_match_dist, c->match_cap))) return rc;
#ifdef SPFE_C1M_TRACE
  CU_OK(c, cudaGetSymbolAddress(&spfe_c1m_trace_ptr, g_c1m_trace));
#endif
  CU_OK(c, cudaDeviceSynchronize());
  return SPFE_OK;
}

int spfe_create(const spfe_config *cfg, spfe_ctx **out) {
  if (out) *out = nullptr;
  if (!cfg || !out || cfg->struct_size != (int32_t)sizeof(spfe_config)) { g_create_error = "spfe_create: bad config pointer / struct_size"; return SPFE_ERR_INVALID; }
  if (cfg->height <= 0 || cfg->width <= 0 || cfg->height % 8 || cfg->width % 8) { g_create_error = "spfe_create: height/width must be positive multiples of 8 (sp_extractor.cpp:70)"; return SPFE_ERR_INVALID; }
  if (cfg->max_keypoints < 1 || cfg->max_batch < 1 || cfg->num_slots < 1 || cfg->nms_radius < 0 || cfg->nms_radius > 8 || cfg->border < 0) { g_create_error = "spfe_create: bad max_keypoints / max_batch / num_slots / nms_radius / border"; return SPFE_ERR_INVALID; }
  if ((size_t)(cfg->height / 8) * (cfg->width / 8) > 33000) { g_create_error = "spfe_create: more than 33000 cells (NMS shared-memory budget)"; return SPFE_ERR_INVALID; }
  if (!cfg->weights_path) { g_create_error = "spfe_create: weights_path is NULL"; return SPFE_ERR_WEIGHTS; }
  spfe_ctx *c = new spfe_ctx();
  c->cfg = *cfg;
  c->weights_path = cfg->weights_path;
  c->cfg.weights_path = c->weights_path.c_str();
  c->H = cfg->height; c->W = cfg->width; c->hc = c->H / 8; c->wc = c->W / 8; c->cells = c->hc * c->wc;
  c->cap = cfg->max_keypoints + 1;
  if (c->cap > c->cells) c->cap = c->cells;
  c->cov = (cfg->flags & SPFE_EMIT_COV) != 0;
  c->heat_host = (cfg->flags & SPFE_EMIT_HEAT) != 0;
  c->heat_inv_host = (cfg->flags & SPFE_EMIT_HEAT_INV) != 0;
  c->exact = (cfg->flags & SPFE_EXACT) != 0;
  c->lazy_heat = (cfg->flags & SPFE_LAZY_HEAT) != 0;
  c->desc_f16 = (cfg->flags & SPFE_DESC_F16) != 0;
  c->heat = c->cov || c->heat_host || c->heat_inv_host || c->lazy_heat;
  c->match_prev = (cfg->flags & SPFE_MATCH_PREV) != 0;
  c->rows_pad = (c->cap + 255) / 256 * 256;
  c->match_nb = c->rows_pad / 256;
  c->match_tiles = (c->cap + 127) / 128;
  {
    const char *e = getenv("SPFE_FUSED_CONV1"), *m = getenv("SPFE_CONV1");
    if (m && !strcmp(m, "ffma")) c->conv1_mode = 1;
    if ((m && !strcmp(m, "unfused")) || (e && e[0] == '0')) c->conv1_mode = 0;
    if (c->exact) c->conv1_mode = 0;  // exact mode: conv1a in fp32 on the CUDA cores, materialised as hi + lo
    c->fused_conv1 = c->conv1_mode != 0;
    const char *pd = getenv("SPFE_PDL");
    c->pdl = pd && pd[0] == '1';
    const char *pr = getenv("SPFE_PAIR");
    c->pair = !(pr && pr[0] == '0');
    const char *p2 = getenv("SPFE_PAIR_CONV2A");
    c->pair_conv2a = c->pair && !(p2 && p2[0] == '0');
    const char *p3 = getenv("SPFE_PAIR_STREAM");
    c->pair_stream = c->pair && !(p3 && p3[0] == '0');
    const char *p1 = getenv("SPFE_PAIR_CONV1");
    c->pair_conv1 = c->pair && !(p1 && p1[0] == '0');
    const char *gr = getenv("SPFE_GRAPH");
    c->use_graph = !(gr && gr[0] == '0') && !c->slot_streams;
    const char *cf = getenv("SPFE_COV_FORCE");
    c->cov_force = cf ? atoi(cf) : 0;
  }
  int rc = create_impl(c);
  if (rc == SPFE_OK) rc = [&]() -> int {
    // the attribute is per function and device, not per context: only ever raise it (several extractors may coexist)
    static std::mutex mu;
    static int nms_granted[64] = {0};  // per device
    std::lock_guard<std::mutex> lock(mu);
    int &nms_smem_max = nms_granted[c->cfg.device_id & 63];
    if (nms_smem_max == 0) nms_smem_max = 48 * 1024;
    // cell arrays (7 bytes per cell) + the sort-key list (8 bytes per cell) when both fit
    c->nms_list_smem = ((c->cells * 7 + 7) & ~7) + c->cells * 8 <= 200 * 1024;
    c->nms_smem = c->nms_list_smem ? ((c->cells * 7 + 7) & ~7) + c->cells * 8 : c->cells * 7;
    if (c->nms_smem > nms_smem_max) {
      CU_OK(c, cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c->nms_smem));
      nms_smem_max = c->nms_smem;
    }
    CU_OK(c, cudaFuncSetAttribute(dust_pose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DUST_SMEM_MAX));
    CU_OK(c, cudaFuncSetAttribute(guided_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GUIDED_SMEM_MAX));
    CU_OK(c, cudaFuncSetAttribute(conv1ab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c1ab::SMEM));
    CU_OK(c, cudaFuncSetAttribute(conv1ab_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c1m::SMEM));
    CU_OK(c, cudaFuncSetAttribute(conv1ab_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, c1m::SMEM));
    CU_OK(c, cudaFuncSetAttribute(cov_flood_kernel<COV_S_WIN, COV_S_QCAP, COV_S_WARPS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, COV_S_SMEM));
    CU_OK(c, cudaFuncSetAttribute(cov_flood_kernel<COV_WIN, COV_QCAP, COV_B_WARPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, COV_B_SMEM));
    CU_OK(c, cudaFuncSetAttribute(cov_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, COV_S_SMEM));
    CU_OK(c, cudaFuncSetAttribute(cov_replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (COV_SEQ_QCAP + COV_SEQ_BITMAP_WORDS) * (int)sizeof(uint32_t)));
    return SPFE_OK;
  }();
  if (rc != SPFE_OK) {
    spfe_destroy(c);
    return rc;
  }
  *out = c;
  return SPFE_OK;
}

void spfe_destroy(spfe_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device_id);
  cudaDeviceSynchronize();
  for (Slot &s : c->slots) {
    if (s.stream && c->slot_streams) cudaStreamDestroy(s.stream);
    for (cudaEvent_t e : {s.ev_in, s.ev_done, s.ev_out, s.ev_fork, s.ev_join, s.ev_graph, s.ev_heat, s.ev_nms}) if (e) cudaEventDestroy(e);
    if (s.graph) cudaGraphExecDestroy(s.graph);
    if (s.ev0) cudaEventDestroy(s.ev0);
    if (s.ev1) cudaEventDestroy(s.ev1);
  }
  while (!c->sets.empty()) set_free(c->sets.back());  // tmp_q / tmp_t and any set the caller did not destroy
  for (cudaEvent_t e : c->dom_ev) cudaEventDestroy(e);
  if (c->match_stream) cudaStreamDestroy(c->match_stream);
  if (c->guided_buf) cudaFree(c->guided_buf);
  if (c->dust_stage) cudaFreeHost(c->dust_stage);
  if (c->guided_stage) cudaFreeHost(c->guided_stage);
  for (cudaStream_t st : {c->compute, c->copy_in, c->copy_out, c->aux, c->copy_late}) if (st) cudaStreamDestroy(st);
  for (void *p : c->dev_allocs) cudaFree(p);
  for (void *p : c->host_allocs) cudaFreeHost(p);
  delete c;
}

static int check_slot(spfe_ctx *c, int32_t slot) {
  if (!c) return SPFE_ERR_INVALID;
  if (slot < 0 || slot >= (int)c->slots.size()) return c->fail(SPFE_ERR_INVALID, "slot out of range");
  return SPFE_OK;
}

// heat_ / heat_inv_: the largest outputs; they are final after heat_norm, long before the batch's last kernel
static int enqueue_d2h_heat(spfe_ctx *c, Slot &s, int B) {
  const size_t px = (size_t)c->H * c->W;
  cudaStream_t st = s.out_stream;
  if (!(c->heat_host || c->heat_inv_host)) return SPFE_OK;
  if (s.out_stream != s.stream) CU_OK(c, cudaStreamWaitEvent(st, s.ev_heat, 0));
  if (c->heat_host) CU_OK(c, cudaMemcpyAsync(s.h_heat, s.heat, B * px * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (c->heat_inv_host) CU_OK(c, cudaMemcpyAsync(s.h_heat_inv, s.heat_inv, B * px * sizeof(float), cudaMemcpyDeviceToHost, st));
  return SPFE_OK;
}

static int enqueue_d2h(spfe_ctx *c, Slot &s, int B) {
  const size_t px = (size_t)c->H * c->W, cells = c->cells, cap = c->cap;
  cudaStream_t st = s.out_stream;
  CU_OK(c, cudaMemcpyAsync(s.h_count, s.count, B * sizeof(int), cudaMemcpyDeviceToHost, st));
  CU_OK(c, cudaMemcpyAsync(s.h_kp_xy, s.kp_xy, B * cap * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU_OK(c, cudaMemcpyAsync(s.h_kp_score, s.kp_score, B * cap * sizeof(float), cudaMemcpyDeviceToHost, st));
  // (descriptors: spfe_wait copies exactly n[b] rows per frame once the counts are here)
  long long bytes = B * (sizeof(int) + cap * 3 * sizeof(float) + cells * (sizeof(int16_t) + 2 * sizeof(float)));
  if (c->heat_host) bytes += B * px * sizeof(float);
  if (c->heat_inv_host) bytes += B * px * sizeof(float);
  if (c->cov) bytes += B * cap * 5 * sizeof(float) + sizeof(int);
  if (c->match_prev) bytes += B * cap * 8 + sizeof(int);
  if (s.eager_desc) bytes += B * cap * 256 * (c->desc_f16 ? 2 : 4);
  s.d2h_bytes = bytes;
  CU_OK(c, cudaMemcpyAsync(s.h_occ, s.occ, B * cells * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
  CU_OK(c, cudaMemcpyAsync(s.h_dense, s.dense_dust, B * cells * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU_OK(c, cudaMemcpyAsync(s.h_semi, s.semi_dust, B * cells * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (s.eager_desc) {  // single-frame graph: all cap rows inside the graph instead of the n-row copy of spfe_wait
    if (c->desc_f16) CU_OK(c, cudaMemcpyAsync(s.h_desc16, s.x16 + (size_t)c->rows_pad * 256, B * cap * 256 * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
    else CU_OK(c, cudaMemcpyAsync(s.h_desc, s.desc, B * cap * 256 * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  if (c->cov) {
    CU_OK(c, cudaMemcpyAsync(s.h_resp, s.resp, B * cap * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU_OK(c, cudaMemcpyAsync(s.h_cov2, s.cov2, B * cap * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU_OK(c, cudaMemcpyAsync(s.h_cov2_inv, s.cov2_inv, B * cap * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU_OK(c, cudaMemcpyAsync(s.h_cov_overflow, s.cov_overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  if (c->match_prev) {
    CU_OK(c, cudaMemcpyAsync(s.h_match, s.match.q2t, B * cap * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_OK(c, cudaMemcpyAsync(s.h_mdist, s.match.dist, B * cap * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU_OK(c, cudaMemcpyAsync(s.h_nprev, s.match.dn, sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  return SPFE_OK;
}

// pageable -> pinned staging of one batch; frames are independent, so large batches are copied by a few threads
static void stage_frames(spfe_ctx *c, Slot &s, const uint8_t *const *grays, int batch, size_t row_stride) {
  const size_t px = (size_t)c->H * c->W;
  auto copy_range = [&](int b0, int b1) {
    for (int b = b0; b < b1; b++) {
      if (row_stride == (size_t)c->W) memcpy(s.h_gray + b * px, grays[b], px);
      else for (int y = 0; y < c->H; y++) memcpy(s.h_gray + b * px + (size_t)y * c->W, grays[b] + y * row_stride, c->W);
    }
  };
  const int nthr = (batch * px >= (size_t)(4u << 20)) ? (batch < 4 ? batch : 4) : 1;
  if (nthr <= 1) return copy_range(0, batch);
  std::vector<std::thread> th;
  for (int t = 1; t < nthr; t++) th.emplace_back(copy_range, batch * t / nthr, batch * (t + 1) / nthr);
  copy_range(0, batch / nthr);
  for (auto &t : th) t.join();
}

static int submit_host(spfe_ctx *c, Slot &s, const uint8_t *src, int batch) {
  int rc;
  // copy-in queue -> compute queue -> copy-out queue, chained by events (one and the same stream with SPFE_SLOT_STREAMS=1)
  CU_OK(c, cudaMemcpyAsync(s.d_gray, src, batch * (size_t)c->H * c->W, cudaMemcpyHostToDevice, s.in_stream));
  if (s.in_stream != s.stream) {
    CU_OK(c, cudaEventRecord(s.ev_in, s.in_stream));
    CU_OK(c, cudaStreamWaitEvent(s.stream, s.ev_in, 0));
  }
  if ((rc = run_pipeline(c, s, batch, nullptr))) return rc;
  CU_OK(c, cudaEventRecord(s.ev_done, s.stream));
  if ((rc = enqueue_d2h_heat(c, s, batch))) return rc;
  if (s.out_stream != s.stream) CU_OK(c, cudaStreamWaitEvent(s.out_stream, s.ev_done, 0));
  if ((rc = enqueue_d2h(c, s, batch))) return rc;
  CU_OK(c, cudaEventRecord(s.ev_out, s.out_stream));
  s.pending = true;
  s.on_host = true;
  return SPFE_OK;
}

// spfe_extract's launch plan (slot 0, one frame, page-locked staging buffer -> every output) captured once into a CUDA
// graph: ~27 kernel launches, ~15 copies and the event chain between the three queues become ONE launch per frame.
static int capture_extract_graph(spfe_ctx *c, Slot &s) {
  cudaGraph_t g = nullptr;
  s.capturing = true;
  s.eager_desc = true;
  const long long l0 = c->launches.load();
  cudaError_t e = cudaStreamBeginCapture(s.in_stream, cudaStreamCaptureModeThreadLocal);
  int rc = e == cudaSuccess ? submit_host(c, s, s.h_gray, 1) : c->fail(SPFE_ERR_CUDA, fmt("cudaStreamBeginCapture: %s", cudaGetErrorString(e)));
  if (e == cudaSuccess) {
    if (!rc && s.out_stream != s.in_stream) {  // join the copy-out queue (and through it the compute queue) back into the origin
      cudaEventRecord(s.ev_out, s.out_stream);
      cudaStreamWaitEvent(s.in_stream, s.ev_out, 0);
    }
    const cudaError_t e2 = cudaStreamEndCapture(s.in_stream, &g);
    if (!rc && e2 != cudaSuccess) rc = c->fail(SPFE_ERR_CUDA, fmt("cudaStreamEndCapture: %s", cudaGetErrorString(e2)));
  }
  s.capturing = false;
  s.eager_desc = false;  // only the graph copies all cap rows; batched submits keep the n-row copy of spfe_wait
  s.pending = false;
  s.graph_d2h = s.d2h_bytes;
  s.graph_launches = c->launches.load() - l0;
  c->launches -= s.graph_launches;  // nothing ran yet
  if (!rc) {
    const cudaError_t e3 = cudaGraphInstantiate(&s.graph, g, 0);
    if (e3 != cudaSuccess) rc = c->fail(SPFE_ERR_CUDA, fmt("cudaGraphInstantiate: %s", cudaGetErrorString(e3)));
  }
  if (g) cudaGraphDestroy(g);
  if (rc) { s.graph = nullptr; cudaGetLastError(); }
  s.graph_thresh = c->cfg.score_thresh;
  return rc;
}

int spfe_submit(spfe_ctx *c, int32_t slot, const uint8_t *const *grays, int32_t batch, size_t row_stride) {
  int rc = check_slot(c, slot);
  if (rc) return rc;
  if (!grays || batch < 1 || batch > c->cfg.max_batch) return c->fail(SPFE_ERR_INVALID, "spfe_submit: bad grays / batch");
  if (row_stride < (size_t)c->W) return c->fail(SPFE_ERR_INVALID, "spfe_submit: row_stride smaller than width");
  Slot &s = c->slots[slot];
  if (s.pending) return c->fail(SPFE_ERR_STATE, "spfe_submit: slot still has an un-waited batch");
  CU_OK(c, cudaSetDevice(c->cfg.device_id));
  for (int b = 0; b < batch; b++)
    if (!grays[b]) return c->fail(SPFE_ERR_EMPTY, "input image is empty");
  stage_frames(c, s, grays, batch, row_stride);
  return submit_host(c, s, s.h_gray, batch);
}

int spfe_submit_pinned(spfe_ctx *c, int32_t slot, const uint8_t *frames, int32_t batch) {
  int rc = check_slot(c, slot);
  if (rc) return rc;
  if (batch < 1 || batch > c->cfg.max_batch) return c->fail(SPFE_ERR_INVALID, "spfe_submit_pinned: bad batch");
  if (!frames) return c->fail(SPFE_ERR_EMPTY, "input image is empty");
  Slot &s = c->slots[slot];
  if (s.pending) return c->fail(SPFE_ERR_STATE, "spfe_submit_pinned: slot still has an un-waited batch");
  CU_OK(c, cudaSetDevice(c->cfg.device_id));
  return submit_host(c, s, frames, batch);
}

void *spfe_host_alloc(size_t bytes) {
  void *p = nullptr;
  return cudaMallocHost(&p, bytes ? bytes : 1) == cudaSuccess ? p : nullptr;
}
void spfe_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

int spfe_wait(spfe_ctx *c, int32_t slot, spfe_frame_out *outs) {
  int rc = check_slot(c, slot);
  if (rc) return rc;
  Slot &s = c->slots[slot];
  if (!s.pending || !s.on_host) return c->fail(SPFE_ERR_STATE, "spfe_wait: nothing submitted on this slot");
  CU_OK(c, cudaEventSynchronize(s.graph_inflight ? s.ev_graph : s.ev_out));
  s.pending = false;
  const size_t px = (size_t)c->H * c->W, cells = c->cells, cap = c->cap;
  if (s.graph_inflight) s.graph_inflight = false;  // (the graph copied the descriptors itself)
  else {  // the descriptors: exactly n[b] rows of every frame (fp32, or fp16 straight from the matcher's copy)
    CU_OK(c, cudaSetDevice(c->cfg.device_id));
    const size_t row = c->desc_f16 ? 256 * sizeof(uint16_t) : 256 * sizeof(float);
    for (int b = 0; b < s.batch; b++) {
      const size_t n = s.h_count[b] > 0 ? (size_t)s.h_count[b] : 0;
      if (n == 0) continue;
      if (c->desc_f16)
        CU_OK(c, cudaMemcpyAsync(s.h_desc16 + b * cap * 256, s.x16 + (size_t)(b + 1) * c->rows_pad * 256, n * row, cudaMemcpyDeviceToHost, c->copy_late));
      else
        CU_OK(c, cudaMemcpyAsync(s.h_desc + b * cap * 256, s.desc + b * cap * 256, n * row, cudaMemcpyDeviceToHost, c->copy_late));
      s.d2h_bytes += (long long)(n * row);
    }
    CU_OK(c, cudaStreamSynchronize(c->copy_late));
  }
  const int cov_bad = c->cov ? s.h_cov_overflow[0] : 0;
  if (cov_bad && !outs)
    return c->fail(SPFE_ERR_STATE, fmt("covariance flood queue overflow in frame %d of the batch (more than 32768 queued pixels in one keypoint basin)", cov_bad - 1));
  if (!outs) return SPFE_OK;
  for (int b = 0; b < s.batch; b++) {
    spfe_frame_out &o = outs[b];
    memset(&o, 0, sizeof o);
    o.n = s.h_count[b];
    o.kp_xy = s.h_kp_xy + b * cap * 2;
    o.kp_score = s.h_kp_score + b * cap;
    if (c->desc_f16) o.desc_f16 = s.h_desc16 + b * cap * 256;
    else o.desc = s.h_desc + b * cap * 256;
    o.occ_grid = s.h_occ + b * cells;
    o.dense_dust = s.h_dense + b * cells;
    o.semi_dust = s.h_semi + b * cells;
    if (c->heat_host) o.heat = s.h_heat + b * px;
    if (c->heat_inv_host) o.heat_inv = s.h_heat_inv + b * px;
    if (c->match_prev) {
      o.n_prev = b == 0 ? s.h_nprev[0] : s.h_count[b - 1];
      o.match_prev = s.h_match + b * cap;
      o.match_dist = s.h_mdist + b * cap;
    }
    if (c->cov) {
      o.kp_response = s.h_resp + b * cap;
      o.cov2 = s.h_cov2 + b * cap * 2;
      o.cov2_inv = s.h_cov2_inv + b * cap * 2;
    }
  }
  if (cov_bad)  // the other frames of the batch are complete and valid
    return c->fail(SPFE_ERR_STATE, fmt("covariance flood queue overflow in frame %d of the batch (more than 32768 queued pixels in one keypoint basin)", cov_bad - 1));
  return SPFE_OK;
}

int64_t spfe_last_d2h_bytes(const spfe_ctx *c, int32_t slot) {
  if (!c || slot < 0 || slot >= (int)c->slots.size()) return SPFE_ERR_INVALID;
  return c->slots[slot].d2h_bytes;
}

int spfe_fetch_heat(spfe_ctx *c, int32_t slot, int32_t frame, float *heat, float *heat_inv) {
  int rc = check_slot(c, slot);
  if (rc) return rc;
  Slot &s = c->slots[slot];
  if (!c->heat) return c->fail(SPFE_ERR_STATE, "spfe_fetch_heat: the heat maps are not computed (set SPFE_LAZY_HEAT, SPFE_EMIT_COV or SPFE_EMIT_HEAT)");
  if (s.pending) return c->fail(SPFE_ERR_STATE, "spfe_fetch_heat: slot still has an un-waited batch");
  if (frame < 0 || frame >= s.batch) return c->fail(SPFE_ERR_STATE, fmt("spfe_fetch_heat: frame %d is not part of the slot's last batch (%d frames)", frame, s.batch));
  if (!heat && !heat_inv) return SPFE_OK;
  std::lock_guard<std::mutex> lock(c->match_mu);
  CU_OK(c, cudaSetDevice(c->cfg.device_id));
  cudaStream_t st = c->match_stream;
  const size_t px = (size_t)c->H * c->W;
  CU_OK(c, cudaStreamWaitEvent(st, s.ev_done, 0));
  // the pipeline's own kernel on one frame: same arithmetic, same bits as the eager SPFE_EMIT_HEAT copy
  heat_norm_kernel<<<dim3(64, 1), 256, 0, st>>>(s.heat_log + frame * px, s.heat_mm + 2 * frame, heat ? c->fetch_heat : nullptr,
                                               c->fetch_heat_inv, nullptr, (int)px);
  c->launches++;
  CU_OK(c, cudaGetLastError());
  if (heat) CU_OK(c, cudaMemcpyAsync(heat, c->fetch_heat, px * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (heat_inv) CU_OK(c, cudaMemcpyAsync(heat_inv, c->fetch_heat_inv, px * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU_OK(c, cudaStreamSynchronize(st));
  return SPFE_OK;
}

int spfe_extract(spfe_ctx *c, const uint8_t *gray, size_t row_stride, spfe_frame_out *out) {
  if (!c) return SPFE_ERR_INVALID;
  if (!gray) return c->fail(SPFE_ERR_EMPTY, "input image is empty");
  if (!out) return c->fail(SPFE_ERR_INVALID, "spfe_extract: out is NULL");
  Slot &s = c->slots[0];
  if (c->use_graph && !s.pending && row_stride >= (size_t)c->W) {
    // first call: plain enqueue (also settles every per-kernel attribute); second call: capture; from then on: replay
    if (s.graph && s.graph_thresh != c->cfg.score_thresh) { cudaGraphExecDestroy(s.graph); s.graph = nullptr; s.extract_calls = 1; }
    if (!s.graph && s.extract_calls == 1) {
      CU_OK(c, cudaSetDevice(c->cfg.device_id));
      CU_OK(c, cudaDeviceSynchronize());  // nothing of this context in flight while the streams are in capture mode
      if (capture_extract_graph(c, s) != SPFE_OK) c->use_graph = false;  // keep working call by call
    }
    if (s.graph) {
      CU_OK(c, cudaSetDevice(c->cfg.device_id));
      const uint8_t *one[1] = {gray};
      stage_frames(c, s, one, 1, row_stride);
      CU_OK(c, cudaGraphLaunch(s.graph, s.in_stream));
      CU_OK(c, cudaEventRecord(s.ev_graph, s.in_stream));
      CU_OK(c, cudaEventRecord(s.ev_done, s.in_stream));  // what spfe_fetch_heat / the dust-pose entries order themselves after
      c->launches += s.graph_launches;
      s.d2h_bytes = s.graph_d2h;
      s.batch = 1;
      s.pending = s.on_host = s.graph_inflight = true;
      return spfe_wait(c, 0, out);
    }
  }
  s.extract_calls++;
  const uint8_t *one[1] = {gray};
  int rc = spfe_submit(c, 0, one, 1, row_stride);
  if (rc) return rc;
  return spfe_wait(c, 0, out);
}

int spfe_submit_device(spfe_ctx *c, int32_t slot, const void *d_gray, int32_t batch) {
  int rc = check_slot(c, slot);
  if (rc) return rc;
  if (!d_gray || batch < 1 || batch > c->cfg.max_batch) return c->fail(SPFE_ERR_INVALID, "spfe_submit_device: bad pointer / batch");
  Slot &s = c->slots[slot];
  CU_OK(c, cudaSetDevice(c->cfg.device_id));
  uint8_t *saved = s.d_gray;
  s.d_gray = const_cast<uint8_t *>(static_cast<const uint8_t *>(d_gray));
  rc = run_pipeline(c, s, batch, nullptr);
  s.d_gray = saved;
  s.pending = false;
  s.on_host = false;
  if (rc) return rc;
  CU_OK(c, cudaEventRecord(s.ev_done, s.stream));
  return SPFE_OK;
}

int spfe_slot_sync(spfe_ctx *c, int32_t slot) {
  int rc = check_slot(c, slot);
  if (rc) return rc;
  CU_OK(c, cudaEventSynchronize(c->slots[slot].ev_done));  // the slot's last submit; other slots may still be running
  return SPFE_OK;
}

// ---- matcher -----------------------------------------------------------------
}  // extern "C"
namespace {
int run_match(spfe_ctx *c, cudaStream_t st, const MatchArgs &a, int Z, int rows) {
  match_init_kernel<<<(Z * a.cap + 255) / 256, 256, 0, st>>>(a.rowbest, a.colbest, Z * a.cap);
  dim3 grid((rows + 63) / 64, (rows + 63) / 64, Z);
  match_dist_kernel<<<grid, 256, 0, st>>>(a);
  match_final_kernel<<<dim3((a.cap + 255) / 256, Z), 256, 0, st>>>(a);
  c->launches += 3;
  CU_OK(c, cudaGetLastError());
  return SPFE_OK;
}
}  // namespace
extern "C" {

// ---- descriptor sets ------------------------------------------------------------------------------------------------
}  // extern "C"
namespace {
int set_alloc(spfe_ctx *c, int capacity, spfe_desc_set **out) {
  spfe_desc_set *s = new spfe_desc_set();
  s->ctx = c;
  s->cap = capacity;
  s->rows_pad = (capacity + 255) / 256 * 256;
  const size_t elems = static_cast<size_t>(s->rows_pad) * 256;
  cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&s->d32), elems * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&s->d16), elems * sizeof(__half));
  if (e == cudaSuccess) e = cudaMemset(s->d16, 0, elems * sizeof(__half));
  if (e == cudaSuccess) e = cudaMemset(s->d32, 0, elems * sizeof(float));
  int rc = e == cudaSuccess ? SPFE_OK : c->fail(SPFE_ERR_CUDA, fmt("descriptor set of %d rows: %s", capacity, cudaGetErrorString(e)));
  if (!rc) rc = make_act_map(c, &s->tmA, s->d16, 256, 8, s->rows_pad / 8, 1, 16);
  if (!rc) rc = make_mat_map(c, &s->tmB, s->d16, 256, s->rows_pad, 256);
  if (rc) {
    if (s->d32) cudaFree(s->d32);
    if (s->d16) cudaFree(s->d16);
    delete s;
    return rc;
  }
  c->sets.push_back(s);
  *out = s;
  return SPFE_OK;
}
void set_free(spfe_desc_set *s) {
  if (!s) return;
  std::vector<spfe_desc_set *> &v = s->ctx->sets;
  for (size_t i = 0; i < v.size(); i++)
    if (v[i] == s) { v.erase(v.begin() + i); break; }
  cudaFree(s->d32);
  cudaFree(s->d16);
  delete s;
}
// (match_mu held) fp32 rows already in s->d32 or at `src` on the device -> d32 + fp16 copy + unit flag
int set_prepare(spfe_ctx *c, spfe_desc_set *s, const float *d_src, const int *d_rows, int n, bool check_unit) {
  cudaStream_t st = c->match_stream;
  s->n = n;
  s->unit = true;
  if (n == 0) return SPFE_OK;
  if (check_unit) CU_OK(c, cudaMemsetAsync(c->set_flag, 0, sizeof(int), st));
  desc_prepare_kernel<<<(n + 7) / 8, 256, 0, st>>>(d_src, d_rows, n, s->d32, s->d16, c->set_flag);
  c->launches++;
  CU_OK(c, cudaGetLastError());
  if (check_unit) {
    CU_OK(c, cudaMemcpyAsync(c->h_set_flag, c->set_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_OK(c, cudaStreamSynchronize(st));
    s->unit = c->h_set_flag[0] == 0;
  }
  return SPFE_OK;
}
int set_upload_locked(spfe_ctx *c, spfe_desc_set *s, const float *rows, int n) {
  if (n > 0) CU_OK(c, cudaMemcpyAsync(s->d32, rows, static_cast<size_t>(n) * 256 * sizeof(float), cudaMemcpyHostToDevice, c->match_stream));
  return set_prepare(c, s, s->d32, nullptr, n, true);
}
// (match_mu held) one direction of the tensor-core nomination: top-3 columns of B per 256-column block for every row of A
int set_nominate(spfe_ctx *c, const spfe_desc_set *A, const spfe_desc_set *B, float2 *cand) {
  ConvArgs a;
  memset(&a, 0, sizeof a);
  a.B = 1; a.H = A->rows_pad / 8; a.W = 8; a.NB = (B->n + 255) / 256;
  a.m_cand = cand; a.m_rows_pad = c->match_cap; a.m_tiles = (A->n + 127) / 128;
  a.m_set = 1; a.m_na = A->n; a.m_nb = B->n; a.m_dir = 0;
  return launch_conv<CfgMatch3>(c, c->match_stream, A->tmA, c->match_layer, a, &B->tmB);
}
int set_rerank(spfe_ctx *c, const spfe_desc_set *A, const spfe_desc_set *B, const float2 *cand, int R, unsigned long long *out1, unsigned long long *out2) {
  SetRerankArgs r;
  r.a_rows = A->d32; r.b_rows = B->d32; r.n_a = A->n; r.n_b = B->n; r.rows_pad = c->match_cap; r.NB = (B->n + 255) / 256; r.dir = 0;
  r.cand = cand; r.out1 = out1; r.out2 = out2;
  if (R == 1) match_rerank_set_kernel<1><<<(A->n + 7) / 8, 256, 0, c->match_stream>>>(r);
  else match_rerank_set_kernel<2><<<(A->n + 7) / 8, 256, 0, c->match_stream>>>(r);
  c->launches++;
  CU_OK(c, cudaGetLastError());
  return SPFE_OK;
}
// exact CUDA-core path for sets whose rows are not unit vectors (the tensor-core score bound does not hold for them)
int set_match_general(spfe_ctx *c, const spfe_desc_set *Q, const spfe_desc_set *T, bool knn2) {
  cudaStream_t st = c->match_stream;
  MatchScratch &m = c->match;
  c->h_match_idx[c->match_cap] = Q->n;
  c->h_match_idx[c->match_cap + 1] = T->n;
  CU_OK(c, cudaMemcpyAsync(m.dn, c->h_match_idx + c->match_cap, 2 * sizeof(int), cudaMemcpyHostToDevice, st));
  MatchArgs a;
  memset(&a, 0, sizeof a);
  a.q = Q->d32; a.nq = m.dn; a.t0 = T->d32; a.nt0 = m.dn + 1;
  a.rowbest = m.rowbest; a.colbest = m.colbest; a.q2t = m.q2t; a.dist = m.dist; a.cap = c->match_cap;
  const int rows = Q->n > T->n ? Q->n : T->n;
  if (!knn2) return run_match(c, st, a, 1, rows);
  dim3 grid((rows + 63) / 64, (rows + 63) / 64, 1);
  match_init_kernel<<<(a.cap + 255) / 256, 256, 0, st>>>(a.rowbest, a.colbest, a.cap);
  match_dist_kernel<<<grid, 256, 0, st>>>(a);                        // pass 1: nearest row of every query
  match_init_kernel<<<(a.cap + 255) / 256, 256, 0, st>>>(c->set_second, c->set_second, a.cap);
  MatchArgs b = a;
  b.excl = m.rowbest;
  b.rowbest = c->set_second;                                         // pass 2: nearest among the others -> second best
  match_dist_kernel<<<grid, 256, 0, st>>>(b);
  c->launches += 4;
  CU_OK(c, cudaGetLastError());
  return SPFE_OK;
}
int check_set(spfe_ctx *c, const spfe_desc_set *s, const char *fn) {
  if (!s || s->ctx != c) return c->fail(SPFE_ERR_INVALID, fmt("%s: descriptor set is NULL or belongs to another context", fn));
  return SPFE_OK;
}
int match_sets_locked(spfe_ctx *c, const spfe_desc_set *Q, const spfe_desc_set *T, int32_t *q2t, float *dist) {
  const int nq = Q->n, nt = T->n;
  if (nq == 0) return SPFE_OK;
  if (nt == 0) {
    for (int i = 0; i < nq; i++) { q2t[i] = -1; if (dist) dist[i] = 0.f; }
    return SPFE_OK;
  }
  cudaStream_t st = c->match_stream;
  MatchScratch &m = c->match;
  int rc;
  if (Q->unit && T->unit) {  // tensor-core nomination in both directions, exact fp32 re-rank, cross-check
    if ((rc = set_nominate(c, Q, T, c->set_cand[0]))) return rc;
    if ((rc = set_nominate(c, T, Q, c->set_cand[1]))) return rc;
    if ((rc = set_rerank(c, Q, T, c->set_cand[0], 1, m.rowbest, nullptr))) return rc;
    if ((rc = set_rerank(c, T, Q, c->set_cand[1], 1, m.colbest, nullptr))) return rc;
    match_cross_kernel<<<(nq + 255) / 256, 256, 0, st>>>(m.rowbest, m.colbest, nq, m.q2t, m.dist);
    c->launches++;
    CU_OK(c, cudaGetLastError());
  } else if ((rc = set_match_general(c, Q, T, false))) {
    return rc;
  }
  CU_OK(c, cudaMemcpyAsync(c->h_match_idx, m.q2t, nq * sizeof(int), cudaMemcpyDeviceToHost, st));
  CU_OK(c, cudaMemcpyAsync(c->h_match_dist, m.dist, nq * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU_OK(c, cudaStreamSynchronize(st));
  memcpy(q2t, c->h_match_idx, nq * sizeof(int));
  if (dist) memcpy(dist, c->h_match_dist, nq * sizeof(float));
  return SPFE_OK;
}
int knn2_sets_locked(spfe_ctx *c, const spfe_desc_set *Q, const spfe_desc_set *T, int32_t *idx, float *dist) {
  const int nq = Q->n, nt = T->n;
  if (nq == 0) return SPFE_OK;
  if (nt == 0) {
    for (int i = 0; i < 2 * nq; i++) { idx[i] = -1; dist[i] = 0.f; }
    return SPFE_OK;
  }
  cudaStream_t st = c->match_stream;
  MatchScratch &m = c->match;
  int rc;
  if (Q->unit && T->unit) {  // ONE nomination pass gives both neighbours (top-3 per block, exact re-rank of the best two)
    if ((rc = set_nominate(c, Q, T, c->set_cand[0]))) return rc;
    if ((rc = set_rerank(c, Q, T, c->set_cand[0], 2, m.rowbest, c->set_second))) return rc;
  } else if ((rc = set_match_general(c, Q, T, true))) {
    return rc;
  }
  // results as [nq][2] indices followed by [nq][2] distances in the (idle) second nominee buffer
  knn2_final_kernel<<<(nq + 255) / 256, 256, 0, st>>>(m.rowbest, c->set_second, nq, reinterpret_cast<int *>(c->set_cand[1]),
                                                      reinterpret_cast<float *>(c->set_cand[1]) + 2 * (size_t)nq);
  c->launches++;
  CU_OK(c, cudaGetLastError());
  CU_OK(c, cudaMemcpyAsync(c->h_match_q, c->set_cand[1], (size_t)nq * 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CU_OK(c, cudaStreamSynchronize(st));
  memcpy(idx, c->h_match_q, (size_t)nq * 2 * sizeof(int));
  memcpy(dist, reinterpret_cast<float *>(c->h_match_q) + 2 * (size_t)nq, (size_t)nq * 2 * sizeof(float));
  return SPFE_OK;
}
}  // namespace
extern "C" {

int spfe_desc_set_create(spfe_ctx *c, int32_t capacity, spfe_desc_set **out) {
  if (out) *out = nullptr;
  if (!c) return SPFE_ERR_INVALID;
  if (!out || capacity < 1 || capacity > c->match_cap) return c->fail(SPFE_ERR_INVALID, fmt("spfe_desc_set_create: capacity must be 1 .. %d rows", c->match_cap));
  std::lock_guard<std::mutex> lock(c->match_mu);
  CU_OK(c, cudaSetDevice(c->cfg.device_id));
  return set_alloc(c, capacity, out);
}

void spfe_desc_set_destroy(spfe_ctx *c, spfe_desc_set *s) {
  if (!c || !s) return;
  std::lock_guard<std::mutex> lock(c->match_mu);
  bool mine = false;
  for (spfe_desc_set *p : c->sets) mine |= p == s;
  if (!mine || s == c->tmp_q || s == c->tmp_t) return;  // not a live set of this context
  cudaSetDevice(c->cfg.device_id);
  cudaStreamSynchronize(c->match_stream);
  set_free(s);
}

int32_t spfe_desc_set_size(const spfe_desc_set *s) { return s ? s->n : SPFE_ERR_INVALID; }

int spfe_desc_set_upload(spfe_ctx *c, spfe_desc_set *s, const float *rows, int32_t n) {
  if (!c) return SPFE_ERR_INVALID;
  int rc = check_set(c, s, "spfe_desc_set_upload");
  if (rc) return rc;
  if (n < 0 || n > s->cap || (n > 0 && !rows)) return c->fail(SPFE_ERR_INVALID, fmt("spfe_desc_set_upload: bad rows / n (capacity %d)", s->cap));
  std::lock_guard<std::mutex> lock(c->match_mu);
  CU_OK(c, cudaSetDevice(c->cfg.device_id));
  return set_upload_locked(c, s, rows, n);
}

int spfe_desc_set_from_frame(spfe_ctx *c, spfe_desc_set *s, int32_t slot, int32_t frame, const int32_t *rows, int32_t n) {
  if (!c) return SPFE_ERR_INVALID;
  int rc = check_set(c, s, "spfe_desc_set_from_frame");
  if (rc) return rc;
  if ((rc = check_slot(c, slot))) return rc;
  Slot &sl = c->slots[slot];
  if (sl.pending) return c->fail(SPFE_ERR_STATE, "spfe_desc_set_from_frame: slot still has an un-waited batch");
  if (frame < 0 || frame >= sl.batch) return c->fail(SPFE_ERR_STATE, fmt("spfe_desc_set_from_frame: frame %d is not part of the slot's last batch (%d frames)", frame, sl.batch));
  std::lock_guard<std::mutex> lock(c->match_mu);
  CU_OK(c, cudaSetDevice(c->cfg.device_id));
  cudaStream_t st = c->match_stream;
  CU_OK(c, cudaStreamWaitEvent(st, sl.ev_done, 0));
  int n_frame = 0;
  if (sl.on_host) n_frame = sl.h_count[frame];
  else {
    CU_OK(c, cudaMemcpyAsync(c->h_set_flag + 1, sl.count + frame, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_OK(c, cudaStreamSynchronize(st));
    n_frame = c->h_set_flag[1];
  }
  const float *src = sl.desc + static_cast<size_t>(frame) * c->cap * 256;
  if (!rows) {  // all key points of the frame
    if (n_frame > s->cap) return c->fail(SPFE_ERR_INVALID, fmt("spfe_desc_set_from_frame: the frame has %d key points, the set holds %d", n_frame, s->cap));
    return set_prepare(c, s, src, nullptr, n_frame, false);  // the extractor's rows are normalised in fp32
  }
  if (n < 0 || n > s->cap) return c->fail(SPFE_ERR_INVALID, "spfe_desc_set_from_frame: bad n");
  for (int i = 0; i < n; i++)
    if (rows[i] < 0 || rows[i] >= n_frame) return c->fail(SPFE_ERR_INVALID, fmt("spfe_desc_set_from_frame: row %d out of range (%d key points)", rows[i], n_frame));
  if (n > 0) {
    memcpy(c->h_set_rows, rows, n * sizeof(int));
    CU_OK(c, cudaMemcpyAsync(c->set_rows, c->h_set_rows, n * sizeof(int), cudaMemcpyHostToDevice, st));
  }
  rc = set_prepare(c, s, src, c->set_rows, n, false);
  if (!rc && n > 0) CU_OK(c, cudaStreamSynchronize(st));  // h_set_rows may be reused by the next call
  return rc;
}

int spfe_match_mutual_nn_sets(spfe_ctx *c, const spfe_desc_set *q, const spfe_desc_set *t, int32_t *q2t, float *dist) {
  if (!c) return SPFE_ERR_INVALID;
  int rc = check_set(c, q, "spfe_match_mutual_nn_sets");
  if (!rc) rc = check_set(c, t, "spfe_match_mutual_nn_sets");
  if (rc) return rc;
  if (q->n > 0 && !q2t) return c->fail(SPFE_ERR_INVALID, "spfe_match_mutual_nn_sets: q2t is NULL");
  std::lock_guard<std::mutex> lock(c->match_mu);
  CU_OK(c, cudaSetDevice(c->cfg.device_id));
  return match_sets_locked(c, q, t, q2t, dist);
}

int spfe_match_knn2_sets(spfe_ctx *c, const spfe_desc_set *q, const spfe_desc_set *t, int32_t *idx, float *dist) {
  if (!c) return SPFE_ERR_INVALID;
  int rc = check_set(c, q, "spfe_match_knn2_sets");
  if (!rc) rc = check_set(c, t, "spfe_match_knn2_sets");
  if (rc) return rc;
  if (q->n > 0 && (!idx || !dist)) return c->fail(SPFE_ERR_INVALID, "spfe_match_knn2_sets: NULL output");
  std::lock_guard<std::mutex> lock(c->match_mu);
  CU_OK(c, cudaSetDevice(c->cfg.device_id));
  return knn2_sets_locked(c, q, t, idx, dist);
}

int spfe_match_mutual_nn(spfe_ctx *c, const float *q, int32_t nq, const float *t, int32_t nt, int32_t *q2t, float *dist) {
  if (!c) return SPFE_ERR_INVALID;
  if (nq < 0 || nt < 0 || (nq > 0 && !q) || (nt > 0 && !t) || (nq > 0 && !q2t)) return c->fail(SPFE_ERR_INVALID, "spfe_match_mutual_nn: bad arguments");
  if (nq > c->match_cap || nt > c->match_cap) return c->fail(SPFE_ERR_INVALID, fmt("spfe_match_mutual_nn: more than %d rows", c->match_cap));
  if (nq == 0) return SPFE_OK;
  if (nt == 0) {
    for (int i = 0; i < nq; i++) { q2t[i] = -1; if (dist) dist[i] = 0.f; }
    return SPFE_OK;
  }
  std::lock_guard<std::mutex> lock(c->match_mu);
  CU_OK(c, cudaSetDevice(c->cfg.device_id));
  int rc;
  if ((rc = set_upload_locked(c, c->tmp_q, q, nq))) return rc;
  if ((rc = set_upload_locked(c, c->tmp_t, t, nt))) return rc;
  return match_sets_locked(c, c->tmp_q, c->tmp_t, q2t, dist);
}

int spfe_match_knn2(spfe_ctx *c, const float *q, int32_t nq, const float *t, int32_t nt, int32_t *idx, float *dist) {
  if (!c) return SPFE_ERR_INVALID;
  if (nq < 0 || nt < 0 || (nq > 0 && (!q || !idx || !dist)) || (nt > 0 && !t)) return c->fail(SPFE_ERR_INVALID, "spfe_match_knn2: bad arguments");
  if (nq > c->match_cap || nt > c->match_cap) return c->fail(SPFE_ERR_INVALID, fmt("spfe_match_knn2: more than %d rows", c->match_cap));
  if (nq == 0) return SPFE_OK;
  if (nt == 0) {
    for (int i = 0; i < 2 * nq; i++) { idx[i] = -1; dist[i] = 0.f; }
    return SPFE_OK;
  }
  std::lock_guard<std::mutex> lock(c->match_mu);
  CU_OK(c, cudaSetDevice(c->cfg.device_id));
  int rc;
  if ((rc = set_upload_locked(c, c->tmp_q, q, nq))) return rc;
  if ((rc = set_upload_locked(c, c->tmp_t, t, nt))) return rc;
  return knn2_sets_locked(c, c->tmp_q, c->tmp_t, idx, dist);
}

}  // extern "C"
namespace {
// Guided search on descriptors that are on the device already (sets) or still on the host (uploaded here).  All the
// small per-call arrays travel in ONE page-locked block (one H2D), the results come back in one D2H.
int guided_run(spfe_ctx *c, const spfe_guided_search *g, const spfe_desc_set *qset, const spfe_desc_set *kset, int32_t *q2kp,
               float *qdist, uint8_t *kp_taken_out, const char *fn) {
  if (!g || g->struct_size != (int32_t)sizeof(spfe_guided_search)) return c->fail(SPFE_ERR_INVALID, fmt("%s: bad struct pointer / struct_size", fn));
  const int m = g->m, n = g->n, cells = g->grid_rows * g->grid_cols;
  if (m < 0 || n < 0 || m >= (1 << 20) || g->grid_rows <= 0 || g->grid_cols <= 0 || (g->mode != SPFE_GUIDED_AREA && g->mode != SPFE_GUIDED_DUST_CELLS))
    return c->fail(SPFE_ERR_INVALID, fmt("%s: bad m / n / grid / mode", fn));
  if (qset && qset->n != m) return c->fail(SPFE_ERR_INVALID, fmt("%s: the query set holds %d rows, m = %d", fn, qset->n, m));
  if (kset && kset->n != n) return c->fail(SPFE_ERR_INVALID, fmt("%s: the key-point set holds %d rows, n = %d", fn, kset->n, n));
  if (m > 0 && ((!qset && !g->qdesc) || !g->qxy || !q2kp || !qdist || !g->occ_grid || (g->mode == SPFE_GUIDED_AREA && !g->qradius)))
    return c->fail(SPFE_ERR_INVALID, fmt("%s: NULL query / output / occ_grid pointer", fn));
  if (n > 0 && ((!kset && !g->kdesc) || (!g->kp_un && (g->mode == SPFE_GUIDED_AREA || g->c2_adaptive > 0.0f))))
    return c->fail(SPFE_ERR_INVALID, fmt("%s: NULL keypoint pointer", fn));
  if (m == 0) {
    if (kp_taken_out && n > 0) { if (g->kp_taken) memcpy(kp_taken_out, g->kp_taken, n); else memset(kp_taken_out, 0, n); }
    return SPFE_OK;
  }
  std::lock_guard<std::mutex> lock(c->match_mu);
  CU_OK(c, cudaSetDevice(c->cfg.device_id));
  cudaStream_t st = c->match_stream;
  // one device block, carved into 256-byte aligned pieces: [staged inputs][results][descriptor uploads + scratch]
  size_t off = 0;
  auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t nn = n > 0 ? n : 1;
  const size_t o_qxy = carve((size_t)m * 8), o_qr = carve((size_t)m * 4), o_qvalid = carve(m), o_qblocks = carve(m),
               o_kpun = carve(nn * 8), o_occ = carve((size_t)cells * 2), o_taken_in = carve(nn), o_kpmin = carve(nn * 4),
               o_ctl = carve(8);  // [overflow flag, ticket of the candidate phase], both 0 at launch
  const size_t in_bytes = off;
  // results: device copies the kernel works on, mirrored at the same offsets of the page-locked block, where the
  // kernel's last CTA writes them directly (mapped host memory)
  const size_t o_q2kp = carve((size_t)m * 4), o_qdist = carve((size_t)m * 4), o_taken = carve(nn), o_over = carve(8);
  const size_t out_end = off;
  const size_t o_qdesc = carve(qset ? 0 : (size_t)m * 1024), o_kdesc = carve(kset ? 0 : nn * 1024),
               o_cand = carve((size_t)m * GUIDED_CAND * 4), o_cdist = carve((size_t)m * GUIDED_CAND * 4), o_ncand = carve((size_t)m * 4), o_dec = carve(m);
  if (off > c->guided_bytes) {
    if (c->guided_buf) cudaFree(c->guided_buf);
    c->guided_buf = nullptr;
    c->guided_bytes = 0;
    CU_OK(c, cudaMalloc(&c->guided_buf, off + off / 2));
    c->guided_bytes = off + off / 2;
  }
  if (out_end > c->guided_stage_bytes) {
    if (c->guided_stage) cudaFreeHost(c->guided_stage);
    c->guided_stage = nullptr;
    c->guided_stage_bytes = 0;
    CU_OK(c, cudaMallocHost(&c->guided_stage, out_end + out_end / 2));
    c->guided_stage_bytes = out_end + out_end / 2;
  }
  uint8_t *base = static_cast<uint8_t *>(c->guided_buf), *hb = static_cast<uint8_t *>(c->guided_stage);
  memcpy(hb + o_qxy, g->qxy, (size_t)m * 8);
  if (g->qradius) memcpy(hb + o_qr, g->qradius, (size_t)m * 4);
  if (g->qvalid) memcpy(hb + o_qvalid, g->qvalid, m);
  if (g->qblocks) memcpy(hb + o_qblocks, g->qblocks, m);
  if (g->kp_un && n > 0) memcpy(hb + o_kpun, g->kp_un, (size_t)n * 8);
  memcpy(hb + o_occ, g->occ_grid, (size_t)cells * 2);
  if (g->kp_taken && n > 0) memcpy(hb + o_taken_in, g->kp_taken, n); else memset(hb + o_taken_in, 0, nn);
  memset(hb + o_kpmin, 0x7F, nn * 4);
  memset(hb + o_ctl, 0, 8);
  CU_OK(c, cudaMemcpyAsync(base, hb, in_bytes, cudaMemcpyHostToDevice, st));
  if (!qset) CU_OK(c, cudaMemcpyAsync(base + o_qdesc, g->qdesc, (size_t)m * 1024, cudaMemcpyHostToDevice, st));
  if (!kset && n > 0) CU_OK(c, cudaMemcpyAsync(base + o_kdesc, g->kdesc, (size_t)n * 1024, cudaMemcpyHostToDevice, st));
  GuidedArgs a;
  a.mode = g->mode; a.m = m; a.n = n; a.grid_rows = g->grid_rows; a.grid_cols = g->grid_cols;
  a.qdesc = qset ? qset->d32 : reinterpret_cast<float *>(base + o_qdesc); a.qxy = reinterpret_cast<float *>(base + o_qxy);
  a.qr = reinterpret_cast<float *>(base + o_qr);
  a.qvalid = g->qvalid ? base + o_qvalid : nullptr; a.qblocks = g->qblocks ? base + o_qblocks : nullptr;
  a.kdesc = kset ? kset->d32 : reinterpret_cast<float *>(base + o_kdesc); a.kp_un = reinterpret_cast<float *>(base + o_kpun);
  a.occ = reinterpret_cast<int16_t *>(base + o_occ); a.taken_in = base + o_taken_in; a.taken = base + o_taken;
  a.kpmin = reinterpret_cast<int *>(base + o_kpmin);
  a.cand = reinterpret_cast<int *>(base + o_cand); a.cdist = reinterpret_cast<float *>(base + o_cdist);
  a.ncand = reinterpret_cast<int *>(base + o_ncand); a.decided = base + o_dec;
  a.q2kp = reinterpret_cast<int *>(base + o_q2kp); a.qdist = reinterpret_cast<float *>(base + o_qdist);
  a.overflow = reinterpret_cast<int *>(base + o_ctl); a.ticket = reinterpret_cast<int *>(base + o_ctl) + 1;
  a.h_q2kp = reinterpret_cast<int *>(hb + o_q2kp); a.h_qdist = reinterpret_cast<float *>(hb + o_qdist);
  a.h_taken = hb + o_taken; a.h_overflow = reinterpret_cast<int *>(hb + o_over);
  a.min_x = g->min_x; a.min_y = g->min_y; a.best_init = g->best_init; a.th_le = g->th_le; a.th_lt = g->th_lt; a.c2 = g->c2_adaptive;
  const size_t g_smem = ((static_cast<size_t>(n) + 15) & ~size_t(15)) + static_cast<size_t>(n) * 4 + 16;
  a.smem_state = g_smem <= static_cast<size_t>(GUIDED_SMEM_MAX);
  guided_kernel<<<(m + GUIDED_THREADS / 32 - 1) / (GUIDED_THREADS / 32), GUIDED_THREADS, a.smem_state ? g_smem : 0, st>>>(a);
  c->launches += 1;
  CU_OK(c, cudaGetLastError());
  CU_OK(c, cudaStreamSynchronize(st));  // the kernel's last CTA wrote the results into the page-locked block
  memcpy(q2kp, hb + o_q2kp, (size_t)m * 4);
  memcpy(qdist, hb + o_qdist, (size_t)m * 4);
  if (kp_taken_out && n > 0) memcpy(kp_taken_out, hb + o_taken, n);
  int over = 0;
  memcpy(&over, hb + o_over, 4);
  memcpy(&c->guided_rounds, hb + o_over + 4, 4);
  if (over) return c->fail(SPFE_ERR_INVALID, fmt("%s: a query has more than %d candidate keypoints (radius too large)", fn, GUIDED_CAND));
  return SPFE_OK;
}
}  // namespace
extern "C" {

int spfe_search_guided(spfe_ctx *c, const spfe_guided_search *g, int32_t *q2kp, float *qdist, uint8_t *kp_taken_out) {
  if (!c) return SPFE_ERR_INVALID;
  return guided_run(c, g, nullptr, nullptr, q2kp, qdist, kp_taken_out, "spfe_search_guided");
}

int32_t spfe_guided_last_rounds(const spfe_ctx *c) { return c ? c->guided_rounds : SPFE_ERR_INVALID; }

int spfe_search_guided_sets(spfe_ctx *c, const spfe_guided_search *g, const spfe_desc_set *qset, const spfe_desc_set *kset,
                            int32_t *q2kp, float *qdist, uint8_t *kp_taken_out) {
  if (!c) return SPFE_ERR_INVALID;
  if ((qset && qset->ctx != c) || (kset && kset->ctx != c)) return c->fail(SPFE_ERR_INVALID, "spfe_search_guided_sets: a descriptor set belongs to another context");
  return guided_run(c, g, qset, kset, q2kp, qdist, kp_taken_out, "spfe_search_guided_sets");
}

// spfe_dust_pose_optimize[_batch] (mode 1) / spfe_dust_linearize (mode 0): `count` problems = one launch of
// dust_pose_kernel with `count` CTAs on the matcher stream.  All inputs are packed into one page-locked staging block
// (one H2D), all results come back in one D2H.
struct DustOut {  // per-problem destinations (any may be null)
  double *pose7 = nullptr;   // in (always) / out (mode 1)
  uint8_t *level = nullptr;  // in / out (mode 0)
  double *err = nullptr, *J = nullptr, *Hb = nullptr;
  float *uv = nullptr;
  uint8_t *visible = nullptr;
  int32_t *n_inlier = nullptr, *n_iter = nullptr;
};

static int dust_pose_run(spfe_ctx *c, const spfe_dust_pose *probs, int count, int mode, const DustOut *outs, const char *fn) {
  if (!c) return SPFE_ERR_INVALID;
  if (count < 0 || count > 4096 || (count > 0 && (!probs || !outs))) return c->fail(SPFE_ERR_INVALID, fmt("%s: bad count / NULL problems", fn));
  if (count == 0) return SPFE_OK;
  struct Lay { size_t xw, pose_in, dust, level_in, pose, hb, res, uv, vis, level, err, J; int rows, cols; const float *d_dust; };
  std::vector<Lay> lay(count);
  size_t off = 0;
  auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 63) / 64 * 64; return o; };
  const size_t o_args = carve((size_t)count * sizeof(DustPoseArgs));
  std::vector<cudaEvent_t> waits;
  for (int i = 0; i < count; i++) {  // validation + input region
    const spfe_dust_pose *p = probs + i;
    if (p->struct_size != (int32_t)sizeof(spfe_dust_pose)) return c->fail(SPFE_ERR_INVALID, fmt("%s: bad struct_size (problem %d)", fn, i));
    const int n = p->n;
    if (n < 0 || n >= (1 << 20) || !outs[i].pose7 || (n > 0 && !p->Xw) || (mode == 1 && p->iterations < 0))
      return c->fail(SPFE_ERR_INVALID, fmt("%s: bad n / iterations / NULL pose or Xw (problem %d)", fn, i));
    if (mode == 0 && (!outs[i].Hb || (n > 0 && (!outs[i].level || !outs[i].err || !outs[i].uv || !outs[i].J))))
      return c->fail(SPFE_ERR_INVALID, fmt("%s: NULL output pointer", fn));
    Lay &l = lay[i];
    l.rows = p->rows; l.cols = p->cols; l.d_dust = nullptr;
    if (p->dust == nullptr) {
      if (p->slot < 0 || p->slot >= (int)c->slots.size()) return c->fail(SPFE_ERR_INVALID, fmt("%s: bad slot (problem %d)", fn, i));
      Slot &s = c->slots[p->slot];
      if (s.pending) return c->fail(SPFE_ERR_STATE, fmt("%s: slot still has an un-waited batch", fn));
      if (p->frame < 0 || p->frame >= s.batch) return c->fail(SPFE_ERR_STATE, fmt("%s: frame %d is not part of the slot's last batch (%d frames)", fn, p->frame, s.batch));
      if ((l.rows && l.rows != c->hc) || (l.cols && l.cols != c->wc)) return c->fail(SPFE_ERR_INVALID, fmt("%s: rows / cols differ from the extractor's H/8 x W/8", fn));
      l.rows = c->hc; l.cols = c->wc;
      l.d_dust = s.dense_dust + (size_t)p->frame * c->cells;
      bool seen = false;
      for (cudaEvent_t e : waits) seen |= e == s.ev_done;
      if (!seen) waits.push_back(s.ev_done);
    }
    if (l.rows < 4 || l.cols < 4 || (size_t)l.rows * l.cols >= (1u << 26)) return c->fail(SPFE_ERR_INVALID, fmt("%s: bad dust map size (problem %d)", fn, i));
    const size_t nn = n > 0 ? n : 1;
    l.xw = carve(nn * 24); l.pose_in = carve(56);
    l.dust = carve(p->dust ? (size_t)l.rows * l.cols * sizeof(float) : 0);
    l.level_in = carve(mode == 0 ? nn : 0);
  }
  const size_t in_bytes = off;
  for (int i = 0; i < count; i++) {  // output region
    const size_t nn = probs[i].n > 0 ? probs[i].n : 1;
    Lay &l = lay[i];
    l.pose = carve(56); l.hb = carve(43 * 8); l.res = carve(8); l.uv = carve(nn * 8); l.vis = carve(nn); l.level = carve(nn);
    l.err = carve(nn * 8); l.J = carve(mode == 0 ? nn * 48 : 0);
  }
  const size_t total = off;
  std::lock_guard<std::mutex> lock(c->match_mu);
  CU_OK(c, cudaSetDevice(c->cfg.device_id));
  cudaStream_t st = c->match_stream;
  if (total > c->guided_bytes) {
    if (c->guided_buf) cudaFree(c->guided_buf);
    c->guided_buf = nullptr;
    c->guided_bytes = 0;
    CU_OK(c, cudaMalloc(&c->guided_buf, total + total / 2));
    c->guided_bytes = total + total / 2;
  }
  if (total > c->dust_stage_bytes) {
    if (c->dust_stage) cudaFreeHost(c->dust_stage);
    c->dust_stage = nullptr;
    c->dust_stage_bytes = 0;
    CU_OK(c, cudaMallocHost(&c->dust_stage, total + total / 2));
    c->dust_stage_bytes = total + total / 2;
  }
  uint8_t *base = static_cast<uint8_t *>(c->guided_buf), *hb = static_cast<uint8_t *>(c->dust_stage);
  DustPoseArgs *args = reinterpret_cast<DustPoseArgs *>(hb + o_args);
  size_t smem = 0;
  for (int i = 0; i < count; i++) {
    const spfe_dust_pose *p = probs + i;
    const Lay &l = lay[i];
    const int n = p->n;
    const size_t map_bytes = (size_t)l.rows * l.cols * sizeof(float);
    if (n > 0) memcpy(hb + l.xw, p->Xw, (size_t)n * 24);
    memcpy(hb + l.pose_in, outs[i].pose7, 56);
    if (p->dust) memcpy(hb + l.dust, p->dust, map_bytes);
    if (mode == 0 && n > 0) memcpy(hb + l.level_in, outs[i].level, n);
    DustPoseArgs a;
    a.dust = p->dust ? reinterpret_cast<const float *>(base + l.dust) : l.d_dust;
    a.rows = l.rows; a.cols = l.cols; a.dust_in_smem = map_bytes <= (size_t)DUST_SMEM_MAX;
    if (a.dust_in_smem && map_bytes > smem) smem = map_bytes;
    a.Xw = reinterpret_cast<const double *>(base + l.xw); a.n = n; a.mode = mode; a.iterations = p->iterations;
    a.fx = p->fx; a.fy = p->fy; a.cx = p->cx; a.cy = p->cy; a.huber = p->huber_delta; a.chi2_inlier = p->chi2_inlier;
    a.pose_in = reinterpret_cast<const double *>(base + l.pose_in);
    a.level_in = mode == 0 ? base + l.level_in : nullptr;
    a.pose = reinterpret_cast<double *>(base + l.pose); a.level = base + l.level; a.err = reinterpret_cast<double *>(base + l.err);
    a.uv = reinterpret_cast<float *>(base + l.uv); a.J = reinterpret_cast<double *>(base + l.J); a.Hb = reinterpret_cast<double *>(base + l.hb);
    a.visible = base + l.vis; a.result = reinterpret_cast<int *>(base + l.res);
    args[i] = a;
  }
  for (cudaEvent_t e : waits) CU_OK(c, cudaStreamWaitEvent(st, e, 0));  // device-resident maps: after the batch that wrote them
  CU_OK(c, cudaMemcpyAsync(base, hb, in_bytes, cudaMemcpyHostToDevice, st));
  dust_pose_kernel<<<count, DP_THREADS, smem, st>>>(reinterpret_cast<const DustPoseArgs *>(base + o_args));
  c->launches += 1;
  CU_OK(c, cudaGetLastError());
  CU_OK(c, cudaMemcpyAsync(hb + in_bytes, base + in_bytes, total - in_bytes, cudaMemcpyDeviceToHost, st));
  CU_OK(c, cudaStreamSynchronize(st));
  int thrown = -1;
  for (int i = 0; i < count; i++) {
    const Lay &l = lay[i];
    const DustOut &o = outs[i];
    const int n = probs[i].n;
    const int *res = reinterpret_cast<const int *>(hb + l.res);
    if (mode == 1) memcpy(o.pose7, hb + l.pose, 56);
    if (o.Hb) memcpy(o.Hb, hb + l.hb, (mode == 0 ? 43 : 3) * sizeof(double));
    if (o.n_inlier) *o.n_inlier = res[1];
    if (o.n_iter) *o.n_iter = res[0];
    if (n > 0) {
      if (o.uv) memcpy(o.uv, hb + l.uv, (size_t)n * 8);
      if (o.level) memcpy(o.level, hb + l.level, n);
      if (o.err) memcpy(o.err, hb + l.err, (size_t)n * 8);
      if (o.J) memcpy(o.J, hb + l.J, (size_t)n * 48);
      if (o.visible) memcpy(o.visible, hb + l.vis, n);
    }
    if (res[0] < 0 && thrown < 0) thrown = i;
  }
  if (thrown >= 0) return c->fail(SPFE_ERR_STATE, fmt("%s: an edge's linearizeOplus projection left the image ( should be omitted), problem %d", fn, thrown));
  return SPFE_OK;
}

int spfe_dust_pose_optimize(spfe_ctx *c, const spfe_dust_pose *p, double *pose7, uint8_t *visible, float *proj_uv,
                            int32_t *n_inlier, int32_t *n_iter, double *stats) {
  if (!c) return SPFE_ERR_INVALID;
  if (!p) return c->fail(SPFE_ERR_INVALID, "spfe_dust_pose_optimize: NULL problem");
  DustOut o;
  o.pose7 = pose7; o.visible = visible; o.uv = proj_uv; o.n_inlier = n_inlier; o.n_iter = n_iter; o.Hb = stats;
  return dust_pose_run(c, p, 1, 1, &o, "spfe_dust_pose_optimize");
}

int spfe_dust_pose_optimize_batch(spfe_ctx *c, const spfe_dust_pose *problems, int32_t count, double *pose7,
                                  uint8_t *const *visible, float *const *proj_uv, int32_t *n_inlier, int32_t *n_iter) {
  if (!c) return SPFE_ERR_INVALID;
  if (count < 0 || (count > 0 && (!problems || !pose7))) return c->fail(SPFE_ERR_INVALID, "spfe_dust_pose_optimize_batch: bad count / NULL problems or poses");
  std::vector<DustOut> outs(count > 0 ? count : 0);
  for (int i = 0; i < count; i++) {
    outs[i].pose7 = pose7 + 7 * (size_t)i;
    outs[i].visible = visible ? visible[i] : nullptr;
    outs[i].uv = proj_uv ? proj_uv[i] : nullptr;
    outs[i].n_inlier = n_inlier ? n_inlier + i : nullptr;
    outs[i].n_iter = n_iter ? n_iter + i : nullptr;
  }
  return dust_pose_run(c, problems, count, 1, outs.data(), "spfe_dust_pose_optimize_batch");
}

int spfe_dust_linearize(spfe_ctx *c, const spfe_dust_pose *p, const double *pose7, uint8_t *level, double *err,
                        float *proj_uv, double *J, double *Hb) {
  if (!c) return SPFE_ERR_INVALID;
  if (!p) return c->fail(SPFE_ERR_INVALID, "spfe_dust_linearize: NULL problem");
  DustOut o;
  o.pose7 = const_cast<double *>(pose7); o.level = level; o.err = err; o.uv = proj_uv; o.J = J; o.Hb = Hb;
  return dust_pose_run(c, p, 1, 0, &o, "spfe_dust_linearize");
}

int spfe_set_score_threshold(spfe_ctx *c, float score_thresh) {
  if (!c) return SPFE_ERR_INVALID;
  if (!(score_thresh >= 0.0f && score_thresh <= 1.0f)) return c->fail(SPFE_ERR_INVALID, "spfe_set_score_threshold: threshold outside [0, 1]");
  c->cfg.score_thresh = score_thresh;  // read by run_pipeline at the next submit
  return SPFE_OK;
}

int spfe_reset_stream(spfe_ctx *c, int32_t slot) {
  int rc = check_slot(c, slot);
  if (rc) return rc;
  CU_OK(c, cudaSetDevice(c->cfg.device_id));
  CU_OK(c, cudaMemsetAsync(c->slots[slot].count_all, 0, sizeof(int), c->slots[slot].stream));
  CU_OK(c, cudaStreamSynchronize(c->slots[slot].stream));  // the next frame may arrive as a graph launch on another queue
  return SPFE_OK;
}

int spfe_timer_start(spfe_ctx *c, int32_t slot) {
  int rc = check_slot(c, slot);
  if (rc) return rc;
  CU_OK(c, cudaSetDevice(c->cfg.device_id));
  CU_OK(c, cudaEventRecord(c->slots[slot].ev0, c->slots[slot].stream));
  return SPFE_OK;
}

int spfe_timer_stop(spfe_ctx *c, int32_t slot, float *ms) {
  int rc = check_slot(c, slot);
  if (rc) return rc;
  if (!ms) return c->fail(SPFE_ERR_INVALID, "spfe_timer_stop: ms is NULL");
  Slot &s = c->slots[slot];
  CU_OK(c, cudaSetDevice(c->cfg.device_id));
  CU_OK(c, cudaEventRecord(s.ev1, s.stream));
  CU_OK(c, cudaEventSynchronize(s.ev1));
  CU_OK(c, cudaEventElapsedTime(ms, s.ev0, s.ev1));
  return SPFE_OK;
}

int64_t spfe_check_weights(const char *path, char *err, size_t errcap) {
  WeightMap wm;
  std::string e;
  if (err && errcap) err[0] = 0;
  if (!path || !load_weights(path, wm, e)) {
    if (err && errcap) snprintf(err, errcap, "%s", path ? e.c_str() : "path is NULL");
    return SPFE_ERR_WEIGHTS;
  }
  int64_t n = 0;
  for (auto &kv : wm) n += (int64_t)kv.second.numel();
  return n;
}

float spfe_l2(const float *a, const float *b) {
  // cv::norm(a, b, NORM_L2) on 1x256 CV_32F rows (sp_matcher.cpp:1636-1640)
  float s = 0.0f;
  for (int i = 0; i < SPFE_DESC_DIM; i++) {
    const float d = a[i] - b[i];
    s += d * d;
  }
  return std::sqrt(s);
}

// ---- introspection --------------------------------------------------------------
int64_t spfe_debug_read(spfe_ctx *c, int32_t slot, const char *name, void *dst, size_t dst_bytes) {
  int rc = check_slot(c, slot);
  if (rc) return rc;
  if (!name || !dst) return c->fail(SPFE_ERR_INVALID, "spfe_debug_read: NULL argument");
  Slot &s = c->slots[slot];
  const size_t B = s.batch > 0 ? s.batch : 1, px = (size_t)c->H * c->W, cells = c->cells, cap = c->cap;
  const size_t xm = c->exact ? 2 : 1;  // exact mode: [hi 64 | lo 64] per 64-channel block
  struct Ent { const char *n; const void *p; size_t bytes; } tab[] = {
      {"conv1a", c->fused_conv1 ? nullptr : s.a1a, B * px * 64 * 2 * xm},        {"conv1b", s.a1b, B * px / 4 * 64 * 2 * xm},
      {"conv2a", s.a2a, B * px / 4 * 64 * 2 * xm},    {"conv2b", s.a2b, B * px / 16 * 64 * 2 * xm},
      {"conv3a", s.a3a, B * px / 16 * 128 * 2 * xm},  {"conv3b", s.a3b, B * cells * 128 * 2 * xm},
      {"conv4a", s.a4a, B * cells * 128 * 2 * xm},    {"conv4b", s.a4b, B * cells * 128 * 2 * xm},
      {"heads", s.heads, B * cells * 512 * 2 * xm},   {"coarse", s.coarse, B * cells * 256 * 2},
      {"score", s.score, B * cells * 4},         {"argmax", s.argmax, B * cells},
      {"semi_dust", s.semi_dust, B * cells * 4}, {"dense_dust", s.dense_dust, B * cells * 4},
      {"heat_log", s.heat_log, B * px * 4},      {"heat", s.heat, B * px * 4},
      {"heat_inv", s.heat_inv, B * px * 4},      {"heat_minmax", s.heat_mm_f, B * 2 * 4},
      {"count", s.count, B * 4},                 {"kp_xy", s.kp_xy, B * cap * 2 * 4},
      {"kp_score", s.kp_score, B * cap * 4},     {"desc", s.desc, B * cap * 256 * 4},
      {"occ_grid", s.occ, B * cells * 2},        {"match_prev", c->match_prev ? s.match.q2t : nullptr, B * cap * 4},
      {"match_dist", c->match_prev ? s.match.dist : nullptr, B * cap * 4},
      {"cov_qlen", s.cov_qlen, B * cap * 4},     {"cov_done", s.cov_done, B * cap * 4},
      {"cov_counters", s.cov_ctr, COV_NCTR * 4},     {"cov_replayed", s.cov_n_replay, B * 2 * 4}};
  for (const Ent &e : tab)
    if (!strcmp(e.n, name)) {
      if (!e.p) return c->fail(SPFE_ERR_STATE, fmt("spfe_debug_read: '%s' is not produced with the current flags", name));
      if (dst_bytes < e.bytes) return c->fail(SPFE_ERR_INVALID, fmt("spfe_debug_read: '%s' needs %zu bytes", name, e.bytes));
      CU_OK(c, cudaSetDevice(c->cfg.device_id));
      CU_OK(c, cudaStreamSynchronize(s.stream));
      CU_OK(c, cudaMemcpy(dst, e.p, e.bytes, cudaMemcpyDeviceToHost));
      return (int64_t)e.bytes;
    }
  return c->fail(SPFE_ERR_INVALID, fmt("spfe_debug_read: unknown tensor '%s'", name));
}

int64_t spfe_launch_count(const spfe_ctx *c) { return c ? (int64_t)c->launches.load() : 0; }

int spfe_dom_timing(spfe_ctx *c, int32_t enable) {
  if (!c) return SPFE_ERR_INVALID;
  CU_OK(c, cudaSetDevice(c->cfg.device_id));
  if (enable && c->dom_ev.empty()) {
    c->dom_ev.resize(128);
    for (cudaEvent_t &e : c->dom_ev) CU_OK(c, cudaEventCreate(&e));
  }
  c->dom_timing = enable != 0;
  c->dom_n = 0;
  return SPFE_OK;
}

int spfe_dom_time(spfe_ctx *c, float *avg_ms, int32_t *count) {
  if (!c) return SPFE_ERR_INVALID;
  if (!avg_ms || !count) return c->fail(SPFE_ERR_INVALID, "spfe_dom_time: NULL output");
  CU_OK(c, cudaSetDevice(c->cfg.device_id));
  const int n = static_cast<int>(c->dom_n < 64 ? c->dom_n : 64);
  double sum = 0;
  for (int i = 0; i < n; i++) {
    float ms = 0.f;
    CU_OK(c, cudaEventSynchronize(c->dom_ev[2 * i + 1]));
    CU_OK(c, cudaEventElapsedTime(&ms, c->dom_ev[2 * i], c->dom_ev[2 * i + 1]));
    sum += ms;
  }
  *avg_ms = n ? static_cast<float>(sum / n) : 0.f;
  *count = n;
  return SPFE_OK;
}

int spfe_profile_device(spfe_ctx *c, int32_t slot, const void *d_gray, int32_t batch, spfe_stage_time *stages, int32_t cap) {
  int rc = check_slot(c, slot);
  if (rc) return rc;
  if (!d_gray || batch < 1 || batch > c->cfg.max_batch || !stages) return c->fail(SPFE_ERR_INVALID, "spfe_profile_device: bad arguments");
  Slot &s = c->slots[slot];
  CU_OK(c, cudaSetDevice(c->cfg.device_id));
  StageTimer tm;
  tm.st = s.stream;
  tm.on = true;
  uint8_t *saved = s.d_gray;
  s.d_gray = const_cast<uint8_t *>(static_cast<const uint8_t *>(d_gray));
  rc = run_pipeline(c, s, batch, &tm);
  s.d_gray = saved;
  if (rc) return rc;
  CU_OK(c, cudaStreamSynchronize(s.stream));
  int n = 0;
  for (size_t i = 1; i < tm.ev.size() && n < cap; i++, n++) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, tm.ev[i - 1], tm.ev[i]);
    snprintf(stages[n].name, sizeof stages[n].name, "%s", tm.names[i].c_str());
    stages[n].ms = ms;
    stages[n].flop = tm.flop[i];
    stages[n].bytes = tm.bytes[i];
  }
  for (cudaEvent_t e : tm.ev) cudaEventDestroy(e);
  return n;
}

}  // extern "C"
