// libbsr.so — host side of the C ABI declared in include/bsr.h: handle, weight packing, workspace
// and the launch plan of Generator.call (/root/reference/model.py:228-290,
// /root/reference/model_with_TSM.py:261-325).  No torch types; CUDA runtime only.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>

#include "../../include/bsr.h"
#include "attention_simple.cuh"
#include "attention_fa.cuh"
#include "common.cuh"
#include "conv_direct.cuh"
#include "conv_tc.cuh"
#include "conv3x3_halo.cuh"
#include "convt_halo.cuh"
#include "glue.cuh"
#include "postprocess.cuh"

using namespace bsr;

namespace {

thread_local std::string g_create_error;

struct Layer {
  std::string name;
  int kh = 0, kw = 0, cin = 0, cout = 0, transposed = 0;
  std::vector<float> w_host, b_host;   // canonical fp32 [tap][cin][cout], [cout]
  float* w_dev = nullptr;              // same, on device (CUDA-core kernels)
  float* b_dev = nullptr;              // bias padded with zeros to a multiple of 16 (+256 slack)
  TcWeights tc;                        // h16 K-major packing + step program + tensor map (conv_tc.cuh)
  TcWeights tc_phase[4];               // wide transposed convs: one plain gather conv per sub-pixel phase
  bool phases_ready = false;
  TcWeights tc_halo;                   // the same layers packed for the fused halo kernel (convt_halo.cuh)
  bool halo_ready = false;
};

struct DebugBuf { float* dev = nullptr; size_t n = 0; };

// One captured micro-batch: the whole launch sequence of forward_mb for a fixed set of pointers / sizes.
struct GraphKey {
  const void *img, *uv, *reg, *gs, *rgb, *m22, *dif;
  int m, frame, share, in_small;
  bool operator<(const GraphKey& o) const {
    return std::tie(img, uv, reg, gs, rgb, m22, dif, m, frame, share, in_small) <
           std::tie(o.img, o.uv, o.reg, o.gs, o.rgb, o.m22, o.dif, o.m, o.frame, o.share, o.in_small);
  }
};
struct GraphEntry { cudaGraphExec_t exec = nullptr; int launches = 0, seen = 0; PlanCounters pc; };

}  // namespace

struct bsr_handle {
  int variant = 0, precision = 0, device = 0, mb = 0;
  bool loaded = false;
  bool debug_keep = false, profile = false;
  int force_direct = 0;      // BSR_FORCE_DIRECT=1: h16 storage but CUDA-core convs/attention (bring-up aid)
  std::string err;
  std::map<std::string, Layer> layers;
  // channel geometry
  int c_first = 0, c_second = 0, ld1 = 0, ld2 = 0;
  size_t es = 2;             // activation element size
  // workspace
  void* arena = nullptr;
  size_t arena_bytes = 0;
  char *X1 = nullptr, *CAT3 = nullptr, *CAT2 = nullptr, *XA = nullptr, *XB = nullptr, *T1 = nullptr, *T2 = nullptr,
       *Y = nullptr, *QK = nullptr, *VT = nullptr, *O = nullptr, *UP3 = nullptr, *F1 = nullptr, *F2 = nullptr,
       *CAT1 = nullptr, *C16 = nullptr, *PIMG = nullptr;
  float *GS32 = nullptr;
  int num_sms = 148;
  float *RAW = nullptr, *DIFGS = nullptr, *UVS = nullptr, *OFF = nullptr, *BMASK = nullptr, *DIFSMALL = nullptr,
        *SH = nullptr;
  char* TAPS = nullptr;        // bilinear tap records of the ShareLayer, [frames][1024][2] x 16 bytes
  bool taps_ready = false;     // computed for the offsets currently in OFF
  int* errflag = nullptr;    // device flag set by kernels whose mbarrier wait timed out
  int* errflag_host = nullptr;   // pinned mirror, refreshed by an async copy at the end of every forward
  cudaEvent_t ev_done = nullptr; // recorded at the end of every forward: the next forward (any stream) waits for it
  cudaStream_t last_stream = nullptr;
  bool have_done = false;
  Knobs kn;
  PlanCounters pc;
  std::map<GraphKey, GraphEntry> graphs;   // captured micro-batches (replayed when the same buffers come back)
  cudaStream_t cap_stream = nullptr;       // capture-only stream (the caller's stream may be the legacy default stream)
  int launches = 0;
  std::map<std::string, DebugBuf> dbg;
  // profiling
  std::vector<cudaEvent_t> ev;
  std::vector<const char*> ev_names;
  size_t ev_used = 0;
  std::vector<std::pair<const char*, float>> times;
  // host-path staging
  // staging of the chunk / host entry points: part of the arena, sized at bsr_create
  char* chunk_stage = nullptr;      // dense img | uv | reg | face planes of one micro-batch (bsr_forward_chunk)
  size_t chunk_stage_bytes = 0;
  bool in_small = false;   // compact host path: uv/reg arrive already resized to 32x32
  bool in_rows = false;    // fp32 host path: uv/reg arrive as the two centre rows of every 8-row band ([n][32][2][256][C])
  char* stage = nullptr;
  size_t stage_bytes = 0;
  int host_step_cap = 0;   // images per host-path chunk the staging was sized for
  // scratch of bsr_postprocess_ucb (grown on demand: not part of the forward path)
  char* pp_buf = nullptr;
  PpPlanes* pp_planes = nullptr;
  int pp_cap = 0;
  cudaStream_t own_stream = nullptr, s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  TmaEncoder tma;
};

namespace {

int configure_tc_kernels() {
  if (int r = configure_tc_kernels_conv()) return r;
  if (int r = configure_tc_kernels_attn_fa()) return r;
  if (!configure_conv3x3_halo()) return -9;
  if (!configure_convt_halo()) return -10;
  return 0;
}

int fail(bsr_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}

#define CK(h, call)                                                                              \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) return fail(h, BSR_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

inline int pad16(int c) { return (c + 15) / 16 * 16; }
constexpr int kLdY = 272;      // 257 res-block channels padded to a multiple of 16: every epilogue chunk is a full vector chunk
inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

// ---------------------------------------------------------------------------------------------
// weight blob (written by blindshadowremoval_b200/convert.py)
//   char magic[8] = "BSRW0001"; int32 variant; int32 n_layers;
//   n_layers x { char name[32]; int32 kh, kw, cin, cout, transposed, pad; uint64 w_off, b_off; }
//   then fp32 payload; offsets are from the start of the blob.
struct BlobEntry {
  char name[32];
  int32_t kh, kw, cin, cout, transposed, pad;
  uint64_t w_off, b_off;
};

struct Step {
  bsr_handle* h;
  cudaStream_t st;
  const char* name;
  Step(bsr_handle* h_, cudaStream_t st_, const char* n) : h(h_), st(st_), name(n) {
    if (h->profile && h->ev_used + 2 <= h->ev.size()) {
      cudaEventRecord(h->ev[h->ev_used], st);
    }
  }
  ~Step() {
    if (h->profile && h->ev_used + 2 <= h->ev.size()) {
      cudaEventRecord(h->ev[h->ev_used + 1], st);
      h->ev_names.push_back(name);
      h->ev_used += 2;
    }
  }
};

template <typename T>
void debug_capture_t(bsr_handle* h, cudaStream_t st, const char* name, const void* buf, int ld, int coff, int C,
                     long long npix) {
  DebugBuf& d = h->dbg[name];
  size_t n = (size_t)npix * C;
  if (d.n != n) {
    if (d.dev) cudaFree(d.dev);
    cudaMalloc(&d.dev, n * sizeof(float));
    d.n = n;
  }
  long long tot = npix * C;
  slice_to_f32_kernel<T><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>((const T*)buf, ld, coff, C, d.dev, npix);
}

void debug_capture(bsr_handle* h, cudaStream_t st, const char* name, const void* buf, int ld, int coff, int C,
                   long long npix, bool is_f32 = false) {
  if (!h->debug_keep) return;
  if (is_f32 || h->precision == BSR_PRECISION_FP32CHECK)
    debug_capture_t<float>(h, st, name, buf, ld, coff, C, npix);
  else
    debug_capture_t<h16>(h, st, name, buf, ld, coff, C, npix);
}

// ---------------------------------------------------------------------------------------------
// one convolution (any of: conv / transposed conv / 1x1) with fused epilogue
struct ConvCall {
  const char* layer;
  const void* in; int in_ld, in_coff; bool in_f32;
  int H, W;           // input spatial size
  int stride;         // 1 or 2 (ignored for transposed)
  EpiParams e;
  EpiExtra x;
};

template <typename TIn, typename T>
void launch_direct(bsr_handle* h, cudaStream_t st, const Layer& L, const ConvCall& c, int n) {
  ConvDirectParams p;
  p.in = c.in; p.in_ld = c.in_ld; p.in_coff = c.in_coff; p.cin = L.cin;
  p.H = c.H; p.W = c.W; p.N = n;
  p.kh = L.kh; p.kw = L.kw; p.transposed = L.transposed;
  if (L.transposed) {
    p.OH = 2 * c.H; p.OW = 2 * c.W; p.stride = 2; p.pad_t = p.pad_l = 0;
  } else {
    p.stride = c.stride;
    p.OH = (c.H + c.stride - 1) / c.stride;
    p.OW = (c.W + c.stride - 1) / c.stride;
    int tot_h = (p.OH - 1) * c.stride + L.kh - c.H; if (tot_h < 0) tot_h = 0;
    int tot_w = (p.OW - 1) * c.stride + L.kw - c.W; if (tot_w < 0) tot_w = 0;
    p.pad_t = tot_h / 2; p.pad_l = tot_w / 2;       // TF SAME: before = total // 2
  }
  p.w = L.w_dev; p.cout = L.cout;
  long long M = (long long)n * p.OH * p.OW;
  int nc = c.e.out_c > L.cout ? c.e.out_c : L.cout;
  dim3 grid((unsigned)((M + CD_TM - 1) / CD_TM), (unsigned)((nc + CD_TN - 1) / CD_TN));
  conv_direct_kernel<TIn, T><<<grid, 256, 0, st>>>(p, c.e);
  h->launches++;
}

int run_conv(bsr_handle* h, cudaStream_t st, ConvCall c, int n) {
  auto it = h->layers.find(c.layer);
  if (it == h->layers.end()) return fail(h, BSR_ESTATE, "layer %s missing from weight blob", c.layer);
  Layer& L = it->second;
  c.e.bias = L.b_dev;
  c.e.cout = L.cout;
  Step step(h, st, L.name.c_str());
  if (h->precision == BSR_PRECISION_FP32CHECK) {
    launch_direct<float, float>(h, st, L, c, n);
    return BSR_OK;
  }
  const bool special = L.tc.ready && (L.tc.kind == TC_HEADS || L.tc.kind == TC_CLR);
  if (!h->force_direct && L.tc.ready && !c.in_f32 && !h->kn.no_halo3 &&
      conv3x3_halo_ok(L.tc, c.in_ld, c.in_coff, c.H, c.W, c.stride, c.e)) {
    // 3x3 / stride 1, 128 -> 128 channels (res conv2): one halo tile per K block instead of nine shifted A tiles
    int rc = launch_conv3x3_halo(h->tma, L.tc, c.in, c.in_ld, c.in_coff, c.H, c.W, n, c.e, h->num_sms, h->errflag, st,
                                 &h->launches, h->kn);
    if (rc != 0) return fail(h, BSR_ECUDA, "halo conv %s: launch failed (%d): %s", c.layer, rc, h->tma.last_error.c_str());
    h->pc.halo3++;
    return BSR_OK;
  }
  if (!h->force_direct && L.tc.ready && !c.in_f32 && !h->kn.no_halo3 && convt_halo_ok(L.tc, c.in_ld, c.in_coff, c.H, c.W, c.e)) {
    // fused 4-phase transposed conv with streamed weights (up1, up2, clr_up2): one 17x9-pixel halo tile per K block
    int rc = launch_convt_halo(h->tma, L.tc, c.in, c.in_ld, c.in_coff, c.H, c.W, n, c.e, h->num_sms, h->errflag, st,
                               &h->launches, h->kn);
    if (rc != 0) return fail(h, BSR_ECUDA, "halo transposed conv %s: launch failed (%d): %s", c.layer, rc, h->tma.last_error.c_str());
    h->pc.halo3++;
    return BSR_OK;
  }
  if (!h->force_direct && L.tc.ready && !c.in_f32 && (!special || c.x.gs_f32 != nullptr)) {
    int rc = launch_conv_tc(h->tma, L.tc, c.in, c.in_ld, c.in_coff, c.H, c.W, c.stride, n, c.e, c.x, -1, h->num_sms,
                            h->errflag, st, &h->launches, h->kn, &h->pc);
    if (rc != 0) return fail(h, BSR_ECUDA, "tensor-core conv %s: launch failed (%d): %s", c.layer, rc,
                             h->tma.last_error.c_str());
    return BSR_OK;
  }
  if (!h->force_direct && L.halo_ready && !c.in_f32 && !h->kn.no_halo3 &&
      convt_halo_ok(L.tc_halo, c.in_ld, c.in_coff, c.H, c.W, c.e)) {
    // wide transposed conv (clr_up1): ONE fused launch on halo tiles instead of one gather conv per sub-pixel phase
    int rc = launch_convt_halo(h->tma, L.tc_halo, c.in, c.in_ld, c.in_coff, c.H, c.W, n, c.e, h->num_sms, h->errflag, st,
                               &h->launches, h->kn);
    if (rc != 0) return fail(h, BSR_ECUDA, "halo transposed conv %s: launch failed (%d): %s", c.layer, rc, h->tma.last_error.c_str());
    h->pc.halo3++;
    return BSR_OK;
  }
  if (!h->force_direct && L.phases_ready && !c.in_f32) {
    for (int ph = 0; ph < 4; ++ph) {
      int rc = launch_conv_tc(h->tma, L.tc_phase[ph], c.in, c.in_ld, c.in_coff, c.H, c.W, c.stride, n, c.e, c.x, ph,
                              h->num_sms, h->errflag, st, &h->launches, h->kn, &h->pc);
      if (rc != 0) return fail(h, BSR_ECUDA, "tensor-core conv %s phase %d: launch failed (%d): %s", c.layer, ph, rc,
                               h->tma.last_error.c_str());
    }
    return BSR_OK;
  }
  // CUDA-core convolutions inside the 16-bit path exist for bring-up only (BSR_FORCE_DIRECT=1, BSR_TC_DISABLE=<layer>);
  // the product path never falls back to them silently
  if (!h->force_direct && !tc_disabled(L.name))
    return fail(h, BSR_EUNSUPPORTED, "layer %s has no tensor-core packing (cin %d cout %d k %dx%d%s); set BSR_FORCE_DIRECT=1 "
                "for the CUDA-core bring-up path", c.layer, L.cin, L.cout, L.kh, L.kw, L.transposed ? " transposed" : "");
  if (c.in_f32) launch_direct<float, h16>(h, st, L, c, n);
  else launch_direct<h16, h16>(h, st, L, c, n);
  return BSR_OK;
}

EpiExtra no_extra() {
  EpiExtra x;
  memset(&x, 0, sizeof x);
  return x;
}

EpiParams epi(void* out, int out_ld, int out_coff, int out_c, int act, int mode = OUT_T) {
  EpiParams e;
  memset(&e, 0, sizeof e);
  e.out = out; e.out_ld = out_ld; e.out_coff = out_coff; e.out_c = out_c; e.act = act; e.out_mode = mode;
  return e;
}

int run_attention(bsr_handle* h, cudaStream_t st, int n, const Layer* wl = nullptr, const EpiParams* ew = nullptr) {
  Step step(h, st, wl ? "attention+w" : "attention");
  if (h->precision == BSR_PRECISION_FP32CHECK) {
    attention_simple_kernel<float><<<dim3(AS_S / AS_Q, n), 256, kAttnSimpleSmem, st>>>(
        (const float*)h->QK, (const float*)h->VT, (float*)h->O, 128);
    h->launches++;
    return BSR_OK;
  }
  if (!h->force_direct) {
    // single-pass kernel (attention_fa.cuh): O = softmax(QK^T) V, or with `wl` the whole rest of the block:
    // out = LeakyReLU(x_in + y + W_w . O + b)
    EpiParams e2;
    if (wl) { e2 = *ew; e2.bias = wl->b_dev; e2.cout = wl->cout; }
    int rc = launch_attention_fa(h->tma, (const h16*)h->QK, (const h16*)h->VT, (h16*)h->O, n, h->num_sms, h->errflag, st, h->kn,
                                 wl ? (const h16*)wl->tc.dev : nullptr, wl ? &e2 : nullptr, h->launches);
    if (rc != 0) return fail(h, BSR_ECUDA, "tensor-core attention launch failed (%d): %s", rc, h->tma.last_error.c_str());
    h->launches++;
    h->pc.attn_fused += wl != nullptr;
    return BSR_OK;
  }
  // BSR_FORCE_DIRECT=1 (bring-up): CUDA-core attention on V^T
  attention_simple_kernel<h16><<<dim3(AS_S / AS_Q, n), 256, kAttnSimpleSmem, st>>>(
      (const h16*)h->QK, (const h16*)h->VT, (h16*)h->O, 128);
  h->launches++;
  return BSR_OK;
}

// ResBottleneck + NonLocalBlock (model.py:98-113, 23-61).  cur: [n,1024,ld] with c_cur live channels.
template <typename T>
int run_res_block(bsr_handle* h, cudaStream_t st, int idx, char* cur, char* nxt, int ld, int c_cur, int n) {
  char nm[5][32];
  snprintf(nm[0], 32, "res%d.conv1", idx);
  snprintf(nm[1], 32, "res%d.conv2", idx);
  snprintf(nm[2], 32, "res%d.conv3", idx);
  snprintf(nm[3], 32, "res%d.qkv", idx);
  snprintf(nm[4], 32, "res%d.w", idx);
  int rc;
  const int ldy = kLdY;
  ConvCall c1{nm[0], cur, ld, 0, false, FEAT, FEAT, 1, epi(h->T1, 128, 0, 128, 1), no_extra()};
  if ((rc = run_conv(h, st, c1, n))) return rc;
  ConvCall c2{nm[1], h->T1, 128, 0, false, FEAT, FEAT, 1, epi(h->T2, 128, 0, 128, 1), no_extra()};
  if ((rc = run_conv(h, st, c2, n))) return rc;
  ConvCall c3{nm[2], h->T2, 128, 0, false, FEAT, FEAT, 1, epi(h->Y, ldy, 0, ldy, 0), no_extra()};
  if ((rc = run_conv(h, st, c3, n))) return rc;
  EpiParams eq = epi(h->QK, 256, 0, 384, 0, OUT_QKV);
  eq.out2 = h->VT;
  eq.spatial = FEAT * FEAT;
  // g leaves the projection conv untransposed ([pix][128]) for the single-pass attention kernel (MN-major B operand of
  // its P V MMA); the fp32 check kernels read V^T
  const bool v_natural = h->precision != BSR_PRECISION_FP32CHECK && !h->force_direct;
  eq.v_natural = v_natural ? 1 : 0;
  ConvCall c4{nm[3], h->Y, ldy, 0, false, FEAT, FEAT, 1, eq, no_extra()};
  if ((rc = run_conv(h, st, c4, n))) return rc;
  if (h->debug_keep && (idx == 0 || idx == 5)) {
    // direct view of the attention core for tests: the projections it reads and its un-fused output O = softmax(QK^T).V
    // (model.py:51-53); these extra launches are not counted
    const int launches0 = h->launches;
    const PlanCounters pc0 = h->pc;
    char nb[3][16];
    snprintf(nb[0], 16, "qk%d", idx); snprintf(nb[1], 16, "vt%d", idx); snprintf(nb[2], 16, "attn_o%d", idx);
    debug_capture(h, st, nb[0], h->QK, 256, 0, 256, (long long)n * 1024);
    if (v_natural) debug_capture(h, st, nb[1], h->VT, 128, 0, 128, (long long)n * 1024);      // V[n][1024][128]
    else debug_capture(h, st, nb[1], h->VT, 1024, 0, 1024, (long long)n * 128);                  // V^T[n][128][1024]
    if ((rc = run_attention(h, st, n))) return rc;
    debug_capture(h, st, nb[2], h->O, 128, 0, 128, (long long)n * 1024);
    h->launches = launches0;
    h->pc = pc0;
  }
  int oc = ld < ldy ? ld : ldy;
  EpiParams ew = epi(nxt, ld, 0, oc, 1);
  // residual widths are the PADDED widths: padding channels of Y / cur are kept at zero, so adding them is exact
  // and keeps every 16-channel chunk on the vectorised path
  ew.res1 = h->Y; ew.res1_ld = ldy; ew.res1_c = ldy;
  ew.res2 = cur; ew.res2_ld = ld; ew.res2_c = ld;
  (void)c_cur;
  // NonLocal output conv + block tail: fused into the attention kernel on the tensor-core path
  const Layer& wl = h->layers[nm[4]];
  const bool fuse_w = h->precision == BSR_PRECISION_BF16 && !h->force_direct && wl.tc.ready && wl.tc.kind == TC_CONV &&
                      wl.tc.bn == 144 && wl.tc.n_tiles == 2 && wl.cin == 128 && oc == kLdY && !h->kn.no_fuse_w && wl.tc.b_total_rows == 288 &&
                      ld % 8 == 0;
  if (fuse_w) {
    if ((rc = run_attention(h, st, n, &wl, &ew))) return rc;
  } else {
    if ((rc = run_attention(h, st, n))) return rc;
    ConvCall c5{nm[4], h->O, 128, 0, false, FEAT, FEAT, 1, ew, no_extra()};
    if ((rc = run_conv(h, st, c5, n))) return rc;
  }
  if (ld > oc) {
    Step step(h, st, "res_tail");
    long long npix = (long long)n * FEAT * FEAT, tot = npix * (ld - oc);
    if (std::is_same<T, h16>::value && ld % 8 == 0 && oc % 8 == 0)
      res_tail_vec_kernel<<<(unsigned)((tot / 8 + 255) / 256), 256, 0, st>>>((const h16*)cur, ld, (h16*)nxt, ld, oc, ld, npix);
    else
      res_tail_kernel<T><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>((const T*)cur, ld, (T*)nxt, ld, oc, ld, npix);
    h->launches++;
  }
  return BSR_OK;
}

template <typename T>
int run_share(bsr_handle* h, cudaStream_t st, char* x, int ld, int C, int coff, int n, int frame, int share) {
  Step step(h, st, "share_layer");
  long long npix = (long long)n * FEAT * FEAT;
  if (!share) {
    long long tot = npix * C;
    share_dup_kernel<T><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>((T*)x, ld, C, coff, npix);
    h->launches++;
    return BSR_OK;
  }
  if ((ld & 3) || (coff & 3)) return fail(h, BSR_EINVAL, "share layer needs 4-channel aligned stride / offset (ld %d, coff %d)", ld, coff);
  const int chunks = n / frame, ldsh = (2 * C + 3) / 4 * 4;
  if (std::is_same<T, h16>::value && ld % 8 == 0 && !h->kn.share_v1) {
    // 16-bit storage: tap records once per forward, eight lanes per cell, shared features in the alignment of their
    // destination (see glue.cuh)
    if (!h->taps_ready) {
      const int n_rec = n * FEAT * FEAT * 2;
      share_taps_kernel<<<(n_rec + 255) / 256, 256, 0, st>>>(h->OFF, (TapRec*)h->TAPS, n_rec);
      h->launches++;
      h->taps_ready = true;
    }
    const int shift = coff & 7, ldsh16 = (shift + 2 * C + 7) / 8 * 8;
    const int cells1 = chunks * FEAT * FEAT, cells2 = n * FEAT * FEAT;        // eight lanes per cell, 32 cells per block
    share_reduce_h16_kernel<<<(cells1 + 31) / 32, 256, (size_t)32 * ldsh16 * 2, st>>>((const h16*)x, ld, C, (const TapRec*)h->TAPS, frame, (h16*)h->SH,
                                                                ldsh16, shift, cells1);
    share_out_h16_kernel<<<(cells2 + 31) / 32, 256, 0, st>>>((const h16*)h->SH, ldsh16, shift, 2 * C, (const TapRec*)h->TAPS, frame,
                                                             (h16*)x, ld, coff, cells2);
    h->launches += 2;
    return BSR_OK;
  }
  const int cells1 = chunks * FEAT * FEAT, cells2 = n * FEAT * FEAT;          // one warp per cell, 8 warps per block
  share_reduce_kernel<T><<<(cells1 + 7) / 8, 256, 0, st>>>((const T*)x, ld, C, h->OFF, frame, (T*)h->SH, ldsh, cells1);
  share_out_kernel<T><<<(cells2 + 7) / 8, 256, 0, st>>>((const T*)h->SH, ldsh, 2 * C, h->OFF, frame, (T*)x, ld, coff, cells2);
  h->launches += 2;
  return BSR_OK;
}

// one micro-batch
template <typename T>
int forward_mb(bsr_handle* h, cudaStream_t st, const float* img, const float* uv, const float* reg, int n, int frame,
               int share, float* gs, float* rgb, float* mask22, float* dif) {
  const bool tsm = h->variant == BSR_VARIANT_TSM;
  int rc;
  const long long px256 = (long long)n * IMG * IMG, px32 = (long long)n * FEAT * FEAT;
  // ---- encoder (model.py:230-233)
  if (h->precision == BSR_PRECISION_BF16 && !h->force_direct && h->layers["conv1"].tc.ready) {
    {
      Step step(h, st, "pack_img");
      const long long tot = (long long)n * (IMG + 1) * (IMG + 8);
      pack_img_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(img, (h16*)h->PIMG, n);
      h->launches++;
    }
    ConvCall cv1{"conv1", h->PIMG, 8, 0, false, IMG, IMG, 1, epi(h->X1, 32, 0, 32, 1), no_extra()};
    if ((rc = run_conv(h, st, cv1, n))) return rc;
  } else {
    ConvCall cv1{"conv1", img, 3, 0, true, IMG, IMG, 1, epi(h->X1, 32, 0, 32, 1), no_extra()};
    if ((rc = run_conv(h, st, cv1, n))) return rc;
  }
  debug_capture(h, st, "x1", h->X1, 32, 0, 32, px256);
  ConvCall d1{"down1", h->X1, 32, 0, false, IMG, IMG, 2, epi(h->CAT3, 128, 64, 64, 1), no_extra()};
  if ((rc = run_conv(h, st, d1, n))) return rc;
  debug_capture(h, st, "x2", h->CAT3, 128, 64, 64, (long long)n * 128 * 128);
  ConvCall d2{"down2", h->CAT3, 128, 64, false, 128, 128, 2, epi(h->CAT2, 160, 96, 64, 1), no_extra()};
  if ((rc = run_conv(h, st, d2, n))) return rc;
  debug_capture(h, st, "x3", h->CAT2, 160, 96, 64, (long long)n * 64 * 64);
  const int ld1 = h->ld1, ld2 = h->ld2;
  ConvCall d3{"down3", h->CAT2, 160, 96, false, 64, 64, 2, epi(h->XA, ld1, 0, 96, 1), no_extra()};
  if ((rc = run_conv(h, st, d3, n))) return rc;
  // ---- uv / registration fields at 32x32 (model.py:237; warp.py:137)
  {
    Step step(h, st, "uv_small");
    int tot = n * FEAT * FEAT * 3;
    if (h->in_small) {
      CK(h, cudaMemcpyAsync(h->UVS, uv, (size_t)tot * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else if (h->in_rows) {
      uv_rows_small_kernel<<<(tot + 255) / 256, 256, 0, st>>>(uv, h->UVS, n);
      h->launches++;
    } else {
      uv_small_kernel<<<(tot + 255) / 256, 256, 0, st>>>(uv, h->UVS, n);
      h->launches++;
    }
    if (tsm) {
      int tot4 = n * FEAT * FEAT * 4;
      if (h->in_small) reg32_off_kernel<<<(tot4 + 255) / 256, 256, 0, st>>>(reg, h->OFF, n);
      else if (h->in_rows) reg_rows_small_kernel<<<(tot4 + 255) / 256, 256, 0, st>>>(reg, h->OFF, n);
      else reg_small_kernel<<<(tot4 + 255) / 256, 256, 0, st>>>(reg, h->OFF, n);
      h->launches++;
      h->taps_ready = false;
    }
  }
  int c_cur = h->c_first;
  if (tsm) {
    if ((rc = run_share<T>(h, st, h->XA, ld1, 96, 96, n, frame, share))) return rc;   // model_with_TSM.py:271
  }
  {
    Step step(h, st, "assemble_uv");
    int uv_off = c_cur - 3;
    if (std::is_same<T, h16>::value && ld1 % 8 == 0) {
      long long tot = px32 * ((ld1 >> 3) - (uv_off >> 3));
      assemble_uv_vec_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>((h16*)h->XA, ld1, h->UVS, uv_off, (int)px32);
    } else {
      long long tot = px32 * (3 + (ld1 - c_cur));
      assemble_uv_kernel<T><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>((T*)h->XA, ld1, h->UVS, uv_off, c_cur, ld1,
                                                                           (int)px32);
    }
    h->launches++;
  }
  debug_capture(h, st, "x_in0", h->XA, ld1, 0, c_cur, px32);
  // ---- res blocks 0-2 (model.py:239-240)
  char *cur = h->XA, *nxt = h->XB;
  static const char* res_names[6] = {"res0", "res1", "res2", "res3", "res4", "res5"};
  for (int i = 0; i < 3; ++i) {
    if ((rc = run_res_block<T>(h, st, i, cur, nxt, ld1, c_cur, n))) return rc;
    if (c_cur < 257) c_cur = 257;
    std::swap(cur, nxt);
    debug_capture(h, st, res_names[i], cur, ld1, 0, c_cur, px32);
  }
  // ---- grey decoder (model.py:243-252)
  ConvCall u1{"up1", cur, ld1, 0, false, FEAT, FEAT, 2, epi(h->CAT2, 160, 0, 96, 1), no_extra()};
  if ((rc = run_conv(h, st, u1, n))) return rc;
  debug_capture(h, st, "up1", h->CAT2, 160, 0, 96, (long long)n * 64 * 64);
  ConvCall u2{"up2", h->CAT2, 160, 0, false, 64, 64, 2, epi(h->CAT3, 128, 0, 64, 1), no_extra()};
  if ((rc = run_conv(h, st, u2, n))) return rc;
  debug_capture(h, st, "up2", h->CAT3, 128, 0, 64, (long long)n * 128 * 128);
  ConvCall u3{"up3", h->CAT3, 128, 0, false, 128, 128, 2, epi(h->UP3, 64, 0, 64, 1), no_extra()};
  if ((rc = run_conv(h, st, u3, n))) return rc;
  debug_capture(h, st, "up3", h->UP3, 64, 0, 64, px256);
  const bool fused_tail = h->precision == BSR_PRECISION_BF16 && !h->force_direct && h->layers["heads"].tc.ready &&
                          h->layers["clr_conv1"].tc.ready;
  if (fused_tail) {
    // conv2|conv3 + tanh/grey composition fused in one kernel (model.py:246-252)
    ConvCall hd{"heads", h->UP3, 64, 0, false, IMG, IMG, 1, epi(nullptr, 0, 0, 2, 0, OUT_F32), no_extra()};
    hd.x.img = img; hd.x.gs_out = gs; hd.x.mask22_out = mask22; hd.x.difgs = h->DIFGS; hd.x.gs_f32 = h->GS32;
    if ((rc = run_conv(h, st, hd, n))) return rc;
  } else {
    ConvCall hd{"heads", h->UP3, 64, 0, false, IMG, IMG, 1, epi(h->RAW, 2, 0, 2, 0, OUT_F32), no_extra()};
    if ((rc = run_conv(h, st, hd, n))) return rc;
    Step step(h, st, "compose");
    compose_kernel<T><<<(unsigned)((px256 + 255) / 256), 256, 0, st>>>(h->RAW, img, gs, mask22, h->DIFGS, (T*)h->CAT1,
                                                                       72, 64, px256);
    h->launches++;
  }
  // ---- hole mask + second-half input (model.py:256-259; model_with_TSM.py:290-294)
  {
    Step step(h, st, "hole");
    int cx = c_cur;
    int uv_off = tsm ? cx + 1 + 2 * cx : cx + 1;
    int cells = (int)px32;
    // same channel stride in both halves (GSC): the mask is applied in place and the 257 feature channels are not copied
    const bool inplace = ld1 == ld2 && !h->kn.no_hole_inplace;
    if (inplace && std::is_same<T, h16>::value && ld2 % 8 == 0)
      hole_inplace_h16_kernel<<<(cells + 255) / 256, 256, 0, st>>>(h->DIFGS, (h16*)cur, ld2, cx, h->UVS, uv_off, uv_off + 3, ld2,
                                                                   h->BMASK, h->DIFSMALL, cells);
    else
      hole_kernel<T><<<(cells + 3) / 4, 128, 0, st>>>(h->DIFGS, (const T*)cur, ld1, (T*)(inplace ? cur : nxt), ld2, cx, h->UVS,
                                                      uv_off, uv_off + 3, ld2, h->BMASK, h->DIFSMALL, cells);
    h->launches++;
    // otherwise the hole kernel wrote into `nxt` viewed with ld2; from here on (cur, nxt) = (that buffer, the other)
    if (!inplace) std::swap(cur, nxt);
  }
  if (tsm) {
    if ((rc = run_share<T>(h, st, cur, ld2, c_cur, c_cur + 1, n, frame, share))) return rc;   // model_with_TSM.py:293
  }
  c_cur = h->c_second;
  debug_capture(h, st, "x_in3", cur, ld2, 0, c_cur, px32);
  debug_capture(h, st, "bmask", h->BMASK, 1, 0, 1, px32, true);
  debug_capture(h, st, "dif_small", h->DIFSMALL, 1, 0, 1, px32, true);
  for (int i = 3; i < 6; ++i) {
    if ((rc = run_res_block<T>(h, st, i, cur, nxt, ld2, c_cur, n))) return rc;
    std::swap(cur, nxt);
    debug_capture(h, st, res_names[i], cur, ld2, 0, c_cur, px32);
  }
  // ---- colour decoder (model.py:264-269, 288)
  ConvCall k1{"clr_up1", cur, ld2, 0, false, FEAT, FEAT, 2, epi(h->F1, 128, 0, 128, 1), no_extra()};
  if ((rc = run_conv(h, st, k1, n))) return rc;
  debug_capture(h, st, "clr_up1", h->F1, 128, 0, 128, (long long)n * 64 * 64);
  ConvCall k2{"clr_up2", h->F1, 128, 0, false, 64, 64, 2, epi(h->F2, 96, 0, 96, 1), no_extra()};
  if ((rc = run_conv(h, st, k2, n))) return rc;
  debug_capture(h, st, "clr_up2", h->F2, 96, 0, 96, (long long)n * 128 * 128);
  if (fused_tail) {
    ConvCall k3{"clr_up3", h->F2, 96, 0, false, 128, 128, 2, epi(h->CAT1, 64, 0, 64, 1), no_extra()};
    if ((rc = run_conv(h, st, k3, n))) return rc;
    debug_capture(h, st, "clr_up3", h->CAT1, 64, 0, 64, px256);
    // clr_conv1 (f on tensor cores + fp32 gs taps) + clr_conv2 + clr_conv3 + final dif in one kernel
    ConvCall kc{"clr_conv1", h->CAT1, 64, 0, false, IMG, IMG, 1, epi(nullptr, 0, 0, 16, 1), no_extra()};
    kc.x.img = img; kc.x.gs_f32 = h->GS32; kc.x.rgb_out = rgb; kc.x.dif_out = dif;
    if ((rc = run_conv(h, st, kc, n))) return rc;
  } else {
    ConvCall k3{"clr_up3", h->F2, 96, 0, false, 128, 128, 2, epi(h->CAT1, 72, 0, 64, 1), no_extra()};
    if ((rc = run_conv(h, st, k3, n))) return rc;
    debug_capture(h, st, "clr_up3", h->CAT1, 72, 0, 64, px256);
    ConvCall kc{"clr_conv1", h->CAT1, 72, 0, false, IMG, IMG, 1, epi(h->C16, 16, 0, 16, 1), no_extra()};
    if ((rc = run_conv(h, st, kc, n))) return rc;
    Step step(h, st, "clr_tail");
    Layer& l2 = h->layers["clr_conv2"];
    Layer& l3 = h->layers["clr_conv3"];
    clr_tail_kernel<T><<<(unsigned)((px256 + 127) / 128), 128, 0, st>>>((const T*)h->C16, 16, l2.w_dev, l2.b_dev,
                                                                        l3.w_dev, l3.b_dev, img, rgb, dif, px256);
    h->launches++;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, BSR_ECUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  return BSR_OK;
}

// RAII: make the handle's device current, restore the caller's device on every exit path.
struct DeviceScope {
  int prev = -1, want;
  bool ok = true;
  explicit DeviceScope(int dev) : want(dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; prev = -1; return; }
    if (prev != want && cudaSetDevice(want) != cudaSuccess) ok = false;
  }
  ~DeviceScope() { if (prev >= 0 && prev != want) cudaSetDevice(prev); }
};

// A watchdog that fired in an EARLIER forward of this handle (its flag copy has landed in pinned memory): report it
// once, before launching anything else on top of garbage.
int pending_device_error(bsr_handle* h) {
  const int flag = *(volatile int*)h->errflag_host;
  if (!flag) return BSR_OK;
  *(volatile int*)h->errflag_host = 0;
  cudaMemsetAsync(h->errflag, 0, sizeof(int), h->last_stream);
  return fail(h, BSR_EDEVICE, "device watchdog: an mbarrier wait timed out in an earlier forward (code %d); its outputs are invalid", flag);
}

// End of every forward: mirror the device error flag into pinned host memory and mark the workspace busy until here.
int finish_forward(bsr_handle* h, cudaStream_t st) {
  CK(h, cudaMemcpyAsync(h->errflag_host, h->errflag, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(h, cudaEventRecord(h->ev_done, st));
  h->last_stream = st;
  h->have_done = true;
  return BSR_OK;
}

void clear_graphs(bsr_handle* h) {
  for (auto& kv : h->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  h->graphs.clear();
}

int forward_common(bsr_handle* h, const float* img, const float* uv, const float* reg, int n, int frame, int share,
                   float* gs, float* rgb, float* mask22, float* dif, cudaStream_t st) {
  if (!h) return BSR_EINVAL;
  if (!h->loaded) return fail(h, BSR_ESTATE, "bsr_load_weights has not been called");
  if (n <= 0 || !img || !uv) return fail(h, BSR_EINVAL, "n must be > 0 and img/uv non-NULL");
  const bool tsm = h->variant == BSR_VARIANT_TSM;
  if (tsm && (!reg || frame <= 0 || n % frame)) return fail(h, BSR_EINVAL, "TSM needs reg and n %% frame == 0");
  if (tsm && frame > h->mb) return fail(h, BSR_EINVAL, "frame %d exceeds micro_batch %d", frame, h->mb);
  if (int rc = pending_device_error(h)) return rc;
  DeviceScope dev_scope(h->device);
  if (!dev_scope.ok) return fail(h, BSR_ECUDA, "cudaSetDevice(%d) failed", h->device);
  // forwards of one handle share the workspace: wait (device side) for the previous one, whatever stream it ran on
  if (h->have_done && h->last_stream != st) CK(h, cudaStreamWaitEvent(st, h->ev_done, 0));
  h->launches = 0;
  h->pc = PlanCounters();
  h->ev_used = 0;
  h->ev_names.clear();
  int step = h->mb;
  if (tsm) step = h->mb / frame * frame;
  for (int i0 = 0; i0 < n; i0 += step) {
    int m = n - i0 < step ? n - i0 : step;
    const size_t o3 = (size_t)i0 * IMG * IMG * 3, o1 = (size_t)i0 * IMG * IMG;
    // uv/reg pixels per image: 32 x 32 (compact), 32 x 2 x 256 (centre rows only), or the full 256 x 256
    const size_t aux_px = h->in_small ? (size_t)i0 * FEAT * FEAT : (h->in_rows ? (size_t)i0 * FEAT * 2 * IMG : o1);
    int rc;
    const float* uvp = uv + aux_px * 3;
    const float* regp = reg ? reg + aux_px * 6 : nullptr;
    float *gsp = gs ? gs + o1 : nullptr, *rgbp = rgb ? rgb + o3 : nullptr, *m22p = mask22 ? mask22 + o3 : nullptr,
          *difp = dif ? dif + o1 : nullptr;
    auto run = [&](cudaStream_t s) {
      return h->precision == BSR_PRECISION_FP32CHECK
                 ? forward_mb<float>(h, s, img + o3, uvp, regp, m, frame, share, gsp, rgbp, m22p, difp)
                 : forward_mb<h16>(h, s, img + o3, uvp, regp, m, frame, share, gsp, rgbp, m22p, difp);
    };
    // CUDA graphs: a micro-batch whose buffers were seen before is captured once (on a private stream) and replayed
    // afterwards - one launch instead of ~50 (debug / profiling handles always launch kernel by kernel)
    if (!h->kn.no_graph && !h->debug_keep && !h->profile) {
      GraphKey key{img + o3, uvp, regp, gsp, rgbp, m22p, difp, m, frame, share, h->in_small ? 1 : (h->in_rows ? 2 : 0)};
      if (h->graphs.size() > 64 && !h->graphs.count(key)) clear_graphs(h);
      GraphEntry& ge = h->graphs[key];
      if (!ge.exec && ++ge.seen >= 2) {
        const int l0 = h->launches;
        const PlanCounters p0 = h->pc;
        cudaGraph_t g = nullptr;
        bool ok = cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
          const int crc = run(h->cap_stream);
          ok = cudaStreamEndCapture(h->cap_stream, &g) == cudaSuccess && crc == BSR_OK && g != nullptr;
        }
        if (ok) ok = cudaGraphInstantiate(&ge.exec, g, 0) == cudaSuccess;
        if (g) cudaGraphDestroy(g);
        if (ok) {
          ge.launches = h->launches - l0;
          ge.pc.resident = h->pc.resident - p0.resident; ge.pc.pinned = h->pc.pinned - p0.pinned;
          ge.pc.staged = h->pc.staged - p0.staged; ge.pc.attn_fused = h->pc.attn_fused - p0.attn_fused;
        } else {                       // capture unsupported here: fall back to plain launches for good
          cudaGetLastError();
          ge.exec = nullptr;
          h->kn.no_graph = 1;
        }
        h->launches = l0;
        h->pc = p0;
      }
      if (ge.exec) {
        CK(h, cudaGraphLaunch(ge.exec, st));
        h->launches += ge.launches;
        h->pc.resident += ge.pc.resident; h->pc.pinned += ge.pc.pinned; h->pc.staged += ge.pc.staged;
        h->pc.attn_fused += ge.pc.attn_fused;
        h->pc.graph_replays++;
        continue;
      }
    }
    rc = run(st);
    if (rc) return rc;
  }
  return finish_forward(h, st);
}

// Host-buffer path: micro-batches are pipelined over three streams (H2D | compute | D2H) with two staging
// slots, so PCIe transfers in both directions overlap the kernels of the neighbouring micro-batches.
//
// Compact mode (img_u8 != NULL): the image arrives as uint8 and is expanded on the compute
// stream, uv/reg arrive at 32x32, and rgb/dif can leave as uint8 / binary16 (converted on the compute stream).
struct HostCompact {
  const unsigned char* img_u8 = nullptr;
  unsigned char* rgb_u8 = nullptr;
  unsigned short* dif_f16 = nullptr;
};
// Staging slots of the pipelined host path for chunks of `step` images: fp32 [img | uv | reg] in, [gs | rgb | mask22 |
// dif] out, and the compact forms [img_u8 | uv32 | reg32] in, [rgb_u8 | dif_f16] out; two slots of each.
struct HostSlots {
  size_t in_slot, out_slot, cin_slot, cout_slot;
  size_t total() const { return 2 * (in_slot + out_slot + cin_slot + cout_slot); }
};
HostSlots host_slots(int step, bool with_reg) {
  const size_t p1 = (size_t)IMG * IMG * sizeof(float), px = (size_t)IMG * IMG, px32 = (size_t)FEAT * FEAT;
  HostSlots s;
  s.in_slot = (size_t)step * p1 * (3 + 3 + (with_reg ? 6 : 0));
  s.out_slot = (size_t)step * p1 * (1 + 3 + 3 + 1);
  s.cin_slot = (size_t)step * (px * 3 + px32 * 3 * sizeof(float) + px32 * 6 * sizeof(float));
  s.cout_slot = (size_t)step * (px * 3 + px * 2);
  return s;
}
struct SmallInputScope {      // forward_mb reads uv/reg as 32x32 maps only while a compact call is in flight
  bsr_handle* h;
  SmallInputScope(bsr_handle* hh, bool small, bool rows) : h(hh) { h->in_small = small; h->in_rows = rows; }
  ~SmallInputScope() { h->in_small = false; h->in_rows = false; }
};

int forward_host(bsr_handle* h, const float* img, const float* uv, const float* reg, int n, int frame, int share,
                 float* gs, float* rgb, float* mask22, float* dif, HostCompact cp = HostCompact()) {
  if (!h) return BSR_EINVAL;
  if (n <= 0) return fail(h, BSR_EINVAL, "n must be > 0");
  if (!h->loaded) return fail(h, BSR_ESTATE, "bsr_load_weights has not been called");
  const bool compact = cp.img_u8 != nullptr;
  if ((!compact && !img) || !uv) return fail(h, BSR_EINVAL, "img/uv must be non-NULL");
  if (!compact && (cp.rgb_u8 || cp.dif_f16)) return fail(h, BSR_EINVAL, "compact outputs need the compact entry point");
  if (h->variant == BSR_VARIANT_TSM && !reg) return fail(h, BSR_EINVAL, "TSM needs reg");
  if (int rc = pending_device_error(h)) return rc;
  // fp32 inputs: upload only the uv / reg rows the model reads (rows 8i+3, 8i+4: one strided 2-D DMA each)
  const bool rows_only = !compact && !h->kn.host_full_uv;
  SmallInputScope small_scope(h, compact, rows_only);
  DeviceScope dev_scope(h->device);
  if (!dev_scope.ok) return fail(h, BSR_ECUDA, "cudaSetDevice(%d) failed", h->device);
  // transfer/compute chunk of the pipelined host path: smaller than the device micro-batch so that PCIe copies and
  // kernels of neighbouring chunks overlap (BSR_HOST_CHUNK overrides).
  // chunk = a multiple of num_sms / 4 images (37 on a B200): attention has 4 work items and res conv2 4 regions per image, the
  // 1x1 convs 8 tiles, so such chunks fill every round of their persistent grids exactly.  Measured on one B200, 256 images
  // per call (images/s, fp32 | compact I/O): 37 -> 24.8 k | 24.9 k, 64 -> 25.1 k | 26.7 k, 74 -> 26.4 k | 28.0 k,
  // 111 -> 21.9 k | 29.2 k, 128 -> 20.2 k | 27.5 k (fp32 transfers are 5x larger: long first / last copies cost more)
  const int q = h->num_sms >= 8 ? h->num_sms / 4 : 32;
  int host_chunk = (compact || n >= 512) ? 3 * q : 2 * q;
  if (h->kn.host_chunk > 0) host_chunk = h->kn.host_chunk;
  if (host_chunk > h->host_step_cap) host_chunk = h->host_step_cap;
  int step = h->mb < host_chunk ? h->mb : host_chunk;
  if (reg) {
    if (frame <= 0 || n % frame) return fail(h, BSR_EINVAL, "TSM needs n %% frame == 0");
    if (frame > h->mb) return fail(h, BSR_EINVAL, "frame %d exceeds micro_batch %d", frame, h->mb);
    if (step < frame) step = frame;
    step = step / frame * frame;
  }
  const size_t p1 = (size_t)IMG * IMG * sizeof(float);           // one single-channel image plane
  const size_t px = (size_t)IMG * IMG, px32 = (size_t)FEAT * FEAT;
  const HostSlots hs = host_slots(step, reg != nullptr);
  const size_t in_slot = hs.in_slot, out_slot = hs.out_slot, cin_slot = hs.cin_slot, cout_slot = hs.cout_slot;
  if (hs.total() > h->stage_bytes)
    return fail(h, BSR_EINVAL, "host path: a chunk of %d images needs %zu bytes of staging, the handle was created with %zu "
                "(chunks are capped at %d images; frame must not exceed that)", step, hs.total(), h->stage_bytes, h->host_step_cap);
  cudaStream_t s_in = h->s_in, s_c = h->own_stream, s_out = h->s_out;
  // the input staging slots are read by the previous forward of this handle until it has finished
  if (h->have_done) CK(h, cudaStreamWaitEvent(s_in, h->ev_done, 0));
  // Chunk schedule: ramp up and down (step/4, step/2, step ... step, step/2, step/4) so that the H2D of the
  // first chunk and the D2H of the last one - the only transfers that cannot overlap compute - are short.
  const int unit = reg ? frame : 1;
  auto round_unit = [&](int v) { v = v / unit * unit; return v < unit ? unit : v; };
  std::vector<int> sched;
  {
    int left = n;
    const int q4 = round_unit(step / 4), q2 = round_unit(step / 2);
    std::vector<int> head, tail;
    if (n >= 3 * step) { head = {q4, q2}; tail = {q2, q4}; }
    for (int v : head) { sched.push_back(v); left -= v; }
    int tail_sum = 0;
    for (int v : tail) tail_sum += v;
    while (left - tail_sum > 0) { int v = left - tail_sum < step ? left - tail_sum : step; sched.push_back(v); left -= v; }
    for (int v : tail) { sched.push_back(v); left -= v; }
  }
  int total_launches = 0, k = 0, i0 = 0;
  PlanCounters pc_total;
  for (size_t ci = 0; ci < sched.size(); i0 += sched[ci], ++ci, ++k) {
    const int m = sched[ci];
    const int slot = k & 1;
    char* sin = (char*)h->stage + slot * in_slot;
    char* sout = (char*)h->stage + 2 * in_slot + slot * out_slot;
    float* d_img = (float*)sin;
    float* d_uv = (float*)(sin + (size_t)step * p1 * 3);
    float* d_reg = reg ? (float*)(sin + (size_t)step * p1 * 6) : nullptr;
    float* d_gs = (float*)sout;
    float* d_rgb = (float*)(sout + (size_t)step * p1);
    float* d_m22 = (float*)(sout + (size_t)step * p1 * 4);
    float* d_dif = (float*)(sout + (size_t)step * p1 * 7);
    const size_t o1 = (size_t)i0 * IMG * IMG, pm = (size_t)m * p1;
    char* cbase = (char*)h->stage + 2 * (in_slot + out_slot);
    unsigned char* c_img = (unsigned char*)(cbase + slot * cin_slot);
    float* c_uv = (float*)(c_img + (size_t)step * px * 3);
    float* c_reg = c_uv + (size_t)step * px32 * 3;
    unsigned char* c_rgb = (unsigned char*)(cbase + 2 * cin_slot + slot * cout_slot);
    unsigned short* c_dif = (unsigned short*)(c_rgb + (size_t)step * px * 3);
    // the input slot is free once the compute that read it (two micro-batches ago) has finished
    if (k >= 2) CK(h, cudaStreamWaitEvent(s_in, h->ev_comp[slot], 0));
    if (compact) {
      const size_t q1 = (size_t)i0 * px32, qm = (size_t)m * px32 * sizeof(float);
      CK(h, cudaMemcpyAsync(c_img, cp.img_u8 + o1 * 3, (size_t)m * px * 3, cudaMemcpyHostToDevice, s_in));
      CK(h, cudaMemcpyAsync(c_uv, uv + q1 * 3, qm * 3, cudaMemcpyHostToDevice, s_in));
      if (reg) CK(h, cudaMemcpyAsync(c_reg, reg + q1 * 6, qm * 6, cudaMemcpyHostToDevice, s_in));
    } else {
      CK(h, cudaMemcpyAsync(d_img, img + o1 * 3, pm * 3, cudaMemcpyHostToDevice, s_in));
      if (rows_only) {
        // source: bands of 8 image rows, of which rows 3 and 4 (contiguous) are copied; the band pitch is uniform
        // across images (256 rows = 32 bands), so one 2-D copy covers the whole chunk
        const size_t row3 = (size_t)IMG * 3 * sizeof(float), row6 = (size_t)IMG * 6 * sizeof(float);
        CK(h, cudaMemcpy2DAsync(d_uv, 2 * row3, (const char*)(uv + o1 * 3) + 3 * row3, 8 * row3, 2 * row3, (size_t)m * FEAT,
                                cudaMemcpyHostToDevice, s_in));
        if (reg)
          CK(h, cudaMemcpy2DAsync(d_reg, 2 * row6, (const char*)(reg + o1 * 6) + 3 * row6, 8 * row6, 2 * row6, (size_t)m * FEAT,
                                  cudaMemcpyHostToDevice, s_in));
      } else {
        CK(h, cudaMemcpyAsync(d_uv, uv + o1 * 3, pm * 3, cudaMemcpyHostToDevice, s_in));
        if (reg) CK(h, cudaMemcpyAsync(d_reg, reg + o1 * 6, pm * 6, cudaMemcpyHostToDevice, s_in));
      }
    }
    CK(h, cudaEventRecord(h->ev_in[slot], s_in));
    CK(h, cudaStreamWaitEvent(s_c, h->ev_in[slot], 0));
    // the output slot is free once its previous D2H (two micro-batches ago) has finished
    if (k >= 2) CK(h, cudaStreamWaitEvent(s_c, h->ev_out[slot], 0));
    int extra_launches = 0;
    if (compact) {
      const long long n16 = (long long)m * px * 3 / 16;
      expand_u8_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, s_c>>>((const uint4*)c_img, (float4*)d_img, n16);
      ++extra_launches;
    }
    const bool want_rgb = rgb || cp.rgb_u8, want_dif = dif || cp.dif_f16;
    int rc = forward_common(h, d_img, compact ? c_uv : d_uv, reg ? (compact ? c_reg : d_reg) : nullptr, m, frame, share,
                            gs ? d_gs : nullptr, want_rgb ? d_rgb : nullptr, mask22 ? d_m22 : nullptr,
                            want_dif ? d_dif : nullptr, s_c);
    if (rc) return rc;
    if (cp.rgb_u8) {
      const long long n16 = (long long)m * px * 3 / 16;
      rgb_to_u8_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, s_c>>>((const float4*)d_rgb, (uint4*)c_rgb, n16);
      ++extra_launches;
    }
    if (cp.dif_f16) {
      const long long n8 = (long long)m * px / 8;
      f32_to_f16_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, s_c>>>((const float4*)d_dif, (uint4*)c_dif, n8);
      ++extra_launches;
    }
    total_launches += h->launches + extra_launches;
    pc_total.resident += h->pc.resident; pc_total.pinned += h->pc.pinned; pc_total.staged += h->pc.staged;
    pc_total.attn_fused += h->pc.attn_fused; pc_total.graph_replays += h->pc.graph_replays; pc_total.halo3 += h->pc.halo3;
    CK(h, cudaEventRecord(h->ev_comp[slot], s_c));
    CK(h, cudaStreamWaitEvent(s_out, h->ev_comp[slot], 0));
    if (gs) CK(h, cudaMemcpyAsync(gs + o1, d_gs, pm, cudaMemcpyDeviceToHost, s_out));
    if (rgb) CK(h, cudaMemcpyAsync(rgb + o1 * 3, d_rgb, pm * 3, cudaMemcpyDeviceToHost, s_out));
    if (mask22) CK(h, cudaMemcpyAsync(mask22 + o1 * 3, d_m22, pm * 3, cudaMemcpyDeviceToHost, s_out));
    if (dif) CK(h, cudaMemcpyAsync(dif + o1, d_dif, pm, cudaMemcpyDeviceToHost, s_out));
    if (cp.rgb_u8) CK(h, cudaMemcpyAsync(cp.rgb_u8 + o1 * 3, c_rgb, (size_t)m * px * 3, cudaMemcpyDeviceToHost, s_out));
    if (cp.dif_f16) CK(h, cudaMemcpyAsync(cp.dif_f16 + o1, c_dif, (size_t)m * px * 2, cudaMemcpyDeviceToHost, s_out));
    CK(h, cudaEventRecord(h->ev_out[slot], s_out));
  }
  CK(h, cudaStreamSynchronize(s_out));
  CK(h, cudaStreamSynchronize(s_c));
  h->launches = total_launches;
  h->pc = pc_total;
  return pending_device_error(h);       // the flag copy of the last chunk has landed (stream synchronised above)
}

}  // namespace

// =============================================================================================
extern "C" {

const char* bsr_version(void) { return "bsr-b200 0.2 (sm_100a, " BSR_ACT_DTYPE_NAME " storage)"; }
const char* bsr_act_dtype(void) { return BSR_ACT_DTYPE_NAME; }
void bsr_convert_h16(const float* in, unsigned short* out, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = f32_to_h16_bits(in[i]);
}

// CRC-32C, reflected polynomial 0x82F63B78, slicing-by-8 (host only)
unsigned int bsr_crc32c(unsigned int crc, const void* data, size_t n) {
  static uint32_t tab[8][256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
      tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) tab[t][i] = (tab[t - 1][i] >> 8) ^ tab[0][tab[t - 1][i] & 0xffu];
    init = true;
  }
  const unsigned char* p = (const unsigned char*)data;
  uint32_t c = ~crc;
  while (n >= 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = tab[7][lo & 0xffu] ^ tab[6][(lo >> 8) & 0xffu] ^ tab[5][(lo >> 16) & 0xffu] ^ tab[4][lo >> 24] ^
        tab[3][hi & 0xffu] ^ tab[2][(hi >> 8) & 0xffu] ^ tab[1][(hi >> 16) & 0xffu] ^ tab[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = (c >> 8) ^ tab[0][(c ^ *p++) & 0xffu];
  return ~c;
}

const char* bsr_last_error(const bsr_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int bsr_create(int variant, int precision, int device, int micro_batch, bsr_handle** out) {
  if (!out) return BSR_EINVAL;
  *out = nullptr;
  if (variant != BSR_VARIANT_GSC && variant != BSR_VARIANT_TSM) return fail(nullptr, BSR_EINVAL, "bad variant %d", variant);
  if (precision != BSR_PRECISION_BF16 && precision != BSR_PRECISION_FP32CHECK)
    return fail(nullptr, BSR_EINVAL, "bad precision %d", precision);
  if (micro_batch <= 0 || micro_batch > 4096) return fail(nullptr, BSR_EINVAL, "micro_batch out of range");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(nullptr, BSR_ECUDA, "no CUDA device available (%s); this library has no CPU path",
                cudaGetErrorString(ce));
  if (device < 0 || device >= ndev) return fail(nullptr, BSR_EINVAL, "device %d out of range (%d devices)", device, ndev);
  cudaDeviceProp prop;
  CK(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(nullptr, BSR_EUNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                prop.major, prop.minor);
  CK(nullptr, cudaSetDevice(device));
  bsr_handle* h = new bsr_handle();
  h->num_sms = prop.multiProcessorCount;
  h->variant = variant; h->precision = precision; h->device = device; h->mb = micro_batch;
  h->es = precision == BSR_PRECISION_FP32CHECK ? 4 : 2;
  const char* ev = getenv("BSR_DEBUG_KEEP");
  h->debug_keep = ev && atoi(ev) != 0;
  ev = getenv("BSR_PROFILE");
  h->profile = ev && atoi(ev) != 0;
  ev = getenv("BSR_FORCE_DIRECT");
  h->force_direct = ev ? atoi(ev) : 0;
  auto env_int = [](const char* k) { const char* v = getenv(k); return v ? atoi(v) : 0; };
  auto env_set = [](const char* k) { return getenv(k) != nullptr ? 1 : 0; };
  h->kn.ablate = env_int("BSR_ABLATE");
  h->kn.no_pdl = env_set("BSR_NO_PDL");
  h->kn.no_tma_store = env_set("BSR_NO_TMA_STORE");
  h->kn.st_bufs = env_int("BSR_ST_BUFS");
  h->kn.no_fuse_w = env_set("BSR_NO_FUSE_W");
  h->kn.host_chunk = env_int("BSR_HOST_CHUNK");
  h->kn.no_graph = env_set("BSR_NO_GRAPH");
  h->kn.host_full_uv = env_set("BSR_HOST_FULL_UV");
  h->kn.no_halo = env_set("BSR_NO_HALO");
  h->kn.no_halo3 = env_set("BSR_NO_HALO3");
  h->kn.unpack_v1 = env_set("BSR_UNPACK_V1");
  h->kn.share_v1 = env_set("BSR_SHARE_V1");
  h->kn.no_hole_inplace = env_set("BSR_NO_HOLE_INPLACE");
  h->c_first = variant == BSR_VARIANT_GSC ? 99 : 291;
  h->c_second = variant == BSR_VARIANT_GSC ? 261 : 877;
  h->ld1 = pad16(h->c_first > 257 ? h->c_first : 257);
  if (h->ld1 < kLdY) h->ld1 = kLdY;
  h->ld2 = pad16(h->c_second);
  if (h->ld2 < kLdY) h->ld2 = kLdY;
  // ---- workspace: one arena, carved with 256-byte alignment
  const size_t es = h->es, mb = micro_batch;
  const size_t ldmax = h->ld1 > h->ld2 ? h->ld1 : h->ld2;
  struct Req { char** p; size_t bytes; };
  std::vector<Req> reqs = {
      {&h->X1, mb * IMG * IMG * 32 * es}, {&h->CAT3, mb * 128 * 128 * 128 * es}, {&h->CAT2, mb * 64 * 64 * 160 * es},
      {&h->XA, mb * 1024 * ldmax * es}, {&h->XB, mb * 1024 * ldmax * es}, {&h->T1, mb * 1024 * 128 * es},
      {&h->T2, mb * 1024 * 128 * es}, {&h->Y, mb * 1024 * kLdY * es}, {&h->QK, mb * 1024 * 256 * es},
      {&h->VT, mb * 1024 * 128 * es}, {&h->O, mb * 1024 * 128 * es}, {&h->UP3, mb * IMG * IMG * 64 * es},
      {&h->F1, mb * 64 * 64 * 128 * es}, {&h->F2, mb * 128 * 128 * 96 * es}, {&h->CAT1, mb * IMG * IMG * 72 * es},
      {&h->C16, mb * IMG * IMG * 16 * es}, {&h->PIMG, mb * (IMG + 1) * (IMG + 8) * 8 * 2},
      {(char**)&h->RAW, mb * IMG * IMG * 2 * 4}, {(char**)&h->GS32, mb * IMG * IMG * 4}, {(char**)&h->DIFGS, mb * IMG * IMG * 4},
      {(char**)&h->UVS, mb * 1024 * 3 * 4}, {(char**)&h->OFF, mb * 1024 * 4 * 4}, {(char**)&h->BMASK, mb * 1024 * 4},
      {(char**)&h->DIFSMALL, mb * 1024 * 4},
      {(char**)&h->SH, variant == BSR_VARIANT_TSM ? mb * 1024 * 584 * 4 : 256},
      {&h->TAPS, variant == BSR_VARIANT_TSM ? mb * 1024 * 2 * 16 : 256},
      {(char**)&h->errflag, 16384}};
  // staging of the host / chunk entry points (include/bsr.h: no allocation inside forward_*): host-path chunks are
  // capped at 128 images (the measured optimum is 64-128), chunk entry at one micro-batch
  h->host_step_cap = micro_batch < 128 ? micro_batch : 128;
  h->stage_bytes = host_slots(h->host_step_cap, variant == BSR_VARIANT_TSM).total();
  h->chunk_stage_bytes = mb * IMG * IMG * 13 * sizeof(float);
  reqs.push_back({&h->stage, h->stage_bytes});
  reqs.push_back({&h->chunk_stage, h->chunk_stage_bytes});
  size_t total = 0;
  for (auto& r : reqs) total += align256(r.bytes);
  if (cudaMalloc(&h->arena, total) != cudaSuccess) {
    cudaGetLastError();
    delete h;
    return fail(nullptr, BSR_ENOMEM, "workspace allocation of %zu bytes failed", total);
  }
  h->arena_bytes = total;
  cudaMemset(h->arena, 0, total);
  size_t off = 0;
  for (auto& r : reqs) { *r.p = (char*)h->arena + off; off += align256(r.bytes); }
  cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming);
  for (int i = 0; i < 2; ++i) {
    cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming);
  }
  if (cudaHostAlloc((void**)&h->errflag_host, 64, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    bsr_destroy(h);
    return fail(nullptr, BSR_ENOMEM, "pinned allocation of the error-flag mirror failed");
  }
  *h->errflag_host = 0;
  if (h->profile) {
    h->ev.resize(512);
    for (auto& e : h->ev) cudaEventCreate(&e);
  }
  if (!h->tma.init()) {
    std::string msg = h->tma.last_error;
    bsr_destroy(h);
    return fail(nullptr, BSR_ECUDA, "cuTensorMapEncodeTiled unavailable: %s", msg.c_str());
  }
  if (int rc = configure_tc_kernels()) {
    bsr_destroy(h);
    return fail(nullptr, BSR_ECUDA, "cudaFuncSetAttribute failed for tensor-core kernels (%d)", rc);
  }
  cudaFuncSetAttribute(attention_simple_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttnSimpleSmem);
  cudaFuncSetAttribute(attention_simple_kernel<h16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttnSimpleSmem);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    bsr_destroy(h);
    return fail(nullptr, BSR_ECUDA, "device init failed: %s", cudaGetErrorString(e));
  }
  *out = h;
  return BSR_OK;
}

int bsr_destroy(bsr_handle* h) {
  if (!h) return BSR_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (auto& kv : h->layers) {
    if (kv.second.w_dev) cudaFree(kv.second.w_dev);
    if (kv.second.b_dev) cudaFree(kv.second.b_dev);
    kv.second.tc.release();
    for (auto& tp : kv.second.tc_phase) tp.release();
    kv.second.tc_halo.release();
  }
  for (auto& kv : h->dbg) if (kv.second.dev) cudaFree(kv.second.dev);
  for (auto& e : h->ev) cudaEventDestroy(e);
  if (h->arena) cudaFree(h->arena);
  if (h->pp_buf) cudaFree(h->pp_buf);
  if (h->pp_planes) cudaFree(h->pp_planes);
  if (h->errflag_host) cudaFreeHost(h->errflag_host);
  clear_graphs(h);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  if (h->ev_done) cudaEventDestroy(h->ev_done);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
    if (h->ev_comp[i]) cudaEventDestroy(h->ev_comp[i]);
    if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
  }
  delete h;
  return BSR_OK;
}

int bsr_load_weights(bsr_handle* h, const void* blob, size_t nbytes) {
  if (!h || !blob) return BSR_EINVAL;
  CK(h, cudaSetDevice(h->device));
  const char* b = (const char*)blob;
  if (nbytes < 16 || memcmp(b, "BSRW0001", 8) != 0) return fail(h, BSR_EINVAL, "bad weight blob magic");
  CK(h, cudaDeviceSynchronize());
  clear_graphs(h);                      // captured launches hold pointers to the weights being replaced
  int32_t variant, n_layers;
  memcpy(&variant, b + 8, 4);
  memcpy(&n_layers, b + 12, 4);
  if (variant != h->variant) return fail(h, BSR_EINVAL, "blob is for variant %d, handle is %d", variant, h->variant);
  if (n_layers <= 0 || 16 + (size_t)n_layers * sizeof(BlobEntry) > nbytes) return fail(h, BSR_EINVAL, "truncated blob");
  for (int i = 0; i < n_layers; ++i) {
    BlobEntry en;
    memcpy(&en, b + 16 + (size_t)i * sizeof(BlobEntry), sizeof en);
    en.name[31] = 0;
    size_t wn = (size_t)en.kh * en.kw * en.cin * en.cout;
    if (en.w_off + wn * 4 > nbytes || en.b_off + (size_t)en.cout * 4 > nbytes)
      return fail(h, BSR_EINVAL, "layer %s: payload outside blob", en.name);
    Layer& L = h->layers[en.name];
    if (L.w_dev) { cudaFree(L.w_dev); L.w_dev = nullptr; }
    if (L.b_dev) { cudaFree(L.b_dev); L.b_dev = nullptr; }
    L.tc.release();
    L.name = en.name;
    L.kh = en.kh; L.kw = en.kw; L.cin = en.cin; L.cout = en.cout; L.transposed = en.transposed;
    L.w_host.assign((const float*)(b + en.w_off), (const float*)(b + en.w_off) + wn);
    L.b_host.assign((const float*)(b + en.b_off), (const float*)(b + en.b_off) + en.cout);
    CK(h, cudaMalloc(&L.w_dev, wn * 4));
    CK(h, cudaMemcpy(L.w_dev, L.w_host.data(), wn * 4, cudaMemcpyHostToDevice));
    size_t bn = ((size_t)en.cout + 15) / 16 * 16 + 1024;
    std::vector<float> bp(bn, 0.f);
    memcpy(bp.data(), L.b_host.data(), (size_t)en.cout * 4);
    CK(h, cudaMalloc(&L.b_dev, bn * 4));
    CK(h, cudaMemcpy(L.b_dev, bp.data(), bn * 4, cudaMemcpyHostToDevice));
    for (auto& tp : L.tc_phase) tp.release();
    L.phases_ready = false;
    L.halo_ready = false;
    if (h->precision == BSR_PRECISION_BF16) {
      std::string why;
      if (!pack_tc_weights(h->tma, L.name, L.kh, L.kw, L.cin, L.cout, L.transposed, L.w_host, L.b_host, &L.tc, &why)) {
        if (!why.empty()) return fail(h, BSR_ECUDA, "packing %s for tensor cores failed: %s", en.name, why.c_str());
        if (L.transposed && !tc_disabled(L.name)) {
          bool ok = true;
          for (int ph = 0; ph < 4 && ok; ++ph)
            ok = pack_tc_weights_phase(h->tma, ph, L.cin, L.cout, L.w_host, &L.tc_phase[ph], &why);
          if (!ok) return fail(h, BSR_ECUDA, "packing %s (per phase) failed: %s", en.name, why.c_str());
          L.phases_ready = true;
          if (L.kh == 3 && L.kw == 3 && L.cout % 16 == 0 && 4 * L.cout <= 512) {
            L.tc_halo.release();
            if (!pack_convt_halo_weights(h->tma, L.cin, L.cout, L.w_host, &L.tc_halo, &why))
              return fail(h, BSR_ECUDA, "packing %s (fused, halo kernel) failed: %s", en.name, why.c_str());
            L.halo_ready = true;
          }
        }
      }
    }
  }
  static const char* required[] = {"conv1", "down1", "down2", "down3", "up1", "up2", "up3", "heads", "clr_up1",
                                   "clr_up2", "clr_up3", "clr_conv1", "clr_conv2", "clr_conv3"};
  for (const char* r : required)
    if (!h->layers.count(r)) return fail(h, BSR_EINVAL, "blob lacks layer %s", r);
  for (int i = 0; i < 6; ++i)
    for (const char* s : {"conv1", "conv2", "conv3", "qkv", "w"}) {
      char nm[32];
      snprintf(nm, 32, "res%d.%s", i, s);
      if (!h->layers.count(nm)) return fail(h, BSR_EINVAL, "blob lacks layer %s", nm);
    }
  if (h->precision == BSR_PRECISION_BF16 && h->layers["clr_conv1"].tc.ready) {
    // weights of the fused colour tail: [9][16] gs taps of clr_conv1 (canonical input channel 64), clr_conv2, clr_conv3
    Layer& c1 = h->layers["clr_conv1"];
    Layer& c2 = h->layers["clr_conv2"];
    Layer& c3 = h->layers["clr_conv3"];
    if (c2.cin != 16 || c2.cout != 16 || c3.cin != 16 || c3.cout != 3) return fail(h, BSR_EINVAL, "unexpected colour tail shapes");
    if (!c1.tc.clr) c1.tc.clr = new ClrWeights();
    ClrWeights& cw = *c1.tc.clr;
    memset(&cw, 0, sizeof cw);
    for (int t = 0; t < 9; ++t)
      for (int o = 0; o < 16; ++o) cw.wg[t * 16 + o] = c1.w_host[((size_t)t * 65 + 64) * 16 + o];
    memcpy(cw.w2, c2.w_host.data(), 256 * 4);
    memcpy(cw.b2, c2.b_host.data(), 16 * 4);
    for (int c = 0; c < 16; ++c)
      for (int o = 0; o < 3; ++o) cw.w3t[o * 16 + c] = c3.w_host[c * 3 + o];       // transposed: [3 out][16 in]
    memcpy(cw.b3, c3.b_host.data(), 3 * 4);
  }
  const int cin_first = h->c_first, cin_second = h->c_second;
  if (h->layers["res0.conv1"].cin != cin_first || h->layers["res3.conv1"].cin != cin_second)
    return fail(h, BSR_EINVAL, "res-stack input widths do not match variant");
  h->loaded = true;
  return BSR_OK;
}

int bsr_forward_gsc(bsr_handle* h, const float* img, const float* uv, int n, float* gs, float* rgb, float* mask22,
                    float* dif, void* cuda_stream) {
  if (!h) return BSR_EINVAL;
  if (h->variant != BSR_VARIANT_GSC) return fail(h, BSR_EINVAL, "handle is not a GSC generator");
  return forward_common(h, img, uv, nullptr, n, 1, 0, gs, rgb, mask22, dif, (cudaStream_t)cuda_stream);
}

int bsr_forward_tsm(bsr_handle* h, const float* img, const float* uv, const float* reg, int n_chunks, int frame,
                    int share, float* gs, float* rgb, float* mask22, float* dif, void* cuda_stream) {
  if (!h) return BSR_EINVAL;
  if (h->variant != BSR_VARIANT_TSM) return fail(h, BSR_EINVAL, "handle is not a TSM generator");
  if (n_chunks <= 0 || frame <= 0) return fail(h, BSR_EINVAL, "n_chunks and frame must be > 0");
  return forward_common(h, img, uv, reg, n_chunks * frame, frame, share, gs, rgb, mask22, dif,
                        (cudaStream_t)cuda_stream);
}

int bsr_forward_gsc_host(bsr_handle* h, const float* img, const float* uv, int n, float* gs, float* rgb,
                         float* mask22, float* dif) {
  if (!h) return BSR_EINVAL;
  if (h->variant != BSR_VARIANT_GSC) return fail(h, BSR_EINVAL, "handle is not a GSC generator");
  return forward_host(h, img, uv, nullptr, n, 1, 0, gs, rgb, mask22, dif);
}

int bsr_forward_tsm_host(bsr_handle* h, const float* img, const float* uv, const float* reg, int n_chunks, int frame,
                         int share, float* gs, float* rgb, float* mask22, float* dif) {
  if (!h) return BSR_EINVAL;
  if (h->variant != BSR_VARIANT_TSM) return fail(h, BSR_EINVAL, "handle is not a TSM generator");
  if (n_chunks <= 0 || frame <= 0 || !reg) return fail(h, BSR_EINVAL, "n_chunks, frame > 0 and reg required");
  return forward_host(h, img, uv, reg, n_chunks * frame, frame, share, gs, rgb, mask22, dif);
}

int bsr_forward_gsc_host_compact(bsr_handle* h, const unsigned char* img_u8, const float* uv32, int n, float* gs,
                                 float* rgb, float* mask22, float* dif, unsigned char* rgb_u8,
                                 unsigned short* dif_f16) {
  if (!h) return BSR_EINVAL;
  if (h->variant != BSR_VARIANT_GSC) return fail(h, BSR_EINVAL, "handle is not a GSC generator");
  if (!img_u8) return fail(h, BSR_EINVAL, "img_u8 must be non-NULL");
  HostCompact cp;
  cp.img_u8 = img_u8; cp.rgb_u8 = rgb_u8; cp.dif_f16 = dif_f16;
  return forward_host(h, nullptr, uv32, nullptr, n, 1, 0, gs, rgb, mask22, dif, cp);
}

int bsr_forward_tsm_host_compact(bsr_handle* h, const unsigned char* img_u8, const float* uv32, const float* reg32,
                                 int n_chunks, int frame, int share, float* gs, float* rgb, float* mask22, float* dif,
                                 unsigned char* rgb_u8, unsigned short* dif_f16) {
  if (!h) return BSR_EINVAL;
  if (h->variant != BSR_VARIANT_TSM) return fail(h, BSR_EINVAL, "handle is not a TSM generator");
  if (!img_u8) return fail(h, BSR_EINVAL, "img_u8 must be non-NULL");
  if (n_chunks <= 0 || frame <= 0) return fail(h, BSR_EINVAL, "n_chunks and frame must be > 0");
  HostCompact cp;
  cp.img_u8 = img_u8; cp.rgb_u8 = rgb_u8; cp.dif_f16 = dif_f16;
  return forward_host(h, nullptr, uv32, reg32, n_chunks * frame, frame, share, gs, rgb, mask22, dif, cp);
}

int bsr_forward_chunk(bsr_handle* h, const float* chunk, int n, int layout, int frame, int share, float* rgb_clipped,
                      float* mask_pred, float* gs, float* mask22, void* cuda_stream) {
  if (!h) return BSR_EINVAL;
  if (!h->loaded) return fail(h, BSR_ESTATE, "bsr_load_weights has not been called");
  if (!chunk || n <= 0 || !rgb_clipped || !mask_pred) return fail(h, BSR_EINVAL, "chunk, rgb_clipped, mask_pred must be non-NULL and n > 0");
  int C, o_uv, o_reg, o_face;
  switch (layout) {      // channel offsets of tf.split(img, [...], 3)
    case BSR_CHUNK_GT: C = 16; o_uv = 6; o_reg = 9; o_face = 15; break;        // [3,3,3,6,1]
    case BSR_CHUNK_SFW: C = 17; o_uv = 7; o_reg = 10; o_face = 16; break;      // [3,3,1,3,6,1]
    case BSR_CHUNK_PLAIN: C = 13; o_uv = 3; o_reg = 6; o_face = 12; break;     // [3,3,6,1]
    default: return fail(h, BSR_EINVAL, "bad chunk layout %d", layout);
  }
  const bool tsm = h->variant == BSR_VARIANT_TSM;
  if (tsm && (frame <= 0 || n % frame)) return fail(h, BSR_EINVAL, "TSM needs n %% frame == 0");
  if (tsm && frame > h->mb) return fail(h, BSR_EINVAL, "frame %d exceeds micro_batch %d", frame, h->mb);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (int rc = pending_device_error(h)) return rc;
  DeviceScope dev_scope(h->device);
  if (!dev_scope.ok) return fail(h, BSR_ECUDA, "cudaSetDevice(%d) failed", h->device);
  int step = h->mb < n ? h->mb : n;
  if (tsm) step = step / frame * frame;
  const size_t px = (size_t)IMG * IMG;      // staging: one micro-batch of 13 dense planes, carved from the arena at create
  // the staging planes are read by the previous forward of this handle until it has finished
  if (h->have_done && h->last_stream != st) CK(h, cudaStreamWaitEvent(st, h->ev_done, 0));
  float* d_img = (float*)h->chunk_stage;
  float* d_uv = d_img + (size_t)step * px * 3;
  float* d_reg = d_uv + (size_t)step * px * 3;
  float* d_face = d_reg + (size_t)step * px * 6;
  int total = 0;
  PlanCounters pc_total;
  for (int i0 = 0; i0 < n; i0 += step) {
    const int m = n - i0 < step ? n - i0 : step;
    const long long mpx = (long long)m * px;
    if (h->kn.unpack_v1)
      unpack_chunk_kernel<<<(unsigned)((mpx * 13 + 255) / 256), 256, 0, st>>>(chunk + (size_t)i0 * px * C, C, o_uv, o_reg, o_face,
                                                                          d_img, d_uv, tsm ? d_reg : nullptr, d_face, mpx);
    else
      unpack_chunk_tile_kernel<<<(unsigned)((mpx + 255) / 256), 256, (size_t)256 * (C | 1) * sizeof(float), st>>>(
          chunk + (size_t)i0 * px * C, C, o_uv, o_reg, o_face, d_img, d_uv, tsm ? d_reg : nullptr, d_face, mpx);
    float* rgb_o = rgb_clipped + (size_t)i0 * px * 3;
    float* mp_o = mask_pred + (size_t)i0 * px;
    int rc = forward_common(h, d_img, d_uv, tsm ? d_reg : nullptr, m, frame, share, gs ? gs + (size_t)i0 * px : nullptr, rgb_o,
                            mask22 ? mask22 + (size_t)i0 * px * 3 : nullptr, mp_o, st);
    if (rc) return rc;
    caller_glue_kernel<<<(unsigned)((mpx + 255) / 256), 256, 0, st>>>(rgb_o, mp_o, d_face, rgb_o, mp_o, mpx);   // in place
    total += h->launches + 2;
    pc_total.resident += h->pc.resident; pc_total.pinned += h->pc.pinned; pc_total.staged += h->pc.staged;
    pc_total.attn_fused += h->pc.attn_fused; pc_total.graph_replays += h->pc.graph_replays; pc_total.halo3 += h->pc.halo3;
  }
  CK(h, cudaGetLastError());
  h->launches = total;
  h->pc = pc_total;
  return finish_forward(h, st);        // the in-place caller glue above is part of this forward
}

int bsr_share_layer(bsr_handle* h, const float* x, const float* reg, int n, int C, int frame, int share, float* out,
                    void* cuda_stream) {
  if (!h) return BSR_EINVAL;
  if (h->variant != BSR_VARIANT_TSM) return fail(h, BSR_EINVAL, "ShareLayer exists in the TSM generator only");
  if (!x || !reg || !out || n <= 0 || frame <= 0 || n % frame) return fail(h, BSR_EINVAL, "x, reg, out non-NULL and n %% frame == 0 required");
  if (n > h->mb) return fail(h, BSR_EINVAL, "n %d exceeds micro_batch %d", n, h->mb);
  if (C < 1 || C > 291) return fail(h, BSR_EINVAL, "C must be in [1, 291] (the generator uses 96 and 291)");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (int rc = pending_device_error(h)) return rc;
  DeviceScope dev_scope(h->device);
  if (!dev_scope.ok) return fail(h, BSR_ECUDA, "cudaSetDevice(%d) failed", h->device);
  if (h->have_done && h->last_stream != st) CK(h, cudaStreamWaitEvent(st, h->ev_done, 0));
  const long long npix = (long long)n * FEAT * FEAT;
  const int ld = h->ld2, coff = (C + 3) / 4 * 4;
  const int tot4 = n * FEAT * FEAT * 4;
  reg_small_kernel<<<(tot4 + 255) / 256, 256, 0, st>>>(reg, h->OFF, n);
  h->launches = 1;
  h->taps_ready = false;
  int rc;
  if (h->precision == BSR_PRECISION_FP32CHECK) {
    pack_act_kernel<float><<<(unsigned)((npix * C + 255) / 256), 256, 0, st>>>(x, C, (float*)h->XA, ld, npix);
    if ((rc = run_share<float>(h, st, h->XA, ld, C, coff, n, frame, share))) return rc;
    slice_to_f32_kernel<float><<<(unsigned)((npix * 2 * C + 255) / 256), 256, 0, st>>>((const float*)h->XA, ld, coff, 2 * C, out, npix);
  } else {
    pack_act_kernel<h16><<<(unsigned)((npix * C + 255) / 256), 256, 0, st>>>(x, C, (h16*)h->XA, ld, npix);
    if ((rc = run_share<h16>(h, st, h->XA, ld, C, coff, n, frame, share))) return rc;
    slice_to_f32_kernel<h16><<<(unsigned)((npix * 2 * C + 255) / 256), 256, 0, st>>>((const h16*)h->XA, ld, coff, 2 * C, out, npix);
  }
  h->launches += 2;
  CK(h, cudaGetLastError());
  return finish_forward(h, st);
}

int bsr_caller_glue(bsr_handle* h, const float* rgb, const float* dif, const float* face, int n, float* rgb_clipped,
                    float* mask_pred, void* cuda_stream) {
  if (!h || n <= 0) return BSR_EINVAL;
  if (mask_pred && (!dif || !face)) return fail(h, BSR_EINVAL, "mask_pred needs dif and face");
  if (rgb_clipped && !rgb) return fail(h, BSR_EINVAL, "rgb_clipped needs rgb");
  long long px = (long long)n * IMG * IMG;
  caller_glue_kernel<<<(unsigned)((px + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(rgb, dif, face, rgb_clipped,
                                                                                         mask_pred, px);
  CK(h, cudaGetLastError());
  return BSR_OK;
}

int bsr_composite(bsr_handle* h, const float* pred, const float* inp, const float* m, size_t n_elems, float* out,
                  void* cuda_stream) {
  if (!h || !pred || !inp || !m || !out) return BSR_EINVAL;
  if (n_elems == 0) return BSR_OK;
  size_t blocks = (n_elems + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  composite_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)cuda_stream>>>(pred, inp, m, out, n_elems);
  CK(h, cudaGetLastError());
  return BSR_OK;
}

int bsr_postprocess_ucb(bsr_handle* h, int n, const float* img, const float* gt, const float* rgb, const float* dif,
                        const int* sizes, const unsigned char* masks, float* final_out, float* detected_out,
                        float* metrics, void* cuda_stream) {
  if (!h) return BSR_EINVAL;
  if (n <= 0 || !img || !gt || !rgb || !dif || !sizes || !masks || !final_out)
    return fail(h, BSR_EINVAL, "n > 0 and img, gt, rgb, dif, sizes, masks, final_out non-NULL required");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  DeviceScope dev_scope(h->device);
  if (!dev_scope.ok) return fail(h, BSR_ECUDA, "cudaSetDevice(%d) failed", h->device);
  if (n > h->pp_cap) {
    CK(h, cudaStreamSynchronize(st));
    if (h->pp_buf) cudaFree(h->pp_buf);
    if (h->pp_planes) cudaFree(h->pp_planes);
    h->pp_buf = nullptr; h->pp_planes = nullptr; h->pp_cap = 0;
    const size_t per = (kPpBytesPerImage + 255) / 256 * 256;
    if (cudaMalloc(&h->pp_buf, per * n) != cudaSuccess || cudaMalloc(&h->pp_planes, sizeof(PpPlanes) * n) != cudaSuccess)
      return fail(h, BSR_ENOMEM, "post-processing scratch of %zu bytes failed", per * n);
    std::vector<PpPlanes> pl(n);
    for (int i = 0; i < n; ++i) {
      char* b = h->pp_buf + per * i;
      PpPlanes& P = pl[i];
      P.tmp = (float*)b; b += (size_t)PP_PIX * 12;
      P.gt = (float*)b; b += (size_t)PP_PIX * 12;
      P.pred = (float*)b; b += (size_t)PP_PIX * 12;
      P.mp = (float*)b; b += (size_t)PP_PIX * 4;
      P.inten = (float*)b; b += (size_t)PP_PIX * 4;
      P.label = (int*)b; b += (size_t)PP_PIX * 4;
      P.csize = (int*)b; b += (size_t)PP_PIX * 4;
      P.chair = (int*)b; b += (size_t)PP_PIX * 4;
      P.stats = (int*)b; b += PS_WORDS * 4;
      P.masks = (unsigned char*)b; b += (size_t)PP_PIX * PP_NMASK;
      P.detected = (unsigned char*)b; b += PP_PIX;
      P.img2 = (unsigned char*)b; b += PP_PIX;
    }
    CK(h, cudaMemcpy(h->pp_planes, pl.data(), sizeof(PpPlanes) * n, cudaMemcpyHostToDevice));
    h->pp_cap = n;
  }
  const dim3 grid(PP_PIX / 256, n);
  PpIn in{img, gt, rgb, dif, masks, sizes};
  pp_reset_kernel<<<grid, 256, 0, st>>>(h->pp_planes);
  pp_resize_kernel<<<grid, 256, 0, st>>>(in, h->pp_planes);
  pp_rules_kernel<<<grid, 256, 0, st>>>(h->pp_planes);
  pp_threshold_kernel<<<grid, 256, 0, st>>>(h->pp_planes);
  pp_ccl_merge_kernel<<<grid, 256, 0, st>>>(h->pp_planes);
  pp_ccl_flatten_kernel<<<grid, 256, 0, st>>>(h->pp_planes);
  pp_cc_max_kernel<<<grid, 256, 0, st>>>(h->pp_planes);
  pp_select_kernel<<<grid, 256, 0, st>>>(h->pp_planes);
  pp_final_kernel<<<grid, 256, 0, st>>>(h->pp_planes, final_out, detected_out);
  if (metrics) {
    pp_ssim_kernel<<<dim3(256, 3, n), 256, 0, st>>>(h->pp_planes);
    pp_metrics_kernel<<<(n + 63) / 64, 64, 0, st>>>(h->pp_planes, metrics, n);
  }
  h->launches = metrics ? 11 : 9;
  CK(h, cudaGetLastError());
  return BSR_OK;
}

int bsr_check(bsr_handle* h) {
  if (!h) return BSR_EINVAL;
  DeviceScope dev_scope(h->device);
  if (h->have_done) CK(h, cudaEventSynchronize(h->ev_done));
  return pending_device_error(h);
}

int bsr_plan_counter(const bsr_handle* h, int which) {
  if (!h) return 0;
  switch (which) {
    case 0: return h->pc.resident;
    case 1: return h->pc.pinned;
    case 2: return h->pc.staged;
    case 3: return h->pc.attn_fused;
    case 4: return h->pc.graph_replays;
    case 5: return h->pc.halo3;
    default: return 0;
  }
}

int bsr_launch_count(const bsr_handle* h) { return h ? h->launches : 0; }
size_t bsr_workspace_bytes(const bsr_handle* h) { return h ? h->arena_bytes : 0; }

int bsr_debug_read(bsr_handle* h, const char* name, float* host_out, size_t capacity, size_t* n_elems) {
  if (!h || !name) return BSR_EINVAL;
  if (!strcmp(name, "errflag")) {      // the raw device flag (sticky until bsr_check / the next forward reports it)
    int flag = 0;
    CK(h, cudaMemcpy(&flag, h->errflag, sizeof(int), cudaMemcpyDeviceToHost));
    if (n_elems) *n_elems = 1;
    if (host_out && capacity >= 1) host_out[0] = (float)flag;
    return BSR_OK;
  }
  if (!h->debug_keep) return fail(h, BSR_ESTATE, "create the handle with BSR_DEBUG_KEEP=1 to keep intermediates");
  if (!strcmp(name, "timers")) {       // BSR_ABLATE=8: 64 launches x 16 cycle counters (see conv_tc.cuh)
    if (n_elems) *n_elems = 1024;
    if (!host_out) return BSR_OK;
    if (capacity < 1024) return fail(h, BSR_EINVAL, "capacity too small");
    std::vector<long long> t(1024);
    CK(h, cudaDeviceSynchronize());
    CK(h, cudaMemcpy(t.data(), (char*)h->errflag + 128, 1024 * 8, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 1024; ++i) host_out[i] = (float)t[i];
    return BSR_OK;
  }
  auto it = h->dbg.find(name);
  if (it == h->dbg.end()) return fail(h, BSR_EINVAL, "no intermediate named %s", name);
  if (n_elems) *n_elems = it->second.n;
  if (!host_out) return BSR_OK;
  if (capacity < it->second.n) return fail(h, BSR_EINVAL, "capacity %zu < %zu", capacity, it->second.n);
  CK(h, cudaDeviceSynchronize());
  CK(h, cudaMemcpy(host_out, it->second.dev, it->second.n * sizeof(float), cudaMemcpyDeviceToHost));
  return BSR_OK;
}

int bsr_layer_times(const bsr_handle* hc, const char** names, float* ms, int capacity) {
  bsr_handle* h = const_cast<bsr_handle*>(hc);
  if (!h || !h->profile) return 0;
  cudaDeviceSynchronize();
  int n = 0;
  for (size_t i = 0; i + 1 < h->ev_used && n < capacity; i += 2, ++n) {
    float t = 0.f;
    cudaEventElapsedTime(&t, h->ev[i], h->ev[i + 1]);
    names[n] = h->ev_names[i / 2];
    ms[n] = t;
  }
  return n;
}

}  // extern "C"
