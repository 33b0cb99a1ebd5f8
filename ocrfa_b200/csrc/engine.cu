// libocrfa_b200.so: host-side engine + C ABI (include/ocrfa_b200.h) of the B200 recognition path.
//
// Holds the recognizer's weights (BatchNorm folded, repacked for the kernels), lays the activations out in a
// caller-provided workspace (NHWC, bf16 or fp32), and issues the kernel sequence of
//   Model.forward            /root/reference/src/ocrfa/modules/blocks.py:440-447
//   ResNet.forward           /root/reference/src/ocrfa/modules/blocks.py:340-391
//   BidirectionalLSTM.forward /root/reference/src/ocrfa/modules/blocks.py:92-102
// on one stream. No CPU fallback exists: every compute entry point needs a CUDA device.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/ocrfa_b200.h"
#include "decode_kernels.cuh"
#include "crop_kernels.cuh"
#include "det_kernels.cuh"
#include "igemm_plan.h"
#include "lstm_tc.cuh"
#include "resize_kernels.cuh"
#include "simt_kernels.cuh"

using namespace ocrfa;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                             \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) return fail(OCRFA_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

// One convolution with its BatchNorm folded in (and, on the bf16 path, the block's 1x1 downsample fused as
// extra K columns). Geometry per /root/reference/src/ocrfa/modules/blocks.py:239-322.
struct ConvW {
  std::string name;
  int cin = 0, cout = 0, kh = 0, kw = 0, sh = 1, ph = 0, pw = 0, dil = 1;
  int cin_pad = 0;     // channels of the input buffer the tensor-core path reads (multiple of 64)
  int ds_cin = 0;      // > 0: fused downsample source channels (bf16 path)
  int ktotal = 0;      // K of the packed weight matrix
  float out_scale = 1.f;   // split path: 2^-k undoing the power-of-two pre-scale of the fp16 hi/lo weights
  void* d_wgt = nullptr;   // bf16 [cout][ktotal] (tc), fp16 hi/lo [cout][2*ktotal'] (split) or float [cout][kh*kw*cin] (simt)
  float* d_bias = nullptr;
};

struct LstmW {
  void* d_wih = nullptr;    // [4096][512]  both directions, gate rows permuted to [tile][gate][16] (fp32 / split paths)
  float* d_bih = nullptr;   // [4096]       b_ih + b_hh, same permutation
  void* d_whh = nullptr;    // fp32: [4096][512]; bf16: [4096][W_ih 512 | W_hh 512]; split: [4096][W_hh hi 512 | lo 512]
  void* d_wlin = nullptr;   // [512][1024]
  float* d_blin = nullptr;  // [512]
  float s_wih = 1.f, s_whh = 1.f, s_wlin = 1.f;  // split path: output scales (2^-k) of the pre-scaled weights
};

struct WsLayout {
  size_t act[4], xproj, hseq, seq[2], feat, cstate, hh, sync, hx, total;
  size_t act_bytes;
};

}  // namespace

// Kernel classes for launch accounting and the optional CUDA-event profile (bench.py's roofline evidence).
enum KernelClass : int {
  kClsStem = 0, kClsIgemmTc = 1, kClsLstmTc = 2, kClsPool = 3, kClsSimtConv = 4, kClsLstmCell = 5, kClsDecode = 6,
  kClsBeam = 7, kNumClasses = 8
};

struct Profiler {
  uint32_t mask = 0;  // bit c set: class c is bracketed by events
  std::vector<cudaEvent_t> pool;
  struct Span { int cls; size_t e0, e1; double flops; };
  std::vector<Span> spans;
  size_t used = 0;
  int64_t launches[kNumClasses] = {0};
  cudaEvent_t next() {
    if (used == pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    return pool[used++];
  }
};

// One captured recognize step (all kernels of forward + decode) keyed by every pointer and size it bakes in.
struct GraphKey {
  const void *img, *widths, *mask; void *labels, *len, *conf, *margin, *ws;
  size_t ws_bytes; int B, H, W, decoder, bw;
  bool operator==(const GraphKey& o) const {
    return img == o.img && widths == o.widths && mask == o.mask && labels == o.labels && len == o.len && conf == o.conf && margin == o.margin &&
           ws == o.ws && ws_bytes == o.ws_bytes && B == o.B && H == o.H && W == o.W && decoder == o.decoder && bw == o.bw;
  }
};
struct GraphEntry {
  GraphKey key; cudaGraphExec_t exec; int64_t launches; int64_t by_class[8]; uint64_t last_use;
};

struct ocrfa_handle {
  std::vector<GraphEntry> graphs;
  int graphs_enabled = 1;  // OCRFA_GRAPHS=0 disables; also cleared after kMaxGraphFailures failed captures
  int graph_failures = 0;
  bool stem_1px = false;   // OCRFA_STEM_1PX (experiment switch), read once in ocrfa_create
  int x3_halo_max_cout = 128;  // OCRFA_X3_HALO_MAXC: f16x3 layers up to this Cout use the halo-staged form (0 = never)
  int x3_kchain = 2;           // K chunks of 64 per hi.hi accumulation in the recognizer's convolutions (igemm_x3.cuh); OCRFA_X3_CHAIN=1: one
  uint64_t graph_tick = 0;
  Profiler prof;
  ocrfa_config cfg;
  int num_sms = 148;
  int network = 0;     // 0 = recognizer (ocrfa_create), 1 = CRAFT detector (ocrfa_craft_create)
  int halo_level = 2;  // OCRFA_HALO: 0 = igemm_tc only, 1 = halo-staged kernel for Cout <= 128, 2 = wherever it tiles well
  bool finalized = false;
  std::map<std::string, HostTensor> host;
  std::map<std::string, ConvW> convs;
  float* d_stem_w = nullptr;  // [9][32]
  float* d_stem_b = nullptr;  // [32]
  StemWeights stem_host;      // the same, passed by value to the bf16 stem kernel (constant-bank operands)
  LstmW lstm[2];
  void* d_pred_w = nullptr;   // [num_class][512]
  float* d_pred_b = nullptr;
  std::vector<void*> allocs;
  int64_t launches = 0;
  // single-pass 16-bit tensor-core path: bf16 or fp16 operands/activations (same kernels, one format flag)
  bool bf16() const { return cfg.precision == OCRFA_PRECISION_BF16 || cfg.precision == OCRFA_PRECISION_F16; }
  bool f16() const { return cfg.precision == OCRFA_PRECISION_F16; }
  bool split() const { return cfg.precision == OCRFA_PRECISION_F16X3; }  // fp16 hi/lo operands, three products
  bool tc() const { return bf16() || split(); }                          // tcgen05 path (vs the CUDA-core fp32 path)
  float s_pred = 1.f;
};

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

WsLayout ws_layout(const ocrfa_handle* h, int B, int H, int W) {
  WsLayout L;
  const size_t es = h->bf16() ? 2 : 4;
  const size_t T = W / 4 + 1;
  L.act_bytes = align_up((size_t)B * H * W * 64 * es, 1024);
  size_t off = 0;
  for (int i = 0; i < 4; ++i) { L.act[i] = off; off += L.act_bytes; }
  L.xproj = off; off += align_up((size_t)B * T * 4096 * 4, 1024);
  L.hseq = off; off += align_up((size_t)B * T * 1024 * es, 1024);
  for (int i = 0; i < 2; ++i) { L.seq[i] = off; off += align_up((size_t)B * T * 512 * es, 1024); }
  L.feat = off; off += align_up((size_t)B * T * 512 * es, 1024);
  L.cstate = off; off += align_up((size_t)2 * B * 512 * 4, 1024);
  L.hh = off; off += align_up((size_t)2 * B * 2048 * 4, 1024);
  L.sync = off; off += 4096;  // group counters of the persistent LSTM launches (zeroed once per forward)
  L.hx = off; off += (size_t)2 * 2 * 2 * 16 * 128 * 64 * 2;  // LSTM h exchange: 2 dirs x 2 m_tiles x 2 parities x <= 16 slabs of 16 KB
  L.total = off;
  return L;
}

const HostTensor* find(const ocrfa_handle* h, const std::string& key) {
  auto it = h->host.find(key);
  return it == h->host.end() ? nullptr : &it->second;
}

template <typename T>
int upload(ocrfa_handle* h, const std::vector<T>& v, void** out) {
  void* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, v.size() * sizeof(T)));
  h->allocs.push_back(d);
  CUDA_TRY(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = d;
  return 0;
}

// Split-precision weight packing: w * 2^k = hi + lo (fp16 pair). k puts the largest |w| in [2^12, 2^13) so that the
// lo halves of all weights within 2^-15 of the largest stay in fp16's normal range; the kernels multiply the fp32
// accumulators by 2^-k (exact).
// float -> the 16-bit storage format of the single-pass path, returned as raw bits in a bf16-typed vector element
inline __nv_bfloat16 to_16(const ocrfa_handle* h, float v) {
  if (!h->f16()) return __float2bfloat16(v);
  const __half hv = __float2half_rn(std::max(-65504.f, std::min(65504.f, v)));
  __nv_bfloat16 r;
  memcpy(&r, &hv, 2);
  return r;
}

int split_scale_exp(double max_abs) {
  if (!(max_abs > 0.0) || !std::isfinite(max_abs)) return 0;
  int e = 0;
  std::frexp(max_abs, &e);  // max_abs = m * 2^e, m in [0.5, 1)
  return 13 - e;
}
inline void split_f16_host(double v, __half* hi, __half* lo) {
  const __half h = __float2half_rn((float)v);
  *hi = h;
  *lo = __float2half_rn((float)(v - (double)__half2float(h)));
}

// BN(eval) folded into the conv: y = conv(x) * s + (beta - mean * s), s = gamma / sqrt(var + eps)
int fold_bn(const ocrfa_handle* h, const std::string& conv, const std::string& bn, std::vector<double>* scale,
            std::vector<double>* shift, const HostTensor** w_out) {
  const HostTensor* w = find(h, conv + ".weight");
  const HostTensor* g = find(h, bn + ".weight");
  const HostTensor* b = find(h, bn + ".bias");
  const HostTensor* m = find(h, bn + ".running_mean");
  const HostTensor* v = find(h, bn + ".running_var");
  if (!w || !g || !b || !m || !v) return fail(OCRFA_ERR_WEIGHTS, "missing weights for %s / %s", conv.c_str(), bn.c_str());
  const size_t co = (size_t)w->shape[0];
  if (g->data.size() != co || b->data.size() != co || m->data.size() != co || v->data.size() != co)
    return fail(OCRFA_ERR_WEIGHTS, "BatchNorm size mismatch for %s", bn.c_str());
  scale->resize(co);
  shift->resize(co);
  for (size_t i = 0; i < co; ++i) {
    const double s = (double)g->data[i] / std::sqrt((double)v->data[i] + 1e-5);
    (*scale)[i] = s;
    (*shift)[i] = (double)b->data[i] - (double)m->data[i] * s;
  }
  *w_out = w;
  return 0;
}

int prep_conv(ocrfa_handle* h, const std::string& prefix, const std::string& conv, const std::string& bn, int kh, int kw,
              int sh, int ph, int pw, const std::string& ds_conv, const std::string& ds_bn) {
  std::vector<double> sc, sf, dsc, dsf;
  const HostTensor* w = nullptr;
  const HostTensor* dw = nullptr;
  int rc = fold_bn(h, prefix + conv, prefix + bn, &sc, &sf, &w);
  if (rc) return rc;
  if (w->shape.size() != 4 || w->shape[2] != kh || w->shape[3] != kw)
    return fail(OCRFA_ERR_WEIGHTS, "unexpected shape for %s", conv.c_str());
  ConvW c;
  c.name = conv;
  c.cout = (int)w->shape[0]; c.cin = (int)w->shape[1];
  c.kh = kh; c.kw = kw; c.sh = sh; c.ph = ph; c.pw = pw;
  const bool fuse_ds = !ds_conv.empty() && h->tc();
  if (fuse_ds) {
    rc = fold_bn(h, prefix + ds_conv, prefix + ds_bn, &dsc, &dsf, &dw);
    if (rc) return rc;
    c.ds_cin = (int)dw->shape[1];
  }
  std::vector<float> bias(c.cout);
  for (int o = 0; o < c.cout; ++o) bias[o] = (float)(sf[o] + (fuse_ds ? dsf[o] : 0.0));
  const int taps = kh * kw;
  if (h->split()) {
    // [cout][tap][hi(cin_pad) | lo(cin_pad)] (+ [hi(ds_pad) | lo(ds_pad)] for the fused downsample), fp16, pre-scaled
    c.cin_pad = ceil_div(c.cin, 64) * 64;
    const int ds_pad = fuse_ds ? ceil_div(c.ds_cin, 64) * 64 : 0;
    c.ktotal = 2 * (taps * c.cin_pad + ds_pad);
    double mx = 0.0;
    for (int o = 0; o < c.cout; ++o) {
      for (size_t i = 0; i < (size_t)c.cin * taps; ++i) mx = std::max(mx, std::fabs((double)w->data[(size_t)o * c.cin * taps + i] * sc[o]));
      if (fuse_ds)
        for (int ci = 0; ci < c.ds_cin; ++ci) mx = std::max(mx, std::fabs((double)dw->data[(size_t)o * c.ds_cin + ci] * dsc[o]));
    }
    const int k = split_scale_exp(mx);
    c.out_scale = (float)std::ldexp(1.0, -k);
    std::vector<__half> pk((size_t)c.cout * c.ktotal, __float2half(0.f));
    for (int o = 0; o < c.cout; ++o) {
      __half* row = &pk[(size_t)o * c.ktotal];
      for (int ci = 0; ci < c.cin; ++ci)
        for (int r = 0; r < kh; ++r)
          for (int s2 = 0; s2 < kw; ++s2) {
            const double v = std::ldexp((double)w->data[(((size_t)o * c.cin + ci) * kh + r) * kw + s2] * sc[o], k);
            const size_t base = (size_t)(r * kw + s2) * 2 * c.cin_pad;
            split_f16_host(v, &row[base + ci], &row[base + c.cin_pad + ci]);
          }
      if (fuse_ds)
        for (int ci = 0; ci < c.ds_cin; ++ci) {
          const double v = std::ldexp((double)dw->data[(size_t)o * c.ds_cin + ci] * dsc[o], k);
          const size_t base = (size_t)taps * 2 * c.cin_pad;
          split_f16_host(v, &row[base + ci], &row[base + ds_pad + ci]);
        }
    }
    rc = upload(h, pk, &c.d_wgt);
  } else if (h->bf16()) {
    c.cin_pad = ceil_div(c.cin, 64) * 64;
    const int ds_pad = fuse_ds ? ceil_div(c.ds_cin, 64) * 64 : 0;
    c.ktotal = taps * c.cin_pad + ds_pad;
    std::vector<__nv_bfloat16> pk((size_t)c.cout * c.ktotal, __float2bfloat16(0.f));
    for (int o = 0; o < c.cout; ++o) {
      for (int ci = 0; ci < c.cin; ++ci)
        for (int r = 0; r < kh; ++r)
          for (int s = 0; s < kw; ++s) {
            const double v = (double)w->data[(((size_t)o * c.cin + ci) * kh + r) * kw + s] * sc[o];
            pk[(size_t)o * c.ktotal + (size_t)(r * kw + s) * c.cin_pad + ci] = to_16(h, (float)v);
          }
      if (fuse_ds)
        for (int ci = 0; ci < c.ds_cin; ++ci)
          pk[(size_t)o * c.ktotal + (size_t)taps * c.cin_pad + ci] =
              to_16(h, (float)((double)dw->data[(size_t)o * c.ds_cin + ci] * dsc[o]));
    }
    rc = upload(h, pk, &c.d_wgt);
  } else {
    c.cin_pad = c.cin;
    c.ktotal = taps * c.cin;
    std::vector<float> pk((size_t)c.cout * c.ktotal);
    for (int o = 0; o < c.cout; ++o)
      for (int ci = 0; ci < c.cin; ++ci)
        for (int r = 0; r < kh; ++r)
          for (int s = 0; s < kw; ++s)
            pk[(size_t)o * c.ktotal + (size_t)(r * kw + s) * c.cin + ci] =
                (float)((double)w->data[(((size_t)o * c.cin + ci) * kh + r) * kw + s] * sc[o]);
    rc = upload(h, pk, &c.d_wgt);
  }
  if (rc) return rc;
  void* db = nullptr;
  rc = upload(h, bias, &db);
  if (rc) return rc;
  c.d_bias = (float*)db;
  h->convs[conv] = c;
  return 0;
}

// Linear layers as 1x1 "convs": weight [out][in] (+ optional row permutation), bias.
int prep_linear(ocrfa_handle* h, const std::vector<float>& w, const std::vector<float>& b, int out, int in, void** d_w,
                float** d_b, float* out_scale = nullptr) {
  int rc;
  if (h->split()) {
    // [out][hi(in) | lo(in)] fp16, pre-scaled by 2^k
    double mx = 0.0;
    for (float v : w) mx = std::max(mx, std::fabs((double)v));
    const int k = split_scale_exp(mx);
    if (out_scale) *out_scale = (float)std::ldexp(1.0, -k);
    std::vector<__half> pk((size_t)out * 2 * in);
    for (int o = 0; o < out; ++o)
      for (int i = 0; i < in; ++i)
        split_f16_host(std::ldexp((double)w[(size_t)o * in + i], k), &pk[(size_t)o * 2 * in + i], &pk[(size_t)o * 2 * in + in + i]);
    rc = upload(h, pk, d_w);
  } else if (h->bf16()) {
    std::vector<__nv_bfloat16> pk(w.size());
    for (size_t i = 0; i < w.size(); ++i) pk[i] = to_16(h, w[i]);
    rc = upload(h, pk, d_w);
  } else {
    rc = upload(h, w, d_w);
  }
  if (rc) return rc;
  void* db = nullptr;
  rc = upload(h, b, &db);
  *d_b = (float*)db;
  (void)out; (void)in;
  return rc;
}

// nn.LSTM gate rows are [i(512) | f(512) | g(512) | o(512)]; the kernels want, per direction, 32 tiles of
// [i(16) f(16) g(16) o(16)] so that one 64-column accumulator tile holds all four gates of 16 hidden units.
inline int lstm_perm_row(int new_row) {  // new row within a direction -> PyTorch row
  const int tile = new_row / 64, gate = (new_row % 64) / 16, u = new_row % 16;
  return gate * 512 + tile * 16 + u;
}

int prep_lstm(ocrfa_handle* h, int li) {
  const std::string pre = "SequenceModeling." + std::to_string(li) + ".";
  std::vector<float> wih((size_t)4096 * 512), whh((size_t)4096 * 512), bsum(4096);
  for (int dir = 0; dir < 2; ++dir) {
    const std::string sfx = dir ? "_reverse" : "";
    const HostTensor* a = find(h, pre + "rnn.weight_ih_l0" + sfx);
    const HostTensor* b = find(h, pre + "rnn.weight_hh_l0" + sfx);
    const HostTensor* c = find(h, pre + "rnn.bias_ih_l0" + sfx);
    const HostTensor* d = find(h, pre + "rnn.bias_hh_l0" + sfx);
    if (!a || !b || !c || !d) return fail(OCRFA_ERR_WEIGHTS, "missing LSTM weights for %s", pre.c_str());
    if (a->data.size() != (size_t)2048 * 512 || b->data.size() != (size_t)2048 * 512 || c->data.size() != 2048 ||
        d->data.size() != 2048)
      return fail(OCRFA_ERR_WEIGHTS, "unexpected LSTM shapes for %s", pre.c_str());
    for (int r = 0; r < 2048; ++r) {
      const int src = lstm_perm_row(r);
      memcpy(&wih[((size_t)dir * 2048 + r) * 512], &a->data[(size_t)src * 512], 512 * sizeof(float));
      memcpy(&whh[((size_t)dir * 2048 + r) * 512], &b->data[(size_t)src * 512], 512 * sizeof(float));
      bsum[dir * 2048 + r] = c->data[src] + d->data[src];
    }
  }
  LstmW& L = h->lstm[li];
  int rc;
  std::vector<float> zero_b(1, 0.f);
  float* dummy = nullptr;
  if (h->bf16()) {
    // fused input projection: resident weights of the persistent kernel are [W_ih | W_hh] (K = 1024)
    std::vector<float> wcat((size_t)4096 * 1024);
    for (int r = 0; r < 4096; ++r) {
      memcpy(&wcat[(size_t)r * 1024], &wih[(size_t)r * 512], 512 * sizeof(float));
      memcpy(&wcat[(size_t)r * 1024 + 512], &whh[(size_t)r * 512], 512 * sizeof(float));
    }
    rc = prep_linear(h, wcat, bsum, 4096, 1024, &L.d_whh, &L.d_bih);
    if (rc) return rc;
  } else {
    rc = prep_linear(h, wih, bsum, 4096, 512, &L.d_wih, &L.d_bih, &L.s_wih);
    if (rc) return rc;
    rc = prep_linear(h, whh, zero_b, 4096, 512, &L.d_whh, &dummy, &L.s_whh);  // split: [hi 512 | lo 512] per row
    if (rc) return rc;
  }
  const HostTensor* lw = find(h, pre + "linear.weight");
  const HostTensor* lb = find(h, pre + "linear.bias");
  if (!lw || !lb || lw->data.size() != (size_t)512 * 1024) return fail(OCRFA_ERR_WEIGHTS, "missing %slinear", pre.c_str());
  return prep_linear(h, lw->data, lb->data, 512, 1024, &L.d_wlin, &L.d_blin, &L.s_wlin);
}

// ------------------------------------------------------------------------------------------------------------
// kernel issue helpers
// ------------------------------------------------------------------------------------------------------------
struct Act {
  void* p; int n, h, w, c;  // NHWC view; c = channels of the buffer
};

// Counts one kernel launch of class `cls`; when that class is being profiled, brackets it with CUDA events on the
// launching stream (begin in the constructor, end in the destructor, i.e. right after the <<<>>>).
struct LaunchScope {
  ocrfa_handle* h; cudaStream_t st; int cls; size_t span = (size_t)-1;
  LaunchScope(ocrfa_handle* h_, int cls_, double flops, cudaStream_t st_) : h(h_), st(st_), cls(cls_) {
    h->launches++;
    h->prof.launches[cls]++;
    if (h->prof.mask & (1u << cls)) {
      const size_t e0 = h->prof.used;
      cudaEventRecord(h->prof.next(), st);
      span = h->prof.spans.size();
      h->prof.spans.push_back({cls, e0, 0, flops});
    }
  }
  ~LaunchScope() {
    if (span != (size_t)-1) {
      h->prof.spans[span].e1 = h->prof.used;
      cudaEventRecord(h->prof.next(), st);
    }
  }
};

// Halo-staged kernel (igemm_halo.cuh) for the 3x3/s1/p1 convs of the bf16 path, optionally with the following 2x2
// max pool fused. Returns 1 when the layer is not a candidate (the caller falls back to igemm_tc [+ pool kernel]).
bool halo_wanted(const ocrfa_handle* h, const ConvW& cw, const Act& in, bool pool) {
  if (!h->bf16() || h->halo_level <= 0) return false;
  const ConvGeom g{in.n, in.h, in.w, cw.cin, cw.cout, cw.kh, cw.kw, cw.sh, cw.ph, cw.pw, cw.dil};
  if (!halo_eligible(g, pool) || (long long)in.n * in.h * in.w >= (1ll << 31)) return false;
  if (cw.cout > 128 && h->halo_level < 2) return false;
  int tw = 0, th = 0;
  halo_choose_tile(in.w, in.h, pool, &tw, &th);
  const double cover = double(ceil_div(in.w, tw) * tw) * double(ceil_div(in.h, th) * th) / (double(in.w) * in.h);
  if (cover > 1.06) return false;  // e.g. 4 x 129 feature maps tile better as {4, 4, 8} boxes across images (igemm_tc)
  // latency shapes: a wide layer with only a few M tiles spreads over more SMs as narrow N tiles of the per-tap kernel
  const int pair_items = ceil_div(ceil_div(in.w, tw) * ceil_div(in.h, th) * in.n, 2);
  if (cw.cout > 128 && pair_items < kPlanSmCount / 8) return false;
  return true;
}

int run_conv_halo(ocrfa_handle* h, const ConvW& cw, const Act& in, const Act* ds_src, const void* residual, void* out,
                  bool relu, bool pool, cudaStream_t st, int ldo = 0) {
  IgemmHaloParams p;
  const ConvGeom g{in.n, in.h, in.w, cw.cin, cw.cout, cw.kh, cw.kw, cw.sh, cw.ph, cw.pw, cw.dil};
  if (cw.ds_cin && !ds_src) return fail(OCRFA_ERR_STATE, "%s expects a fused downsample source", cw.name.c_str());
  int rc = plan_halo(&p, g, in.p, in.c, cw.d_wgt, pool, cw.ds_cin ? ds_src->p : nullptr, cw.ds_cin, cw.ds_cin ? ds_src->c : 0);
  if (rc) return fail(OCRFA_ERR_CUDA, "plan_halo(%s) failed: %d", cw.name.c_str(), rc);
  p.bias = cw.d_bias;
  p.residual = residual; p.ldr = cw.cout;
  p.out = out; p.ldo = ldo > 0 ? ldo : cw.cout; p.relu = relu ? 1 : 0;
  p.f16 = h->f16() ? 1 : 0;
  const double flops = 2.0 * (double)in.n * in.h * in.w * cw.cout * (9.0 * cw.cin + cw.ds_cin);
  LaunchScope ls(h, kClsIgemmTc, flops, st);
  CUDA_TRY(igemm_halo_launch(p, h->num_sms, st));
  return 0;
}

// fp32-grade mode: the halo-staged form of igemm_x3_kernel for the 3x3/s1/p1 layers with few output channels on large
// maps (conv0_2, layer1, conv1: per-tap staging would move more L2 -> SM bytes per MMA than the fabric delivers).
bool x3_halo_wanted(const ocrfa_handle* h, const ConvW& cw, const Act& in, bool pool) {
  if (!h->split() || h->halo_level <= 0) return false;
  const ConvGeom g{in.n, in.h, in.w, cw.cin, cw.cout, cw.kh, cw.kw, cw.sh, cw.ph, cw.pw, cw.dil};
  if (!halo_eligible(g, pool) || (long long)in.n * in.h * in.w >= (1ll << 31)) return false;
  if (cw.cout != 64 && cw.cout != 128 && cw.cout != 256) return false;
  if (cw.cout > h->x3_halo_max_cout || in.c % 64 != 0) return false;
  int tw = 0, th = 0;
  halo_choose_tile(in.w, in.h, pool, &tw, &th);
  const double cover = double(ceil_div(in.w, tw) * tw) * double(ceil_div(in.h, th) * th) / (double(in.w) * in.h);
  if (cover > 1.06) return false;
  int stages = 0;
  igemm_x3_halo_smem_bytes(cw.cout, tw, th, &stages);
  return stages >= 2;
}

int run_conv_x3_halo(ocrfa_handle* h, const ConvW& cw, const Act& in, const Act* ds_src, const void* residual, void* out,
                     bool relu, bool pool, cudaStream_t st) {
  IgemmTcParams p;
  const ConvGeom g{in.n, in.h, in.w, cw.cin, cw.cout, cw.kh, cw.kw, cw.sh, cw.ph, cw.pw, cw.dil};
  int rc = plan_x3_halo(&p, g, in.p, in.c, cw.d_wgt, pool);
  if (rc) return fail(OCRFA_ERR_CUDA, "plan_x3_halo(%s) failed: %d", cw.name.c_str(), rc);
  if (cw.ds_cin) {
    if (!ds_src) return fail(OCRFA_ERR_STATE, "%s expects a fused downsample source", cw.name.c_str());
    rc = plan_add_source(&p, ds_src->p, cw.ds_cin, ds_src->c, cw.d_wgt, cw.ktotal, cw.cout);
    if (rc) return fail(OCRFA_ERR_CUDA, "plan_add_source(%s) failed: %d", cw.name.c_str(), rc);
  }
  p.bias = cw.d_bias;
  p.residual = residual; p.ldr = 2 * cw.cout; p.res_lo = cw.cout;
  p.out = out; p.ldo = 2 * cw.cout; p.out_lo = cw.cout; p.out_f32 = 0; p.relu = relu ? 1 : 0;
  p.out_scale = cw.out_scale;
  p.kchain = h->x3_kchain;
  const double flops = 2.0 * (double)in.n * in.h * in.w * cw.cout * (9.0 * cw.cin + cw.ds_cin);
  LaunchScope ls(h, kClsIgemmTc, flops, st);
  CUDA_TRY(igemm_tc_launch(p, kModeStd, h->num_sms, st));
  return 0;
}

int run_conv_tc(ocrfa_handle* h, const ConvW& cw, const Act& in, const Act* ds_src, const void* residual, void* out,
                bool relu, bool out_f32, int ldo, cudaStream_t st) {
  IgemmTcParams p;
  ConvGeom g{in.n, in.h, in.w, cw.cin, cw.cout, cw.kh, cw.kw, cw.sh, cw.ph, cw.pw, cw.dil};
  const bool split = h->split();
  if (!out_f32 && (ldo % 8) == 0 && halo_wanted(h, cw, in, false))
    return run_conv_halo(h, cw, in, ds_src, residual, out, relu, false, st, ldo);
  if (split && !out_f32 && ldo == cw.cout && x3_halo_wanted(h, cw, in, false))
    return run_conv_x3_halo(h, cw, in, ds_src, residual, out, relu, false, st);
  int rc = plan_conv(&p, g, in.p, in.c, cw.d_wgt, nullptr, split);
  if (rc) return fail(OCRFA_ERR_CUDA, "plan_conv(%s) failed: %d", cw.name.c_str(), rc);
  if (cw.ds_cin) {
    if (!ds_src) return fail(OCRFA_ERR_STATE, "%s expects a fused downsample source", cw.name.c_str());
    rc = plan_add_source(&p, ds_src->p, cw.ds_cin, ds_src->c, cw.d_wgt, cw.ktotal, cw.cout);
    if (rc) return fail(OCRFA_ERR_CUDA, "plan_add_source(%s) failed: %d", cw.name.c_str(), rc);
  }
  p.bias = cw.d_bias;
  p.residual = residual; p.ldr = cw.cout;
  p.out = out; p.ldo = ldo; p.out_f32 = out_f32 ? 1 : 0; p.relu = relu ? 1 : 0;
  p.f16 = h->f16() ? 1 : 0;
  if (split) {
    // fp16 hi/lo pixels: [hi(C) | lo(C)]
    p.out_scale = cw.out_scale;
    p.kchain = h->x3_kchain;
    p.ldr = 2 * cw.cout; p.res_lo = cw.cout;
    if (!out_f32) { p.ldo = 2 * ldo; p.out_lo = ldo; }
  }
  // algorithmic FLOPs: 2 * output pixels * Cout * real K (channel padding and ragged-tile rows not counted)
  const double flops = 2.0 * (double)in.n * p.h_out * p.w_out * cw.cout * ((double)cw.kh * cw.kw * cw.cin + cw.ds_cin);
  LaunchScope ls(h, kClsIgemmTc, flops, st);
  CUDA_TRY(igemm_tc_launch(p, kModeStd, h->num_sms, st));
  return 0;
}

int run_conv_simt(ocrfa_handle* h, const float* wgt, const float* bias, const void* in, long long in_pix_stride, int n,
                  int h_in, int w_in, int cin, int cout, int kh, int kw, int sh, int ph, int pw, const void* residual,
                  void* out, int ldo, bool relu, cudaStream_t st) {
  ConvSimtParams p;
  p.in = in; p.in_pix_stride = in_pix_stride; p.wgt = wgt; p.bias = bias; p.residual = residual; p.ldr = cout;
  p.out = out; p.ldo = ldo; p.n = n; p.h_in = h_in; p.w_in = w_in; p.cin = cin; p.cout = cout; p.kh = kh; p.kw = kw;
  p.sh = sh; p.ph = ph; p.pw = pw;
  p.h_out = (h_in + 2 * ph - kh) / sh + 1; p.w_out = w_in + 2 * pw - kw + 1; p.relu = relu ? 1 : 0;
  const long long pix = (long long)n * p.h_out * p.w_out;
  dim3 grid((unsigned)((pix + 63) / 64), (unsigned)((cout + 63) / 64));
  LaunchScope ls(h, kClsSimtConv, 2.0 * (double)pix * cout * kh * kw * cin, st);
  conv_simt_kernel<float, float><<<grid, 256, 0, st>>>(p);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

template <typename T>
int run_pool(ocrfa_handle* h, const Act& in, void* out, int sh, int sw, int pw, int* ho, int* wo, cudaStream_t st) {
  *ho = (in.h - 2) / sh + 1;
  *wo = (in.w + 2 * pw - 2) / sw + 1;
  const long long total = (long long)in.n * *ho * *wo * (in.c / 8);
  LaunchScope ls(h, kClsPool, 0.0, st);
  maxpool2x2_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const T*)in.p, (T*)out, in.n, in.h, in.w, in.c,
                                                                       sh, sw, pw, *ho, *wo);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int run_pool_split(ocrfa_handle* h, const Act& in, void* out, int sh, int sw, int pw, int* ho, int* wo, cudaStream_t st) {
  *ho = (in.h - 2) / sh + 1;
  *wo = (in.w + 2 * pw - 2) / sw + 1;
  const long long total = (long long)in.n * *ho * *wo * (in.c / 8);
  LaunchScope ls(h, kClsPool, 0.0, st);
  maxpool2x2_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const __half*)in.p, (__half*)out, in.n, in.h,
                                                                           in.w, in.c, sh, sw, pw, *ho, *wo);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
struct BufPool {
  void* buf[4];
  bool used[4] = {false, false, false, false};
  void* get() {
    for (int i = 0; i < 4; ++i)
      if (!used[i]) { used[i] = true; return buf[i]; }
    return nullptr;
  }
  void put(void* p) {
    for (int i = 0; i < 4; ++i)
      if (buf[i] == p) used[i] = false;
  }
};

// d_feat_out != null: also emit the visual features [B,T,512] as fp32 (and stop there when d_logits is null).
// d_feat_in  != null: skip the conv stack and run the sequence head on caller-provided fp32 features.
int forward_impl(ocrfa_handle* h, const float* d_x, const uint8_t* d_u8, const int32_t* d_widths, int B, int H, int W,
                 float* d_logits, void* d_ws, size_t ws_bytes, cudaStream_t st, float* d_feat_out = nullptr,
                 const float* d_feat_in = nullptr) {
  if (!h || !h->finalized) return fail(OCRFA_ERR_STATE, "weights not finalized");
  if (B <= 0 || H < 32 || H % 16 != 0 || W < 4 || W % 4 != 0) return fail(OCRFA_ERR_ARG, "unsupported geometry B=%d H=%d W=%d (H multiple of 16 >= 32, W multiple of 4)", B, H, W);
  if ((!d_x && !d_u8 && !d_feat_in) || (!d_logits && !d_feat_out) || !d_ws) return fail(OCRFA_ERR_ARG, "null pointer");
  const WsLayout L = ws_layout(h, B, H, W);
  if (ws_bytes < L.total) return fail(OCRFA_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, L.total);
  if (((uintptr_t)d_ws & 1023) != 0) return fail(OCRFA_ERR_WORKSPACE, "workspace must be 1024-byte aligned");
  const bool bf = h->bf16();
  const bool sp = h->split();
  const bool tc = bf || sp;
  uint8_t* ws = (uint8_t*)d_ws;
  BufPool pool;
  for (int i = 0; i < 4; ++i) pool.buf[i] = ws + L.act[i];
  int rc;

  auto conv = [&](const char* name, const Act& in, const Act* ds_src, const void* residual, void* out, bool relu) -> int {
    auto it = h->convs.find(name);
    if (it == h->convs.end()) return fail(OCRFA_ERR_STATE, "conv %s not prepared", name);
    const ConvW& cw = it->second;
    if (tc) return run_conv_tc(h, cw, in, ds_src, residual, out, relu, false, cw.cout, st);
    return run_conv_simt(h, (const float*)cw.d_wgt, cw.d_bias, in.p, in.c, in.n, in.h, in.w, cw.cin, cw.cout, cw.kh,
                         cw.kw, cw.sh, cw.ph, cw.pw, residual, out, cw.cout, relu, st);
  };
  auto out_dims = [&](const char* name, const Act& in, int* ho, int* wo, int* co) {
    const ConvW& cw = h->convs.find(name)->second;
    *ho = (in.h + 2 * cw.ph - cw.kh) / cw.sh + 1;
    *wo = in.w + 2 * cw.pw - cw.kw + 1;
    *co = cw.cout;
  };
  // conv + bn + relu into a fresh buffer; releases the input
  auto conv_step = [&](const char* name, Act* a) -> int {
    void* o = pool.get();
    int ho, wo, co;
    out_dims(name, *a, &ho, &wo, &co);
    int r = conv(name, *a, nullptr, nullptr, o, true);
    if (r) return r;
    pool.put(a->p);
    *a = Act{o, a->n, ho, wo, co};
    return 0;
  };
  // BlockB: relu(bn2(conv2(relu(bn1(conv1(x))))) + residual), residual through 1x1 conv+bn when channels change
  auto block = [&](const std::string& pre, bool has_ds, Act* a) -> int {
    void* t = pool.get();
    int ho, wo, co;
    out_dims((pre + "conv1").c_str(), *a, &ho, &wo, &co);
    int r = conv((pre + "conv1").c_str(), *a, nullptr, nullptr, t, true);
    if (r) return r;
    Act ta{t, a->n, ho, wo, co};
    void* o = pool.get();
    if (has_ds && !tc) {
      void* rs = pool.get();
      r = conv((pre + "downsample.0").c_str(), *a, nullptr, nullptr, rs, false);
      if (r) return r;
      r = conv((pre + "conv2").c_str(), ta, nullptr, rs, o, true);
      pool.put(rs);
    } else if (has_ds) {
      r = conv((pre + "conv2").c_str(), ta, a, nullptr, o, true);  // downsample fused as extra K slabs
    } else {
      r = conv((pre + "conv2").c_str(), ta, nullptr, a->p, o, true);
    }
    if (r) return r;
    pool.put(t);
    pool.put(a->p);
    *a = Act{o, a->n, ho, wo, co};
    return 0;
  };
  auto pool_step = [&](int sh, int sw, int pw, Act* a) -> int {
    void* o = pool.get();
    int ho, wo;
    int r = sp ? run_pool_split(h, *a, o, sh, sw, pw, &ho, &wo, st)
               : (bf && h->f16()) ? run_pool<__half>(h, *a, o, sh, sw, pw, &ho, &wo, st)
               : bf ? run_pool<__nv_bfloat16>(h, *a, o, sh, sw, pw, &ho, &wo, st) : run_pool<float>(h, *a, o, sh, sw, pw, &ho, &wo, st);
    if (r) return r;
    pool.put(a->p);
    *a = Act{o, a->n, ho, wo, a->c};
    return 0;
  };

  // conv + bn + relu + 2x2/2 max pool: one halo-staged launch when the layer qualifies, else conv then pool
  auto conv_pool_step = [&](const char* name, Act* a) -> int {
    auto it = h->convs.find(name);
    if (it == h->convs.end()) return fail(OCRFA_ERR_STATE, "conv %s not prepared", name);
    if (halo_wanted(h, it->second, *a, true) || x3_halo_wanted(h, it->second, *a, true)) {
      void* o = pool.get();
      int r = sp ? run_conv_x3_halo(h, it->second, *a, nullptr, nullptr, o, true, true, st)
                 : run_conv_halo(h, it->second, *a, nullptr, nullptr, o, true, true, st);
      if (r) return r;
      pool.put(a->p);
      *a = Act{o, a->n, a->h / 2, a->w / 2, it->second.cout};
      return 0;
    }
    int r = conv_step(name, a);
    if (r) return r;
    return pool_step(2, 2, 0, a);
  };

  int T = W / 4 + 1;
  void* feat = ws + L.feat;
  if (d_feat_in) {
    const long long total = (long long)B * T * 512;
    LaunchScope ls(h, kClsPool, 0.0, st);
    if (sp) cvt_f32_to_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_feat_in, (__half*)feat, (long long)B * T, 512);
    else if (bf && h->f16()) cvt_from_f32_kernel<__half><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_feat_in, (__half*)feat, total);
    else if (bf) cvt_from_f32_kernel<__nv_bfloat16><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_feat_in, (__nv_bfloat16*)feat, total);
    else cvt_from_f32_kernel<float><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_feat_in, (float*)feat, total);
    CUDA_TRY(cudaGetLastError());
  } else {
  // ---- stem: conv0_1 + bn0_1 + relu (blocks.py:341-347)
  const bool stem32 = bf && h->halo_level > 0 && (long long)B * H * W < (1ll << 31) && halo_wanted(h, h->convs.find("conv0_2")->second, Act{nullptr, B, H, W, 32}, false);
  Act a{pool.get(), B, H, W, (tc && !stem32) ? 64 : 32};
  {
    const long long pix = (long long)B * H * W;
    const unsigned grid = (unsigned)((pix + 255) / 256);
    LaunchScope ls(h, kClsStem, 2.0 * (double)pix * 32 * 9, st);
    // four pixels per thread once there are enough pixels to fill the machine; small batches keep one pixel per thread
    // (a latency shape: 4096 pixels at config 1)
    if (stem32 && (W & 3) == 0 && pix >= (1ll << 20) && !h->stem_1px) {
      const unsigned g4 = (unsigned)((pix / 4 + 127) / 128);
      if (h->f16()) stem_conv32x4_kernel<true><<<g4, 128, 0, st>>>(d_x, d_u8, d_widths, h->stem_host, (__nv_bfloat16*)a.p, B, H, W);
      else stem_conv32x4_kernel<false><<<g4, 128, 0, st>>>(d_x, d_u8, d_widths, h->stem_host, (__nv_bfloat16*)a.p, B, H, W);
    } else if (stem32)
      stem_conv32_kernel<<<(unsigned)((pix + 255) / 256), 256, 0, st>>>(d_x, d_u8, d_widths, h->stem_host,
                                                                             (__nv_bfloat16*)a.p, B, H, W, h->f16() ? 1 : 0);
    else if (sp)
      stem_conv_kernel<__half, 128><<<grid, 256, 0, st>>>(d_x, d_u8, d_widths, h->d_stem_w, h->d_stem_b, (__half*)a.p, B, H, W);
    else if (bf && h->f16())
      stem_conv_kernel<__half, 64><<<grid, 256, 0, st>>>(d_x, d_u8, d_widths, h->d_stem_w, h->d_stem_b, (__half*)a.p, B, H, W);
    else if (bf)
      stem_conv_kernel<__nv_bfloat16, 64><<<grid, 256, 0, st>>>(d_x, d_u8, d_widths, h->d_stem_w, h->d_stem_b,
                                                                 (__nv_bfloat16*)a.p, B, H, W);
    else
      stem_conv_kernel<float, 32><<<grid, 256, 0, st>>>(d_x, d_u8, d_widths, h->d_stem_w, h->d_stem_b, (float*)a.p, B, H, W);
    CUDA_TRY(cudaGetLastError());
  }
  if ((rc = conv_pool_step("conv0_2", &a))) return rc;                 // + max_pool1, blocks.py:259
  if ((rc = block("layer1.0.", true, &a))) return rc;
  if ((rc = conv_pool_step("conv1", &a))) return rc;                   // + max_pool2, blocks.py:271-275
  if ((rc = block("layer2.0.", true, &a))) return rc;
  if ((rc = block("layer2.1.", false, &a))) return rc;
  if ((rc = conv_step("conv2", &a))) return rc;
  if ((rc = pool_step(2, 1, 1, &a))) return rc;                       // max_pool3 k2 s(2,1) p(0,1), blocks.py:292
  if ((rc = block("layer3.0.", true, &a))) return rc;
  for (int i = 1; i < 5; ++i)
    if ((rc = block("layer3." + std::to_string(i) + ".", false, &a))) return rc;
  if ((rc = conv_step("conv3", &a))) return rc;
  for (int i = 0; i < 3; ++i)
    if ((rc = block("layer4." + std::to_string(i) + ".", false, &a))) return rc;
  if ((rc = conv_step("conv4_1", &a))) return rc;                      // 2x2 s(2,1) p(0,1), blocks.py:305-313
  if ((rc = conv_step("conv4_2", &a))) return rc;                      // 2x2 s1 p0, blocks.py:314-322

  // ---- AdaptiveAvgPool2d((None,1)) over H' + [B,T,512] layout (blocks.py:441-446)
  T = a.w;
  feat = a.p;
  if (a.h > 1) {
    feat = ws + L.feat;
    const long long total = (long long)B * T * 512;
    LaunchScope ls(h, kClsPool, 0.0, st);
    if (sp) mean_h_split_kernel<<<(unsigned)((total / 8 + 255) / 256), 256, 0, st>>>((const __half*)a.p, (__half*)feat, B, a.h, T, 512);
    else if (bf && h->f16()) mean_h_kernel<__half><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const __half*)a.p, (__half*)feat, B, a.h, T, 512);
    else if (bf) mean_h_kernel<__nv_bfloat16><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const __nv_bfloat16*)a.p, (__nv_bfloat16*)feat, B, a.h, T, 512);
    else mean_h_kernel<float><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const float*)a.p, (float*)feat, B, a.h, T, 512);
    CUDA_TRY(cudaGetLastError());
  }
  }  // conv stack
  if (d_feat_out) {
    const long long total = (long long)B * T * 512;
    {
    LaunchScope ls(h, kClsPool, 0.0, st);
    if (sp) cvt_split_to_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const __half*)feat, d_feat_out, (long long)B * T, 512);
    else if (bf && h->f16()) cvt_to_f32_kernel<__half><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const __half*)feat, d_feat_out, total);
    else if (bf) cvt_to_f32_kernel<__nv_bfloat16><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const __nv_bfloat16*)feat, d_feat_out, total);
    else cvt_to_f32_kernel<float><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const float*)feat, d_feat_out, total);
    CUDA_TRY(cudaGetLastError());
    }
    if (!d_logits) return 0;
  }

  // ---- 2 x BidirectionalLSTM (blocks.py:92-102)
  if (tc) CUDA_TRY(cudaMemsetAsync(ws + L.sync, 0, 4096, st));
  const int rows = B * T;
  float* xproj = (float*)(ws + L.xproj);
  void* hseq = ws + L.hseq;
  float* cst = (float*)(ws + L.cstate);
  float* hh = (float*)(ws + L.hh);
  void* seq_in = feat;
  for (int li = 0; li < 2; ++li) {
    const LstmW& lw = h->lstm[li];
    void* seq_out = ws + L.seq[li];
    if (tc) {
      if (sp) {
        // split path: input projection for both directions as one split igemm:
        // [rows,512] x [4096,512]^T + (b_ih + b_hh) -> fp32 gate pre-activations
        ConvW xp; xp.name = "lstm.xproj"; xp.cin = 512; xp.cout = 4096; xp.kh = xp.kw = 1; xp.sh = 1; xp.cin_pad = 512;
        xp.ktotal = 1024; xp.d_wgt = lw.d_wih; xp.d_bias = lw.d_bih; xp.out_scale = lw.s_wih;
        if ((rc = run_conv_tc(h, xp, Act{seq_in, 1, 1, rows, 512}, nullptr, nullptr, xproj, false, true, 4096, st))) return rc;
      }
      // recurrence: ONE persistent launch per 256-line slice walks all T steps of both directions (lstm_tc.cuh);
      // on the bf16 path the input projection is fused into it (K = [x_t | h_{t-1}])
      const int mode = sp ? kLstmF16Split : (h->f16() ? kLstmF16FusedX : kLstmBf16FusedX);
      const size_t hs_row = sp ? 2048 : 1024;  // elements per (b, t) row of the h sequence
      int slice = 0;
      for (int b0 = 0; b0 < B; b0 += 256, ++slice) {
        const int nb = B - b0 < 256 ? B - b0 : 256;
        LstmTcParams lp;
        memset(&lp, 0, sizeof(lp));
        lp.m_tiles = ceil_div(nb, 128);
        lp.hx = ws + L.hx;
        lp.box_rows = lp.m_tiles > 1 ? 128 : std::max(8, (nb + 7) & ~7);
        int r2 = make_rows64_tmap(&lp.tmX, lp.hx, lstm_hx_rows(mode, lp.m_tiles), lp.box_rows);
        if (r2 == 0) r2 = make_wgt_tmap(&lp.tmW, lw.d_whh, 1024, 4096, 64);
        if (r2 == 0) r2 = make_rows64_store_tmap(&lp.tmP, lp.hx, lstm_hx_rows(mode, lp.m_tiles), lp.box_rows);
        if (r2 == 0 && !sp) r2 = make_hseq_store_tmap(&lp.tmH, (uint16_t*)hseq + (size_t)b0 * T * hs_row, T, nb, lp.box_rows);
        if (r2 == 0 && !sp)
          r2 = make_act_tmap(&lp.tmS, (const __nv_bfloat16*)seq_in + (size_t)b0 * T * 512, 512, T, 1, nb, 1, 1, lp.box_rows, 1);
        if (r2) return fail(OCRFA_ERR_CUDA, "LSTM tensor map failed: %d", r2);
        lp.xproj = sp ? xproj + (size_t)b0 * T * 4096 : nullptr;
        lp.bias = lw.d_bih;
        lp.hseq = (uint16_t*)hseq + (size_t)b0 * T * hs_row;
        lp.acc_scale = lw.s_whh;
        const int ctr = (li * 64 + slice) * 4;
        if (ctr + 4 > 1024) return fail(OCRFA_ERR_ARG, "batch too large for the LSTM counter table");
        lp.sync = (unsigned int*)(ws + L.sync) + ctr;
        lp.dbg = nullptr;
        lp.B = nb; lp.T = T;
        const double kx = sp ? 0.0 : 512.0 * T;  // fused input projection
        LaunchScope ls(h, kClsLstmTc, 2.0 * 2.0 * nb * 2048.0 * (512.0 * (T - 1) + kx), st);
        CUDA_TRY(lstm_tc_launch(lp, mode, st));
      }
      ConvW lin; lin.name = "lstm.linear"; lin.cin = 1024; lin.cout = 512; lin.kh = lin.kw = 1; lin.sh = 1; lin.cin_pad = 1024;
      lin.ktotal = sp ? 2048 : 1024; lin.d_wgt = lw.d_wlin; lin.d_bias = lw.d_blin; lin.out_scale = lw.s_wlin;
      if ((rc = run_conv_tc(h, lin, Act{hseq, 1, 1, rows, 1024}, nullptr, nullptr, seq_out, false, false, 512, st))) return rc;
    } else {
      if ((rc = run_conv_simt(h, (const float*)lw.d_wih, lw.d_bih, seq_in, 512, 1, 1, rows, 512, 4096, 1, 1, 1, 0, 0,
                              nullptr, xproj, 4096, false, st))) return rc;
      for (int s = 0; s < T; ++s) {
        if (s > 0)
          for (int dir = 0; dir < 2; ++dir) {
            const int t_prev = dir == 0 ? s - 1 : T - s;
            const float* hin = (const float*)hseq + (size_t)t_prev * 1024 + dir * 512;
            if ((rc = run_conv_simt(h, (const float*)lw.d_whh + (size_t)dir * 2048 * 512, nullptr, hin, (long long)T * 1024,
                                    1, 1, B, 512, 2048, 1, 1, 1, 0, 0, nullptr, hh + (size_t)dir * B * 2048, 2048, false, st)))
              return rc;
          }
        const int total = 2 * B * 512;
        LaunchScope ls(h, kClsLstmCell, 0.0, st);
        lstm_cell_f32_kernel<<<(total + 255) / 256, 256, 0, st>>>(hh, xproj, cst, (float*)hseq, B, T, s);
        CUDA_TRY(cudaGetLastError());
      }
      if ((rc = run_conv_simt(h, (const float*)lw.d_wlin, lw.d_blin, hseq, 1024, 1, 1, rows, 1024, 512, 1, 1, 1, 0, 0,
                              nullptr, seq_out, 512, false, st))) return rc;
    }
    seq_in = seq_out;
  }

  // ---- Prediction: Linear(512 -> num_class) (blocks.py:438, 441)
  const int C = h->cfg.num_class;
  if (tc) {
    ConvW pr; pr.name = "prediction"; pr.cin = 512; pr.cout = C; pr.kh = pr.kw = 1; pr.sh = 1; pr.cin_pad = 512;
    pr.ktotal = sp ? 1024 : 512; pr.d_wgt = h->d_pred_w; pr.d_bias = h->d_pred_b; pr.out_scale = h->s_pred;
    if ((rc = run_conv_tc(h, pr, Act{seq_in, 1, 1, rows, 512}, nullptr, nullptr, d_logits, false, true, C, st))) return rc;
  } else {
    if ((rc = run_conv_simt(h, (const float*)h->d_pred_w, h->d_pred_b, seq_in, 512, 1, 1, rows, 512, C, 1, 1, 1, 0, 0, nullptr,
                            d_logits, C, false, st))) return rc;
  }
  return 0;
}

#include "craft_engine.inc"

size_t beam_ws_bytes(int B, int T, int C, int bw) {
  (void)C;
  return align_up((size_t)B * 2 * bw * T * sizeof(int16_t), 256);
}

struct RecLayout {
  size_t fwd, img, widths, mask, logits, probs, labels, len, conf, margin, beam, total;
};

RecLayout rec_layout(const ocrfa_handle* h, int B, int H, int W, int bw) {
  RecLayout R;
  const int T = W / 4 + 1, C = h->cfg.num_class;
  size_t off = 0;
  R.fwd = off; off += align_up(ws_layout(h, B, H, W).total, 1024);
  R.img = off; off += align_up((size_t)B * H * W, 1024);
  R.widths = off; off += align_up((size_t)B * 4, 1024);
  R.mask = off; off += align_up((size_t)C, 1024);
  R.logits = off; off += align_up((size_t)B * T * C * 4, 1024);
  R.probs = off; off += align_up((size_t)B * T * C * 4, 1024);
  R.labels = off; off += align_up((size_t)B * T * 4, 1024);
  R.len = off; off += align_up((size_t)B * 4, 1024);
  R.conf = off; off += align_up((size_t)B * 8, 1024);
  R.margin = off; off += align_up((size_t)B * 4, 1024);
  R.beam = off; off += align_up(beam_ws_bytes(B, T, C, bw > 0 ? bw : 1), 1024);
  R.total = off;
  return R;
}

}  // namespace

namespace {
// forward + decode on device buffers, asynchronous on `st` (labels/len/conf stay on the device)
int recognize_device_impl(ocrfa_handle* h, const uint8_t* d_img, const int32_t* d_widths, int B, int H, int W,
                          const uint8_t* d_ignore_mask, int decoder, int beam_width, int32_t* d_labels, int32_t* d_len,
                          double* d_conf, float* d_margin, void* d_ws, size_t ws_bytes, cudaStream_t st);
}  // namespace

// ==============================================================================================================
// C ABI
// ==============================================================================================================
extern "C" {

int ocrfa_abi_version(void) { return 1; }
const char* ocrfa_last_error(void) { return g_err; }

int ocrfa_create(const ocrfa_config* cfg, ocrfa_handle** out) {
  if (!cfg || !out) return fail(OCRFA_ERR_ARG, "null argument");
  if (cfg->input_channel != 1 || cfg->output_channel != 512 || cfg->hidden_size != 512)
    return fail(OCRFA_ERR_ARG, "only the reference's generation1 network (1, 512, 512) is supported");
  if (cfg->num_class < 2 || cfg->num_class > 256) return fail(OCRFA_ERR_ARG, "num_class must be in [2, 256]");
  if (cfg->precision != OCRFA_PRECISION_FP32 && cfg->precision != OCRFA_PRECISION_BF16 &&
      cfg->precision != OCRFA_PRECISION_F16X3 && cfg->precision != OCRFA_PRECISION_F16)
    return fail(OCRFA_ERR_ARG, "unknown precision %d", cfg->precision);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= cfg->device)
    return fail(OCRFA_ERR_CUDA, "no CUDA device %d (this library has no CPU path)", cfg->device);
  CUDA_TRY(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return fail(OCRFA_ERR_CUDA, "sm_100a kernels need a Blackwell B200 (found cc %d.%d)", prop.major, prop.minor);
  ocrfa_handle* h = new ocrfa_handle();
  h->cfg = *cfg;
  h->num_sms = prop.multiProcessorCount;
  if (const char* hl = getenv("OCRFA_HALO")) h->halo_level = atoi(hl);
  h->stem_1px = getenv("OCRFA_STEM_1PX") != nullptr;
  if (const char* e = getenv("OCRFA_X3_HALO_MAXC")) h->x3_halo_max_cout = atoi(e);
  if (const char* e = getenv("OCRFA_X3_CHAIN")) h->x3_kchain = atoi(e) == 2 ? 2 : 1;
  const char* g = getenv("OCRFA_GRAPHS");
  if (g && g[0] == '0') h->graphs_enabled = 0;
  *out = h;
  return 0;
}

int ocrfa_craft_create(int32_t precision, int32_t device, ocrfa_handle** out) {
  if (!out) return fail(OCRFA_ERR_ARG, "null argument");
  if (precision != OCRFA_PRECISION_BF16 && precision != OCRFA_PRECISION_F16 && precision != OCRFA_PRECISION_F16X3)
    return fail(OCRFA_ERR_ARG, "the CRAFT forward runs in the tensor-core modes only (bf16 / f16 / f16x3)");
  ocrfa_config cfg;
  cfg.input_channel = 1; cfg.output_channel = 512; cfg.hidden_size = 512; cfg.num_class = 2;
  cfg.precision = precision; cfg.device = device;
  const int rc = ocrfa_create(&cfg, out);
  if (rc) return rc;
  (*out)->network = 1;
  return 0;
}

size_t ocrfa_craft_workspace_bytes(const ocrfa_handle* h, int32_t N, int32_t H, int32_t W) {
  if (!h || N <= 0 || H <= 0 || W <= 0) return 0;
  return craft_layout(N, H, W, h->split() ? 4 : 2).total;
}

int ocrfa_craft_forward(ocrfa_handle* h, const float* d_x, int32_t N, int32_t H, int32_t W, float* d_y, float* d_feature,
                        void* d_ws, size_t ws_bytes, void* stream) {
  return craft_forward_impl(h, d_x, N, H, W, d_y, d_feature, d_ws, ws_bytes, (cudaStream_t)stream);
}

void ocrfa_destroy(ocrfa_handle* h) {
  if (!h) return;
  for (GraphEntry& e : h->graphs) cudaGraphExecDestroy(e.exec);
  for (cudaEvent_t ev : h->prof.pool) cudaEventDestroy(ev);
  for (void* p : h->allocs) cudaFree(p);
  delete h;
}

int ocrfa_load_weight(ocrfa_handle* h, const char* name, const float* h_data, const int64_t* shape, int32_t ndim) {
  if (!h || !name || !h_data || (ndim > 0 && !shape) || ndim < 0 || ndim > 4) return fail(OCRFA_ERR_ARG, "bad argument");
  if (h->finalized) return fail(OCRFA_ERR_STATE, "weights already finalized");
  std::string key(name);
  if (key.rfind("module.", 0) == 0) key = key.substr(7);  // DataParallel prefix, utils.py:680-685
  const std::string fe = "FeatureExtraction.ConvNet.";
  if (key.rfind(fe, 0) == 0) key = key.substr(fe.size());
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); n *= (size_t)shape[i]; }
  t.data.assign(h_data, h_data + n);
  h->host[key] = std::move(t);
  return 0;
}

int ocrfa_finalize_weights(ocrfa_handle* h, void* stream) {
  (void)stream;
  if (!h) return fail(OCRFA_ERR_ARG, "null handle");
  if (h->finalized) return fail(OCRFA_ERR_STATE, "weights already finalized");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  int rc;
  if (h->network == 1) {
    if ((rc = prep_craft(h))) return rc;
    h->host.clear();
    h->finalized = true;
    return 0;
  }
  // stem: conv0_1 [32][1][3][3] -> [9][32] fp32 with bn0_1 folded
  {
    std::vector<double> sc, sf;
    const HostTensor* w = nullptr;
    if ((rc = fold_bn(h, "conv0_1", "bn0_1", &sc, &sf, &w))) return rc;
    if (w->data.size() != 32 * 9) return fail(OCRFA_ERR_WEIGHTS, "conv0_1 must be [32,1,3,3]");
    std::vector<float> pk(9 * 32), b(32);
    for (int o = 0; o < 32; ++o) {
      for (int k = 0; k < 9; ++k) pk[k * 32 + o] = (float)((double)w->data[o * 9 + k] * sc[o]);
      b[o] = (float)sf[o];
    }
    memcpy(h->stem_host.w, pk.data(), sizeof(h->stem_host.w));
    memcpy(h->stem_host.b, b.data(), sizeof(h->stem_host.b));
    void* d = nullptr;
    if ((rc = upload(h, pk, &d))) return rc;
    h->d_stem_w = (float*)d;
    if ((rc = upload(h, b, &d))) return rc;
    h->d_stem_b = (float*)d;
  }
  if ((rc = prep_conv(h, "", "conv0_2", "bn0_2", 3, 3, 1, 1, 1, "", ""))) return rc;
  const int blocks[4] = {1, 2, 5, 3};
  for (int li = 1; li <= 4; ++li) {
    for (int b = 0; b < blocks[li - 1]; ++b) {
      const std::string pre = "layer" + std::to_string(li) + "." + std::to_string(b) + ".";
      const bool has_ds = find(h, pre + "downsample.0.weight") != nullptr;
      if ((rc = prep_conv(h, "", pre + "conv1", pre + "bn1", 3, 3, 1, 1, 1, "", ""))) return rc;
      if ((rc = prep_conv(h, "", pre + "conv2", pre + "bn2", 3, 3, 1, 1, 1, has_ds ? pre + "downsample.0" : "",
                          has_ds ? pre + "downsample.1" : "")))
        return rc;
      if (has_ds && !h->tc())
        if ((rc = prep_conv(h, "", pre + "downsample.0", pre + "downsample.1", 1, 1, 1, 0, 0, "", ""))) return rc;
    }
    if (li < 4) {
      const std::string c = "conv" + std::to_string(li), bn = "bn" + std::to_string(li);
      if ((rc = prep_conv(h, "", c, bn, 3, 3, 1, 1, 1, "", ""))) return rc;
    }
  }
  if ((rc = prep_conv(h, "", "conv4_1", "bn4_1", 2, 2, 2, 0, 1, "", ""))) return rc;
  if ((rc = prep_conv(h, "", "conv4_2", "bn4_2", 2, 2, 1, 0, 0, "", ""))) return rc;
  // the bf16 conv0_2 reads the stem's 64-channel (32 real + 32 zero) buffer: its cin_pad is already 64
  for (int li = 0; li < 2; ++li)
    if ((rc = prep_lstm(h, li))) return rc;
  const HostTensor* pw = find(h, "Prediction.weight");
  const HostTensor* pb = find(h, "Prediction.bias");
  if (!pw || !pb || pw->data.size() != (size_t)h->cfg.num_class * 512 || pb->data.size() != (size_t)h->cfg.num_class)
    return fail(OCRFA_ERR_WEIGHTS, "Prediction weights missing or not [%d,512]", h->cfg.num_class);
  if ((rc = prep_linear(h, pw->data, pb->data, h->cfg.num_class, 512, &h->d_pred_w, &h->d_pred_b, &h->s_pred))) return rc;
  if (h->tc()) {
    CUDA_TRY(cudaFuncSetAttribute(igemm_tc_kernel<kModeStd>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(igemm_tc_kernel<kModeSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(igemm_tc_kernel<kModePair>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(lstm_tc_persistent_kernel<kLstmBf16FusedX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lstm_tc_smem_bytes()));
    CUDA_TRY(cudaFuncSetAttribute(lstm_tc_persistent_kernel<kLstmF16Split>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lstm_tc_smem_bytes()));
    CUDA_TRY(cudaFuncSetAttribute(lstm_tc_persistent_kernel<kLstmF16FusedX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lstm_tc_smem_bytes()));
    if (!get_tmap_encode()) return fail(OCRFA_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  }
  CUDA_TRY(cudaDeviceSynchronize());
  h->host.clear();
  h->finalized = true;
  return 0;
}

int32_t ocrfa_out_timesteps(int32_t img_h, int32_t img_w) {
  if (img_h < 32 || img_h % 16 != 0 || img_w < 4 || img_w % 4 != 0) return 0;
  return img_w / 4 + 1;
}

size_t ocrfa_workspace_bytes(const ocrfa_handle* h, int32_t B, int32_t H, int32_t W) {
  if (!h || B <= 0 || H <= 0 || W <= 0) return 0;
  return ws_layout(h, B, H, W).total;
}

int ocrfa_preprocess_u8(const uint8_t* d_img, const int32_t* d_widths, int32_t B, int32_t H, int32_t W, int32_t in_stride,
                        float* d_x, void* stream) {
  if (!d_img || !d_x || B <= 0 || H <= 0 || W <= 0 || in_stride < W) return fail(OCRFA_ERR_ARG, "bad argument");
  const long long total = (long long)B * H * ((W + 15) / 16);
  preprocess_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_img, d_widths, d_x, B, H, W, in_stride);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int ocrfa_align_collate_u8(const uint8_t* d_src, const ocrfa_line_desc* d_desc, const int32_t* d_tables, int32_t B,
                           int32_t img_h, int32_t img_w, double contrast_target, uint8_t* d_tmp, uint8_t* d_lut,
                           uint8_t* d_out, void* stream) {
  static_assert(sizeof(ocrfa_line_desc) == sizeof(LineDesc), "ocrfa_line_desc layout");
  if (!d_src || !d_desc || !d_tables || !d_tmp || !d_lut || !d_out || B <= 0 || img_h <= 0 || img_w <= 0)
    return fail(OCRFA_ERR_ARG, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const LineDesc* desc = reinterpret_cast<const LineDesc*>(d_desc);
  contrast_lut_kernel<<<B, 256, 0, st>>>(d_src, desc, contrast_target, d_lut);
  CUDA_TRY(cudaGetLastError());
  resize_h_kernel<<<dim3(64, B), 256, 0, st>>>(d_src, desc, d_tables, d_lut, d_tmp);
  CUDA_TRY(cudaGetLastError());
  resize_v_kernel<<<dim3(64, B), 256, 0, st>>>(d_tmp, desc, d_tables, d_out, img_h, img_w);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int ocrfa_crop_resize_u8(const uint8_t* d_src, const ocrfa_crop_desc* d_desc, const int32_t* d_tables, int32_t n,
                         uint8_t* d_out, void* stream) {
  if (!d_src || !d_desc || !d_tables || !d_out || n <= 0) return fail(OCRFA_ERR_ARG, "bad argument");
  if (n > 65535) return fail(OCRFA_ERR_ARG, "at most 65535 boxes per call");
  dim3 grid(64, (unsigned)n);
  crop_resize_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_src, d_desc, d_tables, n, d_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int ocrfa_warp_perspective_u8(const uint8_t* d_page, int32_t page_h, int32_t page_w, int32_t page_stride,
                              const ocrfa_warp_desc* d_desc, int32_t n, uint8_t* d_out, void* stream) {
  if (!d_page || !d_desc || !d_out || n <= 0 || page_h <= 0 || page_w <= 0 || page_stride < page_w)
    return fail(OCRFA_ERR_ARG, "bad argument");
  if (n > 65535) return fail(OCRFA_ERR_ARG, "at most 65535 boxes per call");
  dim3 grid(32, (unsigned)n);
  warp_perspective_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_page, page_h, page_w, page_stride, d_desc, n, d_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int ocrfa_det_preprocess_u8(const uint8_t* d_img, int32_t src_h, int32_t src_w, int32_t src_stride, int32_t dst_h,
                            int32_t dst_w, int32_t H32, int32_t W32, const int32_t* d_xtab, const int32_t* d_ytab,
                            int32_t mode, const float* mean255, const float* var255, float* d_out, void* stream) {
  if (!d_img || !d_out || !mean255 || !var255 || src_h <= 0 || src_w <= 0 || src_stride < 3 * src_w || dst_h <= 0 ||
      dst_w <= 0 || H32 < dst_h || W32 < dst_w || mode < 0 || mode > 2 || (mode == 0 && (!d_xtab || !d_ytab)))
    return fail(OCRFA_ERR_ARG, "bad argument");
  if ((mode == 1 && (dst_h != src_h || dst_w != src_w)) || (mode == 2 && (2 * dst_h != src_h || 2 * dst_w != src_w)))
    return fail(OCRFA_ERR_ARG, "resize mode does not match the sizes");
  const long long total = (long long)H32 * W32;
  const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 148 * 16);
  det_preprocess_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_img, src_h, src_w, src_stride, dst_h, dst_w, H32, W32,
                                                                 d_xtab, d_ytab, mode, mean255[0], mean255[1], mean255[2],
                                                                 var255[0], var255[1], var255[2], d_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int ocrfa_bgr2gray_u8(const uint8_t* d_img, int32_t img_stride, int32_t H, int32_t W, uint8_t* d_gray,
                      int32_t gray_stride, void* stream) {
  if (!d_img || !d_gray || H <= 0 || W <= 0 || img_stride < 3 * W || gray_stride < W) return fail(OCRFA_ERR_ARG, "bad argument");
  const long long total = (long long)((W + 3) / 4) * H;
  const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 148 * 16);
  bgr2gray_u8_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_img, img_stride, H, W, d_gray, gray_stride);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

namespace {
struct DetLayout {
  size_t parent, rank, flags, block_roots, stats, slot, src_ext, total;
  int n, nb, cap;
};
DetLayout det_layout(int H, int W, int max_rows) {
  DetLayout L;
  L.n = H * W;
  L.nb = (L.n + 1023) / 1024;
  L.cap = L.n / 2 + 2;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  L.parent = take((size_t)L.n * 4);
  L.rank = take((size_t)L.n * 4);
  L.flags = take((size_t)L.n);
  L.block_roots = take((size_t)L.nb * 4);
  L.stats = take((size_t)L.cap * 4 * 6);
  L.slot = take((size_t)L.cap * 4);
  L.src_ext = take((size_t)std::max(max_rows, 1) * 8);
  L.total = off;
  return L;
}
}  // namespace

size_t ocrfa_det_workspace_bytes(int32_t H, int32_t W, int32_t max_rows) {
  if (H <= 0 || W <= 0 || max_rows < 0 || (long long)H * W > (1 << 28)) return 0;
  return det_layout(H, W, max_rows).total;
}

int ocrfa_det_components(const float* d_text, const float* d_link, int32_t pix_stride, int32_t H, int32_t W,
                         float low_text, float link_threshold, float text_threshold, int32_t* d_labels,
                         ocrfa_det_comp* d_comps, int32_t max_comps, int32_t* d_extents, int32_t max_rows,
                         int32_t* d_counts, void* d_ws, size_t ws_bytes, void* stream) {
  if (!d_text || !d_link || !d_comps || !d_extents || !d_counts || !d_ws || pix_stride <= 0 || H <= 0 || W <= 0 ||
      max_comps <= 0 || max_rows <= 0 || (long long)H * W > (1 << 28))
    return fail(OCRFA_ERR_ARG, "bad argument");
  const DetLayout L = det_layout(H, W, max_rows);
  if (ws_bytes < L.total) return fail(OCRFA_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, L.total);
  if (((uintptr_t)d_ws & 255) != 0) return fail(OCRFA_ERR_WORKSPACE, "workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)d_ws;
  int* parent = (int*)(ws + L.parent);
  int* rank = (int*)(ws + L.rank);
  uint8_t* flags = (uint8_t*)(ws + L.flags);
  int* block_roots = (int*)(ws + L.block_roots);
  DetStats s;
  int* sb = (int*)(ws + L.stats);
  s.area = sb; s.minx = sb + L.cap; s.maxx = sb + 2 * (size_t)L.cap; s.miny = sb + 3 * (size_t)L.cap;
  s.maxy = sb + 4 * (size_t)L.cap; s.maxtext = (uint32_t*)(sb + 5 * (size_t)L.cap);
  int* slot = (int*)(ws + L.slot);
  int2* src_ext = (int2*)(ws + L.src_ext);
  const unsigned g256 = (unsigned)((L.n + 255) / 256);
  CUDA_TRY(cudaMemsetAsync(d_counts, 0, 4 * sizeof(int32_t), st));
  det_init_kernel<<<g256, 256, 0, st>>>(d_text, d_link, pix_stride, L.n, low_text, link_threshold, parent, flags);
  det_stats_init_kernel<<<(unsigned)((L.cap + 255) / 256), 256, 0, st>>>(s, L.cap);
  det_merge_kernel<<<g256, 256, 0, st>>>(parent, flags, H, W);
  det_flatten_kernel<<<(unsigned)L.nb, 1024, 0, st>>>(parent, L.n, block_roots);
  det_scan_blocks_kernel<<<1, 1024, 0, st>>>(block_roots, L.nb, d_counts);
  det_rank_kernel<<<(unsigned)L.nb, 1024, 0, st>>>(parent, L.n, block_roots, rank);
  det_label_stats_kernel<<<g256, 256, 0, st>>>(parent, rank, d_text, pix_stride, H, W, d_labels, s);
  det_select_kernel<<<1, 1024, 0, st>>>(s, H, W, text_threshold, d_comps, max_comps, max_rows, slot, d_counts);
  det_rows_init_kernel<<<148, 256, 0, st>>>(src_ext, d_counts);
  det_row_extents_kernel<<<g256, 256, 0, st>>>(parent, rank, flags, H, W, slot, d_comps, src_ext);
  det_dilate_rows_kernel<<<296, 128, 0, st>>>(d_comps, d_counts, src_ext, (int2*)d_extents);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int ocrfa_rot90_u8(const uint8_t* d_src, int32_t h, int32_t w, int32_t k, uint8_t* d_dst, void* stream) {
  if (!d_src || !d_dst || h <= 0 || w <= 0 || k < 1 || k > 3) return fail(OCRFA_ERR_ARG, "bad argument");
  dim3 grid((unsigned)((w + 31) / 32), (unsigned)((h + 31) / 32));
  if (grid.y > 65535) return fail(OCRFA_ERR_ARG, "image too tall");
  rot90_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_src, h, w, k, d_dst);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int ocrfa_forward(ocrfa_handle* h, const float* d_x, int32_t B, int32_t H, int32_t W, float* d_logits, void* d_ws,
                  size_t ws_bytes, void* stream) {
  return forward_impl(h, d_x, nullptr, nullptr, B, H, W, d_logits, d_ws, ws_bytes, (cudaStream_t)stream);
}

int ocrfa_forward_features(ocrfa_handle* h, const float* d_x, int32_t B, int32_t H, int32_t W, float* d_features,
                           void* d_ws, size_t ws_bytes, void* stream) {
  if (!d_x || !d_features) return fail(OCRFA_ERR_ARG, "null pointer");
  return forward_impl(h, d_x, nullptr, nullptr, B, H, W, nullptr, d_ws, ws_bytes, (cudaStream_t)stream, d_features, nullptr);
}

int ocrfa_forward_sequence(ocrfa_handle* h, const float* d_features, int32_t B, int32_t H, int32_t W, float* d_logits,
                           void* d_ws, size_t ws_bytes, void* stream) {
  if (!d_features || !d_logits) return fail(OCRFA_ERR_ARG, "null pointer");
  return forward_impl(h, nullptr, nullptr, nullptr, B, H, W, d_logits, d_ws, ws_bytes, (cudaStream_t)stream, nullptr, d_features);
}

int ocrfa_forward_u8(ocrfa_handle* h, const uint8_t* d_img, const int32_t* d_widths, int32_t B, int32_t H, int32_t W,
                     float* d_logits, void* d_ws, size_t ws_bytes, void* stream) {
  return forward_impl(h, nullptr, d_img, d_widths, B, H, W, d_logits, d_ws, ws_bytes, (cudaStream_t)stream);
}

int ocrfa_ctc_greedy(const float* d_logits, int32_t B, int32_t T, int32_t C, const uint8_t* d_ignore_mask,
                     const uint8_t* d_collapse_ignore, float* d_probs, int32_t* d_idx, float* d_maxp, int32_t* d_labels,
                     int32_t* d_len, double* d_conf, void* stream) {
  if (!d_logits || !d_labels || !d_len || !d_conf || B <= 0 || T <= 0 || C < 2 || C > 32 * kMaxClassesPerLane)
    return fail(OCRFA_ERR_ARG, "bad argument (C must be <= %d)", 32 * kMaxClassesPerLane);
  GreedyParams p{d_logits, d_ignore_mask, d_collapse_ignore, d_probs, d_idx, d_maxp, d_labels, d_len, d_conf, B, T, C};
  const size_t smem = (size_t)T * 8;
  if (smem > 48 * 1024) return fail(OCRFA_ERR_ARG, "T too large for the greedy kernel");
  ctc_greedy_kernel<<<B, kGreedyThreads, smem, (cudaStream_t)stream>>>(p);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int ocrfa_ctc_margin(const float* d_logits, int32_t B, int32_t T, int32_t C, const uint8_t* d_ignore_mask, float* d_margin,
                     void* stream) {
  if (!d_logits || !d_margin || B <= 0 || T <= 0 || C < 2) return fail(OCRFA_ERR_ARG, "bad argument");
  ctc_margin_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(d_logits, d_ignore_mask, d_margin, T, C);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

size_t ocrfa_beam_workspace_bytes(int32_t B, int32_t T, int32_t C, int32_t beam_width) {
  if (B <= 0 || T <= 0 || C <= 0 || beam_width <= 0) return 0;
  return beam_ws_bytes(B, T, C, beam_width);
}

int ocrfa_ctc_collapse(const int32_t* d_idx, const int32_t* d_offsets, int32_t B, int32_t C,
                       const uint8_t* d_collapse_ignore, int32_t* d_labels, int32_t* d_len, void* stream) {
  if (!d_idx || !d_offsets || !d_labels || !d_len || B <= 0 || C < 1) return fail(OCRFA_ERR_ARG, "bad argument");
  ctc_collapse_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(d_idx, d_offsets, B, C, d_collapse_ignore, d_labels, d_len);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int ocrfa_argmax_rows(const float* d_mat, int64_t rows, int32_t C, int32_t* d_idx, void* stream) {
  if (!d_mat || !d_idx || rows <= 0 || C < 1) return fail(OCRFA_ERR_ARG, "bad argument");
  argmax_rows_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, (cudaStream_t)stream>>>(d_mat, rows, C, d_idx);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

static int beam_launch(const float* d_probs, const int32_t* d_span_row0, const int32_t* d_span_len, int32_t B, int32_t T,
                       int32_t C, int32_t beam_width, const uint8_t* d_collapse_ignore, int32_t* d_labels, int32_t* d_len,
                       void* d_ws, size_t ws_bytes, void* stream);

int ocrfa_ctc_beam_spans(const float* d_probs, const int32_t* d_span_row0, const int32_t* d_span_len, int32_t n_spans,
                         int32_t t_max, int32_t C, int32_t beam_width, const uint8_t* d_collapse_ignore, int32_t* d_labels,
                         int32_t* d_len, void* d_ws, size_t ws_bytes, void* stream) {
  if (!d_span_row0 || !d_span_len) return fail(OCRFA_ERR_ARG, "null span arrays");
  return beam_launch(d_probs, d_span_row0, d_span_len, n_spans, t_max, C, beam_width, d_collapse_ignore, d_labels, d_len,
                     d_ws, ws_bytes, stream);
}

int ocrfa_ctc_beam(const float* d_probs, int32_t B, int32_t T, int32_t C, int32_t beam_width,
                   const uint8_t* d_collapse_ignore, int32_t* d_labels, int32_t* d_len, void* d_ws, size_t ws_bytes,
                   void* stream) {
  return beam_launch(d_probs, nullptr, nullptr, B, T, C, beam_width, d_collapse_ignore, d_labels, d_len, d_ws, ws_bytes,
                     stream);
}

// One launch path for both instantiations. n_cand == 0: the best labeling per item (d_labels [B][T], d_len [B]);
// n_cand > 0: the ranked candidate list (d_labels [B][n_cand][T], d_len [B][n_cand], d_count [B]).
static int beam_launch_impl(const float* d_probs, const int32_t* d_span_row0, const int32_t* d_span_len, int32_t B, int32_t T,
                            int32_t C, int32_t beam_width, int32_t n_cand, const uint8_t* d_collapse_ignore,
                            int32_t* d_labels, int32_t* d_len, int32_t* d_count, void* d_ws, size_t ws_bytes, void* stream) {
  if (!d_probs || !d_labels || !d_len || !d_ws || B <= 0 || T <= 0 || C < 2 || C > 32767)
    return fail(OCRFA_ERR_ARG, "bad argument");
  if (beam_width < 1 || beam_width > 64) return fail(OCRFA_ERR_ARG, "beam_width must be in [1, 64]");
  if (n_cand < 0 || n_cand > 64 || (n_cand > 0 && !d_count)) return fail(OCRFA_ERR_ARG, "n_cand must be in [1, 64]");
  const int width = n_cand > beam_width ? n_cand : beam_width;  // state arrays / trellis row (== beam_width when n_cand == 0)
  if (ws_bytes < beam_ws_bytes(B, T, C, width)) return fail(OCRFA_ERR_WORKSPACE, "beam workspace too small");
  BeamParams p;
  p.probs = d_probs; p.labels_out = d_labels; p.len_out = d_len; p.trellis = (uint32_t*)d_ws;
  p.collapse_ignore = d_collapse_ignore;
  p.span_row0 = d_span_row0; p.span_len = d_span_len;
  p.B = B; p.T = T; p.C = C; p.bw = width;
  p.emax = width * (C + 1);
  int hs = 64;
  while (hs < 2 * p.emax) hs <<= 1;
  p.hs = hs;
  p.ht32 = beam_smem_bytes(C, width, p.emax, hs, T, 1) <= 227 * 1024 ? 1 : 0;
  p.keep = beam_width; p.n_out = n_cand;
  p.cand_labels = d_labels; p.cand_len = d_len; p.cand_count = d_count;
  const size_t smem = beam_smem_bytes(C, width, p.emax, hs, T, p.ht32);
  if (smem > 227 * 1024) return fail(OCRFA_ERR_ARG, "beam_width * num_class too large for shared memory (%zu B)", smem);
  static size_t attr_smem_dev[2][kMaxDevices] = {};
  size_t& attr_smem = attr_smem_dev[n_cand > 0 ? 1 : 0][current_device_slot()];
  if (n_cand > 0) {
    if (smem > attr_smem) {
      CUDA_TRY(cudaFuncSetAttribute(ctc_beam_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_smem = smem;
    }
    ctc_beam_kernel<true><<<B, 32, smem, (cudaStream_t)stream>>>(p);
  } else {
    if (smem > attr_smem) {
      CUDA_TRY(cudaFuncSetAttribute(ctc_beam_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_smem = smem;
    }
    ctc_beam_kernel<false><<<B, 32, smem, (cudaStream_t)stream>>>(p);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

static int beam_launch(const float* d_probs, const int32_t* d_span_row0, const int32_t* d_span_len, int32_t B, int32_t T,
                       int32_t C, int32_t beam_width, const uint8_t* d_collapse_ignore, int32_t* d_labels, int32_t* d_len,
                       void* d_ws, size_t ws_bytes, void* stream) {
  return beam_launch_impl(d_probs, d_span_row0, d_span_len, B, T, C, beam_width, 0, d_collapse_ignore, d_labels, d_len,
                          nullptr, d_ws, ws_bytes, stream);
}

size_t ocrfa_beam_candidates_workspace_bytes(int32_t n_items, int32_t t_max, int32_t C, int32_t beam_width, int32_t n_cand) {
  if (n_items <= 0 || t_max <= 0 || C <= 0 || beam_width <= 0 || n_cand <= 0) return 0;
  return beam_ws_bytes(n_items, t_max, C, n_cand > beam_width ? n_cand : beam_width);
}

int ocrfa_ctc_beam_candidates(const float* d_probs, const int32_t* d_span_row0, const int32_t* d_span_len, int32_t n_items,
                              int32_t t_max, int32_t C, int32_t beam_width, int32_t n_cand,
                              const uint8_t* d_collapse_ignore, int32_t* d_cand_labels, int32_t* d_cand_len,
                              int32_t* d_cand_count, void* d_ws, size_t ws_bytes, void* stream) {
  if ((d_span_row0 == nullptr) != (d_span_len == nullptr)) return fail(OCRFA_ERR_ARG, "span arrays come in pairs");
  if (n_cand < 1) return fail(OCRFA_ERR_ARG, "n_cand must be in [1, 64]");
  return beam_launch_impl(d_probs, d_span_row0, d_span_len, n_items, t_max, C, beam_width, n_cand, d_collapse_ignore,
                          d_cand_labels, d_cand_len, d_cand_count, d_ws, ws_bytes, stream);
}

}  // extern "C"

namespace {
constexpr int kMaxGraphFailures = 3;  // a transient capture / instantiate failure does not switch graphs off for good
int recognize_device_body(ocrfa_handle* h, const uint8_t* d_img, const int32_t* d_widths, int B, int H, int W,
                          const uint8_t* d_ignore_mask, int decoder, int beam_width, int32_t* d_labels, int32_t* d_len,
                          double* d_conf, float* d_margin, void* d_ws, size_t ws_bytes, cudaStream_t st) {
  if (!h || !d_img || !d_labels || !d_len || !d_conf || !d_ws) return fail(OCRFA_ERR_ARG, "null pointer");
  if (decoder != 0 && decoder != 1) return fail(OCRFA_ERR_ARG, "decoder must be 0 (greedy) or 1 (beamsearch)");
  const int T = ocrfa_out_timesteps(H, W);
  if (T == 0) return fail(OCRFA_ERR_ARG, "unsupported geometry");
  const int C = h->cfg.num_class;
  const RecLayout R = rec_layout(h, B, H, W, beam_width);
  if (ws_bytes < R.total) return fail(OCRFA_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, R.total);
  uint8_t* ws = (uint8_t*)d_ws;
  float* d_logits = (float*)(ws + R.logits);
  float* d_probs = (float*)(ws + R.probs);
  int rc = forward_impl(h, nullptr, d_img, d_widths, B, H, W, d_logits, ws + R.fwd, R.img - R.fwd, st);
  if (rc) return rc;
  {
    LaunchScope ls(h, kClsDecode, 0.0, st);
    rc = ocrfa_ctc_greedy(d_logits, B, T, C, d_ignore_mask, nullptr, decoder == 1 ? d_probs : nullptr, nullptr, nullptr,
                          d_labels, d_len, d_conf, st);
    if (rc) return rc;
  }
  if (d_margin) {
    LaunchScope ls(h, kClsDecode, 0.0, st);
    rc = ocrfa_ctc_margin(d_logits, B, T, C, d_ignore_mask, d_margin, st);
    if (rc) return rc;
  }
  if (decoder == 1) {
    LaunchScope ls(h, kClsBeam, 0.0, st);
    rc = ocrfa_ctc_beam(d_probs, B, T, C, beam_width, nullptr, d_labels, d_len, ws + R.beam, R.total - R.beam, st);
    if (rc) return rc;
  }
  return 0;
}

// The step is ~40 dependent launches; replaying it as a CUDA graph removes the per-launch CPU work (tensor-map
// encodes, parameter marshalling) and most inter-kernel launch gaps. Captured lazily per (shape, pointers) key.
int recognize_device_impl(ocrfa_handle* h, const uint8_t* d_img, const int32_t* d_widths, int B, int H, int W,
                          const uint8_t* d_ignore_mask, int decoder, int beam_width, int32_t* d_labels, int32_t* d_len,
                          double* d_conf, float* d_margin, void* d_ws, size_t ws_bytes, cudaStream_t st) {
  if (!h) return fail(OCRFA_ERR_ARG, "null handle");
  // Stream capture is illegal on the legacy default stream (and on cudaStreamPerThread's implicit siblings): such callers
  // run eagerly, WITHOUT switching graphs off for later calls on a capturable stream. Callers that want graph replay
  // pass a stream created with cudaStreamCreate / torch.cuda.Stream ("Streams").
  const bool capturable = st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread;
  if (!h->graphs_enabled || !capturable || h->prof.mask != 0 || !h->tc())
    return recognize_device_body(h, d_img, d_widths, B, H, W, d_ignore_mask, decoder, beam_width, d_labels, d_len, d_conf,
                                 d_margin, d_ws, ws_bytes, st);
  const GraphKey key{d_img, d_widths, d_ignore_mask, d_labels, d_len, d_conf, d_margin, d_ws, ws_bytes, B, H, W, decoder, beam_width};
  for (GraphEntry& e : h->graphs)
    if (e.key == key) {
      e.last_use = ++h->graph_tick;
      CUDA_TRY(cudaGraphLaunch(e.exec, st));
      h->launches += e.launches;
      for (int i = 0; i < kNumClasses; ++i) h->prof.launches[i] += e.by_class[i];
      return 0;
    }
  // first sight of this key: run it eagerly once so that every lazy one-time setup (function attributes, driver
  // entry points) happens outside the capture; the captured replay below then recomputes the same outputs
  {
    int rc0 = recognize_device_body(h, d_img, d_widths, B, H, W, d_ignore_mask, decoder, beam_width, d_labels, d_len, d_conf,
                                    d_margin, d_ws, ws_bytes, st);
    if (rc0) return rc0;
  }
  const int64_t l0 = h->launches;
  int64_t c0[kNumClasses];
  for (int i = 0; i < kNumClasses; ++i) c0[i] = h->prof.launches[i];
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    if (++h->graph_failures >= kMaxGraphFailures) h->graphs_enabled = 0;
    return 0;  // the eager run above already produced this call's outputs
  }
  const int rc = recognize_device_body(h, d_img, d_widths, B, H, W, d_ignore_mask, decoder, beam_width, d_labels, d_len,
                                       d_conf, d_margin, d_ws, ws_bytes, st);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(st, &graph);
  cudaGraphExec_t exec = nullptr;
  if (rc != 0 || ce != cudaSuccess || graph == nullptr || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    if (++h->graph_failures >= kMaxGraphFailures) h->graphs_enabled = 0;  // then eager for the rest of the handle's life
    h->launches = l0;
    for (int i = 0; i < kNumClasses; ++i) h->prof.launches[i] = c0[i];
    if (rc != 0) return rc;
    return recognize_device_body(h, d_img, d_widths, B, H, W, d_ignore_mask, decoder, beam_width, d_labels, d_len, d_conf,
                                 d_margin, d_ws, ws_bytes, st);
  }
  cudaGraphDestroy(graph);
  GraphEntry e;
  e.key = key; e.exec = exec; e.launches = h->launches - l0; e.last_use = ++h->graph_tick;
  for (int i = 0; i < kNumClasses; ++i) e.by_class[i] = h->prof.launches[i] - c0[i];
  if (h->graphs.size() >= 16) {  // evict the least recently used
    size_t victim = 0;
    for (size_t i = 1; i < h->graphs.size(); ++i)
      if (h->graphs[i].last_use < h->graphs[victim].last_use) victim = i;
    cudaGraphExecDestroy(h->graphs[victim].exec);
    h->graphs[victim] = e;
  } else {
    h->graphs.push_back(e);
  }
  CUDA_TRY(cudaGraphLaunch(exec, st));
  return 0;
}
}  // namespace

extern "C" {

int ocrfa_recognize_device(ocrfa_handle* h, const uint8_t* d_img, const int32_t* d_widths, int32_t B, int32_t H, int32_t W,
                           const uint8_t* d_ignore_mask, int32_t decoder, int32_t beam_width, int32_t* d_labels,
                           int32_t* d_len, double* d_conf, void* d_ws, size_t ws_bytes, void* stream) {
  return recognize_device_impl(h, d_img, d_widths, B, H, W, d_ignore_mask, decoder, beam_width, d_labels, d_len, d_conf,
                               nullptr, d_ws, ws_bytes, (cudaStream_t)stream);
}

int ocrfa_recognize_device_ex(ocrfa_handle* h, const uint8_t* d_img, const int32_t* d_widths, int32_t B, int32_t H, int32_t W,
                              const uint8_t* d_ignore_mask, int32_t decoder, int32_t beam_width, int32_t* d_labels,
                              int32_t* d_len, double* d_conf, float* d_margin, void* d_ws, size_t ws_bytes, void* stream) {
  return recognize_device_impl(h, d_img, d_widths, B, H, W, d_ignore_mask, decoder, beam_width, d_labels, d_len, d_conf,
                               d_margin, d_ws, ws_bytes, (cudaStream_t)stream);
}

// ---- CUDA-event profile of selected kernel classes (bench.py) -----------------------------------------------
int ocrfa_profile_begin(ocrfa_handle* h, uint32_t class_mask) {
  if (!h) return fail(OCRFA_ERR_ARG, "null handle");
  h->prof.mask = class_mask;
  h->prof.spans.clear();
  h->prof.used = 0;
  for (int i = 0; i < kNumClasses; ++i) h->prof.launches[i] = 0;
  return 0;
}

int ocrfa_profile_end(ocrfa_handle* h, double* ms_by_class, double* flops_by_class, int64_t* launches_by_class) {
  if (!h || !ms_by_class || !flops_by_class || !launches_by_class) return fail(OCRFA_ERR_ARG, "null argument");
  CUDA_TRY(cudaDeviceSynchronize());
  for (int i = 0; i < kNumClasses; ++i) { ms_by_class[i] = 0.0; flops_by_class[i] = 0.0; launches_by_class[i] = h->prof.launches[i]; }
  for (const Profiler::Span& sp : h->prof.spans) {
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->prof.pool[sp.e0], h->prof.pool[sp.e1]));
    ms_by_class[sp.cls] += ms;
    flops_by_class[sp.cls] += sp.flops;
  }
  h->prof.mask = 0;
  h->prof.spans.clear();
  h->prof.used = 0;
  return 0;
}

size_t ocrfa_recognize_workspace_bytes(const ocrfa_handle* h, int32_t B, int32_t H, int32_t W, int32_t beam_width) {
  if (!h || B <= 0 || H <= 0 || W <= 0) return 0;
  return rec_layout(h, B, H, W, beam_width).total;
}

int ocrfa_recognize_host(ocrfa_handle* h, const uint8_t* h_img, const int32_t* h_widths, int32_t B, int32_t H, int32_t W,
                         const uint8_t* h_ignore_mask, int32_t decoder, int32_t beam_width, int32_t* h_labels,
                         int32_t* h_len, double* h_conf, void* d_ws, size_t ws_bytes, void* stream) {
  return ocrfa_recognize_host_ex(h, h_img, h_widths, B, H, W, h_ignore_mask, decoder, beam_width, h_labels, h_len, h_conf,
                                 nullptr, 0u, d_ws, ws_bytes, stream);
}

int ocrfa_recognize_host_ex(ocrfa_handle* h, const uint8_t* h_img, const int32_t* h_widths, int32_t B, int32_t H, int32_t W,
                            const uint8_t* h_ignore_mask, int32_t decoder, int32_t beam_width, int32_t* h_labels,
                            int32_t* h_len, double* h_conf, float* h_margin, uint32_t flags, void* d_ws, size_t ws_bytes,
                            void* stream) {
  if (!h || !h_img || !h_labels || !h_len || !h_conf || !d_ws) return fail(OCRFA_ERR_ARG, "null pointer");
  if (decoder != 0 && decoder != 1) return fail(OCRFA_ERR_ARG, "decoder must be 0 (greedy) or 1 (beamsearch)");
  const int T = ocrfa_out_timesteps(H, W);
  if (T == 0) return fail(OCRFA_ERR_ARG, "unsupported geometry");
  const int C = h->cfg.num_class;
  const RecLayout R = rec_layout(h, B, H, W, beam_width);
  if (ws_bytes < R.total) return fail(OCRFA_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, R.total);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = (uint8_t*)d_ws;
  uint8_t* d_img = ws + R.img;
  int32_t* d_w = (int32_t*)(ws + R.widths);
  uint8_t* d_mask = ws + R.mask;
  int32_t* d_labels = (int32_t*)(ws + R.labels);
  int32_t* d_len = (int32_t*)(ws + R.len);
  double* d_conf = (double*)(ws + R.conf);
  float* d_margin = h_margin ? (float*)(ws + R.margin) : nullptr;
  CUDA_TRY(cudaMemcpyAsync(d_img, h_img, (size_t)B * H * W, cudaMemcpyHostToDevice, st));
  if (h_widths) CUDA_TRY(cudaMemcpyAsync(d_w, h_widths, (size_t)B * 4, cudaMemcpyHostToDevice, st));
  if (h_ignore_mask) CUDA_TRY(cudaMemcpyAsync(d_mask, h_ignore_mask, (size_t)C, cudaMemcpyHostToDevice, st));
  int rc = recognize_device_impl(h, d_img, h_widths ? d_w : nullptr, B, H, W, h_ignore_mask ? d_mask : nullptr, decoder,
                                 beam_width, d_labels, d_len, d_conf, d_margin, d_ws, ws_bytes, st);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(h_labels, d_labels, (size_t)B * T * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(h_len, d_len, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(h_conf, d_conf, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
  if (h_margin) CUDA_TRY(cudaMemcpyAsync(h_margin, d_margin, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  // OCRFA_FLAG_ASYNC: return with the copies still in flight on `stream`; the caller owns the wait (and keeps the
  // host buffers - which should be pinned - and the workspace untouched until then)
  if (!(flags & OCRFA_FLAG_ASYNC)) CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

int64_t ocrfa_launch_count(const ocrfa_handle* h) { return h ? h->launches : 0; }

}  // extern "C"
