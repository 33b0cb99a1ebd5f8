// Standalone bring-up harness for the tcgen05 implicit-GEMM (ocrfa_b200/csrc/igemm_tc.cuh).
// Compares every geometry on the recognition path against a one-thread-per-output CUDA reference and,
// with --perf, times the heavy ResNet layers at the BASELINE config-2 shapes. Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I ocrfa_b200/csrc tools/test_igemm.cu -o build/test_igemm
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "igemm_plan.h"
#include "lstm_tc.cuh"

using namespace ocrfa;

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_));  \
      exit(2);                                                                                 \
    }                                                                                          \
  } while (0)

struct RefParams {
  const __nv_bfloat16* in; int cin_buf;
  const __nv_bfloat16* in2; int cin2_buf, cin2_pad;
  const __nv_bfloat16* wgt; int ktotal;
  const float* bias; const __nv_bfloat16* residual;
  float* out;
  int n, h_in, w_in, cin_pad, cout, kh, kw, sh, ph, pw, ho, wo, relu;
};

__global__ void ref_conv_kernel(RefParams p) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)p.n * p.ho * p.wo * p.cout;
  if (idx >= total) return;
  const int co = idx % p.cout;
  const long long pix = idx / p.cout;
  const int w = pix % p.wo;
  const int h = (pix / p.wo) % p.ho;
  const int n = pix / ((long long)p.wo * p.ho);
  float acc = 0.f;
  int k = 0;
  for (int r = 0; r < p.kh; ++r)
    for (int s = 0; s < p.kw; ++s) {
      const int hi = h * p.sh + r - p.ph, wi = w + s - p.pw;
      const bool ok = hi >= 0 && hi < p.h_in && wi >= 0 && wi < p.w_in;
      const __nv_bfloat16* ip = p.in + (((long long)n * p.h_in + hi) * p.w_in + wi) * p.cin_buf;
      const __nv_bfloat16* wp = p.wgt + (long long)co * p.ktotal + k;
      if (ok)
        for (int c = 0; c < p.cin_pad; ++c) acc += __bfloat162float(ip[c]) * __bfloat162float(wp[c]);
      k += p.cin_pad;
    }
  if (p.in2) {
    const __nv_bfloat16* ip = p.in2 + pix * p.cin2_buf;
    const __nv_bfloat16* wp = p.wgt + (long long)co * p.ktotal + k;
    for (int c = 0; c < p.cin2_pad; ++c) acc += __bfloat162float(ip[c]) * __bfloat162float(wp[c]);
  }
  if (p.bias) acc += p.bias[co];
  if (p.residual) acc += __bfloat162float(p.residual[pix * p.cout + co]);
  if (p.relu) acc = fmaxf(acc, 0.f);
  p.out[idx] = acc;
}

static std::mt19937 rng(1234);
static void fill_bf16(std::vector<__nv_bfloat16>& v, float scale) {
  std::uniform_real_distribution<float> d(-1.f, 1.f);
  for (auto& x : v) x = __float2bfloat16(d(rng) * scale);
}
static void fill_f32(std::vector<float>& v, float scale) {
  std::uniform_real_distribution<float> d(-1.f, 1.f);
  for (auto& x : v) x = d(rng) * scale;
}
template <class T>
static T* to_dev(const std::vector<T>& v) {
  T* d;
  CK(cudaMalloc(&d, v.size() * sizeof(T)));
  CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

static int g_sms = 148;
static int g_fail = 0;

struct Case {
  const char* name;
  ConvGeom g;
  int relu, use_bias, use_res, out_f32, fuse_cin2;  // fuse_cin2 > 0: extra 1x1 source with that many channels
  int tw, th, tn;                                    // 0 = auto
  int perf_iters;
};

static void run_case(const Case& c) {
  const ConvGeom& g = c.g;
  const int ho = g.h_out(), wo = g.w_out();
  const int cin_pad = ceil_div(g.cin, 64) * 64;
  const int cin2_pad = c.fuse_cin2 ? ceil_div(c.fuse_cin2, 64) * 64 : 0;
  const int ktotal = g.kh * g.kw * cin_pad + cin2_pad;
  const size_t npix_in = (size_t)g.n * g.h_in * g.w_in, npix_out = (size_t)g.n * ho * wo;
  std::vector<__nv_bfloat16> h_in(npix_in * cin_pad), h_w((size_t)g.cout * ktotal), h_res, h_in2;
  std::vector<float> h_bias(g.cout);
  fill_bf16(h_in, 1.f);
  fill_bf16(h_w, 1.f / std::sqrt((float)ktotal));
  fill_f32(h_bias, 0.5f);
  // zero the channel padding of the input so padded K columns cannot matter
  if (cin_pad != g.cin)
    for (size_t px = 0; px < npix_in; ++px)
      for (int ch = g.cin; ch < cin_pad; ++ch) h_in[px * cin_pad + ch] = __float2bfloat16(0.f);
  if (c.use_res) { h_res.resize(npix_out * g.cout); fill_bf16(h_res, 1.f); }
  if (c.fuse_cin2) { h_in2.resize(npix_out * cin2_pad); fill_bf16(h_in2, 1.f); }
  __nv_bfloat16 *d_in = to_dev(h_in), *d_w = to_dev(h_w);
  __nv_bfloat16* d_res = c.use_res ? to_dev(h_res) : nullptr;
  __nv_bfloat16* d_in2 = c.fuse_cin2 ? to_dev(h_in2) : nullptr;
  float* d_bias = to_dev(h_bias);
  const size_t out_elems = npix_out * g.cout;
  void* d_out;
  CK(cudaMalloc(&d_out, out_elems * (c.out_f32 ? 4 : 2)));
  CK(cudaMemset(d_out, 0xff, out_elems * (c.out_f32 ? 4 : 2)));
  float* d_ref;
  CK(cudaMalloc(&d_ref, out_elems * 4));

  IgemmTcParams p;
  TileShape forced{c.tw, c.th, c.tn};
  int rc = plan_conv(&p, g, d_in, cin_pad, d_w, c.tw ? &forced : nullptr);
  if (rc == 0 && c.fuse_cin2) rc = plan_add_source(&p, d_in2, c.fuse_cin2, cin2_pad, d_w, ktotal, g.cout);
  else if (rc == 0 && ktotal != g.kh * g.kw * cin_pad) rc = -99;
  if (rc) { printf("[%s] plan failed rc=%d\n", c.name, rc); g_fail++; return; }
  p.bias = c.use_bias ? d_bias : nullptr;
  p.residual = d_res; p.ldr = g.cout;
  p.out = d_out; p.ldo = g.cout; p.out_f32 = c.out_f32; p.relu = c.relu;

  CK(igemm_tc_launch(p, kModeStd, g_sms, 0));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[%s] kernel failed: %s\n", c.name, cudaGetErrorString(e)); exit(3); }

  RefParams r{d_in, cin_pad, d_in2, cin2_pad, cin2_pad, d_w, ktotal, c.use_bias ? d_bias : nullptr, d_res, d_ref,
              g.n, g.h_in, g.w_in, cin_pad, g.cout, g.kh, g.kw, g.sh, g.ph, g.pw, ho, wo, c.relu};
  const long long total = (long long)out_elems;
  ref_conv_kernel<<<(unsigned)((total + 255) / 256), 256>>>(r);
  CK(cudaDeviceSynchronize());

  std::vector<float> ref(out_elems), got(out_elems);
  CK(cudaMemcpy(ref.data(), d_ref, out_elems * 4, cudaMemcpyDeviceToHost));
  if (c.out_f32) {
    CK(cudaMemcpy(got.data(), d_out, out_elems * 4, cudaMemcpyDeviceToHost));
  } else {
    std::vector<__nv_bfloat16> tmp(out_elems);
    CK(cudaMemcpy(tmp.data(), d_out, out_elems * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < out_elems; ++i) got[i] = __bfloat162float(tmp[i]);
  }
  double max_abs = 0, max_ref = 0;
  size_t bad = 0, first_bad = (size_t)-1;
  const float atol = c.out_f32 ? 2e-3f : 2e-2f, rtol = c.out_f32 ? 2e-3f : 1e-2f;
  for (size_t i = 0; i < out_elems; ++i) {
    const float d = std::fabs(got[i] - ref[i]);
    if (!(d <= atol + rtol * std::fabs(ref[i]))) { if (bad == 0) first_bad = i; bad++; }
    if (d > max_abs || std::isnan(d)) max_abs = d;
    max_ref = std::max(max_ref, (double)std::fabs(ref[i]));
  }
  printf("[%s] n=%d in=%dx%d cin=%d cout=%d k=%dx%d sh=%d tile=(%d,%d,%d) bn=%d stages=%d tiles=%d : max_abs_err=%.4g (max|ref|=%.3g) bad=%zu/%zu %s\n",
         c.name, g.n, g.h_in, g.w_in, g.cin, g.cout, g.kh, g.kw, g.sh, p.tw, p.th, p.tn, p.block_n, p.num_stages,
         p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles, max_abs, max_ref, bad, out_elems, bad ? "FAIL" : "ok");
  if (bad) {
    g_fail++;
    const size_t i = first_bad;
    const int co = i % g.cout; const size_t pix = i / g.cout;
    printf("    first bad: pix=%zu (n=%zu h=%zu w=%zu) co=%d got=%f ref=%f\n", pix, pix / ((size_t)ho * wo),
           (pix / wo) % ho, pix % wo, co, got[i], ref[i]);
    // error pattern summary by (row in tile) and column chunk
    size_t shown = 0;
    for (size_t j = first_bad; j < out_elems && shown < 8; ++j) {
      const float d = std::fabs(got[j] - ref[j]);
      if (!(d <= atol + rtol * std::fabs(ref[j]))) { printf("    bad[%zu] pix=%zu co=%zu got=%f ref=%f\n", shown, j / g.cout, j % g.cout, got[j], ref[j]); shown++; }
    }
  }
  if (c.perf_iters > 0 && !bad) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) CK(igemm_tc_launch(p, kModeStd, g_sms, 0));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < c.perf_iters; ++i) CK(igemm_tc_launch(p, kModeStd, g_sms, 0));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= c.perf_iters;
    const double flops = 2.0 * npix_out * g.cout * (double)(g.kh * g.kw * g.cin + c.fuse_cin2);
    printf("    perf: %.3f ms/launch  %.1f TFLOP/s (algorithmic)\n", ms, flops / ms * 1e-9);
  }
  cudaFree(d_in); cudaFree(d_w); cudaFree(d_bias); cudaFree(d_out); cudaFree(d_ref);
  if (d_res) cudaFree(d_res);
  if (d_in2) cudaFree(d_in2);
}

// ---------------------------------------------------------------------------------------------
// LSTM step mode against a host reference
// ---------------------------------------------------------------------------------------------
static void run_lstm_case(int B, int T, int steps, bool persistent = false) {
  if (persistent) steps = T;
  const int H = 512;
  std::vector<__nv_bfloat16> whh((size_t)2 * 2048 * H);  // permuted layout: row = dir*2048 + tile*64 + gate*16 + u
  std::vector<float> xproj((size_t)B * T * 4096);
  fill_bf16(whh, 1.f / std::sqrt((float)H));
  fill_f32(xproj, 1.0f);
  __nv_bfloat16* d_whh = to_dev(whh);
  float* d_xproj = to_dev(xproj);
  __nv_bfloat16* d_hseq; float* d_c;
  CK(cudaMalloc(&d_hseq, (size_t)B * T * 1024 * 2)); CK(cudaMemset(d_hseq, 0, (size_t)B * T * 1024 * 2));
  CK(cudaMalloc(&d_c, (size_t)2 * B * 512 * 4)); CK(cudaMemset(d_c, 0, (size_t)2 * B * 512 * 4));

  IgemmTcParams p;
  memset(&p, 0, sizeof(p));
  p.tw = 1; p.th = 128; p.tn = 1;
  p.tiles_w = 1; p.tiles_h = ceil_div(B, 128); p.tiles_n = 1;
  p.n_img = 1; p.h_out = B; p.w_out = 1;
  p.block_n = 64; p.n_tiles = 64; p.n_valid = 4096;
  p.kchunks = H / 64; p.num_taps = 1; p.taps[0] = TapDesc{0, 0, 0, 0};
  igemm_tc_smem_bytes(p.block_n, &p.num_stages);
  int rc = make_act_tmap(&p.tmA, d_hseq, 1024, T, 1, B, 1, 1, 128, 1);
  p.tmA2 = p.tmA;
  if (rc == 0) rc = make_wgt_tmap(&p.tmB, d_whh, H, 4096, 64);
  if (rc) { printf("[lstm] plan failed %d\n", rc); g_fail++; return; }
  p.lstm_T = T; p.lstm_xproj = d_xproj; p.lstm_c = d_c; p.lstm_hseq = d_hseq;

  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  unsigned int* d_sync = nullptr;
  CK(cudaMalloc(&d_sync, 64));
  LstmTcParams lp;
  memset(&lp, 0, sizeof(lp));
  lp.B = B; lp.T = T; lp.m_tiles = ceil_div(B, 128);
  __nv_bfloat16* d_hx = nullptr;
  const long long hx_rows = (long long)2 * lp.m_tiles * 2 * 8 * 128;
  CK(cudaMalloc(&d_hx, hx_rows * 128)); CK(cudaMemset(d_hx, 0xff, hx_rows * 128));  // NaN-filled: stale reads would show
  if (make_rows64_tmap(&lp.tmX, d_hx, hx_rows, 128)) { printf("[lstm] exchange map failed\n"); g_fail++; return; }
  lp.tmW = p.tmB; lp.xproj = d_xproj; lp.hseq = d_hseq; lp.hx = d_hx; lp.sync = d_sync;
  long long* d_dbg = nullptr;
  CK(cudaMalloc(&d_dbg, 64 * 16 * 8)); CK(cudaMemset(d_dbg, 0, 64 * 16 * 8));
  lp.dbg = (persistent && getenv("LSTM_DBG")) ? d_dbg : nullptr;
  if (persistent) {  // warm-up launch (also exercises relaunch with a re-zeroed counter)
    CK(cudaMemsetAsync(d_sync, 0, 64));
    CK(lstm_tc_launch(lp, 0));
    CK(cudaMemsetAsync(d_hseq, 0, (size_t)B * T * 1024 * 2));
  }
  CK(cudaEventRecord(e0));
  if (persistent) {
    CK(cudaMemsetAsync(d_sync, 0, 64));
    CK(lstm_tc_launch(lp, 0));
  } else {
    for (int s = 0; s < steps; ++s) {
      p.lstm_step = s;
      CK(igemm_tc_launch(p, kModeLstm, g_sms, 0));
    }
  }
  CK(cudaEventRecord(e1));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[lstm] kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  if (lp.dbg) {
    std::vector<long long> ts(64 * 16);
    CK(cudaMemcpy(ts.data(), d_dbg, ts.size() * 8, cudaMemcpyDeviceToHost));
    const char* nm[16] = {"prod:step_begin", "prod:poll_done", "prod:fence_done", "prod:loads_issued", "mma:first_kblock", "mma:last_kblock",
                          "mma:committed", "", "epi:tfull", "epi:tmem_loaded", "epi:math_done", "epi:stored", "epi:threadfence", "epi:bar", "epi:released", ""};
    for (int s = 4; s < 8 && s < T; ++s) {
      const long long base = ts[s * 16 + 0];
      printf("    step %d (cycles since producer step begin):", s);
      for (int k = 0; k < 15; ++k) if (nm[k][0]) printf(" %s=%lld", nm[k], ts[s * 16 + k] - base);
      printf(" | next step begin=%lld\n", ts[(s + 1) * 16 + 0] - base);
    }
  }

  // host reference (float math on bf16-rounded h, mirroring the device's rounding points)
  std::vector<float> hprev((size_t)2 * B * H, 0.f), cst((size_t)2 * B * H, 0.f);
  std::vector<float> href((size_t)B * T * 1024, 0.f);
  std::vector<float> wf(whh.size());
  for (size_t i = 0; i < whh.size(); ++i) wf[i] = __bfloat162float(whh[i]);
  const int bcheck = std::min(B, 3);  // reference only a few batch rows (rows are independent)
  for (int s = 0; s < steps; ++s)
    for (int dir = 0; dir < 2; ++dir) {
      const int t = dir == 0 ? s : T - 1 - s;
      for (int b = 0; b < bcheck; ++b) {
        std::vector<float> hn(H), cn(H);
        for (int j = 0; j < H; ++j) {
          const int tile = j / 16, u = j % 16;
          float gate[4];
          for (int gq = 0; gq < 4; ++gq) {
            const int row = dir * 2048 + tile * 64 + gq * 16 + u;
            float acc = 0.f;
            const float* wr = &wf[(size_t)row * H];
            const float* hp = &hprev[((size_t)dir * B + b) * H];
            for (int k = 0; k < H; ++k) acc += wr[k] * hp[k];
            gate[gq] = acc + xproj[((size_t)b * T + t) * 4096 + row];
          }
          const float gi = 1.f / (1.f + std::exp(-gate[0])), gf = 1.f / (1.f + std::exp(-gate[1]));
          const float gg = std::tanh(gate[2]), go = 1.f / (1.f + std::exp(-gate[3]));
          const float cc = gf * cst[((size_t)dir * B + b) * H + j] + gi * gg;
          cn[j] = cc;
          hn[j] = go * std::tanh(cc);
        }
        for (int j = 0; j < H; ++j) {
          cst[((size_t)dir * B + b) * H + j] = cn[j];
          const float hb = __bfloat162float(__float2bfloat16(hn[j]));
          hprev[((size_t)dir * B + b) * H + j] = hb;
          href[((size_t)b * T + t) * 1024 + dir * 512 + j] = hb;
        }
      }
    }
  std::vector<__nv_bfloat16> hgot((size_t)B * T * 1024);
  CK(cudaMemcpy(hgot.data(), d_hseq, hgot.size() * 2, cudaMemcpyDeviceToHost));
  double max_abs = 0; size_t bad = 0, cnt = 0;
  for (int b = 0; b < bcheck; ++b)
    for (int s = 0; s < steps; ++s)
      for (int dir = 0; dir < 2; ++dir) {
        const int t = dir == 0 ? s : T - 1 - s;
        for (int j = 0; j < H; ++j) {
          const size_t i = ((size_t)b * T + t) * 1024 + dir * 512 + j;
          const float d = std::fabs(__bfloat162float(hgot[i]) - href[i]);
          max_abs = std::max(max_abs, (double)d);
          if (!(d <= 2e-2f)) { if (bad < 5) printf("    lstm bad b=%d t=%d dir=%d j=%d got=%f ref=%f\n", b, t, dir, j, __bfloat162float(hgot[i]), href[i]); bad++; }
          cnt++;
        }
      }
  printf("[lstm%s] B=%d T=%d steps=%d : max_abs_err=%.4g bad=%zu/%zu %s ; %.2f us/step (both dirs)\n", persistent ? "-persistent" : "", B, T,
         steps, max_abs, bad, cnt, bad ? "FAIL" : "ok", ms * 1000.f / steps);
  if (bad) g_fail++;
  cudaFree(d_whh); cudaFree(d_xproj); cudaFree(d_hseq); cudaFree(d_c);
}

int main(int argc, char** argv) {
  bool perf = false;
  const char* only = nullptr;  // --only NAME: run one perf case (or "lstm") alone, e.g. under ncu
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--perf")) perf = true;
    if (!strcmp(argv[i], "--only") && i + 1 < argc) only = argv[++i];
  }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  g_sms = prop.multiProcessorCount;
  printf("device: %s, %d SMs, cc %d.%d\n", prop.name, g_sms, prop.major, prop.minor);

  if (only) {
    std::vector<Case> pc = {
        {"perf_l1_c128", {256, 16, 256, 128, 128, 3, 3, 1, 1, 1}, 1, 1, 1, 0, 0, 0, 0, 0, 2},
        {"perf_l2_c256", {256, 8, 128, 256, 256, 3, 3, 1, 1, 1},  1, 1, 1, 0, 0, 0, 0, 0, 2},
        {"perf_l3_c512", {256, 4, 129, 512, 512, 3, 3, 1, 1, 1},  1, 1, 1, 0, 0, 0, 0, 0, 2},
        {"perf_c0_2",    {256, 32, 512, 64, 64, 3, 3, 1, 1, 1},   1, 1, 0, 0, 0, 0, 0, 0, 2},
    };
    for (const Case& c : pc) if (!strcmp(only, c.name)) run_case(c);
    if (!strcmp(only, "lstm")) run_lstm_case(256, 129, 12);
    if (!strcmp(only, "lstm_persistent")) run_lstm_case(256, 129, 129, true);
    printf("%s (%d failing cases)\n", g_fail ? "FAILED" : "ALL OK", g_fail);
    return g_fail ? 1 : 0;
  }
  std::vector<Case> cases = {
      //  name          n  h   w   cin cout kh kw sh ph pw   relu bias res f32 fuse  tile      perf
      {"gemm_small",   {1, 1, 256, 64, 64, 1, 1, 1, 0, 0},   0, 0, 0, 1, 0, 128, 1, 1, 0},
      {"gemm_k512",    {1, 1, 1024, 512, 256, 1, 1, 1, 0, 0}, 0, 1, 0, 1, 0, 0, 0, 0, 0},
      {"gemm_n185",    {1, 1, 1000, 512, 185, 1, 1, 1, 0, 0}, 0, 1, 0, 1, 0, 0, 0, 0, 0},
      {"gemm_n4096",   {1, 1, 512, 512, 4096, 1, 1, 1, 0, 0}, 0, 1, 0, 1, 0, 0, 0, 0, 0},
      {"c3x3_64",      {2, 32, 512, 64, 64, 3, 3, 1, 1, 1},   1, 1, 0, 0, 0, 0, 0, 0, 0},
      {"c3x3_64_128",  {2, 16, 256, 64, 128, 3, 3, 1, 1, 1},  1, 1, 0, 0, 0, 0, 0, 0, 0},
      {"c3x3_128_res", {2, 16, 256, 128, 128, 3, 3, 1, 1, 1}, 1, 1, 1, 0, 0, 0, 0, 0, 0},
      {"c3x3_256",     {3, 8, 128, 256, 256, 3, 3, 1, 1, 1},  1, 1, 1, 0, 0, 0, 0, 0, 0},
      {"c3x3_512_w129",{9, 4, 129, 512, 512, 3, 3, 1, 1, 1},  1, 1, 1, 0, 0, 0, 0, 0, 0},
      {"c3x3_512_t32", {9, 4, 129, 256, 512, 3, 3, 1, 1, 1},  1, 1, 0, 0, 0, 32, 4, 1, 0},
      {"c3x3_fuse_ds", {5, 4, 129, 512, 512, 3, 3, 1, 1, 1},  1, 1, 0, 0, 256, 0, 0, 0, 0},
      {"c2x2_s2",      {9, 4, 129, 512, 512, 2, 2, 2, 0, 1},  1, 1, 0, 0, 0, 0, 0, 0, 0},
      {"c2x2_s1",      {9, 2, 130, 512, 512, 2, 2, 1, 0, 0},  1, 1, 0, 0, 0, 0, 0, 0, 0},
      {"c2x2_s2_h8",   {3, 8, 513, 512, 512, 2, 2, 2, 0, 1},  1, 1, 0, 0, 0, 0, 0, 0, 0},
      {"c1x1_ds",      {3, 8, 128, 128, 256, 1, 1, 1, 0, 0},  0, 1, 0, 0, 0, 0, 0, 0, 0},
  };
  for (const Case& c : cases) run_case(c);
  run_lstm_case(200, 9, 9);
  run_lstm_case(200, 9, 9, true);
  run_lstm_case(64, 33, 33, true);
  run_lstm_case(256, 17, 17, true);

  if (perf && g_fail == 0) {
    std::vector<Case> pc = {
        {"perf_l1_c128", {256, 16, 256, 128, 128, 3, 3, 1, 1, 1}, 1, 1, 1, 0, 0, 0, 0, 0, 20},
        {"perf_l2_c256", {256, 8, 128, 256, 256, 3, 3, 1, 1, 1},  1, 1, 1, 0, 0, 0, 0, 0, 20},
        {"perf_l3_c512", {256, 4, 129, 512, 512, 3, 3, 1, 1, 1},  1, 1, 1, 0, 0, 0, 0, 0, 20},
        {"perf_l3_t32",  {256, 4, 129, 512, 512, 3, 3, 1, 1, 1},  1, 1, 1, 0, 0, 32, 4, 1, 20},
        {"perf_l3_t1",   {256, 4, 129, 512, 512, 3, 3, 1, 1, 1},  1, 1, 1, 0, 0, 1, 4, 32, 20},
        {"perf_c0_2",    {256, 32, 512, 64, 64, 3, 3, 1, 1, 1},   1, 1, 0, 0, 0, 0, 0, 0, 10},
        {"perf_xproj",   {1, 1, 33024, 512, 4096, 1, 1, 1, 0, 0}, 0, 1, 0, 1, 0, 0, 0, 0, 20},
    };
    for (const Case& c : pc) run_case(c);
    run_lstm_case(256, 129, 129);
    run_lstm_case(256, 129, 129, true);
    run_lstm_case(64, 513, 513, true);
  }
  printf("%s (%d failing cases)\n", g_fail ? "FAILED" : "ALL OK", g_fail);
  return g_fail ? 1 : 0;
}
