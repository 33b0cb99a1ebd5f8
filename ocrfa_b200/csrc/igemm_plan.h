// Host-side planning for the tcgen05 implicit-GEMM: conv geometry -> tap table, M-tile box, tensor maps.
// Geometry follows the reference layers (/root/reference/src/ocrfa/modules/blocks.py:239-322):
// 3x3 s1 p1, 1x1, 2x2 s(2,1) p(0,1) and 2x2 s1 p0; linear layers are 1x1 convs over a [rows,1,1] image.
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "igemm_tc.cuh"
#include "igemm_halo.cuh"

namespace ocrfa {

struct ConvGeom {
  int n, h_in, w_in, cin, cout;
  int kh, kw;  // filter
  int sh;      // stride in H (1 or 2); stride in W is always 1 on this path
  int ph, pw;  // zero padding
  int dil = 1; // dilation (CRAFT's fc6: 3x3, dilation 6, padding 6; stride 1 only)
  int h_out() const { return (h_in + 2 * ph - dil * (kh - 1) - 1) / sh + 1; }
  int w_out() const { return w_in + 2 * pw - dil * (kw - 1); }
};

// Channel count of an activation buffer the TMA box (64 channels) can read: a multiple of 64, or 8/16/32 channels
// (pixel rows of 16/32/64 bytes) when the layer has at most 64 input channels - the box then reaches past the
// tensor's channel extent and TMA zero-fills the rest (the weights carry zeros there).
inline bool act_channels_ok(int cin_buf, int cin_pad) {
  if (cin_buf % 64 == 0) return cin_buf >= cin_pad;
  return cin_pad == 64 && (cin_buf == 8 || cin_buf == 16 || cin_buf == 32);
}

struct TileShape {
  int tw, th, tn;
};

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
constexpr int kPlanSmCount = 148;  // B200

// 128 output pixels per M tile, arranged as a {tw, th, tn} box (powers of two) that wastes the fewest
// MMA rows on the ragged edges of (W, H, N).
inline TileShape choose_tile(int w, int h, int n) {
  TileShape best{128, 1, 1};
  double best_score = 1e30;
  for (int tw = 1; tw <= 128; tw *= 2)
    for (int th = 1; th * tw <= 128; th *= 2) {
      const int tn = 128 / (tw * th);
      const double cover = double(ceil_div(w, tw) * tw) * double(ceil_div(h, th) * th) * double(ceil_div(n, tn) * tn);
      const double waste = cover / (double(w) * h * n) - 1.0;
      // every box is 128 rows of 128 B whatever its shape (measured: a {1,4,32} box runs 4 % faster than {4,1,32}
      // at W = 129, exactly its smaller edge waste), so the score is the waste alone; ties go to the wider box
      const double score = waste;
      if (score < best_score - 1e-12 || (std::fabs(score - best_score) <= 1e-12 && tw > best.tw)) {
        best_score = score;
        best = TileShape{tw, th, tn};
      }
    }
  return best;
}

inline int choose_block_n(int cout) {
  if (cout % 256 == 0) return 256;
  if (cout <= 256) return ceil_div(cout, 32) * 32;
  if (cout % 128 == 0) return 128;
  return 64;
}

// Fills geometry-dependent fields of `p` (maps, taps, tiles). Epilogue pointers are left to the caller.
// `in` : NHWC bf16 [n][h_in][w_in][cin_buf] (cin_buf >= cin, multiple of 64: stem output is channel-padded)
// `wgt`: [cout][kh*kw*cin_pad] bf16 with K ordered (tap-major, channel-minor), cin_pad = round_up(cin, 64)
// split = true: fp16 hi/lo operands - `in` pixels are [hi(cin_buf) | lo(cin_buf)], the weight matrix carries
// [hi(cin_pad) | lo(cin_pad)] per tap (K doubles).
inline int plan_conv(IgemmTcParams* p, const ConvGeom& g, const void* in, int cin_buf, const void* wgt,
                     const TileShape* forced_tile = nullptr, bool split = false) {
  std::memset(p, 0, sizeof(*p));
  const int ho = g.h_out(), wo = g.w_out();
  const int cin_pad = ceil_div(g.cin, 64) * 64;
  if (!act_channels_ok(cin_buf, cin_pad)) return -2;
  if (g.sh != 1 && g.sh != 2) return -3;
  if (g.dil != 1 && g.sh != 1) return -7;
  if (g.sh == 2 && (g.h_in % 2 != 0 || g.ph != 0 || g.kh != 2)) return -4;
  if (g.kh * g.kw > kTcMaxTaps - 1) return -5;
  const TileShape t = forced_tile ? *forced_tile : choose_tile(wo, ho, g.n);
  p->tw = t.tw; p->th = t.th; p->tn = t.tn;
  p->tiles_w = ceil_div(wo, t.tw);
  p->tiles_h = ceil_div(ho, t.th);
  p->tiles_n = ceil_div(g.n, t.tn);
  p->n_img = g.n; p->h_out = ho; p->w_out = wo;
  p->block_n = choose_block_n(g.cout);
  if (split) p->block_n = ceil_div(p->block_n, 64) * 64;  // the split epilogue halves a tile in 32-column groups
  if (split) {
    if (const char* e = std::getenv("OCRFA_X3_BLOCKN")) {  // experiment switch: cap the N tile of the fp32-grade kernel
      const int cap = std::atoi(e);
      if ((cap == 64 || cap == 128) && p->block_n > cap && g.cout % cap == 0) p->block_n = cap;
    }
  }
  // latency shapes (a handful of M tiles, e.g. one 32x128 line): narrower N tiles spread the layer over more SMs -
  // the K loop of a single 128 x 256 tile at K = 4608 is ~20 us on one SM while the other SMs idle
  {
    const int m_items = ceil_div(p->tiles_w * p->tiles_h * p->tiles_n, 2);  // CTA pairs take two M tiles
    const int wide_n = p->block_n;
    while (p->block_n > 64 && p->block_n % 128 == 0 && g.cout % (p->block_n / 2) == 0 &&
           m_items * ceil_div(g.cout, p->block_n) * 2 <= kPlanSmCount / 2)
      p->block_n /= 2;
    // fp32-grade kernel: a layer with 256-column tiles is narrowed to 64 columns or not at all (igemm_x3.cuh pairs the
    // K chunks of a column by its 64-column group, which a 128-column tile cannot do)
    if (split && wide_n == 256 && p->block_n == 128) p->block_n = 256;
  }
  p->n_tiles = ceil_div(g.cout, p->block_n);
  p->n_valid = g.cout;
  p->kchunks = cin_pad / 64;
  p->kchunks2 = 0;
  p->split = split ? 1 : 0;
  p->a_lo = cin_buf;
  p->out_scale = 1.0f;
  // CTA-pair kernel (M = 256 across two SMs) whenever there are at least two M tiles; OCRFA_IGEMM_PAIR=0 disables
  p->pair = (!split && p->tiles_w * p->tiles_h * p->tiles_n >= 2 && (p->block_n & 15) == 0) ? 1 : 0;
  if (const char* e = std::getenv("OCRFA_IGEMM_PAIR")) { if (e[0] == '0') p->pair = 0; }
  igemm_tc_smem_bytes(p->block_n, &p->num_stages, p->pair);
  if (split) {
    // fp32-grade mode: always the fused-stage CTA-pair kernel (igemm_x3.cuh); a lone M tile leaves the peer CTA with
    // an out-of-range tile (TMA zero-fills, nothing is stored). OCRFA_X3=0 selects the older single-CTA kernel (A/B runs).
    p->pair = 1;
    if (const char* e = std::getenv("OCRFA_X3")) { if (e[0] == '0') p->pair = 0; }
    if (p->pair) igemm_x3_smem_bytes(p->block_n, &p->num_stages);
  }
  int nt = 0;
  for (int r = 0; r < g.kh; ++r)
    for (int s = 0; s < g.kw; ++s) {
      TapDesc& d = p->taps[nt++];
      d.src = 0;
      d.dw = static_cast<int16_t>(s * g.dil - g.pw);
      if (g.sh == 1) { d.dr = 0; d.dh = static_cast<int16_t>(r * g.dil - g.ph); }
      else { d.dr = static_cast<int16_t>(r); d.dh = 0; }  // h_in = 2*h_out + r
    }
  p->num_taps = nt;
  p->korder = (!split && g.kh == 3 && g.kw == 3 && g.sh == 1 && g.dil == 1 && g.ph == 1 && g.pw == 1) ? 1 : 0;
  if (const char* e = std::getenv("OCRFA_KORDER")) { if (e[0] == '0') p->korder = 0; }  // experiment switch
  const int R = g.sh;
  const int km = split ? 2 : 1;
  int rc = make_act_tmap(&p->tmA, in, km * cin_buf, g.w_in, R, g.h_in / R, g.n, t.tw, t.th, t.tn);
  if (rc) return rc;
  p->tmA2 = p->tmA;
  rc = make_wgt_tmap(&p->tmB, wgt, km * g.kh * g.kw * cin_pad, g.cout, p->pair ? p->block_n / 2 : p->block_n);
  return rc;
}

// Appends a fused 1x1 (downsample) source: extra K slabs read from `in2` [n][h_out][w_out][cin2_buf].
// The weight matrix must carry cin2_pad (split: 2 * cin2_pad) extra K columns after the main taps; `ktotal` is the
// full K of the weight matrix.
inline int plan_add_source(IgemmTcParams* p, const void* in2, int cin2, int cin2_buf, const void* wgt, int ktotal,
                           int cout) {
  if (p->num_taps >= kTcMaxTaps) return -6;
  const int cin2_pad = ceil_div(cin2, 64) * 64;
  TapDesc& d = p->taps[p->num_taps++];
  d.src = 1; d.dw = 0; d.dr = 0; d.dh = 0;
  p->kchunks2 = cin2_pad / 64;
  p->a_lo2 = cin2_buf;
  int rc = make_act_tmap(&p->tmA2, in2, (p->split ? 2 : 1) * cin2_buf, p->w_out, 1, p->h_out, p->n_img, p->tw, p->th, p->tn);
  if (rc) return rc;
  return make_wgt_tmap(&p->tmB, wgt, ktotal, cout, p->pair ? p->block_n / 2 : p->block_n);
}

// ---------------------------------------------------------------------------------------------
// halo-staged kernel (igemm_halo.cuh): 3x3 / stride 1 / pad 1, Cout a multiple of 32
// ---------------------------------------------------------------------------------------------
inline bool halo_eligible(const ConvGeom& g, bool pool) {
  if (g.kh != 3 || g.kw != 3 || g.sh != 1 || g.ph != 1 || g.pw != 1 || g.dil != 1) return false;
  if (g.cout % 32 != 0) return false;
  if (pool && ((g.h_in & 1) || (g.w_in & 1))) return false;
  return true;
}

// M tile = tw x th pixels of ONE image, tw a multiple of 8 (tap offsets stay swizzle-atom aligned); fewest wasted
// MMA rows first, then the smallest halo overhead (th + 2) / th. The fused pool needs the 2x2 window inside a warp.
inline void halo_choose_tile(int w, int h, bool pool, int* tw_out, int* th_out) {
  double best = 1e30;
  for (int tw = 8; tw <= (pool ? 16 : 128); tw *= 2) {
    const int th = 128 / tw;
    const double cover = double(ceil_div(w, tw) * tw) * double(ceil_div(h, th) * th) / (double(w) * h);
    const double score = cover * 100.0 + double(th + 2) / th;
    if (score < best) { best = score; *tw_out = tw; *th_out = th; }
  }
}

// fp32-grade (f16x3) halo-staged plan for igemm_x3_kernel: 3x3 / s1 / p1, Cout in {64, 128, 256} as ONE N tile.
// `in`: NHWC fp16 hi/lo pixels [hi(cin_buf) | lo(cin_buf)]; `wgt`: [cout][tap][hi(cin_pad) | lo(cin_pad)] (+ fused 1x1
// source appended by plan_add_source). pool: fuse the following 2x2/2 max pool.
inline int plan_x3_halo(IgemmTcParams* p, const ConvGeom& g, const void* in, int cin_buf, const void* wgt, bool pool) {
  std::memset(p, 0, sizeof(*p));
  if (!halo_eligible(g, pool)) return -30;
  if (g.cout != 64 && g.cout != 128 && g.cout != 256) return -31;
  const int cin_pad = ceil_div(g.cin, 64) * 64;
  if (cin_buf % 64 != 0 || cin_buf < cin_pad) return -32;
  halo_choose_tile(g.w_in, g.h_in, pool, &p->tw, &p->th);
  p->tn = 1;
  p->tiles_w = ceil_div(g.w_in, p->tw);
  p->tiles_h = ceil_div(g.h_in, p->th);
  p->tiles_n = g.n;
  p->n_img = g.n; p->h_out = g.h_in; p->w_out = g.w_in;
  p->block_n = g.cout;
  p->n_tiles = 1;
  p->n_valid = g.cout;
  p->kchunks = cin_pad / 64;
  p->kchunks2 = 0;
  p->split = 1; p->pair = 1; p->halo = 1; p->pool = pool ? 1 : 0;
  p->a_lo = cin_buf;
  p->out_scale = 1.0f;
  igemm_x3_halo_smem_bytes(p->block_n, p->tw, p->th, &p->num_stages);
  if (p->num_stages < 2) return -33;
  int nt = 0;
  for (int r = 0; r < 3; ++r)
    for (int s2 = 0; s2 < 3; ++s2) {
      TapDesc& d = p->taps[nt++];
      d.src = 0; d.dw = static_cast<int16_t>(s2 - 1); d.dr = 0; d.dh = static_cast<int16_t>(r - 1);
    }
  p->num_taps = nt;
  int rc = make_act_tmap(&p->tmA, in, 2 * cin_buf, g.w_in, 1, g.h_in, g.n, p->tw, p->th + 2, 1);
  if (rc) return rc;
  p->tmA2 = p->tmA;
  return make_wgt_tmap(&p->tmB, wgt, 2 * 9 * cin_pad, g.cout, p->block_n / 2);
}

// `in`: NHWC bf16 [n][h][w][cin_buf]; cin_buf is a multiple of 64 or, for the 32-channel stem output, 32 (the box
// still asks for 64 channels: TMA zero-fills the out-of-range half, the weights carry zeros there).
// `in2` (optional): fused 1x1 source [n][h][w][cin2_buf]; the weight matrix then has cin2_pad extra K columns.
inline int plan_halo(IgemmHaloParams* p, const ConvGeom& g, const void* in, int cin_buf, const void* wgt, bool pool,
                     const void* in2 = nullptr, int cin2 = 0, int cin2_buf = 0) {
  std::memset(p, 0, sizeof(*p));
  if (!halo_eligible(g, pool)) return -20;
  const int cin_pad = ceil_div(g.cin, 64) * 64;
  if (!act_channels_ok(cin_buf, cin_pad)) return -21;
  const int cin2_pad = in2 ? ceil_div(cin2, 64) * 64 : 0;
  if (in2 && (cin2_buf % 64 != 0 || cin2_buf < cin2_pad)) return -22;
  halo_choose_tile(g.w_in, g.h_in, pool, &p->tw, &p->th);
  p->tiles_w = ceil_div(g.w_in, p->tw);
  p->tiles_h = ceil_div(g.h_in, p->th);
  p->n_img = g.n; p->h_out = g.h_in; p->w_out = g.w_in;
  p->block_n = g.cout <= 256 ? g.cout : (g.cout % 256 == 0 ? 256 : (g.cout % 128 == 0 ? 128 : 64));
  if (g.cout % p->block_n != 0) return -23;
  p->n_tiles = g.cout / p->block_n;
  p->n_valid = g.cout;
  p->kchunks = cin_pad / 64;
  p->kchunks2 = cin2_pad / 64;
  p->pool = pool ? 1 : 0;
  const int total_kb = 9 * p->kchunks + p->kchunks2;
  igemm_halo_smem_bytes(p->block_n, p->tw, p->th, total_kb, p->n_tiles, &p->num_stages, &p->b_resident);
  if (p->num_stages < 2) return -24;
  if (const char* e = std::getenv("OCRFA_HALO_RESIDENT")) {
    if (e[0] == '0' && p->b_resident) {  // debugging aid: force the streamed-weights layout
      const size_t budget = 227 * 1024 - 1024 - kHaloHeader;
      const size_t st = static_cast<size_t>(p->th + 2) * p->tw * 128 + 3 * static_cast<size_t>(p->block_n / 2) * 128;
      p->b_resident = 0;
      p->num_stages = static_cast<int>(budget / st) > kTcMaxStages ? kTcMaxStages : static_cast<int>(budget / st);
    }
  }
  int rc = make_act_tmap(&p->tmA, in, cin_buf, g.w_in, 1, g.h_in, g.n, p->tw, p->th + 2, 1);
  if (rc) return rc;
  p->tmA2 = p->tmA;
  if (in2) {
    rc = make_act_tmap(&p->tmA2, in2, cin2_buf, g.w_in, 1, g.h_in, g.n, p->tw, p->th, 1);
    if (rc) return rc;
  }
  return make_wgt_tmap(&p->tmB, wgt, 9 * cin_pad + cin2_pad, g.cout, p->block_n / 2);
}

}  // namespace ocrfa
