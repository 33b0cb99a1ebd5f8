// tcgen05 implicit-GEMM for the recognizer's convolutions and linear layers (bf16 in, fp32 accumulate).
//
// One kernel covers every dense contraction on the recognition path
//   reference: ResNet convs            /root/reference/src/ocrfa/modules/blocks.py:239-322, 200-223
//              BidirectionalLSTM GEMMs /root/reference/src/ocrfa/modules/blocks.py:89-101
//              Prediction Linear       /root/reference/src/ocrfa/modules/blocks.py:438
// by treating the activation tensor as NHWC bf16 and every filter tap as one K-slab of a GEMM:
//   D[pixel, cout] = sum_{tap} sum_{cin} A[pixel shifted by tap, cin] * Wt[cout, tap, cin]
// A slabs are fetched by TMA as 5-D boxes {64 ch, tw, 1, th, tn} straight out of the NHWC tensor;
// out-of-bounds coordinates (the conv's zero padding) are zero-filled by the TMA unit, so there is no
// im2col buffer and no padded copy in HBM. B slabs are {64, BLOCK_N} boxes of the [Cout][taps*Cin]
// weight matrix. Both land in 128-byte-swizzled shared memory and feed tcgen05.mma (M=128, N<=256,
// K=16) with the accumulator in TMEM; TMEM is double-buffered so the epilogue of tile i overlaps the
// MMAs of tile i+1. Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 =
// epilogue (TMEM -> registers -> bias/residual/ReLU -> global).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace ocrfa {

constexpr int kTcBlockM = 128;
constexpr int kTcBlockK = 64;
constexpr int kTcMaxStages = 8;
constexpr int kTcMaxTaps = 10;
constexpr int kTcThreads = 192;
constexpr int kTcTmemCols = 512;

enum IgemmMode : int { kModeStd = 0, kModeLstm = 1 };

struct TapDesc {
  int16_t dw, dr, dh, src;  // coordinate offsets in W, R (row-parity dim), H; src selects tmA (0) / tmA2 (1)
};

struct alignas(64) IgemmTcParams {
  CUtensorMap tmA;   // 5-D (C, W, R, Ho, N) bf16, box {64, tw, 1, th, tn}, SWIZZLE_128B
  CUtensorMap tmA2;  // optional second activation source (fused 1x1 downsample), same box
  CUtensorMap tmB;   // 2-D (Ktotal, Cout) bf16, box {64, block_n}, SWIZZLE_128B
  TapDesc taps[kTcMaxTaps];
  int num_taps;
  int kchunks;   // 64-channel K chunks per tap of source 0
  int kchunks2;  // ... of source 1
  int n_img, h_out, w_out;
  int tw, th, tn;
  int tiles_w, tiles_h, tiles_n;
  int n_tiles;   // tiles along Cout
  int block_n;   // UMMA N (multiple of 32, <= 256)
  int n_valid;   // real Cout
  int num_stages;
  int relu;
  int out_f32;
  int ldo;       // output row stride (elements)
  int ldr;       // residual row stride (elements)
  const float* bias;              // [n_valid] fp32 or nullptr
  const __nv_bfloat16* residual;  // NHWC bf16 or nullptr
  void* out;                      // NHWC bf16 (or fp32 when out_f32)
  // --- LSTM step mode (kModeLstm): A = h sequence [B][T][1024] bf16, B = permuted W_hh [2*2048][512]
  int lstm_T;
  int lstm_step;                  // 0..T-1 (forward walks t=step, backward walks t=T-1-step)
  const float* lstm_xproj;        // [B*T][4096] fp32, gate pre-activations from the input projection
  float* lstm_c;                  // [2][B][512] fp32 cell state (in/out)
  __nv_bfloat16* lstm_hseq;       // [B][T][1024] bf16 (fwd | bwd)
};

template <int MODE>
__global__ void __launch_bounds__(kTcThreads, 1) igemm_tc_kernel(const __grid_constant__ IgemmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128-byte swizzle atoms (8 rows x 128 B).
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);  // [kTcMaxStages]
  uint64_t* empty_bar = full_bar + kTcMaxStages;           // [kTcMaxStages]
  uint64_t* tfull_bar = empty_bar + kTcMaxStages;          // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* stage_base = smem + 1024;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.num_stages;
  const uint32_t a_bytes = kTcBlockM * kTcBlockK * 2;
  const uint32_t b_bytes = static_cast<uint32_t>(p.block_n) * kTcBlockK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_tiles = m_tiles * p.n_tiles;
  int total_kb = 0;
  for (int t = 0; t < p.num_taps; ++t) total_kb += p.taps[t].src ? p.kchunks2 : p.kchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 128);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmA2);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTcTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const int iw = m_tile % p.tiles_w;
        const int ih = (m_tile / p.tiles_w) % p.tiles_h;
        const int in = m_tile / (p.tiles_w * p.tiles_h);
        const int w0 = iw * p.tw, h0 = ih * p.th, n0 = in * p.tn;
        int kb_global = 0;
        for (int t = 0; t < p.num_taps; ++t) {
          const TapDesc tap = p.taps[t];
          const int chunks = tap.src ? p.kchunks2 : p.kchunks;
          const void* tm = tap.src ? static_cast<const void*>(&p.tmA2) : static_cast<const void*>(&p.tmA);
          for (int cc = 0; cc < chunks; ++cc, ++kb_global) {
            mbar_wait(&empty_bar[stage], phase ^ 1u, 1);
            uint8_t* sa = stage_base + static_cast<size_t>(stage) * stage_bytes;
            uint8_t* sb = sa + a_bytes;
            mbar_expect_tx(&full_bar[stage], stage_bytes);
            if (MODE == kModeLstm) {
              // A rows = batch entries b0..b0+127 of h at the previous timestep of this direction.
              const int dir = n_tile / (p.n_tiles >> 1);
              const int t_prev = dir == 0 ? p.lstm_step - 1 : p.lstm_T - p.lstm_step;  // OOB at step 0 -> zeros
              tma_load_5d(sa, tm, &full_bar[stage], dir * 512 + cc * kTcBlockK, t_prev, 0, h0, 0);
            } else {
              tma_load_5d(sa, tm, &full_bar[stage], cc * kTcBlockK, w0 + tap.dw, tap.dr, h0 + tap.dh, n0);
            }
            tma_load_2d(sb, &p.tmB, &full_bar[stage], kb_global * kTcBlockK, n_tile * p.block_n);
            if (++stage == stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kTcBlockM, static_cast<uint32_t>(p.block_n));
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], aphase ^ 1u, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as) * 256u;
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase, 3);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + static_cast<size_t>(stage) * stage_bytes);
          const uint64_t a_desc = make_sw128_kmajor_desc(sa);
          const uint64_t b_desc = make_sw128_kmajor_desc(sa + a_bytes);
#pragma unroll
          for (int k = 0; k < kTcBlockK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in 16-byte units
            umma_bf16(d_tmem, a_desc + static_cast<uint64_t>(2 * k), b_desc + static_cast<uint64_t>(2 * k), idesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int iw = m_tile % p.tiles_w;
      const int ih = (m_tile / p.tiles_w) % p.tiles_h;
      const int in = m_tile / (p.tiles_w * p.tiles_h);
      const int ww = row % p.tw;
      const int hh = (row / p.tw) % p.th;
      const int nn = row / (p.tw * p.th);
      const int w = iw * p.tw + ww, h = ih * p.th + hh, n = in * p.tn + nn;
      const bool valid = (w < p.w_out) && (h < p.h_out) && (n < p.n_img);
      const size_t pix = (static_cast<size_t>(n) * p.h_out + h) * p.w_out + w;
      const int n_off = n_tile * p.block_n;

      mbar_wait(&tfull_bar[as], aphase, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as) * 256u;

      if (MODE == kModeLstm) {
        // block_n == 64: columns are [gate i,f,g,o][16 hidden units] of one direction.
        uint32_t v0[32], v1[32];
        tmem_ld_32x32b_x32(taddr, v0);
        tmem_ld_32x32b_x32(taddr + 32u, v1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&tempty_bar[as]);
        if (valid) {
          const int b = h;
          const int tiles_per_dir = p.n_tiles >> 1;
          const int dir = n_tile / tiles_per_dir;
          const int t_cur = dir == 0 ? p.lstm_step : p.lstm_T - 1 - p.lstm_step;
          const int j0 = (n_tile % tiles_per_dir) * 16;
          const float* xp = p.lstm_xproj + (static_cast<size_t>(b) * p.lstm_T + t_cur) * 4096 + n_off;
          float* cptr = p.lstm_c + (static_cast<size_t>(dir) * p.h_out + b) * 512 + j0;
          __nv_bfloat16* hptr = p.lstm_hseq + (static_cast<size_t>(b) * p.lstm_T + t_cur) * 1024 + dir * 512 + j0;
          float hout[16];
#pragma unroll
          for (int u = 0; u < 16; u += 4) {
            const float4 xi = __ldg(reinterpret_cast<const float4*>(xp + u));
            const float4 xf = __ldg(reinterpret_cast<const float4*>(xp + 16 + u));
            const float4 xg = __ldg(reinterpret_cast<const float4*>(xp + 32 + u));
            const float4 xo = __ldg(reinterpret_cast<const float4*>(xp + 48 + u));
            float4 c = p.lstm_step == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(cptr + u);
            const float xiv[4] = {xi.x, xi.y, xi.z, xi.w}, xfv[4] = {xf.x, xf.y, xf.z, xf.w};
            const float xgv[4] = {xg.x, xg.y, xg.z, xg.w}, xov[4] = {xo.x, xo.y, xo.z, xo.w};
            float cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float gi = sigmoidf_acc(__uint_as_float(v0[u + e]) + xiv[e]);
              const float gf = sigmoidf_acc(__uint_as_float(v0[16 + u + e]) + xfv[e]);
              const float gg = tanhf(__uint_as_float(v1[u + e]) + xgv[e]);
              const float go = sigmoidf_acc(__uint_as_float(v1[16 + u + e]) + xov[e]);
              cv[e] = gf * cv[e] + gi * gg;
              hout[u + e] = go * tanhf(cv[e]);
            }
            *reinterpret_cast<float4*>(cptr + u) = make_float4(cv[0], cv[1], cv[2], cv[3]);
          }
          uint4 o0, o1;
          o0.x = pack_bf16x2(hout[0], hout[1]);   o0.y = pack_bf16x2(hout[2], hout[3]);
          o0.z = pack_bf16x2(hout[4], hout[5]);   o0.w = pack_bf16x2(hout[6], hout[7]);
          o1.x = pack_bf16x2(hout[8], hout[9]);   o1.y = pack_bf16x2(hout[10], hout[11]);
          o1.z = pack_bf16x2(hout[12], hout[13]); o1.w = pack_bf16x2(hout[14], hout[15]);
          *reinterpret_cast<uint4*>(hptr) = o0;
          *reinterpret_cast<uint4*>(hptr + 8) = o1;
        }
      } else {
        for (int c0 = 0; c0 < p.block_n; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c0), v);
          tmem_ld_wait();
          if (c0 + 32 >= p.block_n) {
            // last read of this accumulator: hand the TMEM buffer back before the global stores
            tc_fence_before();
            mbar_arrive(&tempty_bar[as]);
          }
          const int col0 = n_off + c0;
          if (valid && col0 < p.n_valid) {
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
            const bool full = (col0 + 32 <= p.n_valid);
            if (p.bias != nullptr) {
              if (full) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
                  f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (col0 + i < p.n_valid) f[i] += __ldg(p.bias + col0 + i);
              }
            }
            if (p.residual != nullptr) {
              const __nv_bfloat16* rp = p.residual + pix * p.ldr + col0;
              if (full) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                  const uint4 r4 = __ldg(reinterpret_cast<const uint4*>(rp + i));
                  f[i] += bf16lo_to_f32(r4.x);     f[i + 1] += bf16hi_to_f32(r4.x);
                  f[i + 2] += bf16lo_to_f32(r4.y); f[i + 3] += bf16hi_to_f32(r4.y);
                  f[i + 4] += bf16lo_to_f32(r4.z); f[i + 5] += bf16hi_to_f32(r4.z);
                  f[i + 6] += bf16lo_to_f32(r4.w); f[i + 7] += bf16hi_to_f32(r4.w);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (col0 + i < p.n_valid) f[i] += __bfloat162float(rp[i]);
              }
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.0f);
            }
            if (p.out_f32) {
              float* op = reinterpret_cast<float*>(p.out) + pix * p.ldo + col0;
              if (full && (p.ldo & 3) == 0) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                  *reinterpret_cast<float4*>(op + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (col0 + i < p.n_valid) op[i] = f[i];
              }
            } else {
              __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.ldo + col0;
              if (full) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                  uint4 o;
                  o.x = pack_bf16x2(f[i], f[i + 1]);     o.y = pack_bf16x2(f[i + 2], f[i + 3]);
                  o.z = pack_bf16x2(f[i + 4], f[i + 5]); o.w = pack_bf16x2(f[i + 6], f[i + 7]);
                  *reinterpret_cast<uint4*>(op + i) = o;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (col0 + i < p.n_valid) op[i] = __float2bfloat16_rn(f[i]);
              }
            }
          }
        }
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, kTcTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
inline size_t igemm_tc_smem_bytes(int block_n, int* stages_out) {
  const size_t stage = static_cast<size_t>(kTcBlockM) * kTcBlockK * 2 + static_cast<size_t>(block_n) * kTcBlockK * 2;
  const size_t budget = 227 * 1024 - 2048;  // alignment slack + barrier header
  int stages = static_cast<int>(budget / stage);
  if (stages > kTcMaxStages) stages = kTcMaxStages;
  if (stages_out) *stages_out = stages;
  return 2048 + stages * stage;
}

// cuTensorMapEncodeTiled resolved through the runtime so the library does not link libcuda directly.
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled get_tmap_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(sym);
  }
  return fn;
}

// Activation map: NHWC bf16 tensor seen as (C, W, R, Ho, N) where H = Ho*R (R=2 expresses a stride-2
// walk over H without element strides). `c_total` is the channel count of the underlying buffer.
inline int make_act_tmap(CUtensorMap* tm, const void* base, int c_total, int w, int r, int ho, int n, int tw, int th,
                         int tn) {
  PFN_tmapEncodeTiled enc = get_tmap_encode();
  if (!enc) return -1;
  cuuint64_t dims[5] = {(cuuint64_t)c_total, (cuuint64_t)w, (cuuint64_t)r, (cuuint64_t)ho, (cuuint64_t)n};
  cuuint64_t strides[4] = {(cuuint64_t)c_total * 2, (cuuint64_t)c_total * 2 * w, (cuuint64_t)c_total * 2 * w * r,
                           (cuuint64_t)c_total * 2 * w * r * ho};
  cuuint32_t box[5] = {64, (cuuint32_t)tw, 1, (cuuint32_t)th, (cuuint32_t)tn};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? 0 : -static_cast<int>(rc) - 1000;
}

// Weight map: [cout][ktotal] bf16 row-major.
inline int make_wgt_tmap(CUtensorMap* tm, const void* base, int ktotal, int cout, int block_n) {
  PFN_tmapEncodeTiled enc = get_tmap_encode();
  if (!enc) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)ktotal, (cuuint64_t)cout};
  cuuint64_t strides[1] = {(cuuint64_t)ktotal * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)block_n};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? 0 : -static_cast<int>(rc) - 1000;
}

// Generic row-major bf16 matrix [rows][64] as a 2-D map with {64, box_rows} boxes (LSTM h exchange slabs).
inline int make_rows64_tmap(CUtensorMap* tm, const void* base, long long rows, int box_rows) {
  PFN_tmapEncodeTiled enc = get_tmap_encode();
  if (!enc) return -1;
  cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? 0 : -static_cast<int>(rc) - 1000;
}

inline cudaError_t igemm_tc_launch(const IgemmTcParams& p, int mode, int num_sms, cudaStream_t stream) {
  static bool attr_set[2] = {false, false};
  int stages = 0;
  const size_t smem = igemm_tc_smem_bytes(p.block_n, &stages);
  if (!attr_set[mode]) {
    cudaError_t e = mode == kModeLstm
                        ? cudaFuncSetAttribute(igemm_tc_kernel<kModeLstm>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               227 * 1024)
                        : cudaFuncSetAttribute(igemm_tc_kernel<kModeStd>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[mode] = true;
  }
  const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
  const int grid = total_tiles < num_sms ? total_tiles : num_sms;
  if (mode == kModeLstm)
    igemm_tc_kernel<kModeLstm><<<grid, kTcThreads, smem, stream>>>(p);
  else
    igemm_tc_kernel<kModeStd><<<grid, kTcThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace ocrfa
