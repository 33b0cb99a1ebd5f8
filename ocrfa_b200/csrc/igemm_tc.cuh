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
//
// Split-precision mode (IgemmTcParams::split): every operand is an fp16 hi/lo pair (x = hi + lo, 22 significant
// bits; activations NHWC [hi(C) | lo(C)] per pixel, weights [Cout][tap][hi(C) | lo(C)] pre-scaled by a power of
// two) and every K chunk is multiplied three times - hi.hi + lo.hi + hi.lo: fp32-grade results (the dropped lo.lo
// term is ~2^-22) at one third of the tensor-core rate. Because the tensor core's fp32 accumulate TRUNCATES (a bias
// that grows with the length and magnitude of the running sum - measured: error linear in K), the split kernel
// (kModeSplit, 8 epilogue warps) accumulates a tile in chunks: all lo.hi / hi.lo products first (small), then the
// hi.hi K blocks two at a time into alternating TMEM buffers that the epilogue warps sum in registers with
// round-to-nearest fp32.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace ocrfa {

constexpr int kTcBlockM = 128;
constexpr int kTcBlockK = 64;
constexpr int kTcMaxStages = 8;
constexpr int kTcMaxTaps = 10;
constexpr int kTcThreads = 192;       // bf16 kernel: producer + MMA + 4 epilogue warps
constexpr int kTcThreadsSplit = 320;  // split kernel: producer + MMA + 8 epilogue warps (register accumulators)
constexpr int kTcSplitChunkKb = 2;    // split kernel: hi.hi K blocks per TMEM accumulation chunk
constexpr int kTcTmemCols = 512;
constexpr int kTcSmemHeader = 8192;                   // barriers (256 B) + K-block schedule, then the 1024-aligned stages
constexpr int kTcMaxKb = (kTcSmemHeader - 256) / 32;  // K blocks per tile the schedule table can hold (248)

enum IgemmMode : int { kModeStd = 0, kModeSplit = 1, kModePair = 2 };

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
  int pair;      // 1: CTA-pair kernel (kModePair): M = 256 across two SMs, each CTA holds half of the weight tile
  int split;     // 1: fp16 hi/lo operands, three products per K chunk (see the header comment)
  int f16;       // 1 (non-split): operands and 16-bit outputs are fp16 instead of bf16
  int korder;    // 1 (3x3 stride-1 taps, non-split): K blocks run (dx, chunk, dy) like igemm_halo, so that a layer gives
                 // bit-identical sums whichever of the two kernels the planner picks for a batch size
  int a_lo;      // split: channel offset of the lo half inside source 0's pixel (= its channel count)
  int a_lo2;     // ... inside source 1's pixel
  int out_lo;    // split, !out_f32: channel offset of the lo half inside the output pixel (ldo = 2 * out_lo)
  int res_lo;    // split: channel offset of the lo half inside the residual pixel (ldr = 2 * res_lo)
  float out_scale;  // split: accumulators are multiplied by this power of two (undoes the weight pre-scale)
  int halo;      // split (igemm_x3.cuh): 1 = halo-staged 3x3/s1/p1: tmA boxes are {64, tw, 1, th + 2, 1}, tn = 1, tw % 8 = 0
  int pool;      // split + halo: 1 = fused 2x2 stride-2 max pool, output is [N][H/2][W/2][ldo]
  int kchain;    // split: K chunks of 64 per hi.hi accumulation (1 or 2; 0 = 1), see igemm_x3.cuh
  float comp;    // split: hi.hi chunk sums are multiplied by 1 + comp when they are added up (0 = the default), see igemm_x3.cuh
  long long* dbg;  // split, bring-up profiling only: [256][8] SM-clock stamps of pair 0's leader CTA (tools/test_igemm.cu), or null
  const float* bias;              // [n_valid] fp32 or nullptr
  const void* residual;           // NHWC bf16 (split: fp16 hi/lo) or nullptr
  void* out;                      // NHWC bf16 (split: fp16 hi/lo; fp32 when out_f32)
};

// Epilogue tail shared by both kernels: 32 consecutive output channels of one pixel, fp32 accumulators in f[]
// -> (split: x out_scale) + bias + residual -> ReLU -> store (bf16 / fp16 hi|lo / fp32).
__device__ __forceinline__ void igemm_finish32(const IgemmTcParams& p, float (&f)[32], size_t pix, int col0) {
  const bool full = (col0 + 32 <= p.n_valid);
  if (p.split) {
    const float sc = p.out_scale;
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] *= sc;
  }
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
  if (p.residual != nullptr && p.split) {
    // fp16 hi/lo residual: exact sum of the pair in fp32
    const __half* rp = reinterpret_cast<const __half*>(p.residual) + pix * p.ldr + col0;
    if (full) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        const uint4 rh = __ldg(reinterpret_cast<const uint4*>(rp + i));
        const uint4 rl = __ldg(reinterpret_cast<const uint4*>(rp + p.res_lo + i));
        const uint32_t hw[4] = {rh.x, rh.y, rh.z, rh.w}, lw[4] = {rl.x, rl.y, rl.z, rl.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 hv = __half22float2(*reinterpret_cast<const __half2*>(&hw[j]));
          const float2 lv = __half22float2(*reinterpret_cast<const __half2*>(&lw[j]));
          f[i + 2 * j] += hv.x + lv.x;
          f[i + 2 * j + 1] += hv.y + lv.y;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.n_valid) f[i] += __half2float(rp[i]) + __half2float(rp[p.res_lo + i]);
    }
  } else if (p.residual != nullptr) {
    const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + pix * p.ldr + col0;
    if (full) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        const uint4 r4 = __ldg(reinterpret_cast<const uint4*>(rp + i));
        const bool h = p.f16 != 0;
        f[i] += lo16_to_f32(r4.x, h);     f[i + 1] += hi16_to_f32(r4.x, h);
        f[i + 2] += lo16_to_f32(r4.y, h); f[i + 3] += hi16_to_f32(r4.y, h);
        f[i + 4] += lo16_to_f32(r4.z, h); f[i + 5] += hi16_to_f32(r4.z, h);
        f[i + 6] += lo16_to_f32(r4.w, h); f[i + 7] += hi16_to_f32(r4.w, h);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.n_valid)
          f[i] += p.f16 ? f16bits_to_f32(reinterpret_cast<const uint16_t*>(rp)[i]) : __bfloat162float(rp[i]);
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
  } else if (p.split) {
    __half* op = reinterpret_cast<__half*>(p.out) + pix * p.ldo + col0;
    float lo[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float hv = __half2float(__float2half_rn(f[i]));
      lo[i] = f[i] - hv;
      f[i] = hv;
    }
    if (full) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 o, l;
        __half2 t;
        t = __floats2half2_rn(f[i], f[i + 1]);       o.x = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(f[i + 2], f[i + 3]);   o.y = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(f[i + 4], f[i + 5]);   o.z = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(f[i + 6], f[i + 7]);   o.w = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(lo[i], lo[i + 1]);     l.x = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(lo[i + 2], lo[i + 3]); l.y = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(lo[i + 4], lo[i + 5]); l.z = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(lo[i + 6], lo[i + 7]); l.w = *reinterpret_cast<uint32_t*>(&t);
        *reinterpret_cast<uint4*>(op + i) = o;
        *reinterpret_cast<uint4*>(op + p.out_lo + i) = l;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.n_valid) { op[i] = __float2half_rn(f[i]); op[p.out_lo + i] = __float2half_rn(lo[i]); }
    }
  } else {
    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.ldo + col0;
    if (full) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 o;
        const bool h = p.f16 != 0;
        o.x = pack16x2(f[i], f[i + 1], h);     o.y = pack16x2(f[i + 2], f[i + 3], h);
        o.z = pack16x2(f[i + 4], f[i + 5], h); o.w = pack16x2(f[i + 6], f[i + 7], h);
        *reinterpret_cast<uint4*>(op + i) = o;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.n_valid) {
          if (p.f16) reinterpret_cast<uint16_t*>(op)[i] = static_cast<uint16_t>(pack_f16x2_sat(f[i], 0.f) & 0xffffu);
          else op[i] = __float2bfloat16_rn(f[i]);
        }
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(MODE == kModeSplit ? kTcThreadsSplit : kTcThreads, 1) igemm_tc_kernel(const __grid_constant__ IgemmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128-byte swizzle atoms (8 rows x 128 B).
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);  // [kTcMaxStages]
  uint64_t* empty_bar = full_bar + kTcMaxStages;           // [kTcMaxStages]
  uint64_t* tfull_bar = empty_bar + kTcMaxStages;          // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  int4* kb_table = reinterpret_cast<int4*>(smem + 256);    // [total_kb][2]: per K block {a_c, dw, dr, dh}, {src, b_k, -, -}
  uint8_t* stage_base = smem + kTcSmemHeader;

  // CTA pair (kModePair): the two CTAs of a cluster work on two consecutive M tiles and the same weight tile; CTA r
  // keeps rows [r * block_n/2, (r+1) * block_n/2) of the weight tile. The leader (rank 0) issues the M = 256 MMAs.
  constexpr bool kPair = MODE == kModePair;
  const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.num_stages;
  const int block_n = p.block_n;
  const int b_rows = kPair ? block_n >> 1 : block_n;  // weight rows in THIS CTA's shared memory
  const uint32_t a_bytes = kTcBlockM * kTcBlockK * 2;
  const uint32_t b_bytes = static_cast<uint32_t>(b_rows) * kTcBlockK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_tiles = (kPair ? (m_tiles + 1) >> 1 : m_tiles) * p.n_tiles;
  const int worker = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int num_workers = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int npass = MODE == kModeSplit ? 3 : 1;
  int total_kb = 0;
  for (int t = 0; t < p.num_taps; ++t) total_kb += (p.taps[t].src ? p.kchunks2 : p.kchunks) * npass;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], kPair ? 2 : 1);  // pair: the leader's expect_tx arrive + the peer's remote arrive
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], MODE == kModeStd ? 128 : 256);  // split: 8 epilogue warps; pair: both CTAs' 4 warps
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmA2);
  }
  // K-block schedule: the single producer thread's inner loop must stay a handful of instructions (measured: with
  // the tap/pass bookkeeping inside the loop the producer, not the tensor pipe, set the pace at ~550 cycles per K
  // block), so every thread precomputes one entry of the schedule here.
  for (int kb = threadIdx.x; kb < total_kb; kb += blockDim.x) {
    const int per_pass = total_kb / npass;
    const int pp = kb / per_pass;
    int rem = kb - pp * per_pass;
    // split precision, pass-major: first every lo.hi and hi.lo product (the accumulator is still ~2^-11 of its final
    // size there), then hi.hi. Weight K blocks of a tap are [hi chunks | lo chunks].
    const int pass = npass == 1 ? 0 : (pp == 0 ? 1 : (pp == 1 ? 2 : 0));  // 1: lo.hi, 2: hi.lo, 0: hi.hi
    int kb_base = 0, t = 0, chunks = 0;
    for (; t < p.num_taps; ++t) {
      chunks = p.taps[t].src ? p.kchunks2 : p.kchunks;
      if (rem < chunks) break;
      rem -= chunks;
      kb_base += npass == 1 ? chunks : 2 * chunks;
    }
    if (p.korder && npass == 1 && kb < 9 * p.kchunks) {
      // (dx, chunk, dy) order over the nine taps of source 0 (taps[] is dy-major: tap = dy * 3 + dx)
      const int dxi = kb / (3 * p.kchunks);
      const int r2 = kb - dxi * 3 * p.kchunks;
      const int c = r2 / 3, dyi = r2 - 3 * c;
      t = dyi * 3 + dxi;
      rem = c;
      chunks = p.kchunks;
      kb_base = t * p.kchunks;
    }
    const TapDesc tap = p.taps[t];
    const int a_c = (pass == 1 ? (tap.src ? p.a_lo2 : p.a_lo) : 0) + rem * kTcBlockK;
    const int b_k = (kb_base + (pass == 2 ? chunks : 0) + rem) * kTcBlockK;
    kb_table[2 * kb] = make_int4(a_c, tap.dw, tap.dr, tap.dh);
    kb_table[2 * kb + 1] = make_int4(tap.src, b_k, 0, 0);
  }
  if (warp == 1) {
    if (kPair) { tmem_alloc_pair(tmem_slot, kTcTmemCols); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, kTcTmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = worker; tile < total_tiles; tile += num_workers) {
        const int n_tile = tile % p.n_tiles;
        // pair: an odd tail leaves the peer with m_tile == m_tiles: its image coordinate is out of range, so TMA
        // zero-fills and the epilogue stores nothing
        const int m_tile = kPair ? ((tile / p.n_tiles) << 1) + static_cast<int>(cta_rank) : tile / p.n_tiles;
        const int iw = m_tile % p.tiles_w;
        const int ih = (m_tile / p.tiles_w) % p.tiles_h;
        const int in = m_tile / (p.tiles_w * p.tiles_h);
        const int w0 = iw * p.tw, h0 = ih * p.th, n0 = in * p.tn;
        const int b_col = n_tile * block_n + static_cast<int>(cta_rank) * b_rows;
        for (int kb = 0; kb < total_kb; ++kb) {
          const int4 e0 = kb_table[2 * kb];
          const int4 e1 = kb_table[2 * kb + 1];
          mbar_wait(&empty_bar[stage], phase ^ 1u, 1);
          uint8_t* sa = stage_base + static_cast<size_t>(stage) * stage_bytes;
          const void* tm = e1.x ? static_cast<const void*>(&p.tmA2) : static_cast<const void*>(&p.tmA);
          if (kPair) {
            // both CTAs' loads complete on the LEADER's barrier, which expects the bytes of both
            if (leader) mbar_expect_tx(&full_bar[stage], 2u * stage_bytes);
            tma_load_5d_pair(sa, tm, &full_bar[stage], e0.x, w0 + e0.y, e0.z, h0 + e0.w, n0);
            tma_load_2d_pair(sa + a_bytes, &p.tmB, &full_bar[stage], e1.y, b_col);
            if (!leader) mbar_arrive_remote(&full_bar[stage], 0u);
          } else {
            mbar_expect_tx(&full_bar[stage], stage_bytes);
            tma_load_5d(sa, tm, &full_bar[stage], e0.x, w0 + e0.y, e0.z, h0 + e0.w, n0);
            tma_load_2d(sa + a_bytes, &p.tmB, &full_bar[stage], e1.y, b_col);
          }
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (pair: leader CTA only) ===============
    if (leader && elect_one()) {
      const uint32_t idesc = MODE == kModeSplit ? make_idesc_f16(kTcBlockM, static_cast<uint32_t>(p.block_n))
                             : p.f16           ? make_idesc_f16(kPair ? 2 * kTcBlockM : kTcBlockM, static_cast<uint32_t>(p.block_n))
                                               : make_idesc_bf16(kPair ? 2 * kTcBlockM : kTcBlockM, static_cast<uint32_t>(p.block_n));
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      if (MODE == kModeSplit) {
        // Chunked accumulation: the tensor core's fp32 accumulate truncates, and that bias grows with the length
        // and the magnitude of the running sum. So a tile is accumulated in CHUNKS, each into a fresh TMEM buffer:
        // chunk 0 = every lo.hi / hi.lo product (small), then the hi.hi K blocks kTcSplitChunkKb at a time; the
        // epilogue warps add the chunks in registers with round-to-nearest fp32.
        const int lo_kb = (total_kb / 3) * 2;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
          int kb = 0;
          while (kb < total_kb) {
            const int kb_end = kb < lo_kb ? lo_kb : (kb + kTcSplitChunkKb < total_kb ? kb + kTcSplitChunkKb : total_kb);
            mbar_wait(&tempty_bar[as], aphase ^ 1u, 2);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as) * 256u;
            for (const int k0 = kb; kb < kb_end; ++kb) {
              mbar_wait(&full_bar[stage], phase, 3);
              tc_fence_after();
              const uint32_t sa = smem_u32(stage_base + static_cast<size_t>(stage) * stage_bytes);
              const uint64_t a_desc = make_sw128_kmajor_desc(sa);
              const uint64_t b_desc = make_sw128_kmajor_desc(sa + a_bytes);
#pragma unroll
              for (int k = 0; k < kTcBlockK / 16; ++k)
                umma_bf16(d_tmem, a_desc + static_cast<uint64_t>(2 * k), b_desc + static_cast<uint64_t>(2 * k), idesc,
                          (kb > k0 || k > 0) ? 1u : 0u);
              umma_commit(&empty_bar[stage]);
              if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
            umma_commit(&tfull_bar[as]);
            if (++as == 2) { as = 0; aphase ^= 1u; }
          }
        }
      } else {
        for (int tile = worker; tile < total_tiles; tile += num_workers) {
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
              if (kPair)
                umma_bf16_pair(d_tmem, a_desc + static_cast<uint64_t>(2 * k), b_desc + static_cast<uint64_t>(2 * k), idesc,
                               (kb > 0 || k > 0) ? 1u : 0u);
              else
                umma_bf16(d_tmem, a_desc + static_cast<uint64_t>(2 * k), b_desc + static_cast<uint64_t>(2 * k), idesc,
                          (kb > 0 || k > 0) ? 1u : 0u);
            }
            // frees the smem slot (pair: in both CTAs) once these MMAs have read it
            if (kPair) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
            if (++stage == stages) { stage = 0; phase ^= 1u; }
          }
          // accumulator complete -> epilogue (pair: the epilogue warps of both CTAs)
          if (kPair) umma_commit_pair(&tfull_bar[as]); else umma_commit(&tfull_bar[as]);
          if (++as == 2) { as = 0; aphase ^= 1u; }
        }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    if (MODE == kModeSplit) {
      // 8 warps: warp (q, half) owns rows q*32.. of column half `half`; fp32 register accumulators over the chunks
      const int half = (warp - 2) >> 2;
      const int half_cols = p.block_n >> 1;  // multiple of 32 (block_n is a multiple of 64 on this path)
      const int nchunks = 1 + (total_kb / 3 + kTcSplitChunkKb - 1) / kTcSplitChunkKb;
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
        const int n_off = n_tile * p.block_n + half * half_cols;
        float racc[128];
#pragma unroll
        for (int i = 0; i < 128; ++i) racc[i] = 0.f;
        for (int ch = 0; ch < nchunks; ++ch) {
          mbar_wait(&tfull_bar[as], aphase, 4);
          tc_fence_after();
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as) * 256u +
                                 static_cast<uint32_t>(half * half_cols);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g * 32 < half_cols) {
              uint32_t v[32];
              tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(g * 32), v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) racc[g * 32 + i] += __uint_as_float(v[i]);
            }
          }
          tc_fence_before();
          mbar_arrive(&tempty_bar[as]);
          if (++as == 2) { as = 0; aphase ^= 1u; }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col0 = n_off + g * 32;
          if (g * 32 < half_cols && valid && col0 < p.n_valid) {
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = racc[g * 32 + i];
            igemm_finish32(p, f, pix, col0);
          }
        }
      }
    } else
    for (int tile = worker; tile < total_tiles; tile += num_workers) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = kPair ? ((tile / p.n_tiles) << 1) + static_cast<int>(cta_rank) : tile / p.n_tiles;
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

      {
        for (int c0 = 0; c0 < p.block_n; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c0), v);
          tmem_ld_wait();
          if (c0 + 32 >= p.block_n) {
            // last read of this accumulator: hand the TMEM buffer back before the global stores (pair: to the
            // leader's barrier, which counts the epilogue threads of both CTAs)
            tc_fence_before();
            if (kPair && !leader) mbar_arrive_remote(&tempty_bar[as], 0u);
            else mbar_arrive(&tempty_bar[as]);
          }
          const int col0 = n_off + c0;
          if (valid && col0 < p.n_valid) {
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
            igemm_finish32(p, f, pix, col0);
          }
        }
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();  // pair: the peer's smem/TMEM stay live until the leader's MMAs are done
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if (kPair) tmem_dealloc_pair(tmem_base, kTcTmemCols); else tmem_dealloc(tmem_base, kTcTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
inline size_t igemm_tc_smem_bytes(int block_n, int* stages_out, int pair = 0) {
  const size_t stage = static_cast<size_t>(kTcBlockM) * kTcBlockK * 2 + static_cast<size_t>(pair ? block_n / 2 : block_n) * kTcBlockK * 2;
  const size_t header = 1024 + kTcSmemHeader;  // alignment slack + barriers + K-block schedule
  const size_t budget = 227 * 1024 - header;
  int stages = static_cast<int>(budget / stage);
  if (stages > kTcMaxStages) stages = kTcMaxStages;
  if (stages_out) *stages_out = stages;
  return header + stages * stage;
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

// The same [rows][64] matrix as the DESTINATION of the LSTM's TMA publish: un-swizzled {16 units, box_rows} boxes
// (32 bytes per row), stored from a dense [box_rows][16] shared-memory tile.
inline int make_rows64_store_tmap(CUtensorMap* tm, const void* base, long long rows, int box_rows) {
  PFN_tmapEncodeTiled enc = get_tmap_encode();
  if (!enc) return -1;
  cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? 0 : -static_cast<int>(rc) - 1000;
}

// The LSTM layer output [rows][T * 1024] (16-bit elements) as the destination of the TMA publish: un-swizzled
// {16 units, box_rows} boxes stored from the same dense shared-memory tile as the exchange copy; rows >= `rows` are
// clipped by the map's bounds.
inline int make_hseq_store_tmap(CUtensorMap* tm, const void* base, int T, int rows, int box_rows) {
  PFN_tmapEncodeTiled enc = get_tmap_encode();
  if (!enc) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)T * 1024, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)T * 1024 * 2};
  cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? 0 : -static_cast<int>(rc) - 1000;
}

inline cudaError_t igemm_x3_launch(const IgemmTcParams& p, int num_sms, cudaStream_t stream);  // igemm_x3.cuh

inline cudaError_t igemm_tc_launch(const IgemmTcParams& p, int /*mode*/, int num_sms, cudaStream_t stream) {
  if (p.split && p.pair) return igemm_x3_launch(p, num_sms, stream);  // fp32-grade mode: the fused-stage CTA-pair kernel
  static bool attr_set_dev[kMaxDevices][3] = {};
  bool* attr_set = attr_set_dev[current_device_slot()];
  const int mode = p.split ? kModeSplit : (p.pair ? kModePair : kModeStd);
  int stages = 0;
  const size_t smem = igemm_tc_smem_bytes(p.block_n, &stages, p.pair);
  if (stages != p.num_stages) return cudaErrorInvalidValue;
  if (!attr_set[mode]) {
    const void* fn = mode == kModeSplit  ? reinterpret_cast<const void*>(igemm_tc_kernel<kModeSplit>)
                     : mode == kModePair ? reinterpret_cast<const void*>(igemm_tc_kernel<kModePair>)
                                         : reinterpret_cast<const void*>(igemm_tc_kernel<kModeStd>);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[mode] = true;
  }
  if (p.split && (p.block_n & 63) != 0) return cudaErrorInvalidValue;  // the split epilogue halves the tile in 32-column groups
  if (p.pair && (p.split || (p.block_n & 15) != 0)) return cudaErrorInvalidValue;
  int total_kb = 0;
  for (int t = 0; t < p.num_taps; ++t) total_kb += (p.taps[t].src ? p.kchunks2 : p.kchunks) * (p.split ? 3 : 1);
  if (total_kb > kTcMaxKb) return cudaErrorInvalidValue;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  if (mode == kModePair) {
    // clusters of 2 CTAs (one TPC): 74 pairs on a 148-SM part, persistent over (M-tile pair, N tile) work items
    const int pair_tiles = ((m_tiles + 1) / 2) * p.n_tiles;
    const int pairs = pair_tiles < num_sms / 2 ? pair_tiles : num_sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, igemm_tc_kernel<kModePair>, p);
  }
  const int total_tiles = m_tiles * p.n_tiles;
  const int grid = total_tiles < num_sms ? total_tiles : num_sms;
  if (mode == kModeSplit)
    igemm_tc_kernel<kModeSplit><<<grid, kTcThreadsSplit, smem, stream>>>(p);
  else
    igemm_tc_kernel<kModeStd><<<grid, kTcThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace ocrfa

#include "igemm_x3.cuh"
