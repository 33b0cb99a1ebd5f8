// Split-precision (fp32-grade) tcgen05 implicit GEMM: the "f16x3" mode of every convolution / linear layer.
//
//   reference arithmetic being matched: fp32 convs and GEMMs of
//     /root/reference/src/ocrfa/modules/blocks.py:200-223, 239-322 (ResNet), :89-101 (BiLSTM GEMMs), :438 (Prediction)
//
// Every operand is an fp16 pair x = hi + lo (22 significant bits + the sign of lo; activations NHWC [hi(C) | lo(C)]
// per pixel, weights [Cout][tap][hi(C) | lo(C)] pre-scaled by a power of two) and every algorithmic MAC costs three
// tensor-core products: hi.hi + lo.hi + hi.lo (the dropped lo.lo term is ~2^-22 relative).
//
// Design:
//   * CTA pair (cta_group::2, M = 256 over the two SMs of a TPC): each CTA stages its own 128 pixel rows and HALF of
//     the weight tile, so the weight bytes per SM halve.
//   * ONE shared-memory stage per K chunk of 64 holds all four operand tiles {A_hi, A_lo, W_hi, W_lo} and feeds the
//     three products: 64 KB per 12 K=16 steps instead of 48 KB per 4 - the operands are fetched once, not once per
//     product (L2 -> SM bytes per MMA fall 2.25x against a kernel that streams one product at a time).
//   * The tensor core's fp32 accumulate truncates; measured on B200 (tools/test_igemm.cu --split --positive, relu'd
//     outputs, error against the float64 sum of the same operands, relative to rms): mean -0.8e-8 - 0.56e-8 per chained
//     MMA, i.e. -5.2e-8 for chains of 8, -3.1e-8 for 4, -1.9e-8 for 2 - a BIAS, which adds up linearly over the 29
//     layers. So
//       - the hi.hi products of ONE K chunk (four chained MMAs, K = 64) go to a fresh TMEM accumulator ("main") that
//         the eight epilogue warps drain into fp32 REGISTER accumulators with round-to-nearest adds, chunk after chunk;
//       - the lo.hi / hi.lo products (2^-11 of the result) chain in a second TMEM accumulator ("corr") for the
//         whole tile - their truncation error is 2^-11 of an already small term - and are added once at the end.
//     The MMA issuer orders a stage as corr (8 MMAs), then main: while the epilogue drains chunk k from "main", the
//     tensor pipe is busy with the corr MMAs of chunk k+1, so the drain is hidden behind tensor work:
//       - 256-column tiles (TMEM: main 256 + corr 256 columns): the hi.hi product is issued as two N = 128 halves
//         with their own barriers; the warps of column half h drain half h. A half is needed again only after the
//         other half's MMAs and the next corr MMAs (1280 tensor-pipe cycles).
//       - tiles of <= 128 columns: TWO main buffers (columns 0 and 128) used alternately by successive chunks (the
//         small-Cout layers have short stages; the drain of chunk k also overlaps the hi.hi MMAs of chunk k+1).
//   * halo mode (IgemmTcParams::halo, the 3x3 / stride 1 / pad 1 layers with Cout <= 128 and large maps - conv0_2,
//     layer1, conv1): like igemm_halo.cuh, ONE box of (th + 2) x tw pixels per (dx, 64-channel chunk) - once for the hi
//     and once for the lo halves - feeds the three dy taps (tw is a multiple of 8, so the rows of tap dy start
//     dy * tw * 128 B into the box, a multiple of the 1024-byte swizzle atom); the six weight blocks (three taps x hi, lo)
//     of the group travel with it. A stage then carries 36 MMAs: L2 -> SM bytes per MMA fall from 104 (Cout = 64) /
//     62 (Cout = 128) bytes per tensor-pipe cycle - above what L2 delivers to 148 SMs at once - to 57 / 39. The
//     following 2x2/2 max pool (max_pool1, max_pool2) can be fused into the epilogue.
//   * IgemmTcParams::kchain (1 or 2): K chunks per hi.hi accumulation - the precision / speed knob described at its use.
//   * truncation compensation (IgemmTcParams::comp, default 2^-24): the tensor core adds into its fp32 accumulator by
//     TRUNCATION toward zero, so every chunk sum is a little short and a layer's output is a little smaller than the
//     exactly rounded one - coherently, layer after layer: measured on the recognizer, the visual features come out
//     scaled by 1 - 8.9e-7 (tools/error_budget.py `shrink`), which is 88 % of this mode's whole error against float64.
//     The tile's register sum x is therefore finished as fma(x, 2^-24, x) = the next fp32 number away from zero: with
//     it the shrink is +0.07e-7, the feature error falls from 10.1e-7 to 3.8e-7 of rms (the reference's own fp32
//     evaluation: 4.5e-7) and the logit error from 11.4e-7 to 3.7e-7 (reference: 6.2e-7). One packed FFMA per two
//     accumulators per tile.
// Warp roles (12 warps, three warpgroups): warpgroup 0 = TMA producer (warp 0), MMA issuer + TMEM alloc (warp 1, leader
// CTA) and two idle warps, trimmed to 56 registers with setmaxnreg; warpgroups 1-2 = eight epilogue warps raised to 224
// registers, so the 128 fp32 accumulators + 64 values in flight from TMEM live in registers (at the 168 registers a
// flat 320-thread launch gets, ptxas kept 16 accumulators in local memory and re-read / re-wrote them every chunk).
// The register adds are packed (add.rn.f32x2: two fp32 lanes per instruction).
#pragma once
#include <cstdlib>
#include "igemm_tc.cuh"

namespace ocrfa {

constexpr int kX3Threads = 384;
constexpr int kX3FirstEpiWarp = 4;
constexpr float kX3Comp = 0x1p-24f;  // default of IgemmTcParams::comp: see "truncation compensation" in the header comment (experiment switch: OCRFA_X3_COMP)
constexpr float kX3Comp2 = 0x1.4p-23f;  // ... with kchain = 2 (eight chained MMAs per accumulation): 2.5 x 2^-24
constexpr int kX3CorrCol = 256;  // TMEM column of the corr accumulator (main starts at column 0)
// Tile tail through a per-warp transposition buffer (x3_finish32 below). Measured and NOT kept: the 512-channel layer
// went 1.39-1.47 -> 1.54 ms with it. Stores are fire-and-forget - the warp does not wait for the L1 wavefronts of its
// 16-byte-per-lane accesses - while the transposition adds dependent shared-memory round trips and registers to the
// one phase in which the epilogue warps cannot drain. The code stays as the record of the experiment.
constexpr bool kX3TransposedTail = false;
constexpr int kX3Staging = kX3TransposedTail ? 8 * 2048 : 0;  // per epilogue warp: 32 rows x 64 bytes (behind the stages)

inline size_t igemm_x3_smem_bytes(int block_n, int* stages_out) {
  const size_t stage = 2 * static_cast<size_t>(kTcBlockM) * kTcBlockK * 2 + 2 * static_cast<size_t>(block_n / 2) * kTcBlockK * 2;
  const size_t header = 1024 + kTcSmemHeader + kX3Staging;
  const size_t budget = 227 * 1024 - header;
  int stages = static_cast<int>(budget / stage);
  if (stages > kTcMaxStages) stages = kTcMaxStages;
  if (stages_out) *stages_out = stages;
  return header + stages * stage;
}

// halo mode: a stage = two halo boxes ((th + 2) * tw pixels x 128 B) + six weight blocks
inline size_t igemm_x3_halo_smem_bytes(int block_n, int tw, int th, int* stages_out) {
  const size_t stage = 2 * static_cast<size_t>(th + 2) * tw * 128 + 6 * static_cast<size_t>(block_n / 2) * kTcBlockK * 2;
  const size_t header = 1024 + kTcSmemHeader + kX3Staging;
  const size_t budget = 227 * 1024 - header;
  int stages = static_cast<int>(budget / stage);
  if (stages > kTcMaxStages) stages = kTcMaxStages;
  if (stages_out) *stages_out = stages;
  return header + stages * stage;
}

// 8 consecutive output channels of one pooled pixel: (+bias) -> ReLU -> fp16 hi | lo store (halo mode, fused max pool)
__device__ __forceinline__ void x3_store8_split(const IgemmTcParams& p, float (&g)[8], size_t pix, int col) {
  const float sc = p.out_scale;
#pragma unroll
  for (int i = 0; i < 8; ++i) g[i] *= sc;
  if (p.bias != nullptr) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
    g[0] += b0.x; g[1] += b0.y; g[2] += b0.z; g[3] += b0.w;
    g[4] += b1.x; g[5] += b1.y; g[6] += b1.z; g[7] += b1.w;
  }
  float lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (p.relu) g[i] = fmaxf(g[i], 0.0f);
    const float hv = __half2float(__float2half_rn(g[i]));
    lo[i] = g[i] - hv;
    g[i] = hv;
  }
  uint4 o, l;
  __half2 t;
  t = __floats2half2_rn(g[0], g[1]);   o.x = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2half2_rn(g[2], g[3]);   o.y = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2half2_rn(g[4], g[5]);   o.z = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2half2_rn(g[6], g[7]);   o.w = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2half2_rn(lo[0], lo[1]); l.x = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2half2_rn(lo[2], lo[3]); l.y = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2half2_rn(lo[4], lo[5]); l.z = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2half2_rn(lo[6], lo[7]); l.w = *reinterpret_cast<uint32_t*>(&t);
  __half* op = reinterpret_cast<__half*>(p.out) + pix * p.ldo + col;
  *reinterpret_cast<uint4*>(op) = o;
  *reinterpret_cast<uint4*>(op + p.out_lo) = l;
}

// Tile tail for 32 consecutive output channels of this warp's 32 pixels (lane = pixel): x out_scale, + bias, + residual,
// ReLU, store. Every global access goes through a per-warp transposition buffer (32 rows x 64 B, pieces XOR-swizzled by
// (row >> 1) & 3) so that one load / store instruction covers 8 pixels x 64 contiguous bytes: a 16-byte access per pixel
// row costs one L1 wavefront per lane, and - unlike the single-pass kernels, whose epilogue overlaps the next tile's
// MMAs - these warps are needed again for the next tile's first drain, so the tail is on the tensor pipe's critical path
// (ncu, 512-channel layers before this change: tensor pipe 70 % active). pixk / vk: pixel index / validity of rows
// 8k + (lane >> 2), k = 0..3 (warp shuffles of pix / valid, done once per tile).
__device__ __forceinline__ void x3_finish32(const IgemmTcParams& p, float (&f)[32], uint8_t* stg, int lane, size_t pix,
                                            bool valid, const size_t (&pixk)[4], const bool (&vk)[4], int col0) {
  const bool fast = (col0 + 32 <= p.n_valid) && (!p.out_f32 || (p.ldo & 3) == 0);
  if (!fast) {  // ragged last column group (Prediction's 185 classes): the plain per-pixel tail
    if (valid) igemm_finish32(p, f, pix, col0);
    return;
  }
  const float sc = p.out_scale;
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] *= sc;
  if (p.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
      f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
    }
  }
  const int swz_own = (lane >> 1) & 3;
  const int pc = lane & 3;  // the 16-byte piece this lane moves in the transposed mapping
  uint8_t* own = stg + lane * 64;
  uint8_t* tr = stg + static_cast<uint32_t>(lane >> 2) * 64u + static_cast<uint32_t>((pc ^ ((lane >> 3) & 3)) << 4);
  if (p.residual != nullptr) {
    // fp16 hi | lo residual: both halves are added in fp32 (hi + lo is exact)
    const __half* rbase = reinterpret_cast<const __half*>(p.residual) + col0 + pc * 8;
#pragma unroll
    for (int part = 0; part < 2; ++part) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint4 r4 = make_uint4(0u, 0u, 0u, 0u);
        if (vk[k]) r4 = __ldg(reinterpret_cast<const uint4*>(rbase + pixk[k] * p.ldr + part * p.res_lo));
        *reinterpret_cast<uint4*>(tr + k * 512) = r4;
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 r4 = *reinterpret_cast<const uint4*>(own + ((j ^ swz_own) << 4));
        const uint32_t w4[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w4[e]));
          f[8 * j + 2 * e] += v.x;
          f[8 * j + 2 * e + 1] += v.y;
        }
      }
      __syncwarp();
    }
  }
  if (p.relu) {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.0f);
  }
  if (p.out_f32) {
#pragma unroll
    for (int hp = 0; hp < 2; ++hp) {  // 16 floats = 64 bytes per pixel and pass
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(own + ((j ^ swz_own) << 4)) =
            make_float4(f[16 * hp + 4 * j], f[16 * hp + 4 * j + 1], f[16 * hp + 4 * j + 2], f[16 * hp + 4 * j + 3]);
      __syncwarp();
      float* obase = reinterpret_cast<float*>(p.out) + col0 + 16 * hp + pc * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 o = *reinterpret_cast<const float4*>(tr + k * 512);
        if (vk[k]) *reinterpret_cast<float4*>(obase + pixk[k] * p.ldo) = o;
      }
      __syncwarp();
    }
    return;
  }
  float lo[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float hv = __half2float(__float2half_rn(f[i]));
    lo[i] = f[i] - hv;
    f[i] = hv;
  }
#pragma unroll
  for (int part = 0; part < 2; ++part) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 o;
      __half2 t;
      const float* src = part ? lo : f;
      t = __floats2half2_rn(src[8 * j], src[8 * j + 1]);     o.x = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2half2_rn(src[8 * j + 2], src[8 * j + 3]); o.y = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2half2_rn(src[8 * j + 4], src[8 * j + 5]); o.z = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2half2_rn(src[8 * j + 6], src[8 * j + 7]); o.w = *reinterpret_cast<uint32_t*>(&t);
      *reinterpret_cast<uint4*>(own + ((j ^ swz_own) << 4)) = o;
    }
    __syncwarp();
    __half* obase = reinterpret_cast<__half*>(p.out) + col0 + pc * 8 + part * p.out_lo;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint4 o = *reinterpret_cast<const uint4*>(tr + k * 512);
      if (vk[k]) *reinterpret_cast<uint4*>(obase + pixk[k] * p.ldo) = o;
    }
    __syncwarp();
  }
}

// KIND (compile-time, so that the MMA warp - 56 registers - carries one issue path): 0 = 256-column tiles, 1 = tiles of
// <= 128 columns per tap, 2 = the same halo-staged, 3 = other widths (192: one buffer)
enum X3Kind : int { kX3Wide = 0, kX3Narrow = 1, kX3NarrowHalo = 2, kX3Other = 3 };
template <int KIND>
__global__ void __launch_bounds__(kX3Threads, 1) igemm_x3_kernel(const __grid_constant__ IgemmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);  // [kTcMaxStages]
  uint64_t* empty_bar = full_bar + kTcMaxStages;           // [kTcMaxStages]
  uint64_t* main_full = empty_bar + kTcMaxStages;          // [2] hi.hi chunk (or chunk half) complete (both CTAs, multicast commit)
  uint64_t* main_empty = main_full + 2;                    // [2] drained by the epilogue warps of the pair (leader's copy)
  uint64_t* corr_full = main_empty + 2;                    // every MMA of the tile complete
  uint64_t* corr_empty = corr_full + 1;                    // corr read back by the 16 epilogue warps (leader's copy)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(corr_empty + 1);
  int4* kb_table = reinterpret_cast<int4*>(smem + 256);    // [groups][3]: {a_c, dw, dr, dh}, {src, nsub, a_lo, b_lo}, {b_k of sub 0..2, -}
  uint8_t* stage_base = smem + kTcSmemHeader;

  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.num_stages;
  const int block_n = p.block_n;
  const int b_rows = block_n >> 1;
  const uint32_t a_bytes = kTcBlockM * kTcBlockK * 2;
  const uint32_t b_bytes = static_cast<uint32_t>(b_rows) * kTcBlockK * 2;
  constexpr bool halo = KIND == kX3NarrowHalo;
  const uint32_t tap_bytes = static_cast<uint32_t>(p.tw) * 128u;  // halo: one image row of the box
  const uint32_t a_slot = halo ? static_cast<uint32_t>(p.th + 2) * tap_bytes : a_bytes;  // room of one A box (hi or lo)
  const uint32_t stage_bytes = 2u * a_slot + (halo ? 6u : 2u) * b_bytes;
  // wide: a 256-column tile whose hi.hi product is issued as two N = 128 halves (buffer h = columns {h*64 + 0..63,
  // 128 + h*64 + 0..63} of the tile: an N = 128 pair MMA takes 64 weight rows from each CTA); otherwise two whole-tile
  // buffers used alternately
  constexpr bool wide = KIND == kX3Wide;
  constexpr bool alt = KIND == kX3Narrow || KIND == kX3NarrowHalo;  // two alternating whole-tile buffers (a 192-column tile has room for one)
  // K chunks of 64 per hi.hi accumulation. 1: four chained MMAs, the smallest truncation bias - but every chunk is a
  // 128 x N fp32 read of TMEM, 341 / 4 = 85 B per tensor-pipe cycle and SM, which is what the TMEM read port delivers
  // (measured: the eight epilogue warps busy 69 % of the kernel, tensor pipe 70 % active). 2: eight chained MMAs, half
  // the TMEM traffic (the drain hides), bias -5.2e-8 instead of -3.1e-8 of the rms per layer.
  const int kchain = p.kchain > 1 ? p.kchain : 1;
  const int kmask = kchain - 1;  // kchain is 1 or 2: (x % kchain) == (x & kmask), without the integer division in the issue loop
  const bool stagger = wide && kchain == 2;  // the two column halves close on alternate stages (see the MMA issuer)
  // kchain = 2 pairs the K chunks of an output column as (0,1)(2,3).. when its 64-column group is even and as
  // (0)(1,2)(3,4).. when it is odd - a function of the COLUMN, not of the tile shape, so a layer whose 256-column tiles
  // are staggered gives bit-identical sums when the planner narrows its tiles to 64 columns for a small batch (128-column
  // tiles of such a layer would need both pairings at once: igemm_x3_launch rejects them)
  const bool odd_groups = kchain == 2 && block_n == 64 && p.n_valid >= 256;

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_tiles = ((m_tiles + 1) >> 1) * p.n_tiles;
  const int worker = static_cast<int>(blockIdx.x >> 1);
  const int num_workers = static_cast<int>(gridDim.x >> 1);
  // a GROUP is one shared-memory stage: one K chunk of one tap, or (halo) the three dy taps of one (dx, chunk);
  // a CHUNK is one fresh hi.hi accumulation (K = 64 of one tap)
  int total_kb = 0, total_chunks = 0;
  if (halo) {
    total_kb = 3 * p.kchunks + p.kchunks2;
    total_chunks = 9 * p.kchunks + p.kchunks2;
  } else {
    for (int t = 0; t < p.num_taps; ++t) total_kb += p.taps[t].src ? p.kchunks2 : p.kchunks;
    total_chunks = total_kb;
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 2);  // the leader's expect_tx arrive + the peer's remote arrive
      mbar_init(&empty_bar[s], 1);
    }
    // wide: a buffer half is drained by the 4 warps of that column half in each CTA; else by all 8 warps of each CTA
    for (int b = 0; b < 2; ++b) { mbar_init(&main_full[b], 1); mbar_init(&main_empty[b], wide ? 8 : 16); }
    mbar_init(corr_full, 1);
    mbar_init(corr_empty, 16);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmA2);
  }
  // group schedule. Weight K blocks of a tap are [hi chunks | lo chunks]; taps are dy-major (tap = dy * 3 + dx).
  for (int kb = threadIdx.x; kb < total_kb; kb += blockDim.x) {
    if (halo) {
      if (kb < 3 * p.kchunks) {  // dx-major, chunk-minor; the three sub-blocks are the dy taps
        const int dxi = kb / p.kchunks, c = kb - dxi * p.kchunks;
        kb_table[3 * kb] = make_int4(c * kTcBlockK, dxi - 1, 0, -1);
        kb_table[3 * kb + 1] = make_int4(0, 3, p.a_lo, p.kchunks * kTcBlockK);
        kb_table[3 * kb + 2] = make_int4(((0 * 3 + dxi) * 2 * p.kchunks + c) * kTcBlockK, ((1 * 3 + dxi) * 2 * p.kchunks + c) * kTcBlockK,
                                         ((2 * 3 + dxi) * 2 * p.kchunks + c) * kTcBlockK, 0);
      } else {  // fused 1x1 source
        const int c = kb - 3 * p.kchunks;
        kb_table[3 * kb] = make_int4(c * kTcBlockK, 0, 0, 0);
        kb_table[3 * kb + 1] = make_int4(1, 1, p.a_lo2, p.kchunks2 * kTcBlockK);
        kb_table[3 * kb + 2] = make_int4((9 * 2 * p.kchunks + c) * kTcBlockK, 0, 0, 0);
      }
      continue;
    }
    int rem = kb, kb_base = 0, t = 0, chunks = 0;
    for (; t < p.num_taps; ++t) {
      chunks = p.taps[t].src ? p.kchunks2 : p.kchunks;
      if (rem < chunks) break;
      rem -= chunks;
      kb_base += 2 * chunks;
    }
    const TapDesc tap = p.taps[t];
    kb_table[3 * kb] = make_int4(rem * kTcBlockK, tap.dw, tap.dr, tap.dh);
    kb_table[3 * kb + 1] = make_int4(tap.src, 1, tap.src ? p.a_lo2 : p.a_lo, chunks * kTcBlockK);
    kb_table[3 * kb + 2] = make_int4((kb_base + rem) * kTcBlockK, 0, 0, 0);
  }
  if (warp == 1) { tmem_alloc_pair(tmem_slot, kTcTmemCols); tmem_relinquish_pair(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // bring-up profiling: stamps of the leader CTA of pair 0, slot = running stage / chunk index of the CTA
#ifdef OCRFA_X3_PROFILE  // tools/test_igemm.cu builds with it; the product library carries no stamps
  long long* dbg = (blockIdx.x == 0) ? p.dbg : nullptr;
#define X3_TS(idx, col) do { if (dbg && (idx) < 256 && ((threadIdx.x & 31) == 0)) dbg[(idx) * 8 + (col)] = clock64(); } while (0)
#else
#define X3_TS(idx, col) do { } while (0)
#endif
  if (warp < kX3FirstEpiWarp) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");  // 64 / 224 would use the register file to the last register: tried, the launch fails
  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) =========================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = worker; tile < total_tiles; tile += num_workers) {
        const int n_tile = tile % p.n_tiles;
        // an odd tail leaves the peer with m_tile == m_tiles: its image coordinate is out of range, TMA zero-fills
        const int m_tile = ((tile / p.n_tiles) << 1) + static_cast<int>(cta_rank);
        const int iw = m_tile % p.tiles_w;
        const int ih = (m_tile / p.tiles_w) % p.tiles_h;
        const int in = m_tile / (p.tiles_w * p.tiles_h);
        const int w0 = iw * p.tw, h0 = ih * p.th, n0 = in * p.tn;
        const int b_col = n_tile * block_n + static_cast<int>(cta_rank) * b_rows;
        for (int kb = 0; kb < total_kb; ++kb) {
          const int4 e0 = kb_table[3 * kb];
          const int4 e1 = kb_table[3 * kb + 1];
          const int4 e2 = kb_table[3 * kb + 2];
          mbar_wait(&empty_bar[stage], phase ^ 1u, 1);
          uint8_t* sa = stage_base + static_cast<size_t>(stage) * stage_bytes;
          const void* tm = e1.x ? static_cast<const void*>(&p.tmA2) : static_cast<const void*>(&p.tmA);
          // bytes of this group: two A boxes (halo boxes for the 3x3 source, plain 128-pixel boxes otherwise) + 2 * nsub
          // weight blocks; both CTAs' loads complete on the leader's barrier
          const uint32_t a_box = (halo && e1.x == 0) ? a_slot : a_bytes;
          if (leader) mbar_expect_tx(&full_bar[stage], 2u * (2u * a_box + 2u * static_cast<uint32_t>(e1.y) * b_bytes));
          tma_load_5d_pair(sa, tm, &full_bar[stage], e0.x, w0 + e0.y, e0.z, h0 + e0.w, n0);
          tma_load_5d_pair(sa + a_slot, tm, &full_bar[stage], e0.x + e1.z, w0 + e0.y, e0.z, h0 + e0.w, n0);
          uint8_t* sb = sa + 2u * a_slot;
          tma_load_2d_pair(sb, &p.tmB, &full_bar[stage], e2.x, b_col);
          tma_load_2d_pair(sb + b_bytes, &p.tmB, &full_bar[stage], e2.x + e1.w, b_col);
          if (e1.y == 3) {
            tma_load_2d_pair(sb + 2u * b_bytes, &p.tmB, &full_bar[stage], e2.y, b_col);
            tma_load_2d_pair(sb + 3u * b_bytes, &p.tmB, &full_bar[stage], e2.y + e1.w, b_col);
            tma_load_2d_pair(sb + 4u * b_bytes, &p.tmB, &full_bar[stage], e2.z, b_col);
            tma_load_2d_pair(sb + 5u * b_bytes, &p.tmB, &full_bar[stage], e2.z + e1.w, b_col);
          }
          if (!leader) mbar_arrive_remote(&full_bar[stage], 0u);
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) =====================
    // The WHOLE warp walks the loop with warp-uniform values and only the tcgen05.mma / commit instructions are
    // predicated on the elected lane: inside a one-thread branch every descriptor lives in a vector register and each
    // UTCHMMA pays for four R2UR moves - measured ~60 cycles per issue, 960 of a stage's 1 536 tensor cycles with 16
    // MMAs per stage, i.e. the issuing thread, not the tensor pipe, set the pace (stage period 1 600 / 2 200 cycles).
    if (leader) {
      const bool elected = elect_one();
      const uint32_t idesc = make_idesc_f16(2 * kTcBlockM, static_cast<uint32_t>(block_n));
      const uint32_t idesc_half = make_idesc_f16(2 * kTcBlockM, 128u);
      const uint32_t d_main = tmem_base;
      const uint32_t d_corr = tmem_base + static_cast<uint32_t>(kX3CorrCol);
      int stage = 0;
      uint32_t phase = 0, cph = 0, mph = 0;  // mph: one parity bit per "main" buffer
      int mb = 0;
      int gstage = 0;  // running stage index of this CTA (profiling stamps)
      for (int tile = worker; tile < total_tiles; tile += num_workers) {
        mbar_wait(corr_empty, cph ^ 1u, 2);  // the previous tile's corr has been read back
        cph ^= 1u;
        int unit = 0;  // K chunks of 64 issued so far in this tile; `kchain` consecutive ones share one accumulation
        const int ph = odd_groups ? ((tile % p.n_tiles) & 1) : 0;
        for (int kb = 0; kb < total_kb; ++kb) {
          X3_TS(gstage, 0);
          mbar_wait(&full_bar[stage], phase, 3);
          X3_TS(gstage, 1);
          tc_fence_after();
          const uint32_t sa0 = smem_u32(stage_base + static_cast<size_t>(stage) * stage_bytes);
          const int nsub = halo ? kb_table[3 * kb + 1].y : 1;
          if (stagger) {
            // 256-column tiles with kchain = 2: each column half accumulates TWO K chunks, and the halves close on
            // alternate stages. Issue order per stage: [the half that closes] -> corr -> [the half that starts]. A half
            // handed to the epilogue at the start of stage s is not written again before the end of stage s + 1
            // (2 560 tensor-pipe cycles later), and only one half (64 KB per CTA, ~1 000 cycles of the 64 B/cycle TMEM
            // read port) is being drained at a time - the drain no longer paces the tensor pipe.
            const int u = unit++;
            const int hc = (u & 1) ? 0 : 1;  // half 0 closes on odd chunks, half 1 on even ones
            const uint64_t a_hi = make_sw128_kmajor_desc(sa0);
            const uint64_t a_lo = make_sw128_kmajor_desc(sa0 + a_slot);
            const uint64_t b_hi = make_sw128_kmajor_desc(sa0 + 2u * a_slot);
            const uint64_t b_lo = make_sw128_kmajor_desc(sa0 + 2u * a_slot + b_bytes);
            if (u == 0) {  // the closing half also starts here (a one-chunk accumulation)
              mbar_wait(&main_empty[hc], ((mph >> hc) & 1u) ^ 1u, 4);
              mph ^= 1u << hc;
              tc_fence_after();
            }
            if (elected) {
              const uint64_t b_half = b_hi + static_cast<uint64_t>(hc * (8192 >> 4));
#pragma unroll
              for (int k = 0; k < kTcBlockK / 16; ++k)
                umma_bf16_pair(d_main + static_cast<uint32_t>(hc) * 128u, a_hi + static_cast<uint64_t>(2 * k),
                               b_half + static_cast<uint64_t>(2 * k), idesc_half, (k > 0 || u > 0) ? 1u : 0u);
              umma_commit_pair(&main_full[hc]);
#pragma unroll
              for (int k = 0; k < kTcBlockK / 16; ++k)
                umma_bf16_pair(d_corr, a_lo + static_cast<uint64_t>(2 * k), b_hi + static_cast<uint64_t>(2 * k), idesc,
                               (kb > 0 || k > 0) ? 1u : 0u);
#pragma unroll
              for (int k = 0; k < kTcBlockK / 16; ++k)
                umma_bf16_pair(d_corr, a_hi + static_cast<uint64_t>(2 * k), b_lo + static_cast<uint64_t>(2 * k), idesc, 1u);
            }
            const int hs = hc ^ 1;  // starts a new accumulation: its previous one must have been drained
            mbar_wait(&main_empty[hs], ((mph >> hs) & 1u) ^ 1u, 4);
            mph ^= 1u << hs;
            tc_fence_after();
            if (elected) {
              const uint64_t b_half = b_hi + static_cast<uint64_t>(hs * (8192 >> 4));
#pragma unroll
              for (int k = 0; k < kTcBlockK / 16; ++k)
                umma_bf16_pair(d_main + static_cast<uint32_t>(hs) * 128u, a_hi + static_cast<uint64_t>(2 * k),
                               b_half + static_cast<uint64_t>(2 * k), idesc_half, k > 0 ? 1u : 0u);
              umma_commit_pair(&empty_bar[stage]);
              if (u == total_chunks - 1) umma_commit_pair(&main_full[hs]);
            }
          } else
          for (int j = 0; j < nsub; ++j) {
          // sub-block j: the rows of tap dy = j start j * tap_bytes into the halo boxes; its weight blocks are (hi, lo) pair j
          const uint32_t sa = sa0 + static_cast<uint32_t>(j) * tap_bytes;
          const uint32_t sb = sa0 + 2u * a_slot + static_cast<uint32_t>(2 * j) * b_bytes;
          const uint64_t a_hi = make_sw128_kmajor_desc(sa);
          const uint64_t a_lo = make_sw128_kmajor_desc(sa + a_slot);
          const uint64_t b_hi = make_sw128_kmajor_desc(sb);
          const uint64_t b_lo = make_sw128_kmajor_desc(sb + b_bytes);
          if (alt) {
            // tiles of <= 128 columns: an MMA is shorter than the latency of a dependent accumulate, so the two chains
            // (corr, main) are issued interleaved - same order inside each chain, i.e. the same sums. The buffer the
            // main chain writes was handed back two accumulations ago.
            const bool fresh = unit == 0 || ((unit + ph) & kmask) == 0;
            const bool closes = ((unit + ph) & kmask) == kmask || unit == total_chunks - 1;
            ++unit;
            X3_TS(gstage, 2);
            if (fresh) {
              mbar_wait(&main_empty[mb], ((mph >> mb) & 1u) ^ 1u, 4);
              mph ^= 1u << mb;
              tc_fence_after();
            }
            X3_TS(gstage, 3);
            if (elected) {
#pragma unroll
              for (int k = 0; k < kTcBlockK / 16; ++k) {
                umma_bf16_pair(d_corr, a_lo + static_cast<uint64_t>(2 * k), b_hi + static_cast<uint64_t>(2 * k), idesc,
                               (kb > 0 || j > 0 || k > 0) ? 1u : 0u);
                umma_bf16_pair(d_main + static_cast<uint32_t>(mb) * 128u, a_hi + static_cast<uint64_t>(2 * k),
                               b_hi + static_cast<uint64_t>(2 * k), idesc, (k > 0 || !fresh) ? 1u : 0u);
              }
              if (closes) umma_commit_pair(&main_full[mb]);
#pragma unroll
              for (int k = 0; k < kTcBlockK / 16; ++k)
                umma_bf16_pair(d_corr, a_hi + static_cast<uint64_t>(2 * k), b_lo + static_cast<uint64_t>(2 * k), idesc, 1u);
              if (j == nsub - 1) umma_commit_pair(&empty_bar[stage]);
            }
            if (closes) mb ^= 1;
            continue;
          }
          if (elected) {
#pragma unroll
            for (int k = 0; k < kTcBlockK / 16; ++k)
              umma_bf16_pair(d_corr, a_lo + static_cast<uint64_t>(2 * k), b_hi + static_cast<uint64_t>(2 * k), idesc,
                             (kb > 0 || j > 0 || k > 0) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < kTcBlockK / 16; ++k)
              umma_bf16_pair(d_corr, a_hi + static_cast<uint64_t>(2 * k), b_lo + static_cast<uint64_t>(2 * k), idesc, 1u);
          }
          // a "main" buffer is free again once the epilogue warps have drained its previous contents - they do that
          // while the corr MMAs above keep the tensor pipe busy
          const bool fresh = unit == 0 || ((unit + ph) & kmask) == 0;                     // first K chunk of an accumulation
          const bool closes = ((unit + ph) & kmask) == kmask || unit == total_chunks - 1;  // last one: hand it to the epilogue
          ++unit;
          if (wide) {
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
              if (hb == 0) X3_TS(gstage, 2);
              if (fresh) {
                mbar_wait(&main_empty[hb], ((mph >> hb) & 1u) ^ 1u, 4);
                mph ^= 1u << hb;
                tc_fence_after();
              }
              X3_TS(gstage, 3 + hb);
              // rows hb*64 .. hb*64+63 of each CTA's 128-row weight tile: 8 swizzle atoms = 8192 B further
              const uint64_t b_half = b_hi + static_cast<uint64_t>(hb * (8192 >> 4));
              if (elected) {
#pragma unroll
                for (int k = 0; k < kTcBlockK / 16; ++k)
                  umma_bf16_pair(d_main + static_cast<uint32_t>(hb) * 128u, a_hi + static_cast<uint64_t>(2 * k),
                                 b_half + static_cast<uint64_t>(2 * k), idesc_half, (k > 0 || !fresh) ? 1u : 0u);
                if (hb == 1 && j == nsub - 1) umma_commit_pair(&empty_bar[stage]);  // frees the slot in both CTAs once these MMAs have read it
                if (closes) umma_commit_pair(&main_full[hb]);      // this half of the hi.hi accumulation is complete (both CTAs)
              }
            }
          } else {
            X3_TS(gstage, 2);
            if (fresh) {
              mbar_wait(&main_empty[mb], ((mph >> mb) & 1u) ^ 1u, 4);
              mph ^= 1u << mb;
              tc_fence_after();
            }
            X3_TS(gstage, 3);
            if (elected) {
#pragma unroll
              for (int k = 0; k < kTcBlockK / 16; ++k)
                umma_bf16_pair(d_main + static_cast<uint32_t>(mb) * 128u, a_hi + static_cast<uint64_t>(2 * k),
                               b_hi + static_cast<uint64_t>(2 * k), idesc, (k > 0 || !fresh) ? 1u : 0u);
              if (j == nsub - 1) umma_commit_pair(&empty_bar[stage]);
              if (closes) umma_commit_pair(&main_full[mb]);
            }
            if (closes && alt) mb ^= 1;
          }
          }  // sub-blocks
          ++gstage;
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        if (elected) umma_commit_pair(corr_full);
      }
    }
  }
  } else {
    // ===================================== epilogue (8 warps per CTA) =======================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // warp (q, half): TMEM lane quarter q (rows q*32..), column half `half`
    //   wide tiles: this warp owns "main" buffer `half` (128 TMEM columns = tile columns {half*64.., 128 + half*64..})
    //   else:       columns [half * half_cols, (half + 1) * half_cols) of both buffers
    const int q = warp & 3;
    const int half = (warp - kX3FirstEpiWarp) >> 2;
    const int row = q * 32 + lane;
    const int half_cols = block_n >> 1;  // multiple of 32
    uint32_t cph = 0, mph = 0;  // mph: one parity bit per "main" buffer
    int mb = wide ? half : 0;
    int gchunk = 0;  // running drain index of this CTA (profiling stamps)
    uint8_t* stg = stage_base + static_cast<size_t>(stages) * stage_bytes + static_cast<size_t>(warp - kX3FirstEpiWarp) * 2048;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t t_main = t_lane + (wide ? static_cast<uint32_t>(half) * 128u : static_cast<uint32_t>(half * half_cols));
    for (int tile = worker; tile < total_tiles; tile += num_workers) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = ((tile / p.n_tiles) << 1) + static_cast<int>(cta_rank);
      const int iw = m_tile % p.tiles_w;
      const int ih = (m_tile / p.tiles_w) % p.tiles_h;
      const int in = m_tile / (p.tiles_w * p.tiles_h);
      const int ww = row % p.tw;
      const int hh = (row / p.tw) % p.th;
      const int nn = row / (p.tw * p.th);
      const int w = iw * p.tw + ww, h = ih * p.th + hh, n = in * p.tn + nn;
      const bool valid = (w < p.w_out) && (h < p.h_out) && (n < p.n_img);
      const size_t pix = (static_cast<size_t>(n) * p.h_out + h) * p.w_out + w;
      // transposed store mapping: instruction k covers rows 8k + (lane >> 2)
      size_t pixk[4] = {0, 0, 0, 0};
      bool vk[4] = {false, false, false, false};
      if (kX3TransposedTail) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          pixk[k] = __shfl_sync(0xffffffffu, static_cast<unsigned long long>(pix), 8 * k + (lane >> 2));
          vk[k] = __shfl_sync(0xffffffffu, valid ? 1 : 0, 8 * k + (lane >> 2)) != 0;
        }
      }
      float2 racc[64];  // register accumulator group g (32 columns) = racc[16 g .. 16 g + 15]
#pragma unroll
      for (int i = 0; i < 64; ++i) racc[i] = make_float2(0.f, 0.f);
      // staggered halves (see the issuer): half 0 closes on odd chunks and on the last one, half 1 on even chunks and on the last one
      const int drains = stagger ? (half == 0 ? (total_chunks + 1) / 2 : (total_chunks + 1) / 2 + ((total_chunks & 1) ? 0 : 1))
                                 : (total_chunks + (odd_groups ? (n_tile & 1) : 0) + kmask) >> kmask;
      const bool ts_warp = lane == 0 && q == 0 && (wide || half == 0);  // profiling: one lane of the first warp of each column half
      for (int ch = 0; ch <= drains; ++ch) {
        // ch < drains: one hi.hi accumulation (kchain K chunks) from "main"; ch == drains: the tile's lo.hi + hi.lo sum from "corr"
        const bool last = ch == drains;
        if (last) { mbar_wait(corr_full, cph, 5); cph ^= 1u; }
        else { mbar_wait(&main_full[mb], (mph >> mb) & 1u, 6); mph ^= 1u << mb; }
        if (ts_warp && !last) X3_TS(gchunk, 5 + half);
        tc_fence_after();
        // two TMEM loads (64 columns) in flight per wait
#pragma unroll
        for (int g = 0; g < 4; g += 2) {
          if (g * 32 < half_cols) {
            uint32_t v0[32], v1[32];
            const bool two = (g + 1) * 32 < half_cols;
            // tile column of register group g: wide = {half*64 + 0, +32, 128 + half*64 + 0, +32}, else half*half_cols + 32 g
            uint32_t a0;
            if (last) {
              const uint32_t c0 = wide ? static_cast<uint32_t>((g >> 1) * 128 + half * 64) : static_cast<uint32_t>(half * half_cols + g * 32);
              a0 = t_lane + static_cast<uint32_t>(kX3CorrCol) + c0;
            } else {
              a0 = t_main + (wide ? 0u : static_cast<uint32_t>(mb) * 128u) + static_cast<uint32_t>(g * 32);
            }
            tmem_ld_32x32b_x32(a0, v0);
            if (two) tmem_ld_32x32b_x32(a0 + 32u, v1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              racc[g * 16 + i] = __fadd2_rn(racc[g * 16 + i], make_float2(__uint_as_float(v0[2 * i]), __uint_as_float(v0[2 * i + 1])));
            if (two) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                racc[g * 16 + 16 + i] = __fadd2_rn(racc[g * 16 + 16 + i], make_float2(__uint_as_float(v1[2 * i]), __uint_as_float(v1[2 * i + 1])));
            }
          }
        }
        // hand the buffer back: one arrive per warp (to the leader's barrier) after every lane's reads have landed
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          uint64_t* bar = last ? corr_empty : &main_empty[mb];
          if (leader) mbar_arrive(bar); else mbar_arrive_remote(bar, 0u);
        }
        if (ts_warp && !last && half == 0) X3_TS(gchunk, 7);
        if (!last) ++gchunk;
        if (!last && alt) mb ^= 1;
      }
      // truncation compensation (see kX3Comp): x + x * comp in one rounding - 1 + comp itself is not an fp32 number
      {
        const float2 comp2 = make_float2(p.comp, p.comp);
#pragma unroll
        for (int i = 0; i < 64; ++i) racc[i] = __ffma2_rn(racc[i], comp2, racc[i]);
      }
      if (p.pool) {
        // fused 2x2/2 max pool (halo mode): the window of a pixel lives in lanes {l, l^1, l^tw, l^tw^1} of this warp; the
        // bias add and ReLU commute with the max, so they run once, on the pooled value. Lane (ww & 1, hh & 1) of a window
        // stores 8 of every 32 channels.
        const int sub = (ww & 1) | ((hh & 1) << 1);
        const size_t ppix = (static_cast<size_t>(n) * (p.h_out >> 1) + (h >> 1)) * (p.w_out >> 1) + (w >> 1);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int tcol = wide ? (g >> 1) * 128 + half * 64 + (g & 1) * 32 : half * half_cols + g * 32;
          const int col0 = n_tile * block_n + tcol;
          if (g * 32 < half_cols) {
            float g8[8];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float x = (i & 1) ? racc[g * 16 + (i >> 1)].y : racc[g * 16 + (i >> 1)].x;
              x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 1));
              x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, p.tw));
              if ((i >> 3) == sub) g8[i & 7] = x;
            }
            if (valid && col0 + 8 * sub < p.n_valid) x3_store8_split(p, g8, ppix, col0 + 8 * sub);
          }
        }
      } else
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int tcol = wide ? (g >> 1) * 128 + half * 64 + (g & 1) * 32 : half * half_cols + g * 32;
        const int col0 = n_tile * block_n + tcol;
        if (g * 32 < half_cols && col0 < p.n_valid) {  // warp-uniform: x3_finish32 is a warp-collective
          float f[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) { f[2 * i] = racc[g * 16 + i].x; f[2 * i + 1] = racc[g * 16 + i].y; }
          if (kX3TransposedTail) x3_finish32(p, f, stg, lane, pix, valid, pixk, vk, col0);
          else if (valid) igemm_finish32(p, f, pix, col0);
        }
      }
    }
  }

#undef X3_TS
  tc_fence_before();
  cluster_sync_all();  // the peer's smem/TMEM stay live until the leader's MMAs are done
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTcTmemCols);
  }
}

inline float igemm_x3_default_comp(int kchain) {
  const char* e = getenv("OCRFA_X3_COMP");  // read per launch: tools/error_budget.py scans it inside one process
  return e ? static_cast<float>(atof(e)) : (kchain == 2 ? kX3Comp2 : kX3Comp);
}

inline cudaError_t igemm_x3_launch(const IgemmTcParams& p_in, int num_sms, cudaStream_t stream) {
  IgemmTcParams p = p_in;
  if (p.comp == 0.0f) p.comp = igemm_x3_default_comp(p.kchain);
  static bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[current_device_slot()];
  int stages = 0;
  const size_t smem = p.halo ? igemm_x3_halo_smem_bytes(p.block_n, p.tw, p.th, &stages) : igemm_x3_smem_bytes(p.block_n, &stages);
  if (!p.split || !p.pair || stages != p.num_stages || stages < 2) return cudaErrorInvalidValue;
  if (p.halo && (p.tn != 1 || (p.tw & 7) != 0 || p.tw * p.th != kTcBlockM)) return cudaErrorInvalidValue;
  if (p.pool && (!p.halo || p.residual != nullptr || p.out_f32 || p.tw > 16 || (p.h_out & 1) || (p.w_out & 1) || (p.n_valid & 31)))
    return cudaErrorInvalidValue;
  if ((p.block_n & 63) != 0 || p.block_n > 256) return cudaErrorInvalidValue;  // the epilogue halves the tile in 32-column groups
  if (p.kchain == 2 && p.block_n == 128 && p.n_valid >= 256) return cudaErrorInvalidValue;  // see `odd_groups` in the kernel
  int total_kb = 0;
  for (int t = 0; t < p.num_taps; ++t) total_kb += p.taps[t].src ? p.kchunks2 : p.kchunks;
  if (p.halo) total_kb = 3 * p.kchunks + p.kchunks2;
  if (total_kb > (kTcSmemHeader - 256) / 48) return cudaErrorInvalidValue;  // group schedule: 48 B per group
  const int kind = p.block_n == 256 ? kX3Wide : p.block_n <= 128 ? (p.halo ? kX3NarrowHalo : kX3Narrow) : kX3Other;
  if (p.halo && kind != kX3NarrowHalo) return cudaErrorInvalidValue;
  void (*fn)(IgemmTcParams) = kind == kX3Wide ? igemm_x3_kernel<kX3Wide> : kind == kX3Narrow ? igemm_x3_kernel<kX3Narrow>
                              : kind == kX3NarrowHalo ? igemm_x3_kernel<kX3NarrowHalo> : igemm_x3_kernel<kX3Other>;
  if (!attr_set) {
    for (void (*f)(IgemmTcParams) : {igemm_x3_kernel<kX3Wide>, igemm_x3_kernel<kX3Narrow>, igemm_x3_kernel<kX3NarrowHalo>,
                                     igemm_x3_kernel<kX3Other>}) {
      cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return e;
    }
    attr_set = true;
  }
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int pair_tiles = ((m_tiles + 1) / 2) * p.n_tiles;
  const int pairs = pair_tiles < num_sms / 2 ? pair_tiles : num_sms / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kX3Threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, fn, p);
}

}  // namespace ocrfa
