// tcgen05 implicit GEMM for the 3x3 / stride 1 / pad 1 convolutions of the ResNet front (conv0_2, layer1, conv1,
// layer2; reference: /root/reference/src/ocrfa/modules/blocks.py:200-223, 248-290), halo-staged.
//
// Why a second kernel: igemm_tc_kernel fetches one A box per filter tap, i.e. every input pixel crosses L2 -> SM
// nine times, and one B box per K block per tile. At Cout <= 128 that traffic, not the tensor pipe, sets the pace
// (measured with ncu on the c2 step: conv0_2 21 %, layer1 34-48 % tensor-pipe active while L2 -> SM moves
// ~10-14 TB/s, the LTS cap). This kernel cuts both:
//   * A: ONE box of (th + 2) x tw pixels per (dx, 64-channel chunk) feeds the three dy taps. The M tile is
//     tw x th x 1 pixels with tw a multiple of 8, so the rows of tap dy start dy * tw * 128 B into the box -
//     a multiple of the 1024-byte swizzle atom, hence a plain descriptor start-address offset. A traffic per tile
//     drops from 9 to 3 * (th + 2) / th boxes.
//   * B: when the whole [Cout][K] weight matrix fits beside >= 3 A stages (every Cout <= 128 layer of the path:
//     each CTA of the pair keeps its Cout/2 rows, <= 152 KB) it is loaded ONCE per CTA and stays resident; the
//     per-tile weight traffic disappears. Otherwise the three weight blocks of a group travel with its A box.
//   * optional fused 2x2/2 max pool (max_pool1 after conv0_2, max_pool2 after conv1; blocks.py:259, 271-275) in
//     the epilogue: the 2x2 window of a pixel lives in lanes {l, l^1, l^tw, l^tw^1} of one warp, so two shuffles
//     per value pool it and the conv output never reaches HBM at full resolution.
// Always a CTA pair (cta_group::2, M = 256 over two SMs), TMEM double buffering as igemm_tc; 8 epilogue warps.
#pragma once
#include "igemm_tc.cuh"

namespace ocrfa {

constexpr int kHaloThreads = 320;  // TMA producer + MMA issuer + 8 epilogue warps
constexpr int kHaloTable = 2048;    // barriers (256 B) + group schedule (56 groups)
constexpr int kHaloStaging = 8 * 2048;  // per epilogue warp: 32 rows x 64 B transposition buffer
constexpr int kHaloHeader = kHaloTable + kHaloStaging;  // then the 1024-aligned weight / stage regions

struct alignas(64) IgemmHaloParams {
  CUtensorMap tmA;   // 5-D (C, W, 1, H, N) bf16, box {64, tw, 1, th + 2, 1}, SWIZZLE_128B
  CUtensorMap tmA2;  // fused 1x1 downsample source (C2, W, 1, H, N), box {64, tw, 1, th, 1}
  CUtensorMap tmB;   // 2-D (Ktotal, Cout) bf16, box {64, block_n / 2}
  int kchunks;       // 64-channel chunks of the 3x3 source
  int kchunks2;      // ... of the fused 1x1 source (0: none)
  int n_img, h_out, w_out;
  int tw, th;        // M tile = tw x th pixels of one image (tw * th = 128, tw % 8 = 0)
  int tiles_w, tiles_h;
  int n_tiles;       // tiles along Cout
  int block_n;       // UMMA N (multiple of 32, <= 256)
  int n_valid;       // real Cout (multiple of 32)
  int num_stages;
  int b_resident;    // 1: whole weight matrix resident in shared memory (requires n_tiles == 1)
  int relu;
  int pool;          // 1: fused 2x2 stride-2 max pool, output is [N][H/2][W/2][ldo]
  int f16;           // 1: operands / outputs are fp16 instead of bf16
  int ldo, ldr;
  const float* bias;
  const void* residual;  // NHWC bf16 or nullptr (not with pool)
  void* out;             // NHWC bf16
};

__device__ __forceinline__ void halo_store8(const IgemmHaloParams& p, float (&g)[8], size_t pix, int col) {
  if (p.bias != nullptr) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
    g[0] += b0.x; g[1] += b0.y; g[2] += b0.z; g[3] += b0.w;
    g[4] += b1.x; g[5] += b1.y; g[6] += b1.z; g[7] += b1.w;
  }
  if (p.residual != nullptr) {
    const uint4 r4 = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.residual) + pix * p.ldr + col));
    const bool h = p.f16 != 0;
    g[0] += lo16_to_f32(r4.x, h); g[1] += hi16_to_f32(r4.x, h);
    g[2] += lo16_to_f32(r4.y, h); g[3] += hi16_to_f32(r4.y, h);
    g[4] += lo16_to_f32(r4.z, h); g[5] += hi16_to_f32(r4.z, h);
    g[6] += lo16_to_f32(r4.w, h); g[7] += hi16_to_f32(r4.w, h);
  }
  if (p.relu) {
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = fmaxf(g[i], 0.0f);
  }
  uint4 o;
  const bool h16 = p.f16 != 0;
  o.x = pack16x2(g[0], g[1], h16); o.y = pack16x2(g[2], g[3], h16);
  o.z = pack16x2(g[4], g[5], h16); o.w = pack16x2(g[6], g[7], h16);
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.ldo + col) = o;
}

__global__ void __launch_bounds__(kHaloThreads, 1) igemm_halo_kernel(const __grid_constant__ IgemmHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);  // [kTcMaxStages]
  uint64_t* empty_bar = full_bar + kTcMaxStages;           // [kTcMaxStages]
  uint64_t* tfull_bar = empty_bar + kTcMaxStages;          // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                    // [2]
  uint64_t* bres_bar = tempty_bar + 2;                     // [1] resident weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 1);
  int4* grp_table = reinterpret_cast<int4*>(smem + 256);   // [groups][2]: {a_c, dw, src, n_sub}, {kb0, kb1, kb2, a_bytes}

  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.num_stages;
  const int block_n = p.block_n;
  const int b_rows = block_n >> 1;  // weight rows in THIS CTA's shared memory
  const uint32_t b_bytes = static_cast<uint32_t>(b_rows) * kTcBlockK * 2;
  const uint32_t tap_bytes = static_cast<uint32_t>(p.tw) * 128u;         // one image row of the box
  const uint32_t a_halo_bytes = static_cast<uint32_t>(p.th + 2) * tap_bytes;
  const uint32_t a_flat_bytes = static_cast<uint32_t>(p.th) * tap_bytes;  // = 16 KB
  const int groups = 3 * p.kchunks + p.kchunks2;
  const int total_kb = 9 * p.kchunks + p.kchunks2;
  const bool resident = p.b_resident != 0;
  const uint32_t stage_bytes = a_halo_bytes + (resident ? 0u : 3u * b_bytes);
  uint8_t* stg_base = smem + kHaloTable;
  uint8_t* bres_base = smem + kHaloHeader;
  uint8_t* stage_base = bres_base + (resident ? static_cast<size_t>(total_kb) * b_bytes : 0);

  const int m_tiles = p.tiles_w * p.tiles_h * p.n_img;
  const int total_tiles = ((m_tiles + 1) >> 1) * p.n_tiles;
  const int worker = static_cast<int>(blockIdx.x >> 1);
  const int num_workers = static_cast<int>(gridDim.x >> 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 2);  // the leader's expect_tx arrive + the peer's remote arrive
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 512);  // both CTAs' 8 epilogue warps
    }
    mbar_init(bres_bar, 2);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmA2);
  }
  // group schedule: dx-major, chunk-minor; the three sub-blocks of a group are the dy taps. Weight K order in
  // memory is [tap = (dy+1)*3 + (dx+1)][cin_pad] followed by the fused 1x1 source's chunks.
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    if (g < 3 * p.kchunks) {
      const int dxi = g / p.kchunks, c = g - dxi * p.kchunks;
      grp_table[2 * g] = make_int4(c * kTcBlockK, dxi - 1, 0, 3);
      grp_table[2 * g + 1] = make_int4((0 * 3 + dxi) * p.kchunks + c, (1 * 3 + dxi) * p.kchunks + c,
                                       (2 * 3 + dxi) * p.kchunks + c, static_cast<int>(a_halo_bytes));
    } else {
      const int c = g - 3 * p.kchunks;
      grp_table[2 * g] = make_int4(c * kTcBlockK, 0, 1, 1);
      grp_table[2 * g + 1] = make_int4(9 * p.kchunks + c, 0, 0, static_cast<int>(a_flat_bytes));
    }
  }
  if (warp == 1) { tmem_alloc_pair(tmem_slot, kTcTmemCols); tmem_relinquish_pair(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      if (resident) {
        // the whole weight matrix, once: this CTA's b_rows rows of every K block (both CTAs complete on the leader)
        if (leader) mbar_expect_tx(bres_bar, 2u * static_cast<uint32_t>(total_kb) * b_bytes);
        for (int kb = 0; kb < total_kb; ++kb)
          tma_load_2d_pair(bres_base + static_cast<size_t>(kb) * b_bytes, &p.tmB, bres_bar, kb * kTcBlockK,
                           static_cast<int>(cta_rank) * b_rows);
        if (!leader) mbar_arrive_remote(bres_bar, 0u);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = worker; tile < total_tiles; tile += num_workers) {
        const int n_tile = tile % p.n_tiles;
        // an odd tail leaves the peer with m_tile == m_tiles: its image index is out of range, TMA zero-fills
        const int m_tile = ((tile / p.n_tiles) << 1) + static_cast<int>(cta_rank);
        const int iw = m_tile % p.tiles_w;
        const int ih = (m_tile / p.tiles_w) % p.tiles_h;
        const int in = m_tile / (p.tiles_w * p.tiles_h);
        const int w0 = iw * p.tw, h0 = ih * p.th;
        const int b_col = n_tile * block_n + static_cast<int>(cta_rank) * b_rows;
        for (int g = 0; g < groups; ++g) {
          const int4 e0 = grp_table[2 * g];
          const int4 e1 = grp_table[2 * g + 1];
          mbar_wait(&empty_bar[stage], phase ^ 1u, 1);
          uint8_t* sa = stage_base + static_cast<size_t>(stage) * stage_bytes;
          const uint32_t bytes = static_cast<uint32_t>(e1.w) + (resident ? 0u : static_cast<uint32_t>(e0.w) * b_bytes);
          if (leader) mbar_expect_tx(&full_bar[stage], 2u * bytes);
          if (e0.z) tma_load_5d_pair(sa, &p.tmA2, &full_bar[stage], e0.x, w0, 0, h0, in);
          else tma_load_5d_pair(sa, &p.tmA, &full_bar[stage], e0.x, w0 + e0.y, 0, h0 - 1, in);
          if (!resident) {
            tma_load_2d_pair(sa + a_halo_bytes, &p.tmB, &full_bar[stage], e1.x * kTcBlockK, b_col);
            if (e0.w == 3) {
              tma_load_2d_pair(sa + a_halo_bytes + b_bytes, &p.tmB, &full_bar[stage], e1.y * kTcBlockK, b_col);
              tma_load_2d_pair(sa + a_halo_bytes + 2u * b_bytes, &p.tmB, &full_bar[stage], e1.z * kTcBlockK, b_col);
            }
          }
          if (!leader) mbar_arrive_remote(&full_bar[stage], 0u);
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) =====================
    if (leader && elect_one()) {
      const uint32_t idesc = p.f16 ? make_idesc_f16(2 * kTcBlockM, static_cast<uint32_t>(block_n))
                                   : make_idesc_bf16(2 * kTcBlockM, static_cast<uint32_t>(block_n));
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      if (resident) { mbar_wait(bres_bar, 0u, 5); tc_fence_after(); }
      const uint32_t bres_addr = smem_u32(bres_base);
      for (int tile = worker; tile < total_tiles; tile += num_workers) {
        mbar_wait(&tempty_bar[as], aphase ^ 1u, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as) * 256u;
        uint32_t acc = 0u;
        for (int g = 0; g < groups; ++g) {
          const int4 e0 = grp_table[2 * g];
          const int4 e1 = grp_table[2 * g + 1];
          mbar_wait(&full_bar[stage], phase, 3);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + static_cast<size_t>(stage) * stage_bytes);
          const int kbs[3] = {e1.x, e1.y, e1.z};
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            if (j < e0.w) {
              const uint64_t a_desc = make_sw128_kmajor_desc(sa + static_cast<uint32_t>(j) * tap_bytes);
              const uint64_t b_desc = make_sw128_kmajor_desc(
                  resident ? bres_addr + static_cast<uint32_t>(kbs[j]) * b_bytes : sa + a_halo_bytes + static_cast<uint32_t>(j) * b_bytes);
#pragma unroll
              for (int k = 0; k < kTcBlockK / 16; ++k) {
                umma_bf16_pair(d_tmem, a_desc + static_cast<uint64_t>(2 * k), b_desc + static_cast<uint64_t>(2 * k), idesc, acc);
                acc = 1u;
              }
            }
          }
          umma_commit_pair(&empty_bar[stage]);  // frees the slot in both CTAs once these MMAs have read it
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_pair(&tfull_bar[as]);
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ===================================== epilogue (8 warps) ================================
    // warp (q, half): TMEM lane quarter q (rows q*32..), column half `half` of the tile; the TMEM read of chunk
    // i+1 is in flight while chunk i goes through pool / bias / ReLU / store (small-K layers are epilogue-paced).
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int ww = row % p.tw;
    const int hh = row / p.tw;
    const int ncol = block_n >> 1;  // multiple of 16; chunks of 32 columns, the last may be half valid
    int as = 0;
    uint32_t aphase = 0;
    // pooled output: lane (ww & 1, hh & 1) of a 2x2 window stores 8 of every 32 channels
    const int sub = (ww & 1) | ((hh & 1) << 1);
    uint8_t* stg = stg_base + (warp - 2) * 2048;  // this warp's transposition buffer, pieces XOR-swizzled by (row >> 1) & 3
    const int swz_own = (lane >> 1) & 3;
    for (int tile = worker; tile < total_tiles; tile += num_workers) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = ((tile / p.n_tiles) << 1) + static_cast<int>(cta_rank);
      const int iw = m_tile % p.tiles_w;
      const int ih = (m_tile / p.tiles_w) % p.tiles_h;
      const int n = m_tile / (p.tiles_w * p.tiles_h);
      const int w = iw * p.tw + ww, h = ih * p.th + hh;
      const bool valid = (w < p.w_out) && (h < p.h_out) && (n < p.n_img);
      const int pix = p.pool ? (n * (p.h_out >> 1) + (h >> 1)) * (p.w_out >> 1) + (w >> 1) : (n * p.h_out + h) * p.w_out + w;
      const int n_off = n_tile * block_n + half * ncol;
      // transposed store mapping: instruction k covers rows 8k + (lane >> 2), 16-byte piece (lane & 3) of each
      int pixk[4];
      bool vk[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        pixk[k] = __shfl_sync(0xffffffffu, pix, 8 * k + (lane >> 2));
        vk[k] = __shfl_sync(0xffffffffu, valid ? 1 : 0, 8 * k + (lane >> 2)) != 0;
      }

      mbar_wait(&tfull_bar[as], aphase, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as) * 256u +
                             static_cast<uint32_t>(half * ncol);
      auto finish = [&](uint32_t (&v)[32], int c0) {
        const int col0 = n_off + c0;
        const int lim = ncol - c0;  // valid columns of this chunk (>= 16)
        if (p.pool) {
          float g[8];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float x = __uint_as_float(v[i]);
            x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 1));
            x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, p.tw));
            if ((i >> 3) == sub) g[i & 7] = x;
          }
          if (valid && 8 * sub < lim) halo_store8(p, g, static_cast<size_t>(pix), col0 + 8 * sub);
        } else {
          // bias (+ residual) + ReLU in the thread = pixel layout, then a transposition through shared memory so
          // that every global access covers 64 contiguous bytes of 8 pixels (a 16-byte access per pixel costs one
          // L1 wavefront per lane: measured, the stores - not the MMAs - paced the Cout <= 128 layers)
          float f[32];
          const bool h16 = p.f16 != 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          const int pc = lane & 3;                      // piece this lane moves in the transposed mapping
          const uint32_t t_off = static_cast<uint32_t>(lane >> 2) * 64u + static_cast<uint32_t>((pc ^ ((lane >> 3) & 3)) << 4);
          const bool pc_ok = pc * 8 < lim;
          if (p.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              if (i < lim) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
                f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
              }
            }
          }
          if (p.residual != nullptr) {
            const __nv_bfloat16* rbase = reinterpret_cast<const __nv_bfloat16*>(p.residual) + col0 + pc * 8;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              uint4 r4 = make_uint4(0u, 0u, 0u, 0u);
              if (vk[k] && pc_ok) r4 = __ldg(reinterpret_cast<const uint4*>(rbase + static_cast<size_t>(pixk[k]) * p.ldr));
              *reinterpret_cast<uint4*>(stg + k * 512 + t_off) = r4;
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 r4 = *reinterpret_cast<const uint4*>(stg + lane * 64 + ((j ^ swz_own) << 4));
              f[8 * j] += lo16_to_f32(r4.x, h16);     f[8 * j + 1] += hi16_to_f32(r4.x, h16);
              f[8 * j + 2] += lo16_to_f32(r4.y, h16); f[8 * j + 3] += hi16_to_f32(r4.y, h16);
              f[8 * j + 4] += lo16_to_f32(r4.z, h16); f[8 * j + 5] += hi16_to_f32(r4.z, h16);
              f[8 * j + 6] += lo16_to_f32(r4.w, h16); f[8 * j + 7] += hi16_to_f32(r4.w, h16);
            }
            __syncwarp();
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.0f);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack16x2(f[8 * j], f[8 * j + 1], h16);     o.y = pack16x2(f[8 * j + 2], f[8 * j + 3], h16);
            o.z = pack16x2(f[8 * j + 4], f[8 * j + 5], h16); o.w = pack16x2(f[8 * j + 6], f[8 * j + 7], h16);
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((j ^ swz_own) << 4)) = o;
          }
          __syncwarp();
          __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(p.out) + col0 + pc * 8;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint4 o = *reinterpret_cast<const uint4*>(stg + k * 512 + t_off);
            if (vk[k] && pc_ok) *reinterpret_cast<uint4*>(obase + static_cast<size_t>(pixk[k]) * p.ldo) = o;
          }
          __syncwarp();
        }
      };
      auto release = [&]() {
        // last read of this accumulator: hand the TMEM buffer back (to the leader's barrier) before the stores
        tc_fence_before();
        if (!leader) mbar_arrive_remote(&tempty_bar[as], 0u);
        else mbar_arrive(&tempty_bar[as]);
      };
      uint32_t va[32], vb[32];
      tmem_ld_32x32b_x32(taddr, va);
      for (int c0 = 0; c0 < ncol; c0 += 64) {
        tmem_ld_wait();
        if (c0 + 32 < ncol) tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c0 + 32), vb); else release();
        finish(va, c0);
        if (c0 + 32 < ncol) {
          tmem_ld_wait();
          if (c0 + 64 < ncol) tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c0 + 64), va); else release();
          finish(vb, c0 + 32);
        }
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the peer's smem/TMEM stay live until the leader's MMAs are done
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTcTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// Shared-memory plan: resident weights when they leave >= 3 A stages, else streamed with the A boxes.
inline size_t igemm_halo_smem_bytes(int block_n, int tw, int th, int total_kb, int n_tiles, int* stages_out, int* resident_out) {
  const size_t header = 1024 + kHaloHeader;
  const size_t budget = 227 * 1024 - header;
  const size_t b_bytes = static_cast<size_t>(block_n / 2) * kTcBlockK * 2;
  const size_t a_halo = static_cast<size_t>(th + 2) * tw * 128;
  const size_t res_bytes = static_cast<size_t>(total_kb) * b_bytes;
  int resident = 0, stages = 0;
  if (n_tiles == 1 && res_bytes + 3 * a_halo <= budget) {
    resident = 1;
    stages = static_cast<int>((budget - res_bytes) / a_halo);
  } else {
    stages = static_cast<int>(budget / (a_halo + 3 * b_bytes));
  }
  if (stages > kTcMaxStages) stages = kTcMaxStages;
  if (stages_out) *stages_out = stages;
  if (resident_out) *resident_out = resident;
  return header + (resident ? res_bytes : 0) + stages * (a_halo + (resident ? 0 : 3 * b_bytes));
}

inline cudaError_t igemm_halo_launch(const IgemmHaloParams& p, int num_sms, cudaStream_t stream) {
  static bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[current_device_slot()];
  const int total_kb = 9 * p.kchunks + p.kchunks2;
  const size_t b_bytes = static_cast<size_t>(p.block_n / 2) * kTcBlockK * 2;
  const size_t a_halo = static_cast<size_t>(p.th + 2) * p.tw * 128;
  const size_t smem = 1024 + kHaloHeader + (p.b_resident ? total_kb * b_bytes : 0) +
                      static_cast<size_t>(p.num_stages) * (a_halo + (p.b_resident ? 0 : 3 * b_bytes));
  if (p.num_stages < 2 || p.num_stages > kTcMaxStages || smem > 227 * 1024 || (p.b_resident && p.n_tiles != 1))
    return cudaErrorInvalidValue;
  if (p.tw * p.th != kTcBlockM || (p.tw & 7) != 0 || (p.block_n & 31) != 0 || p.block_n > 256 || (p.n_valid & 31) != 0)
    return cudaErrorInvalidValue;
  if (3 * p.kchunks + p.kchunks2 > (kHaloTable - 256) / 32) return cudaErrorInvalidValue;
  if (p.pool && (p.residual != nullptr || p.tw > 16 || (p.h_out & 1) || (p.w_out & 1))) return cudaErrorInvalidValue;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(igemm_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int m_tiles = p.tiles_w * p.tiles_h * p.n_img;
  const int pair_tiles = ((m_tiles + 1) / 2) * p.n_tiles;
  const int pairs = pair_tiles < num_sms / 2 ? pair_tiles : num_sms / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kHaloThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, igemm_halo_kernel, p);
}

}  // namespace ocrfa
