// Persistent bidirectional-LSTM recurrence on tcgen05 (one launch per layer, all T timesteps, both directions).
//
// reference: nn.LSTM(bidirectional=True, batch_first=True) inside BidirectionalLSTM,
//            /root/reference/src/ocrfa/modules/blocks.py:85-102 (gate order i,f,g,o; h0 = c0 = 0)
//
// The input projection x.W_ih^T + b_ih + b_hh is hoisted into one GEMM (igemm_tc) that leaves fp32 gate
// pre-activations `xproj` [B*T][4096] in HBM (both directions, gate rows permuted to [32 tiles][i,f,g,o][16 units]).
// This kernel then walks the T dependent steps:
//   * CTA (n_tile, m_tile) owns 16 hidden units (64 gate columns) of one direction for 128 batch rows;
//     its W_hh slice [64 x 512] bf16 (64 KB) is loaded ONCE and stays in shared memory for all T steps;
//   * per step the previous hidden state h_{t-1} [128 x 512] bf16 is streamed by TMA from a compact, parity
//     double-buffered exchange buffer (eight contiguous 16 KB slabs, L2 resident) through an 8-deep ring,
//     multiplied on tcgen05 (M=128, N=64) into TMEM, and the epilogue warps apply the gate math with the cell
//     state held in REGISTERS across all steps; h_t is written both to the layer output [B][T][1024] and to the
//     exchange buffer;
//   * the only cross-CTA dependency is "all 32 CTAs of my (direction, m_tile) group wrote h_t": a monotonic
//     global counter per group (release add / acquire poll), no kernel relaunch, no grid-wide barrier;
//   * xproj rows for step s+1 are prefetched into registers while step s's peers finish, so the HBM latency
//     of the 4 MB/step xproj stream is off the critical path.
// All CTAs must be co-resident (<= 128 CTAs, 1 per SM): the host uses a cooperative launch, and every spin is
// bounded (trap on timeout) so a scheduling surprise surfaces as an error, not a hung GPU.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace ocrfa {

constexpr int kLstmThreads = 192;
constexpr int kLstmRing = 8;          // A-ring stages (one per K block of a step)
constexpr int kLstmTmemCols = 128;    // 2 accumulators x 64 columns
constexpr int kLstmGroupCtas = 32;    // CTAs per (direction, m_tile) group = 512 hidden units / 16

struct alignas(64) LstmTcParams {
  CUtensorMap tmX;  // h exchange buffer [2 dirs][m_tiles][2 parities][8 K-blocks][128 rows][64] bf16 as 2-D (64, rows), box {64,128}
  CUtensorMap tmW;  // permuted W_hh [4096][512] bf16, box {64, 64}
  const float* xproj;        // [B*T][4096] fp32
  __nv_bfloat16* hseq;       // [B][T][1024] bf16 (fwd | bwd): the layer output
  __nv_bfloat16* hx;         // exchange buffer (see tmX): h_t of the whole group, one contiguous 16 KB slab per K-block
  unsigned int* sync;        // [2 * m_tiles] group counters, zero before launch
  long long* dbg;            // optional [64 steps][16] SM-clock timestamps of CTA 0 (bring-up profiling), or null
  int B, T, m_tiles;
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ex2.approx-based gate activations (abs error ~1e-6, far below the bf16 rounding of h); saturate cleanly at +-inf.
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

__global__ void __launch_bounds__(kLstmThreads, 1) lstm_tc_persistent_kernel(const __grid_constant__ LstmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem);   // [kLstmRing]
  uint64_t* a_empty = a_full + kLstmRing;                 // [kLstmRing]
  uint64_t* tfull = a_empty + kLstmRing;                  // [2]
  uint64_t* tempty = tfull + 2;                           // [2]
  uint64_t* w_full = tempty + 2;                          // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  uint8_t* w_smem = smem + 1024;                          // 8 K-blocks x (64 rows x 128 B) = 64 KB
  uint8_t* ring = w_smem + 8 * 8192;                      // kLstmRing x 16 KB

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x % 64;
  const int m_tile = blockIdx.x / 64;
  const int dir = n_tile >> 5;
  unsigned int* group_ctr = p.sync + dir * p.m_tiles + m_tile;
  const int T = p.T;
  long long* dbg = (blockIdx.x == 0) ? p.dbg : nullptr;
#define LSTM_TS(step, slot) do { if (dbg && (step) < 64) dbg[(step) * 16 + (slot)] = clock64(); } while (0)

  if (threadIdx.x == 0) {
    for (int s = 0; s < kLstmRing; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 128); }
    mbar_init(w_full, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&p.tmX); tma_prefetch_desc(&p.tmW); }
  if (warp == 1) { tmem_alloc(tmem_slot, kLstmTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      mbar_expect_tx(w_full, 8 * 8192);
      for (int kb = 0; kb < 8; ++kb) tma_load_2d(w_smem + kb * 8192, &p.tmW, w_full, kb * 64, n_tile * 64);
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 0; s < T; ++s) {
        LSTM_TS(s, 0);
        if (s > 0) {
          // wait until every CTA of this (direction, m_tile) group has published h of step s-1
          const unsigned int need = static_cast<unsigned int>(kLstmGroupCtas) * static_cast<unsigned int>(s);
          if (ld_acquire_gpu(group_ctr) < need) {
            const uint64_t t0 = global_timer_ns();
            uint32_t spins = 0;
            while (ld_acquire_gpu(group_ctr) < need) {
              if ((++spins & 0xff) == 0 && global_timer_ns() - t0 > OCRFA_MBAR_TIMEOUT_NS) {
                printf("[ocrfa] lstm group barrier timeout: block %d step %d have %u need %u\n", (int)blockIdx.x, s,
                       ld_acquire_gpu(group_ctr), need);
                __trap();
              }
            }
          }
          LSTM_TS(s, 1);
          fence_proxy_async_all();  // peers' generic-proxy stores of h -> visible to the TMA (async proxy) reads below
        }
        LSTM_TS(s, 2);
        // slabs of h_{s-1}: parity (s-1)&1 of this group's exchange region; at s == 0 an out-of-bounds row
        // coordinate makes the TMA zero-fill (h0 = 0)
        const int xrow0 = s == 0 ? (2 * p.m_tiles * 2 * 8 * 128)
                                 : (((dir * p.m_tiles + m_tile) * 2 + ((s - 1) & 1)) * 8) * 128;
        for (int kb = 0; kb < 8; ++kb) {
          mbar_wait(&a_empty[stage], phase ^ 1u, 11);
          mbar_expect_tx(&a_full[stage], 16384);
          tma_load_2d(ring + stage * 16384, &p.tmX, &a_full[stage], 0, xrow0 + kb * 128);
          if (++stage == kLstmRing) { stage = 0; phase ^= 1u; }
        }
        LSTM_TS(s, 3);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 64);
      mbar_wait(w_full, 0, 12);
      tc_fence_after();
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int s = 0; s < T; ++s) {
        mbar_wait(&tempty[as], aphase ^ 1u, 13);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as) * 64u;
        for (int kb = 0; kb < 8; ++kb) {
          mbar_wait(&a_full[stage], phase, 14);
          if (kb == 0) LSTM_TS(s, 4);
          if (kb == 7) LSTM_TS(s, 5);
          tc_fence_after();
          const uint64_t a_desc = make_sw128_kmajor_desc(smem_u32(ring + stage * 16384));
          const uint64_t b_desc = make_sw128_kmajor_desc(smem_u32(w_smem + kb * 8192));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, a_desc + static_cast<uint64_t>(2 * k), b_desc + static_cast<uint64_t>(2 * k), idesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&a_empty[stage]);
          if (++stage == kLstmRing) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull[as]);
        LSTM_TS(s, 6);
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else {
    // ===================================== epilogue: gate math, c in registers ===============
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int b = m_tile * 128 + row;
    const bool valid = b < p.B;
    const int j0 = (n_tile & 31) * 16;
    const int et = threadIdx.x - 64;  // 0..127 among the epilogue threads
    float c[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) c[u] = 0.f;
    float4 xr[16];  // xproj row slice for the CURRENT step: [i(16) f(16) g(16) o(16)]
    {
      const int t0 = dir == 0 ? 0 : T - 1;
      const float4* xp = reinterpret_cast<const float4*>(p.xproj + (static_cast<size_t>(valid ? b : 0) * T + t0) * 4096 + n_tile * 64);
#pragma unroll
      for (int i = 0; i < 16; ++i) xr[i] = valid ? __ldg(xp + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    int as = 0;
    uint32_t aphase = 0;
    for (int s = 0; s < T; ++s) {
      const int t_cur = dir == 0 ? s : T - 1 - s;
      mbar_wait(&tfull[as], aphase, 15);
      if (et == 0) LSTM_TS(s, 8);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as) * 64u;
      uint32_t v0[32], v1[32];
      tmem_ld_32x32b_x32(taddr, v0);
      tmem_ld_32x32b_x32(taddr + 32u, v1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&tempty[as]);
      if (et == 0) LSTM_TS(s, 9);
      float hout[16];
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4) {
        const float xi[4] = {xr[g4].x, xr[g4].y, xr[g4].z, xr[g4].w};
        const float xf[4] = {xr[4 + g4].x, xr[4 + g4].y, xr[4 + g4].z, xr[4 + g4].w};
        const float xg[4] = {xr[8 + g4].x, xr[8 + g4].y, xr[8 + g4].z, xr[8 + g4].w};
        const float xo[4] = {xr[12 + g4].x, xr[12 + g4].y, xr[12 + g4].z, xr[12 + g4].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int u = g4 * 4 + e;
          const float gi = sigmoid_fast(__uint_as_float(v0[u]) + xi[e]);
          const float gf = sigmoid_fast(__uint_as_float(v0[16 + u]) + xf[e]);
          const float gg = tanh_fast(__uint_as_float(v1[u]) + xg[e]);
          const float go = sigmoid_fast(__uint_as_float(v1[16 + u]) + xo[e]);
          c[u] = gf * c[u] + gi * gg;
          hout[u] = go * tanh_fast(c[u]);
        }
      }
      if (et == 0) LSTM_TS(s, 10);
      if (valid) {
        __nv_bfloat16* hp = p.hseq + (static_cast<size_t>(b) * T + t_cur) * 1024 + dir * 512 + j0;
        uint4 o0, o1;
        o0.x = pack_bf16x2(hout[0], hout[1]);   o0.y = pack_bf16x2(hout[2], hout[3]);
        o0.z = pack_bf16x2(hout[4], hout[5]);   o0.w = pack_bf16x2(hout[6], hout[7]);
        o1.x = pack_bf16x2(hout[8], hout[9]);   o1.y = pack_bf16x2(hout[10], hout[11]);
        o1.z = pack_bf16x2(hout[12], hout[13]); o1.w = pack_bf16x2(hout[14], hout[15]);
        *reinterpret_cast<uint4*>(hp) = o0;
        *reinterpret_cast<uint4*>(hp + 8) = o1;
        // exchange copy for the next step's MMAs (rows >= B are never read back into valid results)
        __nv_bfloat16* xp2 = p.hx + (static_cast<size_t>((((dir * p.m_tiles + m_tile) * 2 + (s & 1)) * 8 + (j0 >> 6)) * 128 + row)) * 64 + (j0 & 63);
        *reinterpret_cast<uint4*>(xp2) = o0;
        *reinterpret_cast<uint4*>(xp2 + 8) = o1;
      }
      // publish: the CTA barrier orders every epilogue thread's h stores before thread 0's gpu-scope release
      // (cumulative), which the peers' acquire polls + proxy fence pair with
      if (et == 0) LSTM_TS(s, 11);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
        LSTM_TS(s, 13);
        red_release_gpu_add(group_ctr, 1u);
        LSTM_TS(s, 14);
      }
      // prefetch the next step's xproj slice while the group converges and the next MMAs run
      if (s + 1 < T) {
        const int t_next = dir == 0 ? s + 1 : T - 2 - s;
        const float4* xp = reinterpret_cast<const float4*>(p.xproj + (static_cast<size_t>(valid ? b : 0) * T + t_next) * 4096 + n_tile * 64);
#pragma unroll
        for (int i = 0; i < 16; ++i) xr[i] = valid ? __ldg(xp + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, kLstmTmemCols);
  }
}

inline size_t lstm_tc_smem_bytes() { return 2048 + 8 * 8192 + static_cast<size_t>(kLstmRing) * 16384; }

// Cooperative launch (all CTAs co-resident). `sync` must hold 2*m_tiles zeroed counters.
inline cudaError_t lstm_tc_launch(const LstmTcParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  const size_t smem = lstm_tc_smem_bytes();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(lstm_tc_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  void* args[] = {const_cast<LstmTcParams*>(&p)};
  return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lstm_tc_persistent_kernel), dim3(64 * p.m_tiles), dim3(kLstmThreads),
                                     args, smem, stream);
}

}  // namespace ocrfa
