// Persistent bidirectional-LSTM recurrence on tcgen05 (one launch per layer, all T timesteps, both directions).
//
// reference: nn.LSTM(bidirectional=True, batch_first=True) inside BidirectionalLSTM,
//            /root/reference/src/ocrfa/modules/blocks.py:85-102 (gate order i,f,g,o; h0 = c0 = 0)
//
// CTA (n_tile, m_tile) owns 16 hidden units (64 gate columns, rows permuted to [32 tiles][i,f,g,o][16 units]) of
// one direction for 128 batch rows. Sixteen 8 KB weight slabs (64 gate columns x 64 K) are loaded ONCE and stay in
// shared memory for all T steps; per step sixteen 16 KB activation slabs (128 rows x 64 K) stream through a TMA
// ring into tcgen05.mma (M=128, N=64, fp32 accumulators double-buffered in TMEM); the epilogue warps apply the gate
// math with the cell state held in REGISTERS across all steps and write h_t to the layer output and to a compact,
// parity double-buffered exchange buffer (contiguous 16 KB slabs, L2 resident) that feeds the next step.
//
// Two arithmetic modes share the pipeline:
//   kLstmBf16FusedX  bf16 operands. K = [x_t | h_{t-1}] (1024): the input projection is FUSED - the resident
//                    weights are [W_ih | W_hh], the eight x_t slabs come straight from the layer input [B][T][512]
//                    and do not depend on the recurrence, so they are fetched and multiplied while the group is
//                    still converging on h_{t-1}; no xproj tensor ever reaches HBM. Bias from shared memory.
//   kLstmF16Split    fp16 hi/lo split operands (fp32-grade): h = h_hi + h_lo, W_hh = W_hi + W_lo (pre-scaled by a
//                    power of two), three products h_hi.W_hi + h_hi.W_lo + h_lo.W_hi per K block; the resident
//                    weights are [W_hi | W_lo]; the input projection (fp32, bias included) comes from the split
//                    igemm; accurate expf/tanhf gate math. The tensor core's fp32 accumulate truncates (error
//                    linear in the number of chained MMAs, see igemm_x3.cuh), so the h_hi.W_hi products of each
//                    64-wide K block go to a fresh TMEM buffer ("main", four in rotation) that the epilogue warps
//                    sum in registers with round-to-nearest adds while the next slabs arrive; the two small
//                    products chain in a separate accumulator ("corr") added once per step.
//
// The only cross-CTA dependency is "all 32 CTAs of my (direction, m_tile) group wrote h_t": a monotonic global
// counter per group (counter add after the slice is at L2 / acquire poll), no kernel relaunch, no grid-wide barrier. All CTAs must be
// co-resident (<= 128 CTAs, 1 per SM): the host uses a cooperative launch, and every spin is bounded (trap on
// timeout) so a scheduling surprise surfaces as an error, not a hung GPU.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace ocrfa {

constexpr int kLstmThreads = 192;
// How a CTA publishes its slice of h_t to the group (fused-x modes):
//   generic  per-thread 16-byte stores + CTA barrier + red.release.gpu (the split-precision mode, and every mode with
//            -DOCRFA_LSTM_GENERIC_PUBLISH). The release is MEMBAR.ALL.GPU + ERRBAR in SASS and waits for the CTA's
//            scattered stores: measured 1.4 k cycles per step.
//   TMA      the slice is staged in shared memory (4 KB, double buffered), one elected thread issues TWO TMA tensor
//            stores from that tile - the exchange copy the peers wait for and the layer-output copy (h sequence) - each
//            in its own bulk group, waits for the FIRST group to complete (cp.async.bulk.wait_group 1), and then bumps the
//            counter with red.RELEASE.gpu: completion makes the bulk writes visible to the issuing thread, the gpu-scope
//            release publishes them (cumulativity), the peers' acquire poll + fence.proxy.async.global order their TMA
//            reads after it. Round 1 used a RELAXED red here on a physical argument (a completed bulk store is at L2);
//            the stress run of tools/test_igemm.cu --lstm-stress caught it: 2 of 100 launches at the c2 geometry read a
//            stale slice. With every global store of the step moved to the async proxy no generic store is outstanding
//            when the release executes, which is what made the fenced path slow before.
#ifndef OCRFA_LSTM_GENERIC_PUBLISH
#define OCRFA_LSTM_TMA_PUBLISH 1
#endif
#ifdef OCRFA_LSTM_TMA_PUBLISH
constexpr int kLstmRing = 5;          // A-ring stages of 16 KB
constexpr int kLstmStageBytes = 8192; // h slice staging: 2 x (128 rows x 16 units x 2 B), alternating by step parity
#else
constexpr int kLstmRing = 6;          // A-ring stages of 16 KB
constexpr int kLstmStageBytes = 0;
#endif
constexpr bool kLstmTmaPublishSplit = false;  // the split-precision mode keeps the generic publish
constexpr int kLstmTmemCols = 128;    // fused-x modes: 2 accumulators x 64 columns
constexpr int kLstmTmemColsSplit = 512;  // split mode: 4 "main" chunk buffers (columns 0..255) + "corr" (column 256)
constexpr float kLstmComp = 0x1p-24f;  // default of LstmTcParams::comp, same rule as kX3Comp in igemm_x3.cuh (experiment switch: OCRFA_LSTM_COMP)
constexpr int kLstmCorrCol = 256;
constexpr int kLstmGroupCtas = 32;    // CTAs per (direction, m_tile) group = 512 hidden units / 16
constexpr int kLstmWSlabs = 16;       // resident weight slabs (8 KB each) = 128 KB

enum LstmMode : int { kLstmBf16FusedX = 0, kLstmF16Split = 1, kLstmF16FusedX = 2 };  // FusedX modes differ only in the 16-bit format

// slabs of the exchange buffer per (direction, m_tile, parity): 8 K-blocks of h (bf16) or 8 hi + 8 lo (split)
__host__ __device__ constexpr int lstm_hx_slabs(int mode) { return mode == kLstmF16Split ? 16 : 8; }
__host__ __device__ inline long long lstm_hx_rows(int mode, int m_tiles) {
  return static_cast<long long>(2) * m_tiles * 2 * lstm_hx_slabs(mode) * 128;
}

struct alignas(64) LstmTcParams {
  CUtensorMap tmX;  // exchange buffer [2 dirs][m_tiles][2 parities][slabs][128 rows][64] as 2-D (64, rows), box {64,128}
  CUtensorMap tmW;  // resident weights [4096][1024] (bf16: [W_ih | W_hh]; split: [W_hi | W_lo]), box {64, 64}
  CUtensorMap tmS;  // bf16 mode: layer input [B][T][512] bf16 as the 5-D activation map, box {64,1,1,128,1}
  CUtensorMap tmP;  // TMA publish: the exchange buffer again, un-swizzled, box {16 units, box_rows}
  CUtensorMap tmH;  // TMA publish: the layer output [B][T * 1024] as 2-D (T * 1024, B), un-swizzled, box {16 units, box_rows}
  const float* xproj;        // split mode: [B*T][4096] fp32 gate pre-activations (bias included)
  const float* bias;         // bf16 mode: [4096] b_ih + b_hh in the permuted order
  void* hseq;                // layer output: bf16 [B][T][1024] (fwd | bwd), or fp16 [B][T][hi 1024 | lo 1024]
  void* hx;                  // exchange buffer (see tmX)
  unsigned int* sync;        // [2 * m_tiles] group counters, zero before launch
  long long* dbg;            // optional [64 steps][16] SM-clock timestamps of CTA 0 (bring-up profiling), or null
  float acc_scale;           // split mode: 2^-k undoing the power-of-two pre-scale of W_hh
  float comp;                // split mode: truncation compensation, sums become x + x * comp (0 = the default), see igemm_x3.cuh
  int B, T, m_tiles;
  int box_rows;              // batch rows per TMA slab: 128, or B rounded up to 8 when one M tile holds the whole batch
                             // (a step is paced by the bytes each SM ingests: 64 lines load half of them)
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
#ifdef OCRFA_LSTM_RELAXED_PUBLISH
  // MEASUREMENT ONLY (tools/test_igemm.cu, never the product library): drops the gpu-scope fence of the release
  // (MEMBAR.ALL.GPU + ERRBAR in SASS) to see what it costs per step; without it the exchange stores are not ordered
  // before the counter update.
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
// The h slabs are GLOBAL memory written through the generic proxy by the peer CTAs and read back by TMA (async proxy):
// the fence is restricted to the global state space - measured 36 cycles against 770 for the all-spaces form, which
// sat on the critical path of every step.
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void red_relaxed_gpu_add(unsigned int* p, unsigned int v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread except the most recent one have completed (their writes are visible to this thread)
__device__ __forceinline__ void tma_store_wait_all_but_last() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ex2.approx-based gate activations (abs error ~1e-6, far below the bf16 rounding of h); saturate cleanly at +-inf.
#ifdef OCRFA_LSTM_TANH_APPROX
// one MUFU op per activation (tanh.approx.f32, rel. error 2^-11) instead of two (ex2 + rcp): the gate math of a
// step is MUFU-throughput bound (16 units x 5 activations x 128 rows per CTA)
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(tanh_fast(0.5f * x), 0.5f, 0.5f); }
#else
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }
#endif

// fp32-grade gate activations of the split mode, BRANCH-FREE: the gate math of a step is 80 activations per thread with
// one warp per SM sub-partition, i.e. bound by the latency of each dependent chain unless the compiler can interleave the
// 16 hidden units - which the branches inside libdevice's tanhf / the IEEE division prevented (measured: 7 100 cycles
// of a 21 700-cycle step). Accuracy: exp ~1 ulp (Cody-Waite reduction + the degree-5 kernel of Cephes expf), reciprocal
// < 1 ulp (MUFU.RCP + one Newton step), tanh by the odd polynomial of Cephes tanhf below 0.625.
__device__ __forceinline__ float exp_bf(float x) {
  x = fminf(fmaxf(x, -87.0f), 88.0f);
  const float n = rintf(x * 1.44269504088896341f);
  float r = fmaf(n, -0.693359375f, x);         // ln2 split as in Cephes: C1 = 0.693359375, C2 = -2.12194440e-4
  r = fmaf(n, 2.12194440e-4f, r);
  float q = 1.9875691500e-4f;
  q = fmaf(q, r, 1.3981999507e-3f);
  q = fmaf(q, r, 8.3334519073e-3f);
  q = fmaf(q, r, 4.1665795894e-2f);
  q = fmaf(q, r, 1.6666665459e-1f);
  q = fmaf(q, r, 5.0000001201e-1f);
  q = fmaf(q * r, r, r) + 1.0f;
  return __int_as_float(__float_as_int(q) + (static_cast<int>(n) << 23));  // q in [0.7, 1.42], result stays normal
}
__device__ __forceinline__ float rcp_bf(float d) {  // d in [1, 2^127)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return fmaf(r, fmaf(-d, r, 1.0f), r);
}
__device__ __forceinline__ float sigmoid_bf(float x) { return rcp_bf(1.0f + exp_bf(-x)); }
__device__ __forceinline__ float tanh_bf(float x) {
  const float a = fabsf(x);
  const float big = fmaf(-2.0f, rcp_bf(1.0f + exp_bf(2.0f * a)), 1.0f);
  const float z = x * x;
  float q = -5.70498872745e-3f;
  q = fmaf(q, z, 2.06390887954e-2f);
  q = fmaf(q, z, -5.37397155531e-2f);
  q = fmaf(q, z, 1.33314422036e-1f);
  q = fmaf(q, z, -3.33332819422e-1f);
  const float small = fmaf(q * z, x, x);
  return a < 0.625f ? small : copysignf(big, x);
}

template <int MODE> __device__ __forceinline__ float lstm_sigmoid(float x) {
  return MODE == kLstmF16Split ? sigmoid_bf(x) : sigmoid_fast(x);
}
template <int MODE> __device__ __forceinline__ float lstm_tanh(float x) {
  return MODE == kLstmF16Split ? tanh_bf(x) : tanh_fast(x);
}

__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// x = hi + lo with hi = fp16(x), lo = fp16(x - hi): 22 significant bits for |x| in the fp16 normal range
__device__ __forceinline__ void split_f16(float x, float* hi, float* lo) {
  const float h = __half2float(__float2half_rn(x));
  *hi = h;
  *lo = x - h;
}

template <int MODE_T>
__global__ void __launch_bounds__(kLstmThreads, 1) lstm_tc_persistent_kernel(const __grid_constant__ LstmTcParams p) {
  // the fused-x modes share every code path except the operand format (bf16 / fp16)
  constexpr int MODE = MODE_T == kLstmF16FusedX ? kLstmBf16FusedX : MODE_T;
  constexpr bool kHalf16 = MODE_T == kLstmF16FusedX;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem);   // [kLstmRing]
  uint64_t* a_empty = a_full + kLstmRing;                 // [kLstmRing]
  uint64_t* tfull = a_empty + kLstmRing;                  // [2]
  uint64_t* tempty = tfull + 2;                           // [2]
  uint64_t* w_full = tempty + 2;                          // [1]
  uint64_t* main_full = w_full + 1;                       // [4] split mode: hi.hi chunk complete
  uint64_t* main_empty = main_full + 4;                   // [4] split mode: chunk drained by the four epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(main_empty + 4);
  float* sbias = reinterpret_cast<float*>(smem + 512);    // [64] (bf16 mode)
  uint8_t* w_smem = smem + 1024;                          // kLstmWSlabs x (64 rows x 128 B) = 128 KB
  uint8_t* ring = w_smem + kLstmWSlabs * 8192;            // kLstmRing x 16 KB
  uint8_t* hstage = ring + kLstmRing * 16384;             // kLstmStageBytes (TMA publish)
  (void)hstage;

  constexpr int kHxSlabs = lstm_hx_slabs(MODE);
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
    for (int s = 0; s < 4; ++s) { mbar_init(&main_full[s], 1); mbar_init(&main_empty[s], 4); }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmX);
    tma_prefetch_desc(&p.tmW);
    if (MODE == kLstmBf16FusedX) tma_prefetch_desc(&p.tmS);
#ifdef OCRFA_LSTM_TMA_PUBLISH
    tma_prefetch_desc(&p.tmP);
    tma_prefetch_desc(&p.tmH);
#endif
  }
  constexpr uint32_t kTmemCols = MODE == kLstmF16Split ? kLstmTmemColsSplit : kLstmTmemCols;
  if (warp == 1) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  if (MODE == kLstmBf16FusedX && threadIdx.x >= 64 && threadIdx.x < 128)
    sbias[threadIdx.x - 64] = __ldg(p.bias + n_tile * 64 + (threadIdx.x - 64));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      mbar_expect_tx(w_full, kLstmWSlabs * 8192);
      for (int i = 0; i < kLstmWSlabs; ++i) tma_load_2d(w_smem + i * 8192, &p.tmW, w_full, i * 64, n_tile * 64);
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t slab_bytes = static_cast<uint32_t>(p.box_rows) * 128u;  // rows >= box_rows of a ring slot are never read back
      const int oob_row = static_cast<int>(lstm_hx_rows(MODE, p.m_tiles));
      for (int s = 0; s < T; ++s) {
        LSTM_TS(s, 0);
        if (MODE == kLstmBf16FusedX) {
          // x_t slabs: independent of the recurrence, so they are in flight (and multiplied) while the group converges
          const int t_cur = dir == 0 ? s : T - 1 - s;
          for (int kb = 0; kb < 8; ++kb) {
            mbar_wait(&a_empty[stage], phase ^ 1u, 10);
            mbar_expect_tx(&a_full[stage], slab_bytes);
            tma_load_5d(ring + stage * 16384, &p.tmS, &a_full[stage], kb * 64, t_cur, 0, m_tile * 128, 0);
            if (++stage == kLstmRing) { stage = 0; phase ^= 1u; }
          }
        }
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
        const int xrow0 = s == 0 ? oob_row : (((dir * p.m_tiles + m_tile) * 2 + ((s - 1) & 1)) * kHxSlabs) * 128;
        for (int i = 0; i < kHxSlabs; ++i) {
          mbar_wait(&a_empty[stage], phase ^ 1u, 11);
          mbar_expect_tx(&a_full[stage], slab_bytes);
          tma_load_2d(ring + stage * 16384, &p.tmX, &a_full[stage], 0, xrow0 + i * 128);
          if (++stage == kLstmRing) { stage = 0; phase ^= 1u; }
        }
        LSTM_TS(s, 3);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (elect_one()) {
      const uint32_t idesc = (MODE == kLstmF16Split || kHalf16) ? make_idesc_f16(128, 64) : make_idesc_bf16(128, 64);
      mbar_wait(w_full, 0, 12);
      tc_fence_after();
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0, mphase = 0;  // mphase: one parity bit per "main" buffer (split mode)
      for (int s = 0; s < T; ++s) {
        if (MODE == kLstmF16Split) {
          // corr in TMEM column 256 (one buffer: step s+1's slabs exist only after the epilogue of step s has read it)
          mbar_wait(&tempty[0], aphase ^ 1u, 13);
          aphase ^= 1u;
          tc_fence_after();
          const uint32_t d_corr = tmem_base + static_cast<uint32_t>(kLstmCorrCol);
          for (int i = 0; i < 16; ++i) {
            mbar_wait(&a_full[stage], phase, 14);
            if (i == 8) LSTM_TS(s, 4);
            if (i == 15) LSTM_TS(s, 5);
            tc_fence_after();
            const uint64_t a_desc = make_sw128_kmajor_desc(smem_u32(ring + stage * 16384));
            // slabs 0..7 = h_hi: corr += h_hi.W_lo, main[i & 3] = h_hi.W_hi; slabs 8..15 = h_lo: corr += h_lo.W_hi
            const uint64_t bc_desc = make_sw128_kmajor_desc(smem_u32(w_smem + (i < 8 ? i + 8 : i - 8) * 8192));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_corr, a_desc + static_cast<uint64_t>(2 * k), bc_desc + static_cast<uint64_t>(2 * k), idesc,
                        (i > 0 || k > 0) ? 1u : 0u);
            if (i < 8) {
              const int mb = i & 3;
              mbar_wait(&main_empty[mb], ((mphase >> mb) & 1u) ^ 1u, 16);
              mphase ^= 1u << mb;
              tc_fence_after();
              const uint64_t bm_desc = make_sw128_kmajor_desc(smem_u32(w_smem + i * 8192));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base + static_cast<uint32_t>(mb) * 64u, a_desc + static_cast<uint64_t>(2 * k),
                          bm_desc + static_cast<uint64_t>(2 * k), idesc, k > 0 ? 1u : 0u);
              umma_commit(&a_empty[stage]);
              umma_commit(&main_full[mb]);
            } else {
              umma_commit(&a_empty[stage]);
            }
            if (++stage == kLstmRing) { stage = 0; phase ^= 1u; }
          }
          umma_commit(&tfull[0]);
          LSTM_TS(s, 6);
          continue;
        }
        mbar_wait(&tempty[as], aphase ^ 1u, 13);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as) * 64u;
        uint32_t accumulate = 0;
        for (int i = 0; i < 16; ++i) {
          mbar_wait(&a_full[stage], phase, 14);
          if (i == 8) LSTM_TS(s, 4);
          if (i == 15) LSTM_TS(s, 5);
          tc_fence_after();
          const uint64_t a_desc = make_sw128_kmajor_desc(smem_u32(ring + stage * 16384));
          // bf16: A slab i (x K-blocks 0..7, then h K-blocks 0..7) x weight slab i ([W_ih | W_hh]).
          // split: slabs 0..7 = h_hi x (W_hi, W_lo); slabs 8..15 = h_lo x W_hi.
          const int w0 = MODE == kLstmF16Split ? (i & 7) : i;
          const int nw = (MODE == kLstmF16Split && i < 8) ? 2 : 1;
          for (int j = 0; j < nw; ++j) {
            const uint64_t b_desc = make_sw128_kmajor_desc(smem_u32(w_smem + (w0 + 8 * j) * 8192));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16(d_tmem, a_desc + static_cast<uint64_t>(2 * k), b_desc + static_cast<uint64_t>(2 * k), idesc, accumulate);
              accumulate = 1u;
            }
          }
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
    float4 xr[16];  // split mode: xproj row slice for the CURRENT step: [i(16) f(16) g(16) o(16)]
    if (MODE == kLstmF16Split) {
      const int t0 = dir == 0 ? 0 : T - 1;
      const float4* xp = reinterpret_cast<const float4*>(p.xproj + (static_cast<size_t>(valid ? b : 0) * T + t0) * 4096 + n_tile * 64);
#pragma unroll
      for (int i = 0; i < 16; ++i) xr[i] = valid ? __ldg(xp + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float acc_scale = p.acc_scale;
    const size_t hx_region = static_cast<size_t>(dir * p.m_tiles + m_tile) * 2;
    int as = 0;
    uint32_t aphase = 0;
    uint32_t mphase = 0;
    for (int s = 0; s < T; ++s) {
      const int t_cur = dir == 0 ? s : T - 1 - s;
      uint32_t v0[32], v1[32];
      if (MODE == kLstmF16Split) {
        // sum the eight hi.hi chunks in registers (round-to-nearest fp32) as they complete, then add corr
        float racc[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) racc[i] = 0.f;
        const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        for (int i = 0; i <= 8; ++i) {
          const int mb = i & 3;
          if (i < 8) { mbar_wait(&main_full[mb], (mphase >> mb) & 1u, 17); mphase ^= 1u << mb; }
          else { mbar_wait(&tfull[0], aphase, 15); if (et == 0) LSTM_TS(s, 8); }
          tc_fence_after();
          const uint32_t taddr = tlane + (i < 8 ? static_cast<uint32_t>(mb) * 64u : static_cast<uint32_t>(kLstmCorrCol));
          tmem_ld_32x32b_x32(taddr, v0);
          tmem_ld_32x32b_x32(taddr + 32u, v1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) { racc[j] += __uint_as_float(v0[j]); racc[32 + j] += __uint_as_float(v1[j]); }
          tc_fence_before();
          if (i < 8) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&main_empty[mb]);
          } else {
            mbar_arrive(&tempty[0]);
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v0[j] = __float_as_uint(fmaf(racc[j], p.comp, racc[j]));
          v1[j] = __float_as_uint(fmaf(racc[32 + j], p.comp, racc[32 + j]));
        }
      } else {
      mbar_wait(&tfull[as], aphase, 15);
      if (et == 0) LSTM_TS(s, 8);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as) * 64u;
      tmem_ld_32x32b_x32(taddr, v0);
      tmem_ld_32x32b_x32(taddr + 32u, v1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&tempty[as]);
      }
      if (et == 0) LSTM_TS(s, 9);
      float hout[16];
      // a warp whose 32 rows are all beyond the batch skips the gate math: the activations are MUFU bound and shared
      // by the four epilogue warps (64 lines: half of the SM's transcendental work would be on padding rows)
      const bool warp_live = (m_tile * 128 + q * 32) < p.B;
#pragma unroll
      for (int u = 0; u < 16; ++u) hout[u] = 0.f;
      if (warp_live)
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4) {
        float4 ai, af, ag, ao;
        if (MODE == kLstmF16Split) {
          ai = xr[g4]; af = xr[4 + g4]; ag = xr[8 + g4]; ao = xr[12 + g4];
        } else {
          const float4* sb = reinterpret_cast<const float4*>(sbias);
          ai = sb[g4]; af = sb[4 + g4]; ag = sb[8 + g4]; ao = sb[12 + g4];
        }
        const float xi[4] = {ai.x, ai.y, ai.z, ai.w}, xf[4] = {af.x, af.y, af.z, af.w};
        const float xg[4] = {ag.x, ag.y, ag.z, ag.w}, xo[4] = {ao.x, ao.y, ao.z, ao.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int u = g4 * 4 + e;
          float pi = __uint_as_float(v0[u]), pf = __uint_as_float(v0[16 + u]);
          float pg = __uint_as_float(v1[u]), po = __uint_as_float(v1[16 + u]);
          if (MODE == kLstmF16Split) { pi *= acc_scale; pf *= acc_scale; pg *= acc_scale; po *= acc_scale; }
          const float gi = lstm_sigmoid<MODE>(pi + xi[e]);
          const float gf = lstm_sigmoid<MODE>(pf + xf[e]);
          const float gg = lstm_tanh<MODE>(pg + xg[e]);
          const float go = lstm_sigmoid<MODE>(po + xo[e]);
          c[u] = gf * c[u] + gi * gg;
          hout[u] = go * lstm_tanh<MODE>(c[u]);
        }
      }
      if (et == 0) LSTM_TS(s, 10);
      // pack h_t; the exchange copy goes out first (it is what the peers wait for), the layer output after the release
      uint4 o0, o1, l0, l1;
      if (MODE == kLstmF16Split) {
        float hi[16], lo[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) split_f16(hout[u], &hi[u], &lo[u]);
        o0.x = pack_f16x2(hi[0], hi[1]);   o0.y = pack_f16x2(hi[2], hi[3]);
        o0.z = pack_f16x2(hi[4], hi[5]);   o0.w = pack_f16x2(hi[6], hi[7]);
        o1.x = pack_f16x2(hi[8], hi[9]);   o1.y = pack_f16x2(hi[10], hi[11]);
        o1.z = pack_f16x2(hi[12], hi[13]); o1.w = pack_f16x2(hi[14], hi[15]);
        l0.x = pack_f16x2(lo[0], lo[1]);   l0.y = pack_f16x2(lo[2], lo[3]);
        l0.z = pack_f16x2(lo[4], lo[5]);   l0.w = pack_f16x2(lo[6], lo[7]);
        l1.x = pack_f16x2(lo[8], lo[9]);   l1.y = pack_f16x2(lo[10], lo[11]);
        l1.z = pack_f16x2(lo[12], lo[13]); l1.w = pack_f16x2(lo[14], lo[15]);
      } else {
        o0.x = pack16x2(hout[0], hout[1], kHalf16);   o0.y = pack16x2(hout[2], hout[3], kHalf16);
        o0.z = pack16x2(hout[4], hout[5], kHalf16);   o0.w = pack16x2(hout[6], hout[7], kHalf16);
        o1.x = pack16x2(hout[8], hout[9], kHalf16);   o1.y = pack16x2(hout[10], hout[11], kHalf16);
        o1.z = pack16x2(hout[12], hout[13], kHalf16); o1.w = pack16x2(hout[14], hout[15], kHalf16);
        l0 = o0; l1 = o1;
      }
#ifdef OCRFA_LSTM_TMA_PUBLISH
      constexpr bool kTmaPublish = MODE != kLstmF16Split || kLstmTmaPublishSplit;
#else
      constexpr bool kTmaPublish = false;
#endif
      if (kTmaPublish) {
        // stage the CTA's [rows][16 units] slice (rows beyond the batch carry whatever the padding lanes computed: the
        // exchange rows they land in never reach a valid result, and the layer-output map ends at row B), hand it to the
        // async proxy, two TMA stores in two bulk groups, wait for the exchange group, release the counter. The tile of
        // step s-2 (same parity) is free: its groups completed before this thread's release of step s-1, which every
        // thread's tfull wait of step s transitively follows (the group counter gates the MMAs that signal tfull).
        uint8_t* hst = hstage + (s & 1) * 4096;
        uint4* sp4 = reinterpret_cast<uint4*>(hst) + row * 2;
        sp4[0] = o0;
        sp4[1] = o1;
        fence_proxy_async_smem();
        if (et == 0) LSTM_TS(s, 11);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          LSTM_TS(s, 13);
          const int slab_row0 = static_cast<int>(((hx_region + (s & 1)) * kHxSlabs + (j0 >> 6)) * 128);
          tma_store_2d(&p.tmP, hst, j0 & 63, slab_row0);
          tma_store_commit();
          tma_store_2d(&p.tmH, hst, t_cur * 1024 + dir * 512 + j0, m_tile * 128);
          tma_store_commit();
          tma_store_wait_all_but_last();
          red_release_gpu_add(group_ctr, 1u);
          LSTM_TS(s, 14);
        }
      } else {
        if (valid) {
          // rows >= B of the exchange buffer are never read back into valid results
          uint16_t* hx = reinterpret_cast<uint16_t*>(p.hx);
          const size_t slab = (hx_region + (s & 1)) * kHxSlabs + (j0 >> 6);
          uint16_t* xp2 = hx + (slab * 128 + row) * 64 + (j0 & 63);
          *reinterpret_cast<uint4*>(xp2) = o0;
          *reinterpret_cast<uint4*>(xp2 + 8) = o1;
          if (MODE == kLstmF16Split) {
            uint16_t* xp3 = xp2 + static_cast<size_t>(8) * 128 * 64;  // lo slabs follow the 8 hi slabs
            *reinterpret_cast<uint4*>(xp3) = l0;
            *reinterpret_cast<uint4*>(xp3 + 8) = l1;
          }
        }
        // publish: the CTA barrier orders every epilogue thread's exchange stores before thread 0's gpu-scope release
        // (cumulative), which the peers' acquire polls + proxy fence pair with
        if (et == 0) LSTM_TS(s, 11);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          LSTM_TS(s, 13);
          red_release_gpu_add(group_ctr, 1u);
          LSTM_TS(s, 14);
        }
      }
      if (valid && !kTmaPublish) {
        uint16_t* hs = reinterpret_cast<uint16_t*>(p.hseq);
        if (MODE == kLstmF16Split) {
          uint16_t* hp = hs + (static_cast<size_t>(b) * T + t_cur) * 2048 + dir * 512 + j0;
          *reinterpret_cast<uint4*>(hp) = o0;
          *reinterpret_cast<uint4*>(hp + 8) = o1;
          *reinterpret_cast<uint4*>(hp + 1024) = l0;
          *reinterpret_cast<uint4*>(hp + 1032) = l1;
        } else {
          uint16_t* hp = hs + (static_cast<size_t>(b) * T + t_cur) * 1024 + dir * 512 + j0;
          *reinterpret_cast<uint4*>(hp) = o0;
          *reinterpret_cast<uint4*>(hp + 8) = o1;
        }
      }
      // split mode: prefetch the next step's xproj slice while the group converges and the next MMAs run
      if (MODE == kLstmF16Split && s + 1 < T) {
        const int t_next = dir == 0 ? s + 1 : T - 2 - s;
        const float4* xp = reinterpret_cast<const float4*>(p.xproj + (static_cast<size_t>(valid ? b : 0) * T + t_next) * 4096 + n_tile * 64);
#pragma unroll
        for (int i = 0; i < 16; ++i) xr[i] = valid ? __ldg(xp + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (MODE == kLstmF16Split) {
        aphase ^= 1u;
      } else {
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
#ifdef OCRFA_LSTM_TMA_PUBLISH
    if (MODE != kLstmF16Split && et == 0) tma_store_wait_all();  // the last layer-output store reads shared memory
#endif
  }
#undef LSTM_TS

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

inline size_t lstm_tc_smem_bytes() {
  return 2048 + kLstmWSlabs * 8192 + static_cast<size_t>(kLstmRing) * 16384 + kLstmStageBytes;
}

// Cooperative launch (all CTAs co-resident). `sync` must hold 2*m_tiles zeroed counters.
inline cudaError_t lstm_tc_launch(const LstmTcParams& p_in, int mode, cudaStream_t stream) {
  LstmTcParams p = p_in;
  if (p.comp == 0.0f) {
    const char* e = getenv("OCRFA_LSTM_COMP");
    p.comp = e ? static_cast<float>(atof(e)) : kLstmComp;
  }
  static bool attr_set_dev[kMaxDevices][3] = {};
  bool* attr_set = attr_set_dev[current_device_slot()];
  const size_t smem = lstm_tc_smem_bytes();
  void* fn = mode == kLstmF16Split    ? reinterpret_cast<void*>(lstm_tc_persistent_kernel<kLstmF16Split>)
             : mode == kLstmF16FusedX ? reinterpret_cast<void*>(lstm_tc_persistent_kernel<kLstmF16FusedX>)
                                      : reinterpret_cast<void*>(lstm_tc_persistent_kernel<kLstmBf16FusedX>);
  if (!attr_set[mode]) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set[mode] = true;
  }
  void* args[] = {&p};
  return cudaLaunchCooperativeKernel(fn, dim3(64 * p.m_tiles), dim3(kLstmThreads), args, smem, stream);
}

}  // namespace ocrfa
