// CUDA-core kernels of the recognition path: everything that is not a large dense contraction, plus the
// fp32 "parity mode" contraction (CUDA-core implicit GEMM, used when logits must match the reference's fp32
// arithmetic to 1e-3 and as the on-device cross-check of the tcgen05 path).
//
//   preprocess   /root/reference/src/ocrfa/utils.py:541-550, 592-602   (normalise + replicate-pad)
//   stem         /root/reference/src/ocrfa/modules/blocks.py:239-247   (conv0_1 + bn0_1 + relu, Cin = 1)
//   max pools    /root/reference/src/ocrfa/modules/blocks.py:259, 271-275, 292
//   H' mean      /root/reference/src/ocrfa/modules/blocks.py:433, 441-446
//   LSTM cell    /root/reference/src/ocrfa/modules/blocks.py:89 (nn.LSTM gate math, order i,f,g,o)
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace ocrfa {

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)); }
template <typename T> struct Is16F16 { static constexpr bool value = false; };
template <> struct Is16F16<__half> { static constexpr bool value = true; };

// ---------------------------------------------------------------------------------------------
// preprocess: uint8 lines (already at model height) -> float32 [B,1,H,W] in [-1,1], right-padded by
// replicating the last valid column. Memory-bound: 1 B read + 4 B written per pixel; 16 pixels/thread.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float normalize_u8(uint32_t v) {
  // ToTensor (/255) then sub_(0.5).div_(0.5), in fp32 exactly as the reference computes it
  return (__fdiv_rn(static_cast<float>(v), 255.0f) - 0.5f) / 0.5f;
}

__global__ void preprocess_u8_kernel(const uint8_t* __restrict__ in, const int32_t* __restrict__ widths,
                                     float* __restrict__ out, int B, int H, int W, int in_stride) {
  // one thread = 16 consecutive pixels of one row; W is a multiple of 4 on this path
  const int chunks_per_row = (W + 15) / 16;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * H * chunks_per_row;
  if (gid >= total) return;
  const int ch = gid % chunks_per_row;
  const long long row = gid / chunks_per_row;  // b*H + y
  const int b = row / H;
  const int wv = widths ? min(max(widths[b], 1), W) : W;
  const uint8_t* src = in + row * in_stride;
  float* dst = out + row * W;
  const int x0 = ch * 16;
  const float edge = normalize_u8(src[wv - 1]);
  if (x0 + 16 <= wv && (in_stride & 15) == 0) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + x0));
    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 o;
      o.x = normalize_u8(w4[q] & 0xff);
      o.y = normalize_u8((w4[q] >> 8) & 0xff);
      o.z = normalize_u8((w4[q] >> 16) & 0xff);
      o.w = normalize_u8(w4[q] >> 24);
      *reinterpret_cast<float4*>(dst + x0 + q * 4) = o;
    }
  } else {
    for (int x = x0; x < min(x0 + 16, W); ++x) dst[x] = x < wv ? normalize_u8(src[x]) : edge;
  }
}

// ---------------------------------------------------------------------------------------------
// stem: conv0_1 (1 -> 32, 3x3, pad 1) + folded BN + ReLU. Input fp32 [B,H,W] or uint8 (+normalise),
// output NHWC with OUT_C channels (32 real; the bf16 path pads to 64 so the next conv's K slab is 128 B).
// One thread = one pixel, 32 channels in registers; weights [9][32] + bias[32] in shared memory.
// ---------------------------------------------------------------------------------------------
template <typename TOut, int OUT_C>
__global__ void __launch_bounds__(256) stem_conv_kernel(const float* __restrict__ x, const uint8_t* __restrict__ x_u8,
                                                        const int32_t* __restrict__ widths,
                                                        const float* __restrict__ wgt /*[9][32]*/,
                                                        const float* __restrict__ bias /*[32]*/, TOut* __restrict__ out,
                                                        int B, int H, int W) {
  constexpr bool kSplit = OUT_C == 128;  // TOut = __half, split-precision activations
  __shared__ float sw[9 * 32];
  __shared__ float sb[32];
  for (int i = threadIdx.x; i < 9 * 32; i += blockDim.x) sw[i] = wgt[i];
  if (threadIdx.x < 32) sb[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)B * H * W) return;
  const int xw = pix % W;
  const int y = (pix / W) % H;
  const int b = pix / ((long long)W * H);
  float v[9];
  const int wv = (x_u8 && widths) ? min(max(widths[b], 1), W) : W;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int yy = y + r - 1, xx = xw + s - 1;
      float t = 0.f;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        if (x_u8) t = normalize_u8(x_u8[((long long)b * H + yy) * W + min(xx, wv - 1)]);
        else t = __ldg(x + ((long long)b * H + yy) * W + xx);
      }
      v[r * 3 + s] = t;
    }
  TOut* op = out + pix * OUT_C;
#pragma unroll
  for (int c0 = 0; c0 < 32; c0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 9; ++k) a = fmaf(v[k], sw[k * 32 + c0 + j], a);
      acc[j] = fmaxf(a + sb[c0 + j], 0.f);
    }
    if (kSplit) {
      // fp16 hi/lo pair per channel: pixel = [hi(64) | lo(64)], channels 32..63 of both halves are zero padding
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __half2 hv = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
        const float2 hf = __half22float2(hv);
        const __half2 lv = __floats2half2_rn(acc[2 * j] - hf.x, acc[2 * j + 1] - hf.y);
        hw[j] = *reinterpret_cast<const uint32_t*>(&hv);
        lw[j] = *reinterpret_cast<const uint32_t*>(&lv);
      }
      __half* hp = reinterpret_cast<__half*>(op);
      *reinterpret_cast<uint4*>(hp + c0) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      *reinterpret_cast<uint4*>(hp + 64 + c0) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    } else if (sizeof(TOut) == 2) {
      uint4 o;
      constexpr bool h16 = Is16F16<TOut>::value;  // (the split path's __half output took the branch above)
      o.x = pack16x2(acc[0], acc[1], h16); o.y = pack16x2(acc[2], acc[3], h16);
      o.z = pack16x2(acc[4], acc[5], h16); o.w = pack16x2(acc[6], acc[7], h16);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(op) + c0) = o;
    } else {
      float* fo = reinterpret_cast<float*>(op) + c0;
      *reinterpret_cast<float4*>(fo) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(fo + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
  if (kSplit) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    uint16_t* bp = reinterpret_cast<uint16_t*>(op);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      *reinterpret_cast<uint4*>(bp + 32 + i * 8) = z;
      *reinterpret_cast<uint4*>(bp + 96 + i * 8) = z;
    }
  } else if (OUT_C > 32) {
    // zero channel padding (bf16 path): 32 x 2 B = 64 B
    const uint4 z = make_uint4(0, 0, 0, 0);
    __nv_bfloat16* bp = reinterpret_cast<__nv_bfloat16*>(op) + 32;
#pragma unroll
    for (int i = 0; i < (OUT_C - 32) / 8; ++i) *reinterpret_cast<uint4*>(bp + i * 8) = z;
  }
}

// ---------------------------------------------------------------------------------------------
// stem, bf16 tensor-core path: same arithmetic, 32-channel NHWC bf16 output (64 B per pixel, no channel padding:
// conv0_2's TMA box zero-fills channels 32..63). One thread = one pixel; the 9 x 32 weights and 32 biases are
// a kernel argument, i.e. constant-bank FFMA operands (288 FFMA per pixel and no weight loads); uint8 inputs are
// normalised through a 256-entry table built with the reference's fp32 ops; each warp transposes its 32 pixels
// x 64 B through shared memory so that every store instruction writes 512 contiguous bytes.
// HBM bound on paper (1 B read + 64 B written per pixel); in practice FFMA-issue bound.
// ---------------------------------------------------------------------------------------------
struct StemWeights {
  float w[9 * 32];  // [tap][channel], BN scale folded
  float b[32];      // BN shift
};

__global__ void __launch_bounds__(256) stem_conv32_kernel(const float* __restrict__ x, const uint8_t* __restrict__ x_u8,
                                                          const int32_t* __restrict__ widths,
                                                          const __grid_constant__ StemWeights sw,
                                                          __nv_bfloat16* __restrict__ out, int B, int H, int W, int f16) {
  __shared__ float lut[256];
  __shared__ uint4 stg[8][32 * 4];  // per warp: 32 pixels x 4 pieces of 16 B, piece index XOR-swizzled by (pixel >> 1) & 3
  lut[threadIdx.x] = normalize_u8(threadIdx.x);
  __syncthreads();
  const unsigned total = static_cast<unsigned>(B) * H * W;  // < 2^31 (checked by the caller)
  const unsigned pix = blockIdx.x * 256u + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.f;
  if (pix < total) {
    const unsigned rowi = pix / static_cast<unsigned>(W);  // b * H + y
    const int xw = static_cast<int>(pix - rowi * W);
    const int b = static_cast<int>(rowi / static_cast<unsigned>(H));
    const int y = static_cast<int>(rowi - static_cast<unsigned>(b) * H);
    const int wv = (x_u8 && widths) ? min(max(__ldg(widths + b), 1), W) : W;
    float v[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int yy = y + r - 1;
      const bool row_ok = yy >= 0 && yy < H;
      const size_t roff = (static_cast<size_t>(b) * H + (row_ok ? yy : y)) * W;
#pragma unroll
      for (int s2 = 0; s2 < 3; ++s2) {
        const int xx = xw + s2 - 1;
        float t = 0.f;
        if (row_ok && xx >= 0 && xx < W) {
          if (x_u8) t = lut[__ldg(x_u8 + roff + min(xx, wv - 1))];
          else t = __ldg(x + roff + xx);
        }
        v[r * 3 + s2] = t;
      }
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 9; ++k) a = fmaf(v[k], sw.w[k * 32 + c], a);
      acc[c] = fmaxf(a + sw.b[c], 0.f);
    }
  }
  const int swz = (lane >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 o;
    const bool h16 = f16 != 0;
    o.x = pack16x2(acc[8 * j], acc[8 * j + 1], h16);     o.y = pack16x2(acc[8 * j + 2], acc[8 * j + 3], h16);
    o.z = pack16x2(acc[8 * j + 4], acc[8 * j + 5], h16); o.w = pack16x2(acc[8 * j + 6], acc[8 * j + 7], h16);
    stg[warp][lane * 4 + (j ^ swz)] = o;
  }
  __syncwarp();
  // instruction k writes pixels 8k .. 8k+7 of this warp, 64 B each: 512 contiguous bytes
  const unsigned warp_pix0 = blockIdx.x * 256u + warp * 32u;
  uint4* op = reinterpret_cast<uint4*>(out + static_cast<size_t>(warp_pix0) * 32);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int pr = 8 * k + (lane >> 2);  // pixel of the warp, piece lane & 3
    if (warp_pix0 + pr < total) op[pr * 4 + (lane & 3)] = stg[warp][pr * 4 + ((lane & 3) ^ ((pr >> 1) & 3))];
  }
}

// Same arithmetic, four horizontally adjacent pixels per thread (W is a multiple of 4 on every path that reaches the
// stem). The one-pixel kernel above is issue bound on everything AROUND its 288 FFMAs (about 1000 instructions per
// pixel: uniform-register weight loads, index math, the 16-bit format select); here the 3 x 6 input window, the weight
// loads and the index math are shared by four pixels (about 430 instructions per pixel), the output format is a
// template parameter, and each warp stages its 128 pixels x 64 B (8 KB, contiguous in NHWC) through shared memory with a
// conflict-free swizzle so that every store instruction writes 512 contiguous bytes. Bit-identical results: per pixel
// the same fmaf chain in the same order.
template <bool F16>
__global__ void __launch_bounds__(128) stem_conv32x4_kernel(const float* __restrict__ x, const uint8_t* __restrict__ x_u8,
                                                            const int32_t* __restrict__ widths,
                                                            const __grid_constant__ StemWeights sw,
                                                            __nv_bfloat16* __restrict__ out, int B, int H, int W) {
  __shared__ float lut[256];
  __shared__ uint4 stg[4][128 * 4];  // per warp: slot (pixel) x 4 pieces of 16 B
  lut[threadIdx.x] = normalize_u8(threadIdx.x);
  lut[threadIdx.x + 128] = normalize_u8(threadIdx.x + 128);
  __syncthreads();
  const unsigned qpr = static_cast<unsigned>(W) >> 2;             // quads per row
  const unsigned total_q = static_cast<unsigned>(B) * H * qpr;    // < 2^29 (checked by the caller)
  const unsigned q = blockIdx.x * 128u + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint4* my = &stg[warp][0];
  const int slot_flip = (lane >> 2) & 1, piece_flip = lane & 3;
  if (q < total_q) {
    const unsigned rowi = q / qpr;  // b * H + y
    const int xq = static_cast<int>(q - rowi * qpr) * 4;
    const int b = static_cast<int>(rowi / static_cast<unsigned>(H));
    const int y = static_cast<int>(rowi - static_cast<unsigned>(b) * H);
    const int wv = (x_u8 && widths) ? min(max(__ldg(widths + b), 1), W) : W;
    // column offsets and validity are the same for the three rows: columns xq-1 .. xq+4, the uint8 path reads the last
    // valid column for x >= widths[b] (NormalizePAD's replicate pad), outside the line the conv's zero padding applies
    int coff[6];
    bool cok[6];
#pragma unroll
    for (int s2 = 0; s2 < 6; ++s2) {
      const int xx = xq + s2 - 1;
      cok[s2] = xx >= 0 && xx < W;
      coff[s2] = x_u8 ? min(max(xx, 0), wv - 1) : min(max(xx, 0), W - 1);
    }
    float v[3][6];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int yy = y + r - 1;
      const bool row_ok = yy >= 0 && yy < H;
      const size_t roff = (static_cast<size_t>(b) * H + (row_ok ? yy : y)) * W;
#pragma unroll
      for (int s2 = 0; s2 < 6; ++s2) {
        float t = x_u8 ? lut[__ldg(x_u8 + roff + coff[s2])] : __ldg(x + roff + coff[s2]);
        v[r][s2] = (row_ok && cok[s2]) ? t : 0.f;
      }
    }
#pragma unroll
    for (int cg = 0; cg < 4; ++cg) {
      float acc[4][8];
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        const int c = cg * 8 + c8;
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          float a = 0.f;
#pragma unroll
          for (int k = 0; k < 9; ++k) a = fmaf(v[k / 3][px + k % 3], sw.w[k * 32 + c], a);
          acc[px][c8] = fmaxf(a + sw.b[c], 0.f);
        }
      }
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        uint4 o;
        o.x = pack16x2(acc[px][0], acc[px][1], F16); o.y = pack16x2(acc[px][2], acc[px][3], F16);
        o.z = pack16x2(acc[px][4], acc[px][5], F16); o.w = pack16x2(acc[px][6], acc[px][7], F16);
        my[(lane * 4 + (px ^ slot_flip)) * 4 + (cg ^ piece_flip)] = o;
      }
    }
  }
  __syncwarp();
  // the warp's 32 quads are 128 consecutive pixels = 8 KB contiguous; instruction k writes pixels 8k .. 8k+7
  const size_t warp_pix0 = (static_cast<size_t>(blockIdx.x) * 128u + warp * 32u) * 4u;
  const size_t total_pix = static_cast<size_t>(total_q) * 4u;
  uint4* op = reinterpret_cast<uint4*>(out + warp_pix0 * 32);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int pr = 8 * k + (lane >> 2), piece = lane & 3;  // pixel of the warp, 16-byte piece of its 64 B
    const int owner = pr >> 2, px = pr & 3;               // the lane that computed it and its position in that lane's quad
    if (warp_pix0 + pr < total_pix)
      op[pr * 4 + piece] = my[(owner * 4 + (px ^ ((owner >> 2) & 1))) * 4 + (piece ^ (owner & 3))];
  }
}

// ---------------------------------------------------------------------------------------------
// generic CUDA-core implicit GEMM (fp32 accumulate): conv kh x kw, stride (sh,1), zero pad (ph,pw), NHWC.
// 64 pixels x 64 output channels per 256-thread block, 4x4 outputs per thread, K staged 16 at a time.
// ---------------------------------------------------------------------------------------------
struct ConvSimtParams {
  const void* in; long long in_pix_stride;  // elements between consecutive input pixels (>= cin)
  const float* wgt;                          // [cout][kh*kw*cin] fp32, K = (tap-major, channel-minor)
  const float* bias;                         // [cout] or null
  const void* residual; int ldr;
  void* out; int ldo;
  int n, h_in, w_in, cin, cout, kh, kw, sh, ph, pw, h_out, w_out, relu;
};

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvSimtParams p) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long total_pix = (long long)p.n * p.h_out * p.w_out;
  const long long pix0 = (long long)blockIdx.x * 64;
  const int co0 = blockIdx.y * 64;
  const int ktotal = p.kh * p.kw * p.cin;

  // loader role: thread loads 4 consecutive K elements of row (tid/4)
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const long long lpix = pix0 + lrow;
  const bool lpix_ok = lpix < total_pix;
  int ln = 0, lho = 0, lwo = 0;
  if (lpix_ok) {
    lwo = lpix % p.w_out;
    lho = (lpix / p.w_out) % p.h_out;
    ln = lpix / ((long long)p.w_out * p.h_out);
  }
  const int lco = co0 + lrow;
  const TIn* in = reinterpret_cast<const TIn*>(p.in);

  // fp64 accumulators: products of two floats are exact in double, so the only rounding left is the final
  // conversion - the parity mode then differs from the reference by the REFERENCE's own fp32 summation error only.
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (int r = 0; r < p.kh; ++r)
    for (int s = 0; s < p.kw; ++s) {
      const int hi = lho * p.sh + r - p.ph, wi = lwo + s - p.pw;
      const bool a_ok = lpix_ok && hi >= 0 && hi < p.h_in && wi >= 0 && wi < p.w_in;
      const TIn* ap = in + (((long long)ln * p.h_in + hi) * p.w_in + wi) * p.in_pix_stride;
      const float* bp = p.wgt + (long long)lco * ktotal + (r * p.kw + s) * p.cin;
      for (int c0 = 0; c0 < p.cin; c0 += 16) {
        float a4[4] = {0.f, 0.f, 0.f, 0.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
        if (a_ok) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (c0 + lk + e < p.cin) a4[e] = to_f32<TIn>(ap[c0 + lk + e]);
        }
        if (lco < p.cout) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (c0 + lk + e < p.cin) b4[e] = __ldg(bp + c0 + lk + e);
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 4; ++e) { As[lk + e][lrow] = a4[e]; Bs[lk + e][lrow] = b4[e]; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
          const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
          const double a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
      }
    }

  TOut* out = reinterpret_cast<TOut*>(p.out);
  const TOut* res = reinterpret_cast<const TOut*>(p.residual);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long pix = pix0 + ty * 4 + i;
    if (pix >= total_pix) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co >= p.cout) continue;
      double vd = acc[i][j];
      if (p.bias) vd += (double)__ldg(p.bias + co);
      if (res) vd += (double)to_f32<TOut>(res[pix * p.ldr + co]);
      float v = (float)vd;
      if (p.relu) v = fmaxf(v, 0.f);
      out[pix * p.ldo + co] = from_f32<TOut>(v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// 2x2 max pool, stride (sh, sw), padding (0, pw) with -inf semantics, NHWC, 8 channels per thread.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxpool2x2_kernel(const T* __restrict__ in, T* __restrict__ out, int n, int h_in, int w_in, int c,
                                  int sh, int sw, int pw, int h_out, int w_out) {
  const int cv = c / 8;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n * h_out * w_out * cv;
  if (gid >= total) return;
  const int cg = gid % cv;
  const long long pix = gid / cv;
  const int wo = pix % w_out;
  const int ho = (pix / w_out) % h_out;
  const int b = pix / ((long long)w_out * h_out);
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int hi = ho * sh + r, wi = wo * sw + s - pw;
      if (hi < h_in && wi >= 0 && wi < w_in) {
        const T* ip = in + (((long long)b * h_in + hi) * w_in + wi) * c + cg * 8;
        if (sizeof(T) == 2) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(ip));
          constexpr bool h16 = Is16F16<T>::value;
          m[0] = fmaxf(m[0], lo16_to_f32(v.x, h16)); m[1] = fmaxf(m[1], hi16_to_f32(v.x, h16));
          m[2] = fmaxf(m[2], lo16_to_f32(v.y, h16)); m[3] = fmaxf(m[3], hi16_to_f32(v.y, h16));
          m[4] = fmaxf(m[4], lo16_to_f32(v.z, h16)); m[5] = fmaxf(m[5], hi16_to_f32(v.z, h16));
          m[6] = fmaxf(m[6], lo16_to_f32(v.w, h16)); m[7] = fmaxf(m[7], hi16_to_f32(v.w, h16));
        } else {
          const float4 v0 = __ldg(reinterpret_cast<const float4*>(ip));
          const float4 v1 = __ldg(reinterpret_cast<const float4*>(ip) + 1);
          m[0] = fmaxf(m[0], v0.x); m[1] = fmaxf(m[1], v0.y); m[2] = fmaxf(m[2], v0.z); m[3] = fmaxf(m[3], v0.w);
          m[4] = fmaxf(m[4], v1.x); m[5] = fmaxf(m[5], v1.y); m[6] = fmaxf(m[6], v1.z); m[7] = fmaxf(m[7], v1.w);
        }
      }
    }
  T* op = out + pix * c + cg * 8;
  if (sizeof(T) == 2) {
    uint4 o;
    constexpr bool h16 = Is16F16<T>::value;
    o.x = pack16x2(m[0], m[1], h16); o.y = pack16x2(m[2], m[3], h16);
    o.z = pack16x2(m[4], m[5], h16); o.w = pack16x2(m[6], m[7], h16);
    *reinterpret_cast<uint4*>(op) = o;
  } else {
    float* fo = reinterpret_cast<float*>(op);
    *reinterpret_cast<float4*>(fo) = make_float4(m[0], m[1], m[2], m[3]);
    *reinterpret_cast<float4*>(fo + 4) = make_float4(m[4], m[5], m[6], m[7]);
  }
}

// ---------------------------------------------------------------------------------------------
// mean over H' (AdaptiveAvgPool2d((None,1)) after the permute): [B,H',T,C] -> [B,T,C]
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void mean_h_kernel(const T* __restrict__ in, T* __restrict__ out, int b, int hp, int t, int c) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)b * t * c;
  if (gid >= total) return;
  const int ch = gid % c;
  const int tt = (gid / c) % t;
  const int bb = gid / ((long long)c * t);
  float s = 0.f;
  for (int h = 0; h < hp; ++h) s += to_f32<T>(in[(((long long)bb * hp + h) * t + tt) * c + ch]);
  out[gid] = from_f32<T>(s / static_cast<float>(hp));
}

template <typename T>
__global__ void cvt_to_f32_kernel(const T* __restrict__ in, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = to_f32<T>(in[i]);
}
template <typename T>
__global__ void cvt_from_f32_kernel(const float* __restrict__ in, T* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = from_f32<T>(in[i]);
}

// ---------------------------------------------------------------------------------------------
// split-precision (fp16 hi/lo) variants: a pixel is [hi(c) | lo(c)]; hi + lo is exact in fp32.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_split8(const __half* px, int c, int ch, float (&v)[8]) {
  const uint4 h = __ldg(reinterpret_cast<const uint4*>(px + ch));
  const uint4 l = __ldg(reinterpret_cast<const uint4*>(px + c + ch));
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[j]));
    const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[j]));
    v[2 * j] = hf.x + lf.x;
    v[2 * j + 1] = hf.y + lf.y;
  }
}
__device__ __forceinline__ void store_split8(__half* px, int c, int ch, const float (&v)[8]) {
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half2 hv = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    const float2 hf = __half22float2(hv);
    const __half2 lv = __floats2half2_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
    hw[j] = *reinterpret_cast<const uint32_t*>(&hv);
    lw[j] = *reinterpret_cast<const uint32_t*>(&lv);
  }
  *reinterpret_cast<uint4*>(px + ch) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  *reinterpret_cast<uint4*>(px + c + ch) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

__global__ void maxpool2x2_split_kernel(const __half* __restrict__ in, __half* __restrict__ out, int n, int h_in, int w_in,
                                        int c, int sh, int sw, int pw, int h_out, int w_out) {
  const int cv = c / 8;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n * h_out * w_out * cv;
  if (gid >= total) return;
  const int cg = gid % cv;
  const long long pix = gid / cv;
  const int wo = pix % w_out;
  const int ho = (pix / w_out) % h_out;
  const int b = pix / ((long long)w_out * h_out);
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int hi = ho * sh + r, wi = wo * sw + s - pw;
      if (hi < h_in && wi >= 0 && wi < w_in) {
        float v[8];
        load_split8(in + (((long long)b * h_in + hi) * w_in + wi) * 2 * c, c, cg * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], v[i]);
      }
    }
  store_split8(out + pix * 2 * c, c, cg * 8, m);
}

// [B,H',T,C] split -> [B,T,C] split; the sum of H' <= 4 values and the division are fp32 like the reference's
__global__ void mean_h_split_kernel(const __half* __restrict__ in, __half* __restrict__ out, int b, int hp, int t, int c) {
  const int cv = c / 8;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)b * t * cv;
  if (gid >= total) return;
  const int cg = gid % cv;
  const int tt = (gid / cv) % t;
  const int bb = gid / ((long long)cv * t);
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  for (int h = 0; h < hp; ++h) {
    float v[8];
    load_split8(in + (((long long)bb * hp + h) * t + tt) * 2 * c, c, cg * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] += v[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = s[i] / static_cast<float>(hp);
  store_split8(out + ((long long)bb * t + tt) * 2 * c, c, cg * 8, s);
}

// rows of [hi(c) | lo(c)] fp16 <-> rows of c fp32
__global__ void cvt_split_to_f32_kernel(const __half* __restrict__ in, float* __restrict__ out, long long rows, int c) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * c) return;
  const long long r = i / c;
  const int ch = i % c;
  out[i] = __half2float(in[r * 2 * c + ch]) + __half2float(in[r * 2 * c + c + ch]);
}
__global__ void cvt_f32_to_split_kernel(const float* __restrict__ in, __half* __restrict__ out, long long rows, int c) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * c) return;
  const long long r = i / c;
  const int ch = i % c;
  const float v = in[i];
  const __half h = __float2half_rn(v);
  out[r * 2 * c + ch] = h;
  out[r * 2 * c + c + ch] = __float2half_rn(v - __half2float(h));
}

// ---------------------------------------------------------------------------------------------
// fp32-mode LSTM cell: gates = xproj[b,t,dir,:] + hh[dir][b][:] (both in the permuted [tile][gate][16] order)
// ---------------------------------------------------------------------------------------------
__global__ void lstm_cell_f32_kernel(const float* __restrict__ hh /*[2][B][2048]*/, const float* __restrict__ xproj,
                                     float* __restrict__ c /*[2][B][512]*/, float* __restrict__ hseq /*[B][T][1024]*/,
                                     int B, int T, int step) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= 2 * B * 512) return;
  const int j = gid % 512;
  const int b = (gid / 512) % B;
  const int dir = gid / (512 * B);
  const int t = dir == 0 ? step : T - 1 - step;
  const int tile = j >> 4, u = j & 15;
  const float* hp = hh + ((long long)dir * B + b) * 2048 + tile * 64 + u;
  const float* xp = xproj + ((long long)b * T + t) * 4096 + dir * 2048 + tile * 64 + u;
  const bool first = step == 0;
  const float gi = sigmoidf_acc(xp[0] + (first ? 0.f : hp[0]));
  const float gf = sigmoidf_acc(xp[16] + (first ? 0.f : hp[16]));
  const float gg = tanhf(xp[32] + (first ? 0.f : hp[32]));
  const float go = sigmoidf_acc(xp[48] + (first ? 0.f : hp[48]));
  const float cprev = first ? 0.f : c[gid];
  const float cn = gf * cprev + gi * gg;
  c[gid] = cn;
  hseq[((long long)b * T + t) * 1024 + dir * 512 + j] = go * tanhf(cn);
}

// ---------------------------------------------------------------------------------------------
// CRAFT detector glue kernels (reference: OCRFA / Vgg16BN, /root/reference/src/ocrfa/utils.py:48-98 and
// /root/reference/src/ocrfa/modules/blocks.py:27-82). 16-bit NHWC activations (bf16 or fp16, `f16` flag), 8 channels
// per thread; inputs / outputs may be channel slices of wider buffers (ld_in / ld_out = channels of the buffer), which
// is how the U-net's torch.cat never costs a copy.
// ---------------------------------------------------------------------------------------------
// x [N,3,H,W] float32 (already mean/variance normalised, detectors.py:38-40) -> NHWC with 8 channels (3 real + 5 zero)
__global__ void nchw3_to_nhwc8_kernel(const float* __restrict__ x, uint4* __restrict__ out, int n, int h, int w, int f16) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long hw = (long long)h * w;
  if (pix >= (long long)n * hw) return;
  const long long b = pix / hw, r = pix - b * hw;
  const float* xp = x + b * 3 * hw + r;
  const float c0 = __ldg(xp), c1 = __ldg(xp + hw), c2 = __ldg(xp + 2 * hw);
  uint4 o;
  o.x = pack16x2(c0, c1, f16 != 0); o.y = pack16x2(c2, 0.f, f16 != 0); o.z = 0u; o.w = 0u;
  out[pix] = o;
}

// max pool k x k (k = 1: plain copy), stride s, -inf padding p, optional ReLU on the result (max and ReLU commute, so
// "ReLU then pool" of the reference is computed as pool then ReLU). Covers VGG's 2x2/2 pools, slice5's 3x3/1 pad 1
// pool and the "ReLU of a tensor that is also needed before its ReLU" copies.
__global__ void pool16_kernel(const uint16_t* __restrict__ in, int ld_in, uint16_t* __restrict__ out, int ld_out, int n,
                              int h_in, int w_in, int c, int k, int stride, int pad, int relu, int f16) {
  const int cv = c / 8;
  const int h_out = (h_in + 2 * pad - k) / stride + 1, w_out = (w_in + 2 * pad - k) / stride + 1;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)n * h_out * w_out * cv) return;
  const int cg = gid % cv;
  const long long pix = gid / cv;
  const int wo = pix % w_out, ho = (pix / w_out) % h_out;
  const int b = pix / ((long long)w_out * h_out);
  const bool hf = f16 != 0;
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
  for (int r = 0; r < k; ++r)
    for (int s2 = 0; s2 < k; ++s2) {
      const int hi = ho * stride + r - pad, wi = wo * stride + s2 - pad;
      if (hi < 0 || hi >= h_in || wi < 0 || wi >= w_in) continue;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (((long long)b * h_in + hi) * w_in + wi) * ld_in + cg * 8));
      m[0] = fmaxf(m[0], lo16_to_f32(v.x, hf)); m[1] = fmaxf(m[1], hi16_to_f32(v.x, hf));
      m[2] = fmaxf(m[2], lo16_to_f32(v.y, hf)); m[3] = fmaxf(m[3], hi16_to_f32(v.y, hf));
      m[4] = fmaxf(m[4], lo16_to_f32(v.z, hf)); m[5] = fmaxf(m[5], hi16_to_f32(v.z, hf));
      m[6] = fmaxf(m[6], lo16_to_f32(v.w, hf)); m[7] = fmaxf(m[7], hi16_to_f32(v.w, hf));
    }
  if (relu) {
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], 0.f);
  }
  uint4 o;
  o.x = pack16x2(m[0], m[1], hf); o.y = pack16x2(m[2], m[3], hf); o.z = pack16x2(m[4], m[5], hf); o.w = pack16x2(m[6], m[7], hf);
  *reinterpret_cast<uint4*>(out + pix * ld_out + cg * 8) = o;
}

// F.interpolate(y, size = 2x, mode="bilinear", align_corners=False) (utils.py:84, 88, 92): source index
// max(0, (dst + 0.5) / 2 - 0.5), neighbours clamped at the border, fp32 blend in PyTorch's order
// h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11).
__global__ void upsample2x_bilinear16_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int ld_out, int n,
                                             int h_in, int w_in, int c, int f16) {
  const int cv = c / 8;
  const int h_out = 2 * h_in, w_out = 2 * w_in;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)n * h_out * w_out * cv) return;
  const int cg = gid % cv;
  const long long pix = gid / cv;
  const int wo = pix % w_out, ho = (pix / w_out) % h_out;
  const int b = pix / ((long long)w_out * h_out);
  const bool hf = f16 != 0;
  const float fy = fmaxf(0.f, (ho + 0.5f) * 0.5f - 0.5f), fx = fmaxf(0.f, (wo + 0.5f) * 0.5f - 0.5f);
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = min(y0 + 1, h_in - 1), x1 = min(x0 + 1, w_in - 1);
  const float ly1 = fy - y0, lx1 = fx - x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  auto ld = [&](int yy, int xx) {
    return __ldg(reinterpret_cast<const uint4*>(in + (((long long)b * h_in + yy) * w_in + xx) * c + cg * 8));
  };
  const uint4 v00 = ld(y0, x0), v01 = ld(y0, x1), v10 = ld(y1, x0), v11 = ld(y1, x1);
  const uint32_t a[4] = {v00.x, v00.y, v00.z, v00.w}, bq[4] = {v01.x, v01.y, v01.z, v01.w};
  const uint32_t cq[4] = {v10.x, v10.y, v10.z, v10.w}, dq[4] = {v11.x, v11.y, v11.z, v11.w};
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float lo = ly0 * (lx0 * lo16_to_f32(a[i], hf) + lx1 * lo16_to_f32(bq[i], hf)) +
                     ly1 * (lx0 * lo16_to_f32(cq[i], hf) + lx1 * lo16_to_f32(dq[i], hf));
    const float hi = ly0 * (lx0 * hi16_to_f32(a[i], hf) + lx1 * hi16_to_f32(bq[i], hf)) +
                     ly1 * (lx0 * hi16_to_f32(cq[i], hf) + lx1 * hi16_to_f32(dq[i], hf));
    o[i] = pack16x2(lo, hi, hf);
  }
  *reinterpret_cast<uint4*>(out + pix * ld_out + cg * 8) = make_uint4(o[0], o[1], o[2], o[3]);
}

// NHWC 16-bit [N,H,W,ld] (first c channels) -> NCHW float32 [N,c,H,W] (the `feature` output of OCRFA.forward)
__global__ void nhwc16_to_nchw_f32_kernel(const uint16_t* __restrict__ in, int ld_in, float* __restrict__ out, int n, int h,
                                          int w, int c, int f16) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long hw = (long long)h * w;
  if (gid >= (long long)n * c * hw) return;
  const long long r = gid % hw;
  const int ch = (gid / hw) % c;
  const long long b = gid / (hw * c);
  const uint32_t v = in[(b * hw + r) * ld_in + ch];
  out[gid] = f16 ? f16bits_to_f32(v) : __uint_as_float(v << 16);
}

// ---------------------------------------------------------------------------------------------
// The same glue for the fp32-grade (f16x3) CRAFT forward: activations are fp16 hi/lo pairs, pixel = [hi(ld) | lo(ld)]
// with ld = channels of the buffer (a channel slice [c0, c0 + c) of a concat buffer has its hi halves at c0 and its lo
// halves at ld + c0); values are reconstructed as hi + lo in fp32, combined, and split again.
// ---------------------------------------------------------------------------------------------
__global__ void nchw3_to_nhwc8_split_kernel(const float* __restrict__ x, __half* __restrict__ out, int n, int h, int w) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long hw = (long long)h * w;
  if (pix >= (long long)n * hw) return;
  const long long b = pix / hw, r = pix - b * hw;
  const float* xp = x + b * 3 * hw + r;
  const float v[8] = {__ldg(xp), __ldg(xp + hw), __ldg(xp + 2 * hw), 0.f, 0.f, 0.f, 0.f, 0.f};
  store_split8(out + pix * 16, 8, 0, v);
}

__global__ void pool_split_kernel(const __half* __restrict__ in, int ld_in, __half* __restrict__ out, int ld_out, int n,
                                  int h_in, int w_in, int c, int k, int stride, int pad, int relu) {
  const int cv = c / 8;
  const int h_out = (h_in + 2 * pad - k) / stride + 1, w_out = (w_in + 2 * pad - k) / stride + 1;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)n * h_out * w_out * cv) return;
  const int cg = gid % cv;
  const long long pix = gid / cv;
  const int wo = pix % w_out, ho = (pix / w_out) % h_out;
  const int b = pix / ((long long)w_out * h_out);
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
  for (int r = 0; r < k; ++r)
    for (int s2 = 0; s2 < k; ++s2) {
      const int hi = ho * stride + r - pad, wi = wo * stride + s2 - pad;
      if (hi < 0 || hi >= h_in || wi < 0 || wi >= w_in) continue;
      float v[8];
      load_split8(in + (((long long)b * h_in + hi) * w_in + wi) * 2 * ld_in, ld_in, cg * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], v[i]);
    }
  if (relu) {
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], 0.f);
  }
  store_split8(out + pix * 2 * ld_out, ld_out, cg * 8, m);
}

__global__ void upsample2x_bilinear_split_kernel(const __half* __restrict__ in, __half* __restrict__ out, int ld_out, int n,
                                                 int h_in, int w_in, int c) {
  const int cv = c / 8;
  const int h_out = 2 * h_in, w_out = 2 * w_in;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)n * h_out * w_out * cv) return;
  const int cg = gid % cv;
  const long long pix = gid / cv;
  const int wo = pix % w_out, ho = (pix / w_out) % h_out;
  const int b = pix / ((long long)w_out * h_out);
  const float fy = fmaxf(0.f, (ho + 0.5f) * 0.5f - 0.5f), fx = fmaxf(0.f, (wo + 0.5f) * 0.5f - 0.5f);
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = min(y0 + 1, h_in - 1), x1 = min(x0 + 1, w_in - 1);
  const float ly1 = fy - y0, lx1 = fx - x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  float v00[8], v01[8], v10[8], v11[8], o[8];
  load_split8(in + (((long long)b * h_in + y0) * w_in + x0) * 2 * c, c, cg * 8, v00);
  load_split8(in + (((long long)b * h_in + y0) * w_in + x1) * 2 * c, c, cg * 8, v01);
  load_split8(in + (((long long)b * h_in + y1) * w_in + x0) * 2 * c, c, cg * 8, v10);
  load_split8(in + (((long long)b * h_in + y1) * w_in + x1) * 2 * c, c, cg * 8, v11);
#pragma unroll
  for (int i = 0; i < 8; ++i)  // PyTorch's order: h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11), no contraction
    o[i] = __fadd_rn(__fmul_rn(ly0, __fadd_rn(__fmul_rn(lx0, v00[i]), __fmul_rn(lx1, v01[i]))),
                     __fmul_rn(ly1, __fadd_rn(__fmul_rn(lx0, v10[i]), __fmul_rn(lx1, v11[i]))));
  store_split8(out + pix * 2 * ld_out, ld_out, cg * 8, o);
}

__global__ void nhwc_split_to_nchw_f32_kernel(const __half* __restrict__ in, int ld_in, float* __restrict__ out, int n, int h,
                                              int w, int c) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long hw = (long long)h * w;
  if (gid >= (long long)n * c * hw) return;
  const long long r = gid % hw;
  const int ch = (gid / hw) % c;
  const long long b = gid / (hw * c);
  const __half* px = in + (b * hw + r) * 2 * ld_in;
  out[gid] = __half2float(px[ch]) + __half2float(px[ld_in + ch]);
}

}  // namespace ocrfa
