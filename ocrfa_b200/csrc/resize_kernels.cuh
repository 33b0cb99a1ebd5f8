// AlignCollate on the GPU for lines that are NOT yet at model height:
//   contrast stretch   /root/reference/src/ocrfa/utils.py:450-463 (np.percentile 10/90, linear interpolation)
//   PIL BICUBIC resize /root/reference/src/ocrfa/utils.py:592-598 (Pillow Resample.c, 8-bit fixed-point, 2 passes)
// Bit-exact integer/byte work: the per-position fixed-point coefficient tables are computed on the host exactly as
// Pillow's precompute_coeffs + normalize_coeffs_8bpc do (ocrfa_b200/resize.py); the kernels only gather, multiply-
// accumulate in int32 and clip - and build the per-line contrast LUT from a 256-bin histogram in float64.
#pragma once
#include "common.cuh"

namespace ocrfa {

struct LineDesc {
  long long src_off;  // byte offset of the uint8 [src_h][src_w] line in the packed source buffer
  long long tmp_off;  // byte offset of the horizontal-pass intermediate [src_h][out_w]
  int src_h, src_w, out_w;
  int kh_off, kh_size, bh_off;  // horizontal tables inside `tables`: coef int32 [out_w][kh_size], bounds int32 [out_w][2]
  int kv_off, kv_size, bv_off;  // vertical tables: coef [img_h][kv_size], bounds [img_h][2]
  int contrast;                 // 1 = run adjust_contrast_grey on this line first
};

constexpr int kResizePrecisionBits = 32 - 8 - 2;

// np.percentile(img, q) of uint8 data from its histogram (method "linear"): virtual index (n-1)*q/100,
// value = lerp(sorted[floor], sorted[floor+1], frac) with numpy's two-sided lerp.
__device__ inline double percentile_from_hist(const unsigned int* cum /*inclusive cumulative counts*/, long long n,
                                              double quantile) {
  const double virt = static_cast<double>(n - 1) * quantile;
  const double fl = floor(virt);
  long long prev = static_cast<long long>(fl);
  long long next = prev + 1;
  if (prev < 0) prev = 0;
  if (prev > n - 1) prev = n - 1;
  if (next > n - 1) next = n - 1;
  const double t = virt - fl;
  int a = 0, b = 0;
  for (int v = 0; v < 256; ++v) if (static_cast<long long>(cum[v]) > prev) { a = v; break; }
  for (int v = 0; v < 256; ++v) if (static_cast<long long>(cum[v]) > next) { b = v; break; }
  const double diff = static_cast<double>(b - a);
  return t >= 0.5 ? static_cast<double>(b) - diff * (1.0 - t) : static_cast<double>(a) + diff * t;
}

// One CTA per line: LUT[256] = identity, or the contrast stretch when (high-low)/max(10, high+low) < target.
__global__ void __launch_bounds__(256) contrast_lut_kernel(const uint8_t* __restrict__ src, const LineDesc* __restrict__ desc,
                                                           double target, uint8_t* __restrict__ lut) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int cum[256];
  __shared__ double s_low, s_ratio;
  __shared__ int s_apply;
  const LineDesc d = desc[blockIdx.x];
  const int tid = threadIdx.x;
  uint8_t* out = lut + static_cast<size_t>(blockIdx.x) * 256;
  if (!d.contrast) { out[tid] = static_cast<uint8_t>(tid); return; }
  hist[tid] = 0;
  __syncthreads();
  const long long n = static_cast<long long>(d.src_h) * d.src_w;
  const uint8_t* p = src + d.src_off;
  for (long long i = tid; i < n; i += 256) atomicAdd(&hist[p[i]], 1u);
  __syncthreads();
  if (tid == 0) {
    unsigned int run = 0;
    for (int v = 0; v < 256; ++v) { run += hist[v]; cum[v] = run; }
    const double high = percentile_from_hist(cum, n, 90.0 / 100.0);
    const double low = percentile_from_hist(cum, n, 10.0 / 100.0);
    const double contrast = (high - low) / fmax(10.0, high + low);
    s_apply = contrast < target ? 1 : 0;
    s_low = low;
    s_ratio = 200.0 / fmax(10.0, high - low);
  }
  __syncthreads();
  if (s_apply) {
    double v = (static_cast<double>(tid) - s_low + 25.0) * s_ratio;  // (img - low + 25) * ratio in float64
    v = fmax(0.0, fmin(255.0, v));
    out[tid] = static_cast<uint8_t>(v);  // astype(np.uint8): truncation
  } else {
    out[tid] = static_cast<uint8_t>(tid);
  }
}

__device__ __forceinline__ uint8_t clip8_fixed(int acc) {
  const int v = acc >> kResizePrecisionBits;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass (+ contrast LUT): tmp[line][y][xx] for y < src_h, xx < out_w. grid.y = line.
__global__ void resize_h_kernel(const uint8_t* __restrict__ src, const LineDesc* __restrict__ desc,
                                const int32_t* __restrict__ tables, const uint8_t* __restrict__ lut,
                                uint8_t* __restrict__ tmp) {
  const LineDesc d = desc[blockIdx.y];
  const long long total = static_cast<long long>(d.src_h) * d.out_w;
  const uint8_t* l = lut + static_cast<size_t>(blockIdx.y) * 256;
  const uint8_t* s = src + d.src_off;
  uint8_t* o = tmp + d.tmp_off;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xx = static_cast<int>(i % d.out_w);
    const int y = static_cast<int>(i / d.out_w);
    const uint8_t* row = s + static_cast<long long>(y) * d.src_w;
    if (d.kh_size == 0) { o[i] = l[row[xx]]; continue; }  // width unchanged: copy through the LUT
    const int xmin = tables[d.bh_off + xx * 2], n = tables[d.bh_off + xx * 2 + 1];
    const int32_t* k = tables + d.kh_off + xx * d.kh_size;
    int acc = 1 << (kResizePrecisionBits - 1);
    for (int x = 0; x < n; ++x) acc += static_cast<int>(l[row[xmin + x]]) * k[x];
    o[i] = clip8_fixed(acc);
  }
}

// vertical pass: out[line][yy][xx] for yy < img_h, xx < out_w (row stride img_w); columns >= out_w are zeroed.
__global__ void resize_v_kernel(const uint8_t* __restrict__ tmp, const LineDesc* __restrict__ desc,
                                const int32_t* __restrict__ tables, uint8_t* __restrict__ out, int img_h, int img_w) {
  const LineDesc d = desc[blockIdx.y];
  const long long total = static_cast<long long>(img_h) * img_w;
  const uint8_t* s = tmp + d.tmp_off;
  uint8_t* o = out + static_cast<size_t>(blockIdx.y) * img_h * img_w;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xx = static_cast<int>(i % img_w);
    const int yy = static_cast<int>(i / img_w);
    if (xx >= d.out_w) { o[i] = 0; continue; }
    if (d.kv_size == 0) { o[i] = s[static_cast<long long>(yy) * d.out_w + xx]; continue; }  // height unchanged
    const int ymin = tables[d.bv_off + yy * 2], n = tables[d.bv_off + yy * 2 + 1];
    const int32_t* k = tables + d.kv_off + yy * d.kv_size;
    int acc = 1 << (kResizePrecisionBits - 1);
    for (int y = 0; y < n; ++y) acc += static_cast<int>(s[static_cast<long long>(ymin + y) * d.out_w + xx]) * k[y];
    o[i] = clip8_fixed(acc);
  }
}

}  // namespace ocrfa
