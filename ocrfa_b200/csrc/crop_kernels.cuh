// Crop + resize of text boxes out of a page image: the step that feeds the recognizer
//   reference: get_image_list / compute_ratio_and_resize, /root/reference/src/ocrfa/utils.py:1361-1418
// The reference calls cv2.resize(..., interpolation=Image.Resampling.LANCZOS): the PIL enum has value 1, which cv2
// reads as INTER_LINEAR, so this is OpenCV's 8-bit bilinear resize, reproduced bit for bit:
//   horizontal  H = S[x0] * a0 + S[x1] * a1            (a = round(coef * 2048) as int16, coef in float32)
//   vertical    D = (((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2
// with OpenCV's special cases: equal sizes copy, an exact 2x shrink in both axes is INTER_AREA ((s00+s01+s10+s11+2)>>2).
// The per-axis tables (source index, two fixed-point weights) depend only on (source size, destination size) and
// come from the host (ocrfa_b200/crops.py restates OpenCV's float32 coefficient arithmetic).
// Memory bound: <= 4 source bytes read (L1/L2 hits) and 1 byte written per output pixel.
#pragma once
#include <stdint.h>
#include "../../include/ocrfa_b200.h"

namespace ocrfa {

enum : int { kCropLinear = 0, kCropCopy = 1, kCropArea2 = 2 };

// one thread = 4 consecutive output pixels of one row (one 32-bit store when the line start is 4-byte aligned)
__global__ void __launch_bounds__(256) crop_resize_u8_kernel(const uint8_t* __restrict__ page,
                                                             const ocrfa_crop_desc* __restrict__ desc,
                                                             const int32_t* __restrict__ tables, int n,
                                                             uint8_t* __restrict__ out) {
  const int b = blockIdx.y;
  if (b >= n) return;
  const ocrfa_crop_desc d = desc[b];
  const int quads = (d.dst_w + 3) >> 2;
  const long long total = (long long)quads * d.dst_h;
  const uint8_t* src = page + d.src_off;
  const int page_stride = d.src_stride;
  uint8_t* dst = out + d.out_off;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const int dy = (int)(g / quads);
    const int dx0 = (int)(g - (long long)dy * quads) * 4;
    uint32_t v[4];
    if (d.mode == kCropCopy) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = dx0 + i < d.dst_w ? src[(size_t)dy * page_stride + dx0 + i] : 0u;
    } else if (d.mode == kCropArea2) {
      const uint8_t* r0 = src + (size_t)(2 * dy) * page_stride;
      const uint8_t* r1 = r0 + page_stride;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int x = 2 * (dx0 + i);
        v[i] = dx0 + i < d.dst_w ? ((uint32_t)r0[x] + r0[x + 1] + r1[x] + r1[x + 1] + 2u) >> 2 : 0u;
      }
    } else {
      // tables: x: [dst_w] {x0, a0 | a1 << 16}, y: [dst_h] {y0 (unclamped), b0 | b1 << 16}
      const int2 ty = reinterpret_cast<const int2*>(tables + d.ytab_off)[dy];
      const int ya = min(max(ty.x, 0), d.src_h - 1), yb = min(max(ty.x + 1, 0), d.src_h - 1);
      const int b0 = (int16_t)(ty.y & 0xffff), b1 = (int16_t)(ty.y >> 16);
      const uint8_t* r0 = src + (size_t)ya * page_stride;
      const uint8_t* r1 = src + (size_t)yb * page_stride;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[i] = 0u;
        if (dx0 + i < d.dst_w) {
          const int2 tx = reinterpret_cast<const int2*>(tables + d.xtab_off)[dx0 + i];
          const int xa = tx.x, xb = min(tx.x + 1, d.src_w - 1);
          const int a0 = (int16_t)(tx.y & 0xffff), a1 = (int16_t)(tx.y >> 16);
          const int h0 = r0[xa] * a0 + r0[xb] * a1;
          const int h1 = r1[xa] * a0 + r1[xb] * a1;
          const int val = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
          v[i] = (uint32_t)min(max(val, 0), 255);
        }
      }
    }
    uint8_t* op = dst + (size_t)dy * d.dst_w + dx0;
    if (dx0 + 4 <= d.dst_w && ((reinterpret_cast<uintptr_t>(op) & 3) == 0)) {
      *reinterpret_cast<uint32_t*>(op) = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
    } else {
      for (int i = 0; i < 4 && dx0 + i < d.dst_w; ++i) op[i] = (uint8_t)v[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// four_point_transform (utils.py:1218-1231) = cv2.warpPerspective(image, M, (mw, mh)): INTER_LINEAR, BORDER_CONSTANT 0.
// OpenCV's arithmetic, restated: the inverse map is evaluated in float64 per 64-column block (X0 = M0*x + M1*y + M2 at
// the block's first column, then (X0 + M0*x1) * (32 / (W0 + M6*x1)) per pixel, no fused multiply-adds), rounded to
// nearest-even into 1/32-pixel fixed point; the four neighbours are blended with the integer weights
// (32-fy)(32-fx)*32 ... (sum 32768) and (acc + 16384) >> 15. Pixels outside the page read 0.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) warp_perspective_u8_kernel(const uint8_t* __restrict__ page, int page_h, int page_w,
                                                                  int page_stride, const ocrfa_warp_desc* __restrict__ desc,
                                                                  int n, uint8_t* __restrict__ out) {
  const int b = blockIdx.y;
  if (b >= n) return;
  const ocrfa_warp_desc d = desc[b];
  const long long total = (long long)d.dst_w * d.dst_h;
  uint8_t* dst = out + d.out_off;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(g / d.dst_w);
    const int x = (int)(g - (long long)y * d.dst_w);
    const int xb = (x / d.block_w) * d.block_w;  // OpenCV evaluates the row start at the block's first column
    const double x1 = (double)(x - xb), xs = (double)xb, ys = (double)y;
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(d.m[0], xs), __dmul_rn(d.m[1], ys)), d.m[2]);
    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(d.m[3], xs), __dmul_rn(d.m[4], ys)), d.m[5]);
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(d.m[6], xs), __dmul_rn(d.m[7], ys)), d.m[8]);
    double W = __dadd_rn(W0, __dmul_rn(d.m[6], x1));
    W = W != 0.0 ? __ddiv_rn(32.0, W) : 0.0;
    double fX = __dmul_rn(__dadd_rn(X0, __dmul_rn(d.m[0], x1)), W);
    double fY = __dmul_rn(__dadd_rn(Y0, __dmul_rn(d.m[3], x1)), W);
    fX = fmax(-2147483648.0, fmin(2147483647.0, fX));
    fY = fmax(-2147483648.0, fmin(2147483647.0, fY));
    const int X = __double2int_rn(fX), Y = __double2int_rn(fY);
    const int sx = min(max(X >> 5, -32768), 32767), sy = min(max(Y >> 5, -32768), 32767);
    const int ax = X & 31, ay = Y & 31;
    const int w00 = (32 - ay) * (32 - ax) * 32, w01 = (32 - ay) * ax * 32, w10 = ay * (32 - ax) * 32, w11 = ay * ax * 32;
    auto px = [&](int yy, int xx) -> int {
      return (yy >= 0 && yy < page_h && xx >= 0 && xx < page_w) ? (int)page[(size_t)yy * page_stride + xx] : 0;
    };
    const int acc = px(sy, sx) * w00 + px(sy, sx + 1) * w01 + px(sy + 1, sx) * w10 + px(sy + 1, sx + 1) * w11;
    dst[g] = (uint8_t)min(max((acc + (1 << 14)) >> 15, 0), 255);
  }
}

// make_rotated_img_list (utils.py:1599-1612) for right angles: scipy.ndimage.rotate(img, 90 k, reshape=True) on a uint8
// image equals np.rot90(img, k) bit for bit (checked against scipy in tests/test_reader.py), i.e. a pure permutation:
//   k = 1: out[i][j] = in[j][w-1-i]   (h x w -> w x h)      k = 2: out[i][j] = in[h-1-i][w-1-j]
//   k = 3: out[i][j] = in[h-1-j][i]   (h x w -> w x h)
// 32 x 32 tiles through shared memory so that both the reads and the writes are row-contiguous.
__global__ void __launch_bounds__(256) rot90_u8_kernel(const uint8_t* __restrict__ src, int h, int w, int k,
                                                       uint8_t* __restrict__ dst) {
  __shared__ uint8_t tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  for (int r = ty; r < 32; r += 8) {
    const int y = y0 + r, x = x0 + tx;
    tile[r][tx] = (y < h && x < w) ? src[(size_t)y * w + x] : (uint8_t)0;
  }
  __syncthreads();
  if (k == 2) {
    for (int r = ty; r < 32; r += 8) {
      const int y = y0 + r, x = x0 + tx;
      if (y < h && x < w) dst[(size_t)(h - 1 - y) * w + (w - 1 - x)] = tile[r][tx];
    }
    return;
  }
  // transposed tile: thread (tx, r) writes the element read as tile[tx][r] = in[y0 + tx][x0 + r]
  for (int r = ty; r < 32; r += 8) {
    const int y = y0 + tx, x = x0 + r;
    if (y < h && x < w) {
      const size_t o = k == 1 ? (size_t)(w - 1 - x) * h + y : (size_t)x * h + (h - 1 - y);
      dst[o] = tile[tx][r];
    }
  }
}

}  // namespace ocrfa
