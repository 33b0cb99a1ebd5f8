// Text-detection glue around the CRAFT forward: the pixel work of the reference's detector pre- and post-processing.
//
//   det_preprocess_kernel   resize_aspect_ratio + normalize_mean_variance + HWC->CHW
//                           /root/reference/src/ocrfa/utils.py:479-484, 497-524; detectors.py:33-43
//   det_* (labeling)        get_det_boxes_core up to the point where cv2.minAreaRect takes over
//                           /root/reference/src/ocrfa/utils.py:106-160
//
// get_det_boxes_core, per connected component k of (text > low_text) | (link > link_threshold), 4-connectivity:
//   skip if area < 10 or max(textmap over k) < text_threshold; segmap = pixels of k minus (link & !text);
//   dilate segmap inside the ROI [sy:ey, sx:ex] with a (1+niter)^2 rectangle; minAreaRect of its non-zero pixels.
// The reference does this with O(components x H x W) NumPy passes on the host. Here one pass over the map labels it
// (union-find with atomicMin hooks; root = the component's first pixel in raster order, so ranking the roots gives
// cv2.connectedComponentsWithStats' label numbers), one pass accumulates cv2's stats + the per-label score maximum
// (warp-aggregated atomics), and the dilation is done on ROW EXTENTS: a rectangle dilation of a row set is
// min/max arithmetic on each row's (min x, max x), and the minimum-area rectangle of a pixel set depends only on its
// convex hull, i.e. on those row extents. The host receives, per surviving component, its stats and (ey - sy)
// dilated row extents - a few hundred bytes instead of H x W masks.
// Integer / compare work, HBM-bound on paper (12 B read + 4 B written per pixel); at CRAFT map sizes (368 x 640) it is
// launch-latency bound: 9 small launches.
#pragma once
#include <stdint.h>
#include <limits.h>
#include "../../include/ocrfa_b200.h"

namespace ocrfa {

// ------------------------------------------------------------------------------------------------------------------
// detector preprocessing: uint8 HWC page -> cv2.resize(INTER_LINEAR) -> zero canvas (multiples of 32) -> (v - mean*255)
// / (variance*255) in float32 -> float32 CHW. Same fixed-point bilinear arithmetic as crop_resize_u8_kernel, 3 channels.
// One thread per canvas pixel; reads hit L1/L2, writes are three coalesced float planes.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) det_preprocess_kernel(const uint8_t* __restrict__ img, int src_h, int src_w,
                                                             int src_stride, int dst_h, int dst_w, int H32, int W32,
                                                             const int32_t* __restrict__ xtab,
                                                             const int32_t* __restrict__ ytab, int mode, float m0, float m1,
                                                             float m2, float v0, float v1, float v2,
                                                             float* __restrict__ out) {
  const long long total = (long long)H32 * W32;
  const size_t plane = (size_t)H32 * W32;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(g / W32), x = (int)(g - (long long)y * W32);
    int c[3] = {0, 0, 0};
    if (y < dst_h && x < dst_w) {
      if (mode == 1) {  // equal sizes: cv2.resize copies
        const uint8_t* p = img + (size_t)y * src_stride + 3 * x;
        c[0] = p[0]; c[1] = p[1]; c[2] = p[2];
      } else if (mode == 2) {  // exact 2x shrink: OpenCV switches INTER_LINEAR to INTER_AREA
        const uint8_t* r0 = img + (size_t)(2 * y) * src_stride + 6 * x;
        const uint8_t* r1 = r0 + src_stride;
#pragma unroll
        for (int k = 0; k < 3; ++k) c[k] = (r0[k] + r0[k + 3] + r1[k] + r1[k + 3] + 2) >> 2;
      } else {
        const int2 ty = reinterpret_cast<const int2*>(ytab)[y];
        const int2 tx = reinterpret_cast<const int2*>(xtab)[x];
        const int ya = min(max(ty.x, 0), src_h - 1), yb = min(max(ty.x + 1, 0), src_h - 1);
        const int b0 = (int16_t)(ty.y & 0xffff), b1 = (int16_t)(ty.y >> 16);
        const int xa = tx.x, xb = min(tx.x + 1, src_w - 1);
        const int a0 = (int16_t)(tx.y & 0xffff), a1 = (int16_t)(tx.y >> 16);
        const uint8_t* r0 = img + (size_t)ya * src_stride;
        const uint8_t* r1 = img + (size_t)yb * src_stride;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int h0 = r0[3 * xa + k] * a0 + r0[3 * xb + k] * a1;
          const int h1 = r1[3 * xa + k] * a0 + r1[3 * xb + k] * a1;
          c[k] = min(max((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2, 0), 255);
        }
      }
    }
    out[g] = __fdiv_rn(__fsub_rn((float)c[0], m0), v0);
    out[plane + g] = __fdiv_rn(__fsub_rn((float)c[1], m1), v1);
    out[2 * plane + g] = __fdiv_rn(__fsub_rn((float)c[2], m2), v2);
  }
}

// reformat_input's grey page (utils.py:1545-1547: cv2.cvtColor(image, cv2.COLOR_BGR2GRAY) on a uint8 [H, W, 3] array):
// OpenCV's 15-bit fixed point, (c0 * 3735 + c1 * 19235 + c2 * 9798 + 16384) >> 15 with c0 the FIRST channel in memory.
// 3 bytes read, 1 written per pixel; one thread makes 4 pixels (12 contiguous bytes in, one 32-bit store out).
__global__ void __launch_bounds__(256) bgr2gray_u8_kernel(const uint8_t* __restrict__ img, int img_stride, int H, int W,
                                                          uint8_t* __restrict__ gray, int gray_stride) {
  const int quads = (W + 3) >> 2;
  const long long total = (long long)quads * H;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(g / quads), x0 = (int)(g - (long long)y * quads) * 4;
    const uint8_t* p = img + (size_t)y * img_stride + 3 * x0;
    uint8_t* o = gray + (size_t)y * gray_stride + x0;
    uint32_t v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      v[i] = x0 + i < W ? (uint32_t)(p[3 * i] * 3735 + p[3 * i + 1] * 19235 + p[3 * i + 2] * 9798 + (1 << 14)) >> 15 : 0u;
    if (x0 + 4 <= W && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) {
      *reinterpret_cast<uint32_t*>(o) = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
    } else {
      for (int i = 0; i < 4 && x0 + i < W; ++i) o[i] = (uint8_t)v[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// connected components
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(const int* L, int a) {
  int p;
  while ((p = __ldcg(L + a)) != a) a = p;
  return a;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  while (true) {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a == b) return;
    if (a > b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(L + b, a);  // hook the larger root under the smaller one
    if (old == b) return;
    b = old;                              // b was hooked elsewhere meanwhile: merge that tree with a instead
  }
}

__device__ __forceinline__ uint32_t float_order(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_unorder(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// cv2.threshold(map, t, 1, THRESH_BINARY) on float32 compares `v > (float)t`. flags: bit 0 text, bit 1 link.
__global__ void __launch_bounds__(256) det_init_kernel(const float* __restrict__ text, const float* __restrict__ link,
                                                       int pix_stride, int n, float low_text, float link_thr,
                                                       int* __restrict__ parent, uint8_t* __restrict__ flags) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int f = (text[(size_t)p * pix_stride] > low_text ? 1 : 0) | (link[(size_t)p * pix_stride] > link_thr ? 2 : 0);
  flags[p] = (uint8_t)f;
  parent[p] = f ? p : -1;
}

__global__ void __launch_bounds__(256) det_merge_kernel(int* __restrict__ parent, const uint8_t* __restrict__ flags, int H,
                                                        int W) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W || !flags[p]) return;
  const int y = p / W, x = p - y * W;
  const bool left = x > 0 && flags[p - 1];
  const bool up = y > 0 && flags[p - W];
  if (left) uf_union(parent, p, p - 1);
  // with left, up-left and up all set, up is already tied to this pixel through them
  if (up && !(left && flags[p - W - 1])) uf_union(parent, p, p - W);
}

// flattens the forest and counts the roots of each 1024-pixel block
__global__ void __launch_bounds__(1024) det_flatten_kernel(int* __restrict__ parent, int n, int* __restrict__ block_roots) {
  const int p = blockIdx.x * 1024 + threadIdx.x;
  int is_root = 0;
  if (p < n) {
    const int q = parent[p];
    if (q >= 0) {
      const int r = uf_find(parent, p);
      is_root = r == p;
      if (!is_root) parent[p] = r;  // roots never change here, so concurrent finds stay valid
    }
  }
  const int c = __syncthreads_count(is_root);
  if (threadIdx.x == 0) block_roots[blockIdx.x] = c;
}

// exclusive scan of the per-block root counts (one CTA; nb <= a few thousand); counts[0] = cv2's nLabels (roots + 1)
__global__ void __launch_bounds__(1024) det_scan_blocks_kernel(int* __restrict__ block_roots, int nb, int* __restrict__ counts) {
  __shared__ int warp_sum[32];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int start = 0; start < nb; start += 1024) {
    const int i = start + threadIdx.x;
    const int v = i < nb ? block_roots[i] : 0;
    int s = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, d);
      if (lane >= d) s += t;
    }
    if (lane == 31) warp_sum[warp] = s;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += t;
      }
      warp_sum[lane] = w;
    }
    __syncthreads();
    const int excl = base + (warp ? warp_sum[warp - 1] : 0) + s - v;
    if (i < nb) block_roots[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) base = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[0] = base + 1;
}

// rank[p] = 1-based label of root p = 1 + number of roots before it in raster order
__global__ void __launch_bounds__(1024) det_rank_kernel(const int* __restrict__ parent, int n,
                                                        const int* __restrict__ block_off, int* __restrict__ rank) {
  __shared__ int warp_sum[32];
  const int p = blockIdx.x * 1024 + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool is_root = p < n && parent[p] == p;
  const unsigned bal = __ballot_sync(0xffffffffu, is_root);
  if (lane == 0) warp_sum[warp] = __popc(bal);
  __syncthreads();
  if (warp == 0) {
    int w = warp_sum[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += t;
    }
    warp_sum[lane] = w;
  }
  __syncthreads();
  if (is_root) rank[p] = block_off[blockIdx.x] + (warp ? warp_sum[warp - 1] : 0) + __popc(bal & ((1u << lane) - 1)) + 1;
}

struct DetStats {  // per label, capacity cap = H*W/2 + 2 (a 4-connected checkerboard is the worst case)
  int* area;
  int* minx;
  int* maxx;
  int* miny;
  int* maxy;
  uint32_t* maxtext;
};

__global__ void __launch_bounds__(256) det_stats_init_kernel(DetStats s, int cap) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cap) return;
  s.area[k] = 0;
  s.minx[k] = INT_MAX;
  s.miny[k] = INT_MAX;
  s.maxx[k] = -1;
  s.maxy[k] = -1;
  s.maxtext[k] = 0u;  // below float_order of every float
}

// labels (cv2 numbering) + cv2's CC_STAT_* + the per-label maximum of textmap; lanes of a warp that carry the same
// label (the common case: a warp covers 32 consecutive pixels of a row) combine before touching global memory.
__global__ void __launch_bounds__(256) det_label_stats_kernel(const int* __restrict__ parent, const int* __restrict__ rank,
                                                              const float* __restrict__ text, int pix_stride, int H, int W,
                                                              int32_t* __restrict__ labels, DetStats s) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int k = 0;
  if (p < H * W) {
    const int r = parent[p];
    if (r >= 0) k = rank[r];
    if (labels) labels[p] = k;
  }
  const unsigned active = __ballot_sync(0xffffffffu, k > 0);
  if (k > 0) {
    const int y = p / W, x = p - y * W;
    const unsigned m = __match_any_sync(active, k);
    const int mnx = __reduce_min_sync(m, x), mxx = __reduce_max_sync(m, x);
    const int mny = __reduce_min_sync(m, y), mxy = __reduce_max_sync(m, y);
    const uint32_t mt = __reduce_max_sync(m, float_order(text[(size_t)p * pix_stride]));
    if (lane == __ffs(m) - 1) {
      atomicAdd(s.area + k, __popc(m));
      atomicMin(s.minx + k, mnx);
      atomicMax(s.maxx + k, mxx);
      atomicMin(s.miny + k, mny);
      atomicMax(s.maxy + k, mxy);
      atomicMax(s.maxtext + k, mt);
    }
  }
}

// Size / score filters, niter and the dilation ROI (utils.py:127-153), compacted in label order (= the order of the
// reference's `det` list). One CTA. counts: [0] nLabels, [1] kept components, [2] extent rows, [3] overflow flag.
__global__ void __launch_bounds__(1024) det_select_kernel(DetStats s, int H, int W, float text_threshold,
                                                          ocrfa_det_comp* __restrict__ comps, int max_comps, int max_rows,
                                                          int* __restrict__ slot, int* __restrict__ counts) {
  __shared__ int warp_a[32], warp_b[32];
  __shared__ int base_c, base_r;
  if (threadIdx.x == 0) { base_c = 0; base_r = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_labels = counts[0];
  for (int start = 1; start < n_labels; start += 1024) {
    const int k = start + threadIdx.x;
    int keep = 0, rows = 0;
    ocrfa_det_comp c;
    if (k < n_labels) {
      const int size = s.area[k];
      const float mt = float_unorder(s.maxtext[k]);
      if (size >= 10 && !(mt < text_threshold)) {
        const int x = s.minx[k], y = s.miny[k], w = s.maxx[k] - x + 1, h = s.maxy[k] - y + 1;
        // int(math.sqrt(size * min(w, h) / (w * h)) * 2): int32 products (NumPy scalars), float64 division and sqrt
        const int num = (int)((unsigned)size * (unsigned)min(w, h));
        const int den = (int)((unsigned)w * (unsigned)h);
        const int niter = (int)(sqrt((double)num / (double)den) * 2.0);
        int sx = x - niter, ex = x + w + niter + 1, sy = y - niter, ey = y + h + niter + 1;
        if (sx < 0) sx = 0;
        if (sy < 0) sy = 0;
        if (ex >= W) ex = W;
        if (ey >= H) ey = H;
        keep = 1;
        rows = ey - sy;
        c.label = k; c.area = size; c.left = x; c.top = y; c.width = w; c.height = h; c.niter = niter;
        c.sx = sx; c.sy = sy; c.ex = ex; c.ey = ey; c.max_text = mt;
      }
    }
    int a = keep, b = rows;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int ta = __shfl_up_sync(0xffffffffu, a, d), tb = __shfl_up_sync(0xffffffffu, b, d);
      if (lane >= d) { a += ta; b += tb; }
    }
    if (lane == 31) { warp_a[warp] = a; warp_b[warp] = b; }
    __syncthreads();
    if (warp == 0) {
      int wa = warp_a[lane], wb = warp_b[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int ta = __shfl_up_sync(0xffffffffu, wa, d), tb = __shfl_up_sync(0xffffffffu, wb, d);
        if (lane >= d) { wa += ta; wb += tb; }
      }
      warp_a[lane] = wa; warp_b[lane] = wb;
    }
    __syncthreads();
    const int ci = base_c + (warp ? warp_a[warp - 1] : 0) + a - keep;
    const int ri = base_r + (warp ? warp_b[warp - 1] : 0) + b - rows;
    if (k < n_labels) {
      int sl = -1;
      if (keep) {
        if (ci < max_comps && ri + rows <= max_rows) {
          c.row_off = ri;
          comps[ci] = c;
          sl = ci;
        } else {
          counts[3] = 1;
        }
      }
      slot[k] = sl;
    }
    __syncthreads();
    if (threadIdx.x == 1023) { base_c = ci + keep; base_r = ri + rows; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { counts[1] = min(base_c, max_comps); counts[2] = min(base_r, max_rows); }
}

__global__ void __launch_bounds__(256) det_rows_init_kernel(int2* __restrict__ src_ext, const int* __restrict__ counts) {
  const int n = counts[2];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) src_ext[i] = make_int2(INT_MAX, -1);
}

// row extents of segmap = (labels == k) minus the link-only pixels (utils.py:137-146), before the dilation
__global__ void __launch_bounds__(256) det_row_extents_kernel(const int* __restrict__ parent, const int* __restrict__ rank,
                                                              const uint8_t* __restrict__ flags, int H, int W,
                                                              const int* __restrict__ slot,
                                                              const ocrfa_det_comp* __restrict__ comps,
                                                              int2* __restrict__ src_ext) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int row = -1, x = 0;
  if (p < H * W) {
    const int r = parent[p];
    if (r >= 0 && flags[p] != 2) {  // flags == 2: link_score == 1 and text_score == 0
      const int sl = slot[rank[r]];
      if (sl >= 0) {
        const int y = p / W;
        x = p - y * W;
        row = comps[sl].row_off + (y - comps[sl].sy);
      }
    }
  }
  const unsigned active = __ballot_sync(0xffffffffu, row >= 0);
  if (row >= 0) {
    const unsigned m = __match_any_sync(active, row);
    const int mn = __reduce_min_sync(m, x), mx = __reduce_max_sync(m, x);
    if (lane == __ffs(m) - 1) {
      atomicMin(&src_ext[row].x, mn);
      atomicMax(&src_ext[row].y, mx);
    }
  }
}

// cv2.dilate(segmap[sy:ey, sx:ex], ones((1+niter, 1+niter))) on row extents: anchor a = (1+niter)/2, so a destination
// pixel is set when a source pixel lies in [x - a, x + niter - a] x [y - a, y + niter - a]; pixels outside the ROI are
// ignored (constant border, isolated ROI) and the result is clipped to the ROI.
__global__ void __launch_bounds__(128) det_dilate_rows_kernel(const ocrfa_det_comp* __restrict__ comps,
                                                              const int* __restrict__ counts,
                                                              const int2* __restrict__ src_ext, int2* __restrict__ out_ext) {
  const int n = counts[1];
  for (int ci = blockIdx.x; ci < n; ci += gridDim.x) {
    const ocrfa_det_comp c = comps[ci];
    const int rows = c.ey - c.sy, a = (1 + c.niter) / 2;
    for (int r = threadIdx.x; r < rows; r += blockDim.x) {
      const int lo = max(r - a, 0), hi = min(r + c.niter - a, rows - 1);
      int mn = INT_MAX, mx = -1;
      for (int q = lo; q <= hi; ++q) {
        const int2 e = src_ext[c.row_off + q];
        mn = min(mn, e.x);
        mx = max(mx, e.y);
      }
      if (mx >= 0) {
        mn = max(mn - (c.niter - a), c.sx);
        mx = min(mx + a, c.ex - 1);
      }
      out_ext[c.row_off + r] = make_int2(mn, mx);
    }
  }
}

}  // namespace ocrfa
